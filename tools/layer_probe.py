"""One conv layer of the path exactly as the step launches it (channels-last fp32,
folded-BN affine, optional residual + ReLU), timed with the L2 flushed between
launches -- the target of `ncu --set full --import-source on -k regex:conv_halo`.

    python tools/layer_probe.py <name> [reps]

names: see LAYERS (dims, n, cin, cout, spatial, k, stride, pad, residual)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from preworld_b200 import ops

LAYERS = {
    'v32': (3, 1, 32, 32, (16, 200, 200), 3, 1, 1, True),       # BasicBlock3D conv2 + shortcut sum
    'v32n': (3, 1, 32, 32, (16, 200, 200), 3, 1, 1, False),
    'v3264': (3, 1, 32, 64, (16, 200, 200), 3, 1, 1, False),
    'v64': (3, 1, 64, 64, (16, 200, 200), 3, 1, 1, False),
    'p64_256': (2, 18, 64, 256, (64, 176), 1, 1, 0, True),      # ResNet layer1 conv3 + identity
    'p128_512': (2, 12, 128, 512, (32, 88), 1, 1, 0, True),
    'p256_1024': (2, 12, 256, 1024, (16, 44), 1, 1, 0, True),
    'p256_64': (2, 18, 256, 64, (64, 176), 1, 1, 0, False),
    'p1024_256': (2, 12, 1024, 256, (16, 44), 1, 1, 0, False),
    'i256': (2, 12, 256, 256, (16, 44), 3, 1, 1, False),
    'i64': (2, 18, 64, 64, (64, 176), 3, 1, 1, False),
    'v128': (3, 1, 128, 128, (4, 50, 50), 3, 1, 1, True),
    'v64h': (3, 1, 64, 64, (8, 100, 100), 3, 1, 1, True),
}


def main():
    name = sys.argv[1]
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    dims, n, cin, cout, sp, k, stride, pad, with_res = LAYERS[name]
    dev = torch.device('cuda', 0)
    g = torch.Generator().manual_seed(1)
    w = torch.randn(cout, cin, *([k] * dims), generator=g) / (cin * k ** dims) ** .5
    scale = torch.rand(cout, generator=g) + 0.5
    bias = torch.randn(cout, generator=g) * 0.1
    bn = (scale.to(dev), bias.to(dev), torch.zeros(cout, device=dev), torch.ones(cout, device=dev), 0.0)
    pc = ops.PackedConv(w.to(dev), None, bn, stride=stride, padding=pad)
    x = torch.randn(n, *sp, cin, generator=g).relu().to(dev)
    osp = tuple((s + 2 * pad - k) // stride + 1 for s in sp)
    res = torch.randn(n, *osp, cout, generator=g).to(dev) if with_res else None
    out = ops.conv(x, pc, act='relu', residual=res)
    torch.cuda.synchronize()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.conv(x, pc, act='relu', residual=res, out=out)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    us = ts[len(ts) // 2]
    px = out[..., 0].numel()
    flops = 2.0 * cin * k ** dims * cout * px
    bytes_ = 4.0 * px * (cin / stride ** dims + cout * (2 if with_res else 1))
    print(f'{name}: {us:.1f} us  {flops / us / 1e6:.1f} TFLOP/s  {bytes_ / us / 1e3:.0f} GB/s '
          f'(min {ts[0]:.1f} max {ts[-1]:.1f})', flush=True)


if __name__ == '__main__':
    main()
