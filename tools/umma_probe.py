"""Quick GPU probe of the tensor-core conv paths (conv_halo.cu, conv_umma.cu)
vs the fp32 SIMT kernel and a CPU fp32 reference, with per-launch timing
(run under `timeout -s KILL`: a pipeline bug shows up as a hang).

    python tools/umma_probe.py [small|full|all]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.nn.functional as F
from preworld_b200 import ops

SMALL = [
    # dims, n, cin, cout, spatial, k, stride, pad, dil
    (2, 1, 32, 32, (8, 16), 1, 1, 0, 1),
    (2, 2, 64, 64, (16, 24), 1, 1, 0, 1),
    (2, 2, 64, 256, (17, 23), 1, 1, 0, 1),
    (2, 2, 64, 64, (17, 23), 3, 1, 1, 1),
    (2, 2, 128, 96, (16, 44), 3, 1, 6, 6),
    (2, 2, 256, 512, (17, 23), 1, 2, 0, 1),
    (2, 2, 128, 128, (16, 20), 3, 2, 1, 1),
    (3, 1, 32, 32, (8, 20, 24), 3, 1, 1, 1),
    (3, 1, 64, 64, (8, 20, 20), 3, 2, 1, 1),
    (3, 1, 64, 48, (7, 21, 19), 3, 1, 1, 1),
    (3, 1, 224, 32, (4, 8, 8), 1, 1, 0, 1),
    (3, 1, 32, 16, (6, 10, 12), 3, 1, 1, 1),
    (3, 2, 128, 128, (4, 10, 10), 3, 1, 1, 1),
    (2, 2, 64, 32, (19, 45), 3, 1, 2, 2),
    (3, 1, 32, 16, (5, 9, 37), 3, 1, 1, 1),
    (3, 1, 64, 80, (3, 11, 70), 3, 1, 1, 1),
]
FULL = [
    (3, 1, 32, 32, (16, 200, 200), 3, 1, 1, 1),
    (3, 1, 32, 64, (16, 200, 200), 3, 1, 1, 1),
    (3, 1, 64, 64, (16, 200, 200), 3, 1, 1, 1),
    (3, 1, 32, 16, (16, 200, 200), 3, 1, 1, 1),
    (3, 1, 64, 64, (8, 100, 100), 3, 1, 1, 1),
    (3, 1, 128, 128, (4, 50, 50), 3, 1, 1, 1),
    (3, 1, 224, 32, (16, 200, 200), 1, 1, 0, 1),
    (2, 6, 256, 256, (16, 44), 3, 1, 1, 1),
    (2, 6, 64, 256, (64, 176), 1, 1, 0, 1),
    (2, 6, 64, 64, (64, 176), 3, 1, 1, 1),
    (2, 6, 128, 512, (32, 88), 1, 1, 0, 1),
    (2, 6, 256, 1024, (16, 44), 1, 1, 0, 1),
    (2, 6, 1024, 256, (16, 44), 1, 1, 0, 1),
    (2, 6, 512, 512, (8, 22), 3, 1, 1, 1),
    (2, 6, 128, 128, (32, 88), 3, 1, 1, 1),
    (2, 6, 256, 64, (64, 176), 1, 1, 0, 1),
]


def run(case, check=True, reps=5, only=None):
    dims, n, cin, cout, sp, k, stride, pad, dil = case
    g = torch.Generator().manual_seed(1)
    x = torch.randn(n, cin, *sp, generator=g)
    w = torch.randn(cout, cin, *([k] * dims), generator=g) / (cin * k ** dims) ** .5
    conv = F.conv2d if dims == 2 else F.conv3d
    want = conv(x, w, None, stride, pad, dil) if check else None
    pc = ops.PackedConv(w.cuda(), None, None, stride=stride, padding=pad,
                        dilation=dil)
    perm = (0, *range(2, 2 + dims), 1)
    x_cl = x.permute(*perm).contiguous().cuda()
    flops = 2.0 * cin * k ** dims * cout
    res = {}
    for name, umma, halo, fold in (('simt', False, False, False),
                                   ('umma', True, False, False),
                                   ('halo', True, True, False),
                                   ('fold', True, True, True)):
        if only and name not in only:
            continue
        ops.USE_UMMA, ops.USE_HALO, ops.USE_FOLD = umma, halo, fold
        got = ops.conv(x_cl, pc)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            ops.conv(x_cl, pc, out=got)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / reps
        tf = flops * got[..., 0].numel() / us / 1e6
        if check:
            err = ((ops.to_logical(got).cpu() - want).abs().max().item()
                   / want.abs().max().item())
            res[name] = f'{err:.1e} {us:7.1f}us {tf:6.1f}TF'
        else:
            res[name] = f'{us:7.1f}us {tf:6.1f}TF'
    print(case, res, flush=True)


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else 'small'
    torch.manual_seed(0)
    if which == 'one':                       # ncu target: one FULL case
        run(FULL[int(sys.argv[2])], check=False, reps=1, only=('fold',))
        return
    if which in ('small', 'all'):
        for case in SMALL:
            run(case)
    if which in ('full', 'all'):
        for case in FULL:
            run(case, check=len(sys.argv) > 2, reps=10, only=('halo', 'fold'))


if __name__ == '__main__':
    main()
