"""Quick GPU probe of the tcgen05 conv path vs the fp32 SIMT kernel and a CPU
fp32 reference (run under `timeout -s KILL`: a pipeline bug shows up as a hang)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.nn.functional as F
from preworld_b200 import ops

CASES = [
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
    (3, 1, 224, 32, (4, 8, 8), 1, 1, 0, 1),
    (3, 1, 32, 16, (6, 10, 12), 3, 1, 1, 1),
]


def main():
    torch.manual_seed(0)
    for case in CASES:
        dims, n, cin, cout, sp, k, stride, pad, dil = case
        x = torch.randn(n, cin, *sp)
        w = torch.randn(cout, cin, *([k] * dims)) / (cin * k ** dims) ** .5
        conv = F.conv2d if dims == 2 else F.conv3d
        want = conv(x, w, None, stride, pad, dil)
        pc = ops.PackedConv(w.cuda(), None, None, stride=stride, padding=pad,
                            dilation=dil)
        perm = (0, *range(2, 2 + dims), 1)
        x_cl = x.permute(*perm).contiguous().cuda()
        res = {}
        for name, flag in (('simt', False), ('umma', True)):
            ops.USE_UMMA = flag
            t0 = time.time()
            got = ops.to_logical(ops.conv(x_cl, pc)).cpu()
            torch.cuda.synchronize()
            err = (got - want).abs().max().item() / want.abs().max().item()
            res[name] = err
        print(case, {k: f'{v:.2e}' for k, v in res.items()}, flush=True)


if __name__ == '__main__':
    main()
