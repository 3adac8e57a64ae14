"""Times the voxel-loss kernels (csrc/losses.cu) on the BASELINE grid (200x200x16,
18 classes) and, for scale, the reference formulation in plain torch on the same GPU.

    python tools/loss_probe.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from preworld_b200 import losses, ops
from oracle import loss_ref           # (checker, for the torch-on-GPU comparison only)


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


def main():
    pred, target, cam, cw = loss_ref.seeded_case(5, shape=(1, 18, 200, 200, 16))
    cwz = torch.cat([cw, torch.zeros(1)]).cuda()
    rows = pred.cuda().permute(0, 2, 3, 4, 1).reshape(-1, 18).contiguous()
    t = target.reshape(-1).to(torch.uint8).cuda()
    c = cam.reshape(-1).to(torch.uint8).cuda()
    v = rows.shape[0]
    stats, _ = ops.voxel_loss_stats(rows, t, c, cwz, 17)
    us_s = timed(lambda: ops.voxel_loss_stats(rows, t, c, cwz, 17))
    us_g = timed(lambda: ops.voxel_loss_grad(rows, t, c, cwz, 17, stats, 1., 1., 1.))
    by_s = v * (18 * 4 + 2)
    by_g = v * (18 * 8 + 2)
    print(f'stats {us_s:.1f} us  {by_s / us_s / 1e3:.0f} GB/s   grad {us_g:.1f} us  '
          f'{by_g / us_g / 1e3:.0f} GB/s')
    us_l = timed(lambda: ops.lovasz_softmax_rows(rows, True, t, c, 17), reps=10)
    print(f'lovasz (keys + radix sort + per-class scan, with gradient) {us_l:.0f} us')
    pd, td, cd = pred.cuda(), target.cuda(), cam.cuda()
    us_lt = timed(lambda: loss_ref.lovasz_softmax(torch.softmax(pd, 1), td, 17, cd), reps=3)
    print(f'lovasz, reference formulation in torch on the same GPU (forward only): {us_lt:.0f} us')
    us_t = timed(lambda: (loss_ref.ce_ssc_loss(pd, td, cwz, 255),
                          loss_ref.sem_scal_loss(pd, td, 255, cd),
                          loss_ref.geo_scal_loss(pd, td, 255, 17, cd)), reps=3)
    print(f'reference formulation in torch on the same GPU (forward only): {us_t:.0f} us')


if __name__ == '__main__':
    main()
