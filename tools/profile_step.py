"""ncu driver: warm up, then run ONE region between cudaProfilerStart/Stop
(use `ncu --profile-from-start off`).

    --part step   one full forward step of the bench workload (launch list)
    --part hot    the hot kernels once each on full-size tensors: the
                  pre_process_net BasicBlock3D (fused conv1|shortcut + conv2),
                  one lift, one cost volume, one ResNet layer2 bottleneck
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402
from preworld_b200 import ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--part', default='step', choices=['step', 'hot'])
    args = ap.parse_args()
    dev = torch.device('cuda', 0)
    wl = bench.Workload('finetune', dev, n_variants=2).to_device()
    model, dev_samples = wl.model, wl.dev_samples
    step = wl.step_resident

    for i in range(3):
        step(i)
    torch.cuda.synchronize()

    if args.part == 'step':
        torch.cuda.profiler.start()
        step(0)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return

    # ---- hot kernels on full-size tensors ---------------------------------
    g = torch.Generator().manual_seed(0)
    vol = torch.randn(1, 16, 200, 200, 32, generator=g).to(dev)
    pre = model.pre_process_net
    imgs, s2e, e2g, intr, pr, pt, bda = dev_samples[0]
    pi = model.prepare_inputs(dev_samples[0], stereo=True)
    vt = model.img_view_transformer
    depth = torch.rand(6, 88, 16, 44, generator=g).softmax(1).to(dev)
    feat = torch.randn(6, 16, 44, 32, generator=g).to(dev)
    cam = ops.lift_camera_params(pi[1][0], pi[3][0], pi[4][0], pi[5][0])
    xs, ys, ds = vt._frustum_axes(vt.frustum, dev)
    grid = tuple(int(v) for v in vt.grid_size)
    curr = torch.relu(torch.randn(6, 64, 176, 256, generator=g)).to(dev)
    prev = torch.relu(torch.randn(6, 64, 176, 256, generator=g)).to(dev)
    cvcam = ops.cv_camera_params(pi[7][0], pi[3][0], pi[4][0], pi[5][0])
    cxs, cys, cds = vt._frustum_axes(vt.cv_frustum, dev)
    x2 = torch.randn(12, 32, 88, 512, generator=g).to(dev)
    bb = model.img_backbone
    blk = bb.packs()['layers'][1][1]

    def hot():
        with torch.no_grad():
            pre(ops.to_logical(vol))
            ops.lift_fused(depth, feat, cam, bda.reshape(1, 9).contiguous(),
                           xs, ys, ds, vt.grid_lower_bound.tolist(),
                           vt.grid_interval.tolist(), 1, 6, grid)
            ops.cost_volume(curr, prev, cvcam, cxs, cys, cds, 5.0, (256, 704))
            type(bb.layer2[1]).run(blk, x2)

    hot()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    hot()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


if __name__ == '__main__':
    main()
