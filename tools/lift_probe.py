"""Times pw_lift_fused (binned two-pass lift + fallback launch) and the
accelerate=True split at the BASELINE config's size with the synthetic rig's
geometry.  Under `ncu --metrics gpu__time_duration.sum` the three launches of one
call show up separately (lift_bin_kernel, lift_pool_bins_kernel, the early-exit
lift_fused_kernel)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from preworld_b200 import build_model, model_cfg, ops
from preworld_b200 import synthetic as S

dev = torch.device('cuda', 0)
model = build_model(model_cfg('finetune', 'r50', (256, 704))).eval()
s0 = tuple(t.to(dev) for t in S.make_img_inputs(1, (256, 704), seed=0))
pi = model.prepare_inputs(s0, stereo=True)
vt = model.img_view_transformer
g = torch.Generator().manual_seed(0)
depth = torch.rand(6, 88, 16, 44, generator=g).softmax(1).to(dev)
feat = torch.randn(6, 16, 44, 32, generator=g).to(dev)
cam = ops.lift_camera_params(pi[1][0], pi[3][0], pi[4][0], pi[5][0])
xs, ys, ds = vt._frustum_axes(vt.frustum, dev)
grid = tuple(int(v) for v in vt.grid_size)
bda = s0[6].reshape(1, 9).contiguous()
args = (depth, feat, cam, bda, xs, ys, ds, vt.grid_lower_bound.tolist(),
        vt.grid_interval.tolist(), 1, 6, grid)
out = ops.lift_fused(*args)
torch.cuda.synchronize()
alg = 4 * (6 * 88 * 16 * 44 + 6 * 16 * 44 * 32 + 640000 * 32)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ts = []
for i in range(10):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.lift_fused(*args, out=out); e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
print('fused, us per call (L2 flushed):', [round(t, 1) for t in ts])
print('algorithmic MB', alg / 1e6, 'GB/s at median', alg / sorted(ts)[len(ts) // 2] / 1e3)
ws2 = ops.lift_prepare(cam, bda, xs, ys, ds, vt.grid_lower_bound.tolist(),
                       vt.grid_interval.tolist(), 1, 6, grid)
out2 = ops.lift_pool(depth, feat, ws2, 1, 6, grid)
torch.cuda.synchronize()
print('pool == fused:', bool(torch.equal(out2, out)), ' non-empty voxels:',
      int((out.abs().sum(-1) > 0).sum()))
for fn, name in ((lambda: ops.lift_fused(*args, out=out), 'fused'),
                 (lambda: ops.lift_pool(depth, feat, ws2, 1, 6, grid, out=out2), 'pool')):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(20):
        fn()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / 20
    print(f'{name}: {us:.1f} us/call back-to-back, {alg / us / 1e3:.0f} GB/s algorithmic')
