"""Times pw_lift_fused at the bench size and prints the phase stamps of CTA 0."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from preworld_b200 import ops

dev = torch.device('cuda', 0)
cfg, model, samples = bench.build_workload(1)
model = model.to(dev)
s0 = tuple(t.to(dev) for t in samples[0])
pi = model.prepare_inputs(s0, stereo=True)
vt = model.img_view_transformer
g = torch.Generator().manual_seed(0)
depth = torch.rand(6, 88, 16, 44, generator=g).softmax(1).to(dev)
feat = torch.randn(6, 16, 44, 32, generator=g).to(dev)
cam = ops.lift_camera_params(pi[1][0], pi[3][0], pi[4][0], pi[5][0])
xs, ys, ds = vt._frustum_axes(vt.frustum, dev)
grid = tuple(int(v) for v in vt.grid_size)
bda = s0[6].reshape(1, 9).contiguous()
args = (depth, feat, cam, bda, xs, ys, ds, vt.grid_lower_bound.tolist(), vt.grid_interval.tolist(), 1, 6, grid)
out = ops.lift_fused(*args)
torch.cuda.synchronize()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ts = []
for i in range(10):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.lift_fused(*args, out=out); e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
ws = list(ops._ws_cache.values())[0]
st = ws[32:32 + 64].view(torch.int64).cpu().tolist()
print('us per call (L2 flushed):', [round(t, 1) for t in ts])
print('phase stamps (us since kernel start):', [round((x - st[0]) / 1e3, 2) for x in st[:8]])
print('queued voxels:', int(ws[:12].view(torch.int32)[2]), 'list entries:', int(ws[:12].view(torch.int32)[1]))
alg = 4 * (6 * 88 * 16 * 44 + 6 * 16 * 44 * 32 + 640000 * 32)
print('algorithmic MB', alg / 1e6, 'GB/s at median', alg / sorted(ts)[len(ts) // 2] / 1e3)

# accelerate=True split: lists built once, then pool only
ws2 = ops.lift_prepare(cam, bda, xs, ys, ds, vt.grid_lower_bound.tolist(),
                       vt.grid_interval.tolist(), 1, 6, grid)
out2 = ops.lift_pool(depth, feat, ws2, 1, 6, grid)
torch.cuda.synchronize()
print('pool == fused:', bool(torch.equal(out2, out)))
ts2 = []
for i in range(10):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.lift_pool(depth, feat, ws2, 1, 6, grid, out=out2); e1.record()
    torch.cuda.synchronize()
    ts2.append(e0.elapsed_time(e1) * 1e3)
st = ws2[32:32 + 64].view(torch.int64).cpu().tolist()
print('pool-only us per call (L2 flushed):', [round(t, 1) for t in ts2])
print('pool-only phase stamps:', [round((x - st[0]) / 1e3, 2) for x in st[4:8]])
print('pool-only GB/s at median', alg / sorted(ts2)[len(ts2) // 2] / 1e3)
# back-to-back (no host gap): 20 calls inside one event pair
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
