"""Runs the stereo cost-volume kernel alone at the BASELINE config's size (12
images: 6 cameras x 2 lifted frames, 64x176 stereo pixels, 256 channels, 88 bins)
with the synthetic rig's geometry -- the target of an `ncu -k regex:cost_volume`
capture (tools/gpu_profile.sh) and a quick CUDA-event timing."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from preworld_b200 import build_model, model_cfg, ops
from preworld_b200 import synthetic as S


def main():
    dev = 'cuda'
    model = build_model(model_cfg('finetune', 'r50', (256, 704))).eval()
    inputs = tuple(t.to(dev) for t in S.make_img_inputs(1, (256, 704), seed=0))
    pi = model.prepare_inputs(inputs, stereo=True)
    vt = model.img_view_transformer
    k2s = torch.cat([pi[7][0], pi[7][1]], 0)
    cat = lambda ts: torch.cat([ts[0], ts[1]], 0)
    cam = ops.cv_camera_params(k2s, cat(pi[3]), cat(pi[4]), cat(pi[5]))
    fr = vt.cv_frustum.to(dev)
    xs, ys, ds = fr[0, 0, :, 0].contiguous(), fr[0, :, 0, 1].contiguous(), fr[:, 0, 0, 2].contiguous()
    g = torch.Generator(device=dev).manual_seed(0)
    curr = torch.relu(torch.randn(12, 64, 176, 256, device=dev, generator=g))
    prev = torch.relu(torch.randn(12, 64, 176, 256, device=dev, generator=g))
    for _ in range(3):
        out = ops.cost_volume(curr, prev, cam, xs, ys, ds, 5.0, (256, 704), pad_to=96)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        out = ops.cost_volume(curr, prev, cam, xs, ys, ds, 5.0, (256, 704), pad_to=96)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    mb = 4e-6 * 12 * 64 * 176 * (2 * 256 + 88)
    print(f'cost volume, 12 images: {ms * 1e3:.1f} us, algorithmic {mb:.0f} MB -> {mb / ms:.0f} GB/s; '
          f'checksum {out.sum().item():.3f}')


if __name__ == '__main__':
    main()
