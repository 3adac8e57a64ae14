"""Secondary GPU baselines on the same B200 (NOT the oracle):

  1. the reference's own bev_pool_v2 CUDA kernel, compiled UNMODIFIED for sm_100a
     (oracle/build_ref.sh -> oracle/_ref), timed the way the model pays for it
     (QuickCumsumCuda.forward, ops/bev_pool_v2/bev_pool.py:17-41,86-92: zero-fill + kernel +
     permute(0,4,1,2,3).contiguous()), the bare kernel, and the rank preparation in torch
     (voxel_pooling_prepare_v2, view_transformer.py:203-261) -- against pw_lift_fused
     (everything), pw_lift_pool (accelerate=True) and the drop-in pw_bev_pool_v2;
  2. the conv stages of the reference path as the reference runs them on a GPU: the oracle's
     functional restatement (oracle/torch_ref.py, pinned to the reference files) with the
     state dict on the device, i.e. cuDNN / cuBLAS, strict fp32 and torch's default TF32 --
     against the same stages of this library.

    python tools/baseline_gpu.py > gpurun_out/baseline_gpu.json   (on the GPU box)
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.nn.functional as F

from oracle import gpu_ref, torch_ref
from preworld_b200 import build_model, model_cfg, ops
from preworld_b200 import synthetic as S

DEV = torch.device('cuda', 0)
FLUSH = None


def timed(fn, reps=10, flush=True):
    """median us per call, CUDA events, L2 flushed between calls"""
    global FLUSH
    if FLUSH is None:
        FLUSH = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        if flush:
            FLUSH.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return float(np.median(ts))


def lift_baseline(model, inputs):
    vt = model.img_view_transformer
    pc = torch_ref.PathConfig(model_cfg('finetune', 'r50', (256, 704)))
    geo = pc.geo
    pi = model.prepare_inputs(tuple(t.to(DEV) for t in inputs), stereo=True)
    s2k, intr, pr, pt, bda = pi[1][0], pi[3][0], pi[4][0], pi[5][0], inputs[6].to(DEV)
    B, N = s2k.shape[:2]
    D, H, W = geo.frustum.shape[:3]
    C = 32
    g = torch.Generator().manual_seed(0)
    depth = torch.rand(B * N, D, H, W, generator=g).softmax(1).to(DEV)
    feat = torch.randn(B * N, H, W, C, generator=g).to(DEV)
    gz, gy, gx = [int(v) for v in (geo.grid_size[2], geo.grid_size[1], geo.grid_size[0])]

    # --- reference: rank preparation in torch on the GPU (view_transformer.py:203-261) ---
    geo_dev = torch_ref.LiftGeometry(geo.grid_config, geo.input_size, geo.downsample, C)
    for k in ('lower', 'interval', 'grid_size', 'frustum'):
        setattr(geo_dev, k, getattr(geo_dev, k).to(DEV))

    def prepare():
        coor = torch_ref.get_lidar_coor(geo_dev, s2k, intr, pr, pt, bda)
        return _prepare_v2_dev(geo_dev, coor)

    rb, rd, rf, st, ln = prepare()
    out_ref = torch.empty((B, gz, gy, gx, C), device=DEV)

    def ref_kernel():
        gpu_ref.bev_pool_v2_forward(depth, feat, out_ref, rd, rf, rb, ln, st)

    def ref_quickcumsum():                       # bev_pool.py:17-41,86-92
        out = depth.new_zeros((B, gz, gy, gx, C))
        gpu_ref.bev_pool_v2_forward(depth, feat, out, rd, rf, rb, ln, st)
        return out.permute(0, 4, 1, 2, 3).contiguous()

    # --- this library -----------------------------------------------------------------
    cam = ops.lift_camera_params(s2k, intr, pr, pt)
    xs, ys, ds = vt._frustum_axes(vt.frustum, DEV)
    grid = tuple(int(v) for v in vt.grid_size)
    bda9 = bda.reshape(B, 9).contiguous().float()
    lo, iv = vt.grid_lower_bound.tolist(), vt.grid_interval.tolist()
    out = ops.lift_fused(depth, feat, cam, bda9, xs, ys, ds, lo, iv, B, N, grid)
    ws = ops.lift_prepare(cam, bda9, xs, ys, ds, lo, iv, B, N, grid)
    out_b = torch.zeros((B, gz, gy, gx, C), device=DEV)

    def ours_dropin():                           # same contract as the reference kernel
        ops.bev_pool_v2_(depth, feat, rd, rf, rb, st, ln, out_b)

    ref_full = ref_quickcumsum()
    same = bool(torch.equal(ref_full, out.permute(0, 4, 1, 2, 3)))
    alg_mb = 4 * (B * N * D * H * W + B * N * H * W * C + B * gz * gy * gx * C) / 1e6
    res = dict(
        algorithmic_MB=alg_mb, kept_points=int(rb.numel()), intervals=int(st.numel()),
        reference_equals_ours_bit_for_bit=same,
        reference_prepare_torch_us=timed(prepare),
        reference_bare_kernel_us=timed(ref_kernel),
        reference_zero_fill_kernel_permute_us=timed(ref_quickcumsum),
        ours_lift_fused_us=timed(lambda: ops.lift_fused(depth, feat, cam, bda9, xs, ys, ds, lo, iv,
                                                        B, N, grid, out=out)),
        ours_lift_pool_accelerate_us=timed(lambda: ops.lift_pool(depth, feat, ws, B, N, grid, out=out)),
        ours_drop_in_bev_pool_v2_kernel_us=timed(ours_dropin))
    res['reference_total_us'] = res['reference_prepare_torch_us'] + res['reference_zero_fill_kernel_permute_us']
    for k in list(res):
        if k.endswith('_us'):
            res[k.replace('_us', '_GBs_algorithmic')] = alg_mb / res[k] * 1e3
    return res


def _prepare_v2_dev(geo, coor):
    """torch_ref.voxel_pooling_prepare_v2 (view_transformer.py:203-261) with every tensor on
    coor's device (the oracle's version creates its index tensors on the CPU)."""
    dev = coor.device
    B, N, D, H, W, _ = coor.shape
    num_points = B * N * D * H * W
    ranks_depth = torch.arange(num_points, dtype=torch.int, device=dev)
    ranks_feat = torch.arange(num_points // D, dtype=torch.int, device=dev) \
        .reshape(B, N, 1, H, W).expand(B, N, D, H, W).flatten()
    coor = ((coor - geo.lower) / geo.interval).long().view(num_points, 3)
    batch_idx = torch.arange(B, device=dev).reshape(B, 1).expand(B, num_points // B) \
        .reshape(num_points, 1).to(coor)
    coor = torch.cat((coor, batch_idx), 1)
    gs = geo.grid_size
    kept = (coor[:, 0] >= 0) & (coor[:, 0] < gs[0]) & (coor[:, 1] >= 0) & (coor[:, 1] < gs[1]) & \
           (coor[:, 2] >= 0) & (coor[:, 2] < gs[2])
    coor, ranks_depth, ranks_feat = coor[kept], ranks_depth[kept], ranks_feat[kept]
    ranks_bev = coor[:, 3] * (gs[2] * gs[1] * gs[0])
    ranks_bev += coor[:, 2] * (gs[1] * gs[0])
    ranks_bev += coor[:, 1] * gs[0] + coor[:, 0]
    order = torch.sort(ranks_bev, stable=True).indices
    ranks_bev, ranks_depth, ranks_feat = ranks_bev[order], ranks_depth[order], ranks_feat[order]
    kept = torch.ones(ranks_bev.shape[0], dtype=torch.bool, device=dev)
    kept[1:] = ranks_bev[1:] != ranks_bev[:-1]
    interval_starts = torch.where(kept)[0].int()
    interval_lengths = torch.zeros_like(interval_starts)
    interval_lengths[:-1] = interval_starts[1:] - interval_starts[:-1]
    interval_lengths[-1] = ranks_bev.shape[0] - interval_starts[-1]
    return (ranks_bev.int().contiguous(), ranks_depth.int().contiguous(),
            ranks_feat.int().contiguous(), interval_starts.int().contiguous(),
            interval_lengths.int().contiguous())


def conv_stage_baseline(model, sd_dev, inputs):
    pc = torch_ref.PathConfig(model_cfg('finetune', 'r50', (256, 704)))
    dev_inputs = tuple(t.to(DEV) for t in inputs)
    imgs = dev_inputs[0]                                       # [1,18,3,256,704]
    B = imgs.shape[0]
    frames = imgs.view(B, 6, 3, 3, 256, 704)                   # camera-major, 3 frames each
    lifted = frames[:, :, :2].reshape(B, 12, 3, 256, 704)       # key + adjacent: full encoder
    ref_only = frames[:, :, 2].reshape(B, 6, 3, 256, 704)       # extra frame: stem + layer1
    g = torch.Generator().manual_seed(1)
    vol64 = torch.randn(1, 64, 16, 200, 200, generator=g).to(DEV)
    vol32 = torch.randn(1, 32, 16, 200, 200, generator=g).to(DEV)
    out = {}

    def ref_image():
        torch_ref.image_encoder(sd_dev, pc, lifted)
        torch_ref.extract_stereo_ref_feat(sd_dev, pc, ref_only)

    def ref_voxel():
        for _ in range(2):                                     # pre_process_net per lifted frame
            torch_ref.custom_resnet3d(sd_dev, 'pre_process_net', vol32, pc.pre_layers, pc.pre_stride)
        f = torch_ref.custom_resnet3d(sd_dev, 'img_bev_encoder_backbone', vol64, pc.enc_layers,
                                      pc.enc_stride)
        x = torch_ref.lss_fpn3d(sd_dev, 'img_bev_encoder_neck', f)
        x = torch_ref.conv_module(sd_dev, 'final_conv', x, 1, 1, norm=False)
        return torch_ref.occ_head(sd_dev, 'occupancy_head', x.permute(0, 1, 4, 3, 2))

    per_frame = [frames[:, :, f].contiguous() for f in range(3)]   # key, adjacent, reference-only

    def ours_image():
        model.encode_frames(per_frame)

    vol32_l = ops.to_logical(vol32.permute(0, 2, 3, 4, 1).contiguous())
    vol64_l = ops.to_logical(vol64.permute(0, 2, 3, 4, 1).contiguous())

    def ours_voxel():
        for _ in range(2):
            model.pre_process_net(vol32_l)
        x = model.bev_encoder(vol64_l)
        x = ops.conv(ops.from_logical(x), model.packs()['final'], 'relu')
        return model._occ_pair_from_head(x)

    with torch.no_grad():
        for name, tf32 in (('cudnn_fp32', False), ('cudnn_tf32_torch_default', True)):
            old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            torch.backends.cudnn.benchmark = True
            out[f'reference_{name}_image_side_ms'] = timed(ref_image, 5) / 1e3
            out[f'reference_{name}_voxel_side_ms'] = timed(ref_voxel, 5) / 1e3
            torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
        out['ours_image_side_ms'] = timed(ours_image, 5) / 1e3
        out['ours_voxel_side_ms'] = timed(ours_voxel, 5) / 1e3
    out['stages'] = ('image side: ResNet-50 + CustomFPN on the 12 lifted images, stem + layer1 on the '
                     '6 reference-only ones; voxel side: 2 x pre_process_net, CustomResNet3D, '
                     'LSSFPN3D, final_conv, OccHead (+ argmax on our side) on 16x200x200 volumes')
    return out


def main():
    torch.cuda.set_device(0)
    model = build_model(model_cfg('finetune', 'r50', (256, 704))).eval()
    S.lively_init_(model, 0)
    sd_dev = {k: v.detach().clone().to(DEV) for k, v in model.state_dict().items()}
    model = model.to(DEV)
    inputs = S.make_img_inputs(1, (256, 704), seed=0)
    res = dict(gpu=torch.cuda.get_device_name(0))
    with torch.no_grad():
        if gpu_ref.available():
            res['lift'] = lift_baseline(model, inputs)
        else:
            res['lift'] = 'oracle/_ref/libbev_pool_v2_ref.so missing (build: oracle/build_ref.sh)'
        res['conv_stages'] = conv_stage_baseline(model, sd_dev, inputs)
    print(json.dumps(res, indent=1))


if __name__ == '__main__':
    main()
