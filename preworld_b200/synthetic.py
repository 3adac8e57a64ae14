"""Synthetic workloads for the camera->voxel forward path.

There is no dataset in the build/bench environment, so every test and bench
uses the tensor *contracts* of the reference's loaders with synthetic values:

* ``img_inputs`` 7-tuple produced by ``PrepareImageInputs`` (reference
  mmdet3d/datasets/pipelines/loading.py:1124-1134): images are camera-major
  ``[cam0:(key, adj, ref), cam1: ...]`` (loading.py:1037-1104) while the pose
  tensors are frame-major (loading.py:1111-1122).
* rays ``[B, R, 16]`` as laid out by mmdet3d/datasets/ray.py:49-56
  (0-1 pixel, 2 depth, 3 semantic, 4:7 origin, 7:10 direction, 10:13 viewdir,
  13:16 rgb).
* ego states ``[B, 1, 21]`` (preworld_temporal_traj.py:119-121).

The 6-camera rig is nuScenes-like (our choice, not in the reference): see
``camera_rig``.  The R50 / 256x704 model dict is a *derived* config: the
reference ships only Swin-B @ 512x1408 (configs/preworld/nuscenes/
bevstereo-occ.py:16,45-67) but supports mmdet ResNet + CustomFPN in code
(detectors/bevdet.py:577-588, necks/fpn.py:10-11).
"""
import copy
import math

import torch

YAWS_DEG = (55.0, 0.0, -55.0, -110.0, 180.0, 110.0)


def camera_rig(num_cams=6, dtype=torch.float32):
    """sensor2ego [N,4,4] and intrinsics [N,3,3] of the synthetic rig."""
    cam2ego_axes = torch.tensor([[0., 0., 1.], [-1., 0., 0.], [0., -1., 0.]],
                                dtype=torch.float64)
    s2e = torch.zeros(num_cams, 4, 4, dtype=torch.float64)
    for i in range(num_cams):
        yaw = math.radians(YAWS_DEG[i % len(YAWS_DEG)])
        rz = torch.tensor([[math.cos(yaw), -math.sin(yaw), 0.],
                           [math.sin(yaw), math.cos(yaw), 0.],
                           [0., 0., 1.]], dtype=torch.float64)
        s2e[i, :3, :3] = rz @ cam2ego_axes
        s2e[i, :3, 3] = torch.tensor([1.5 * math.cos(yaw),
                                      1.5 * math.sin(yaw), 1.5])
        s2e[i, 3, 3] = 1.
    K = torch.tensor([[1266., 0., 816.], [0., 1266., 491.], [0., 0., 1.]],
                     dtype=torch.float64)
    return s2e.to(dtype), K.expand(num_cams, 3, 3).contiguous().to(dtype)


def make_img_inputs(batch=1, input_size=(256, 704), num_frames=3, num_cams=6,
                    seed=0, device='cpu'):
    """The reference's ``img_inputs`` 7-tuple with synthetic values."""
    H, W = input_size
    g = torch.Generator().manual_seed(seed)
    imgs = torch.randn(batch, num_cams * num_frames, 3, H, W, generator=g)
    s2e, K = camera_rig(num_cams)
    # frame-major pose tensors
    sensor2egos = s2e.repeat(num_frames, 1, 1)[None].repeat(batch, 1, 1, 1)
    ego2globals = torch.eye(4).repeat(batch, num_frames * num_cams, 1, 1)
    for f in range(num_frames):
        ego2globals[:, f * num_cams:(f + 1) * num_cams, 0, 3] = -0.5 * f
    intrins = K.repeat(num_frames, 1, 1)[None].repeat(batch, 1, 1, 1)
    scale = H / 256.0 * 0.44  # sample_augmentation test branch,
    # loading.py:988-1000: resize = W/1600, crop so the bottom stays
    post_rots = torch.eye(3).repeat(batch, num_frames * num_cams, 1, 1)
    post_rots[..., 0, 0] = scale
    post_rots[..., 1, 1] = scale
    post_trans = torch.zeros(batch, num_frames * num_cams, 3)
    post_trans[..., 1] = -(900 * scale - H)
    bda = torch.eye(3).repeat(batch, 1, 1)
    out = (imgs, sensor2egos, ego2globals, intrins, post_rots, post_trans, bda)
    return tuple(t.to(device) for t in out)


def make_rays(img_inputs, num_rays=38400, seed=1):
    """rays [B,R,16] through random pixels of the 6 key-frame cameras
    (geometry as reference mmdet3d/datasets/ray.py:34-45)."""
    imgs, sensor2egos, _, intrins, post_rots, post_trans, _ = img_inputs
    B = imgs.shape[0]
    H, W = imgs.shape[-2:]
    N = 6
    g = torch.Generator().manual_seed(seed)
    rays = torch.zeros(B, num_rays, 16)
    for b in range(B):
        cam = torch.randint(0, N, (num_rays,), generator=g)
        u = torch.rand(num_rays, generator=g) * (W - 1)
        v = torch.rand(num_rays, generator=g) * (H - 1)
        pix = torch.stack([u, v, torch.ones_like(u)], -1)
        s2e = sensor2egos[b, :N].cpu()[cam]
        K = intrins[b, :N].cpu()[cam]
        pr = post_rots[b, :N].cpu()[cam]
        pt = post_trans[b, :N].cpu()[cam]
        p = torch.linalg.solve(pr, (pix - pt).unsqueeze(-1))
        d_cam = torch.linalg.solve(K, p).squeeze(-1)
        d_ego = (s2e[:, :3, :3] @ d_cam.unsqueeze(-1)).squeeze(-1)
        d_ego = d_ego / d_ego.norm(dim=-1, keepdim=True)
        rays[b, :, 0] = u
        rays[b, :, 1] = v
        rays[b, :, 2] = 1 + 51 * torch.rand(num_rays, generator=g)
        rays[b, :, 3] = torch.randint(0, 17, (num_rays,), generator=g).float()
        rays[b, :, 4:7] = s2e[:, :3, 3]
        rays[b, :, 7:10] = d_ego
        rays[b, :, 10:13] = d_ego
        rays[b, :, 13:16] = torch.randn(num_rays, 3, generator=g)
    return rays.to(imgs.device)


def make_ego_states(batch=1, seed=2, device='cpu'):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(batch, 1, 21, generator=g).to(device)


def r50_model_cfg(model_cfg, input_size=(256, 704), depth=50):
    """Derive the ResNet-50/101 @ 256x704 model dict BASELINE.json names from
    a shipped (Swin-B) preworld model dict.  Everything downstream of the view
    transformer is left untouched."""
    cfg = copy.deepcopy(model_cfg)
    cfg['img_backbone'] = dict(
        type='ResNet', depth=depth, num_stages=4, out_indices=(0, 2, 3),
        frozen_stages=-1, norm_cfg=dict(type='BN', requires_grad=True),
        norm_eval=False, with_cp=False, style='pytorch')
    cfg['img_neck'] = dict(
        type='CustomFPN', in_channels=[1024, 2048], out_channels=256,
        num_outs=1, start_level=0, out_ids=[0])
    cfg['img_view_transformer']['in_channels'] = 256
    cfg['img_view_transformer']['input_size'] = tuple(input_size)
    return cfg


def _key_generator(key, seed):
    """Per-tensor generator seeded by (state_dict key, seed) so the synthetic
    weights depend only on the key names -- the reference's modules and ours
    get bit-identical values regardless of module construction order."""
    import zlib
    g = torch.Generator()
    g.manual_seed((zlib.crc32(key.encode()) * 2654435761 + seed) % (2 ** 63))
    return g


_RESIDUAL_TAIL = ('bn3', 'bn2', 'conv2.bn')


@torch.no_grad()
def lively_init_(module, seed=0, residual_gain=0.25):
    """Deterministic, key-addressed random init that keeps activations O(1)
    through the ReLU stacks so the synthetic depth distributions and argmax
    grids are not degenerate: He-normal conv/linear weights, N(0, 0.1) biases,
    non-identity norm statistics (running_mean~N(0,0.1),
    running_var~U(0.5,1.5), gamma~U(0.5,1.5), beta~N(0,0.1)) and a damped gamma
    on the last norm of every residual branch (otherwise ~30 residual adds blow
    the scale up by 1e6 and every softmax saturates)."""
    for name, m in module.named_modules():
        if isinstance(m, (torch.nn.modules.conv._ConvNd, torch.nn.Linear)):
            fan_in = m.weight[0].numel()
            g = _key_generator(name + '.weight', seed)
            m.weight.copy_(torch.randn(m.weight.shape, generator=g)
                           * math.sqrt(2.0 / fan_in))
            if m.bias is not None:
                g = _key_generator(name + '.bias', seed)
                m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)
            if name.endswith('depth_net.depth_conv.4'):
                m.weight.mul_(0.15)   # depth logits ~N(0,2): soft depth bins
        elif isinstance(m, torch.nn.modules.batchnorm._BatchNorm):
            n = m.num_features
            g = _key_generator(name + '.bn', seed)
            m.running_mean.copy_(torch.randn(n, generator=g) * 0.1)
            m.running_var.copy_(torch.rand(n, generator=g) + 0.5)
            m.weight.copy_(torch.rand(n, generator=g) + 0.5)
            m.bias.copy_(torch.randn(n, generator=g) * 0.1)
            if name.endswith(_RESIDUAL_TAIL):
                m.weight.mul_(residual_gain)
        elif isinstance(m, torch.nn.LayerNorm):                 # Swin image side
            n = m.normalized_shape[0]
            g = _key_generator(name + '.ln', seed)
            m.weight.copy_(torch.rand(n, generator=g) + 0.5)
            m.bias.copy_(torch.randn(n, generator=g) * 0.1)
    for name, p in module.named_parameters():
        if name.endswith('relative_position_bias_table'):
            g = _key_generator(name, seed)
            p.copy_(torch.randn(p.shape, generator=g) * 0.5)
    return module
