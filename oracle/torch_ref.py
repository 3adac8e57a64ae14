"""TEST INFRASTRUCTURE ONLY -- CPU fp32 restatement of the reference's
camera->voxel forward path, written functionally over a reference-keyed
``state_dict`` so that it can travel to the GPU box (where /root/reference does
not exist).  Every function cites the reference file:line it follows
(getterupper/PreWorld @ 0b0e021).  Never imported by ``preworld_b200``.

Parity pin: ``oracle/make_golden.py`` runs the reference's own .py files
verbatim (oracle/ref_shim.py) on the same seeded inputs in the build container
and checks this restatement against them stage by stage; the resulting
fixtures live in tests/golden/.  The reference itself pins only bev_pool_v2
(bev_pool.py:145-176) -- that KAT is checked in tests/test_oracle.py.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from . import c_ref

EPS = 1e-5


# ---------------------------------------------------------------- primitives
def _bn(sd, p, x):
    return F.batch_norm(x, sd[p + '.running_mean'], sd[p + '.running_var'],
                        sd[p + '.weight'], sd[p + '.bias'], False, 0.0, EPS)


def _conv(sd, p, x, stride=1, padding=0, dilation=1):
    w = sd[p + '.weight']
    f = F.conv3d if w.dim() == 5 else F.conv2d
    return f(x, w, sd.get(p + '.bias'), stride, padding, dilation)


def _linear(sd, p, x):
    return F.linear(x, sd[p + '.weight'], sd.get(p + '.bias'))


def conv_module(sd, p, x, stride=1, padding=0, norm=True, act=True):
    """mmcv 1.6.0 ConvModule: conv -> bn -> relu."""
    x = _conv(sd, p + '.conv', x, stride, padding)
    if norm:
        x = _bn(sd, p + '.bn', x)
    return F.relu(x) if act else x


# ------------------------------------------------- mmdet ResNet (third party)
RESNET_BLOCKS = {50: (3, 4, 6, 3), 101: (3, 4, 23, 3)}


def _bottleneck(sd, p, x, stride):
    """mmdet 2.24 resnet.py Bottleneck, style='pytorch' (stride on conv2)."""
    out = F.relu(_bn(sd, p + '.bn1', _conv(sd, p + '.conv1', x)))
    out = F.relu(_bn(sd, p + '.bn2', _conv(sd, p + '.conv2', out, stride, 1)))
    out = _bn(sd, p + '.bn3', _conv(sd, p + '.conv3', out))
    if p + '.downsample.0.weight' in sd:
        x = _bn(sd, p + '.downsample.1',
                _conv(sd, p + '.downsample.0', x, stride))
    return F.relu(out + x)


def resnet_stem(sd, p, x):
    x = F.relu(_bn(sd, p + '.bn1', _conv(sd, p + '.conv1', x, 2, 3)))
    return F.max_pool2d(x, 3, 2, 1)


def resnet_layer(sd, p, x, i, depth):
    for b in range(RESNET_BLOCKS[depth][i]):
        stride = 2 if (b == 0 and i > 0) else 1
        x = _bottleneck(sd, f'{p}.layer{i + 1}.{b}', x, stride)
    return x


def resnet(sd, p, x, depth=50, out_indices=(0, 2, 3)):
    x = resnet_stem(sd, p, x)
    outs = []
    for i in range(4):
        x = resnet_layer(sd, p, x, i, depth)
        if i in out_indices:
            outs.append(x)
    return outs


def custom_fpn(sd, p, inputs):
    """necks/fpn.py:154-203 with num_outs=1, out_ids=[0], no norm/act,
    nearest upsampling."""
    lat = [_conv(sd, f'{p}.lateral_convs.{i}.conv', x)
           for i, x in enumerate(inputs)]
    for i in range(len(lat) - 1, 0, -1):
        lat[i - 1] = lat[i - 1] + F.interpolate(
            lat[i], size=lat[i - 1].shape[2:], mode='nearest')
    return _conv(sd, f'{p}.fpn_convs.0.conv', lat[0], 1, 1)


# ------------------------------------------------------------ view transformer
class LiftGeometry:
    """Constants of LSSViewTransformer.__init__ (view_transformer.py:40-112)."""

    def __init__(self, grid_config, input_size, downsample, out_channels=32):
        self.grid_config = grid_config
        self.input_size = input_size
        self.downsample = downsample
        self.out_channels = out_channels
        xyz = [grid_config[k] for k in ('x', 'y', 'z')]
        self.lower = torch.Tensor([c[0] for c in xyz])
        self.interval = torch.Tensor([c[2] for c in xyz])
        self.grid_size = torch.Tensor([(c[1] - c[0]) / c[2] for c in xyz])
        self.frustum = self.create_frustum(downsample)
        self.cv_frustum = self.create_frustum(4)
        self.D = self.frustum.shape[0]

    def create_frustum(self, downsample):
        """view_transformer.py:84-112 (sid=False)."""
        H_in, W_in = self.input_size
        Hf, Wf = H_in // downsample, W_in // downsample
        d = torch.arange(*self.grid_config['depth'], dtype=torch.float) \
            .view(-1, 1, 1).expand(-1, Hf, Wf)
        D = d.shape[0]
        x = torch.linspace(0, W_in - 1, Wf, dtype=torch.float) \
            .view(1, 1, Wf).expand(D, Hf, Wf)
        y = torch.linspace(0, H_in - 1, Hf, dtype=torch.float) \
            .view(1, Hf, 1).expand(D, Hf, Wf)
        return torch.stack((x, y, d), -1)


def get_lidar_coor(geo, sensor2ego, cam2imgs, post_rots, post_trans, bda):
    """view_transformer.py:114-153."""
    B, N = sensor2ego.shape[:2]
    points = geo.frustum - post_trans.view(B, N, 1, 1, 1, 3)
    points = torch.inverse(post_rots).view(B, N, 1, 1, 1, 3, 3) \
        .matmul(points.unsqueeze(-1))
    points = torch.cat(
        (points[..., :2, :] * points[..., 2:3, :], points[..., 2:3, :]), 5)
    combine = sensor2ego[:, :, :3, :3].matmul(torch.inverse(cam2imgs))
    points = combine.view(B, N, 1, 1, 1, 3, 3).matmul(points).squeeze(-1)
    points = points + sensor2ego[:, :, :3, 3].view(B, N, 1, 1, 1, 3)
    points = bda.view(B, 1, 1, 1, 1, 3, 3).matmul(
        points.unsqueeze(-1)).squeeze(-1)
    return points


def voxel_pooling_prepare_v2(geo, coor):
    """view_transformer.py:203-261.  The reference's ``argsort`` is not
    stable, so its order inside a voxel is unspecified; the oracle fixes it to
    ascending point index (stable sort)."""
    B, N, D, H, W, _ = coor.shape
    num_points = B * N * D * H * W
    ranks_depth = torch.arange(num_points, dtype=torch.int)
    ranks_feat = torch.arange(num_points // D, dtype=torch.int) \
        .reshape(B, N, 1, H, W).expand(B, N, D, H, W).flatten()
    coor = ((coor - geo.lower) / geo.interval).long().view(num_points, 3)
    batch_idx = torch.arange(B).reshape(B, 1).expand(B, num_points // B) \
        .reshape(num_points, 1).to(coor)
    coor = torch.cat((coor, batch_idx), 1)
    gs = geo.grid_size
    kept = (coor[:, 0] >= 0) & (coor[:, 0] < gs[0]) & \
           (coor[:, 1] >= 0) & (coor[:, 1] < gs[1]) & \
           (coor[:, 2] >= 0) & (coor[:, 2] < gs[2])
    if kept.sum() == 0:
        return None, None, None, None, None
    coor, ranks_depth, ranks_feat = coor[kept], ranks_depth[kept], \
        ranks_feat[kept]
    ranks_bev = coor[:, 3] * (gs[2] * gs[1] * gs[0])   # float32, as reference
    ranks_bev += coor[:, 2] * (gs[1] * gs[0])
    ranks_bev += coor[:, 1] * gs[0] + coor[:, 0]
    order = torch.sort(ranks_bev, stable=True).indices
    ranks_bev, ranks_depth, ranks_feat = \
        ranks_bev[order], ranks_depth[order], ranks_feat[order]
    kept = torch.ones(ranks_bev.shape[0], dtype=torch.bool)
    kept[1:] = ranks_bev[1:] != ranks_bev[:-1]
    interval_starts = torch.where(kept)[0].int()
    interval_lengths = torch.zeros_like(interval_starts)
    interval_lengths[:-1] = interval_starts[1:] - interval_starts[:-1]
    interval_lengths[-1] = ranks_bev.shape[0] - interval_starts[-1]
    return (ranks_bev.int().contiguous(), ranks_depth.int().contiguous(),
            ranks_feat.int().contiguous(), interval_starts.int().contiguous(),
            interval_lengths.int().contiguous())


def bev_pool_v2(depth, feat, ranks_depth, ranks_feat, ranks_bev,
                bev_feat_shape, interval_starts, interval_lengths,
                channels_last=False):
    """ops/bev_pool_v2/bev_pool.py:17-41,86-92 over the C restatement of
    bev_pool_cuda.cu:21-48.  depth [B,N,D,H,W], feat [B,N,H,W,C]."""
    out = c_ref.bev_pool_v2_fwd(
        depth.contiguous().float().numpy().ravel(),
        feat.contiguous().float().numpy().reshape(-1, feat.shape[-1]),
        ranks_depth.numpy(), ranks_feat.numpy(), ranks_bev.numpy(),
        interval_starts.numpy(), interval_lengths.numpy(),
        int(np.prod(bev_feat_shape[:-1])))
    out = torch.from_numpy(out).view(*bev_feat_shape)
    if channels_last:
        return out
    return out.permute(0, 4, 1, 2, 3).contiguous()


def voxel_pooling_v2(geo, coor, depth, feat):
    """view_transformer.py:176-201 (collapse_z=False)."""
    rb, rd, rf, st, ln = voxel_pooling_prepare_v2(geo, coor)
    B = depth.shape[0]
    gz, gy, gx = int(geo.grid_size[2]), int(geo.grid_size[1]), \
        int(geo.grid_size[0])
    if rb is None:
        return torch.zeros(B, feat.shape[2], gz, gy, gx)
    feat = feat.permute(0, 1, 3, 4, 2)
    shape = (B, gz, gy, gx, feat.shape[-1])
    return bev_pool_v2(depth, feat, rd, rf, rb, shape, st, ln)


def get_mlp_input(sensor2ego, intrin, post_rot, post_tran, bda):
    """view_transformer.py:713-734."""
    B, N = sensor2ego.shape[:2]
    bda = bda.view(B, 1, 3, 3).repeat(1, N, 1, 1)
    mlp_input = torch.stack([
        intrin[:, :, 0, 0], intrin[:, :, 1, 1], intrin[:, :, 0, 2],
        intrin[:, :, 1, 2], post_rot[:, :, 0, 0], post_rot[:, :, 0, 1],
        post_tran[:, :, 0], post_rot[:, :, 1, 0], post_rot[:, :, 1, 1],
        post_tran[:, :, 1], bda[:, :, 0, 0], bda[:, :, 0, 1],
        bda[:, :, 1, 0], bda[:, :, 1, 1], bda[:, :, 2, 2]], dim=-1)
    return torch.cat([mlp_input,
                      sensor2ego[:, :, :3, :].reshape(B, N, -1)], dim=-1)


# ------------------------------------------------------------------ DepthNet
def _mlp(sd, p, x):
    return _linear(sd, p + '.fc2', F.relu(_linear(sd, p + '.fc1', x)))


def _se(sd, p, x, x_se):
    x_se = _conv(sd, p + '.conv_expand',
                 F.relu(_conv(sd, p + '.conv_reduce', x_se)))
    return x * torch.sigmoid(x_se)


def _basic_block(sd, p, x):
    """mmdet 2.24 BasicBlock as instantiated at view_transformer.py:515-517
    (downsample = biased 1x1 conv, :505-506)."""
    out = F.relu(_bn(sd, p + '.bn1', _conv(sd, p + '.conv1', x, 1, 1)))
    out = _bn(sd, p + '.bn2', _conv(sd, p + '.conv2', out, 1, 1))
    if p + '.downsample.weight' in sd:
        x = _conv(sd, p + '.downsample', x)
    return F.relu(out + x)


def _aspp(sd, p, x):
    """view_transformer.py:355-418 (dropout is an eval no-op)."""
    outs = []
    for i, dil in enumerate((1, 6, 12, 18)):
        q = f'{p}.aspp{i + 1}'
        pad = 0 if i == 0 else dil
        y = _conv(sd, q + '.atrous_conv', x, 1, pad, dil)
        outs.append(F.relu(_bn(sd, q + '.bn', y)))
    g = F.adaptive_avg_pool2d(x, (1, 1))
    g = F.relu(_bn(sd, p + '.global_avg_pool.2',
                   _conv(sd, p + '.global_avg_pool.1', g)))
    outs.append(F.interpolate(g, size=x.shape[2:], mode='bilinear',
                              align_corners=True))
    x = torch.cat(outs, dim=1)
    return F.relu(_bn(sd, p + '.bn1', _conv(sd, p + '.conv1', x)))


def gen_cv_grid(metas, B, N, D, H, W, hi, wi):
    """view_transformer.py:546-574."""
    frustum = metas['frustum']
    points = frustum - metas['post_trans'].view(B, N, 1, 1, 1, 3)
    points = torch.inverse(metas['post_rots']).view(B, N, 1, 1, 1, 3, 3) \
        .matmul(points.unsqueeze(-1))
    points = torch.cat(
        (points[..., :2, :] * points[..., 2:3, :], points[..., 2:3, :]), 5)
    rots = metas['k2s_sensor'][:, :, :3, :3].contiguous()
    trans = metas['k2s_sensor'][:, :, :3, 3].contiguous()
    combine = rots.matmul(torch.inverse(metas['intrins']))
    points = combine.view(B, N, 1, 1, 1, 3, 3).matmul(points)
    points = points + trans.view(B, N, 1, 1, 1, 3, 1)
    neg_mask = points[..., 2, 0] < 1e-3
    points = metas['intrins'].view(B, N, 1, 1, 1, 3, 3).matmul(points)
    points = points[..., :2, :] / points[..., 2:3, :]
    points = metas['post_rots'][..., :2, :2].view(B, N, 1, 1, 1, 2, 2) \
        .matmul(points).squeeze(-1)
    points = points + metas['post_trans'][..., :2].view(B, N, 1, 1, 1, 2)
    px = points[..., 0] / (wi - 1.0) * 2.0 - 1.0
    py = points[..., 1] / (hi - 1.0) * 2.0 - 1.0
    px[neg_mask] = -2
    py[neg_mask] = -2
    return torch.stack([px, py], dim=-1).view(B * N, D * H, W, 2)


def calculate_cost_volume(metas, bias):
    """view_transformer.py:576-604."""
    prev, curr = metas['cv_feat_list']
    group_size = 4
    _, c, hf, wf = curr.shape
    hi, wi = hf * 4, wf * 4
    B, N, _ = metas['post_trans'].shape
    D, H, W, _ = metas['frustum'].shape
    grid = gen_cv_grid(metas, B, N, D, H, W, hi, wi).to(curr.dtype)
    prev = prev.view(B * N, -1, H, W)
    curr = curr.view(B * N, -1, H, W)
    cost = 0
    for fid in range(curr.shape[1] // group_size):
        sl = slice(fid * group_size, (fid + 1) * group_size)
        wrap_prev = F.grid_sample(prev[:, sl], grid, align_corners=True,
                                  padding_mode='zeros')
        tmp = curr[:, sl].unsqueeze(2) - wrap_prev.view(B * N, -1, D, H, W)
        cost = cost + tmp.abs().sum(dim=1)
    if not bias == 0:
        invalid = wrap_prev[:, 0, ...].view(B * N, D, H, W) == 0
        cost[invalid] = cost[invalid] + bias
    return (-cost).softmax(dim=1)


def depth_net(sd, p, x, mlp_input, stereo_metas, depth_channels, cv_bias):
    """view_transformer.py:606-638 (stereo=True, use_dcn=False, use_aspp)."""
    mlp_input = F.batch_norm(
        mlp_input.reshape(-1, mlp_input.shape[-1]),
        sd[p + '.bn.running_mean'], sd[p + '.bn.running_var'],
        sd[p + '.bn.weight'], sd[p + '.bn.bias'], False, 0.0, EPS)
    x = F.relu(_bn(sd, p + '.reduce_conv.1',
                   _conv(sd, p + '.reduce_conv.0', x, 1, 1)))
    context_se = _mlp(sd, p + '.context_mlp', mlp_input)[..., None, None]
    context = _se(sd, p + '.context_se', x, context_se)
    context = _conv(sd, p + '.context_conv', context)
    depth_se = _mlp(sd, p + '.depth_mlp', mlp_input)[..., None, None]
    depth = _se(sd, p + '.depth_se', x, depth_se)
    if stereo_metas['cv_feat_list'][0] is None:
        BN, _, H, W = x.shape
        s = float(stereo_metas['downsample']) / stereo_metas['cv_downsample']
        cost = torch.zeros(BN, depth_channels, int(H * s), int(W * s))
    else:
        cost = calculate_cost_volume(stereo_metas, cv_bias)
    for i in (0, 2):   # cost_volumn_net: 2x (3x3 s2 conv + BN)
        cost = _bn(sd, f'{p}.cost_volumn_net.{i + 1}',
                   _conv(sd, f'{p}.cost_volumn_net.{i}', cost, 2, 1))
    depth = torch.cat([depth, cost], dim=1)
    for i in range(3):
        depth = _basic_block(sd, f'{p}.depth_conv.{i}', depth)
    depth = _aspp(sd, p + '.depth_conv.3', depth)
    depth = _conv(sd, p + '.depth_conv.4', depth)
    return torch.cat([depth, context], dim=1)


def view_transformer_forward(sd, p, geo, inputs, stereo_metas, cv_bias=5.0):
    """LSSViewTransformerBEVDepth.forward, view_transformer.py:791-804."""
    x, sensor2keyego, ego2global, intrin, post_rot, post_tran, bda, mlp_in = \
        inputs
    B, N, C, H, W = x.shape
    x = depth_net(sd, p + '.depth_net', x.view(B * N, C, H, W), mlp_in,
                  stereo_metas, geo.D, cv_bias)
    depth = x[:, :geo.D].softmax(dim=1)
    tran_feat = x[:, geo.D:geo.D + geo.out_channels]
    coor = get_lidar_coor(geo, sensor2keyego, intrin, post_rot, post_tran, bda)
    bev = voxel_pooling_v2(geo, coor, depth.view(B, N, geo.D, H, W),
                           tran_feat.view(B, N, geo.out_channels, H, W))
    return bev, depth


# --------------------------------------------------------------- 3-D encoder
def _basic_block3d(sd, p, x, stride):
    """backbones/resnet.py:88-123; the shortcut is a full 3^3 conv + BN
    (:147-162)."""
    identity = x
    if p + '.downsample.conv.weight' in sd:
        identity = conv_module(sd, p + '.downsample', x, stride, 1, act=False)
    x = conv_module(sd, p + '.conv1', x, stride, 1)
    x = conv_module(sd, p + '.conv2', x, 1, 1, act=False)
    return F.relu(x + identity)


def custom_resnet3d(sd, p, x, num_layer, stride):
    """backbones/resnet.py:126-184."""
    feats = []
    for i, (nl, s) in enumerate(zip(num_layer, stride)):
        for b in range(nl):
            x = _basic_block3d(sd, f'{p}.layers.{i}.{b}', x, s if b == 0 else 1)
        feats.append(x)
    return feats


def lss_fpn3d(sd, p, feats):
    """necks/lss_fpn.py:132-148."""
    x8, x16, x32 = feats
    x16 = F.interpolate(x16, scale_factor=2, mode='trilinear',
                        align_corners=True)
    x32 = F.interpolate(x32, scale_factor=4, mode='trilinear',
                        align_corners=True)
    return conv_module(sd, p + '.conv', torch.cat([x8, x16, x32], dim=1))


def occ_head(sd, p, voxel_feat):
    """heads/occupancy_head.py:124-177 with num_level=1, use_deblock=False,
    soft_weights=True.  softmax over the single soft-weight channel is
    identically 1 and the same-size trilinear interpolate is the identity, so
    both are evaluated exactly as the reference does only in make_golden.py;
    here they reduce to a multiplication by 1.0."""
    x = F.relu(_bn(sd, p + '.occ_convs.0.1',
                   _conv(sd, p + '.occ_convs.0.0', voxel_feat, 1, 1)))
    x = F.relu(_bn(sd, p + '.occ_pred_conv.1',
                   _conv(sd, p + '.occ_pred_conv.0', x)))
    return _conv(sd, p + '.occ_pred_conv.3', x)


# ------------------------------------------------------------------ detector
def prepare_inputs(inputs, num_frame=3, temporal_frame=2):
    """detectors/bevdet_occ.py:88-139 (stereo=True)."""
    B, N, C, H, W = inputs[0].shape
    N = N // num_frame
    imgs = inputs[0].view(B, N, num_frame, C, H, W)
    imgs = [t.squeeze(2) for t in torch.split(imgs, 1, 2)]
    sensor2egos, ego2globals, intrins, post_rots, post_trans, bda = inputs[1:7]
    sensor2egos = sensor2egos.view(B, num_frame, N, 4, 4)
    ego2globals = ego2globals.view(B, num_frame, N, 4, 4)
    keyego2global = ego2globals[:, 0, 0, ...].unsqueeze(1).unsqueeze(1)
    global2keyego = torch.inverse(keyego2global.double())
    sensor2keyegos = (global2keyego @ ego2globals.double()
                      @ sensor2egos.double()).float()
    tf = temporal_frame
    curr2adj = torch.inverse(ego2globals[:, 1:tf + 1].double()
                             @ sensor2egos[:, 1:tf + 1].double()) \
        @ ego2globals[:, :tf].double() @ sensor2egos[:, :tf].double()
    curr2adj = [p.squeeze(1) for p in torch.split(curr2adj.float(), 1, 1)]
    curr2adj.extend([None] * (num_frame - tf))
    extra = [sensor2keyegos, ego2globals,
             intrins.view(B, num_frame, N, 3, 3),
             post_rots.view(B, num_frame, N, 3, 3),
             post_trans.view(B, num_frame, N, 3)]
    extra = [[p.squeeze(1) for p in torch.split(t, 1, 1)] for t in extra]
    return (imgs, *extra, bda, curr2adj)


class PathConfig:
    """Structural hyper-parameters the oracle needs from the model dict."""

    def __init__(self, model_cfg):
        vt = model_cfg['img_view_transformer']
        self.geo = LiftGeometry(vt['grid_config'], tuple(vt['input_size']),
                                vt['downsample'], vt['out_channels'])
        self.cv_bias = vt['depthnet_cfg'].get('bias', 0.0)
        self.backbone_cfg = dict(model_cfg['img_backbone'])
        self.neck_cfg = dict(model_cfg['img_neck'])
        self.swin = self.backbone_cfg['type'] == 'SwinTransformer'
        self.depth = self.backbone_cfg.get('depth')
        self.out_indices = tuple(self.backbone_cfg['out_indices'])
        enc = model_cfg['img_bev_encoder_backbone']
        self.enc_layers, self.enc_stride = enc['num_layer'], enc['stride']
        pre = model_cfg['pre_process']
        self.pre_layers, self.pre_stride = pre['num_layer'], pre['stride']
        self.num_frame = model_cfg.get('num_adj', 1) + 1 + 1
        self.num_classes = model_cfg.get('num_classes', 18)
        self.if_post_finetune = model_cfg.get('if_post_finetune', False)
        self.test_threshold = model_cfg.get('test_threshold', 8.5)


def image_encoder(sd, pc, img):
    """detectors/bevdet.py:34-50 (stereo=True)."""
    B, N, C, H, W = img.shape
    if pc.swin:                      # the shipped image side (oracle/swin_ref.py)
        from . import swin_ref
        outs = swin_ref.backbone_forward(
            swin_ref.backbone_from_state_dict(sd, 'img_backbone', pc.backbone_cfg),
            img.view(B * N, C, H, W))
        x = swin_ref.neck_forward(
            swin_ref.neck_from_state_dict(sd, 'img_neck', pc.neck_cfg), outs[1:])
        return x.view(B, N, *x.shape[1:]), outs[0]
    x = resnet(sd, 'img_backbone', img.view(B * N, C, H, W), pc.depth,
               pc.out_indices)
    stereo_feat, x = x[0], x[1:]
    x = custom_fpn(sd, 'img_neck', x)
    return x.view(B, N, *x.shape[1:]), stereo_feat


def extract_stereo_ref_feat(sd, pc, img):
    """detectors/bevdet.py:573-588 (mmdet ResNet branch)."""
    B, N, C, H, W = img.shape
    if pc.swin:                      # bevdet.py:589-604
        from . import swin_ref
        return swin_ref.stage0_forward(
            swin_ref.backbone_from_state_dict(sd, 'img_backbone', pc.backbone_cfg),
            img.view(B * N, C, H, W))
    x = resnet_stem(sd, 'img_backbone', img.view(B * N, C, H, W))
    return resnet_layer(sd, 'img_backbone', x, 0, pc.depth)


def extract_img_feat(sd, pc, img_inputs, stages=None):
    """detectors/bevdet_occ.py:167-269 (with_prev, no depth gt, no BEV
    alignment).  Returns the encoded volume [B,32,Z,Y,X]."""
    imgs, s2k, e2g, intrins, post_rots, post_trans, bda, curr2adj = img_inputs
    geo = pc.geo
    bev_list, feat_prev_iv = [], None
    for fid in range(pc.num_frame - 1, -1, -1):
        extra_ref = fid == pc.num_frame - 1
        if extra_ref:
            feat_prev_iv = extract_stereo_ref_feat(sd, pc, imgs[fid])
            continue
        mlp_input = get_mlp_input(s2k[0], intrins[fid], post_rots[fid],
                                  post_trans[fid], bda)
        x, stereo_feat = image_encoder(sd, pc, imgs[fid])
        metas = dict(k2s_sensor=curr2adj[fid], intrins=intrins[fid],
                     post_rots=post_rots[fid], post_trans=post_trans[fid],
                     frustum=geo.cv_frustum, cv_downsample=4,
                     downsample=geo.downsample,
                     cv_feat_list=[feat_prev_iv, stereo_feat])
        bev, depth = view_transformer_forward(
            sd, 'img_view_transformer', geo,
            [x, s2k[fid], e2g[fid], intrins[fid], post_rots[fid],
             post_trans[fid], bda, mlp_input], metas, pc.cv_bias)
        if stages is not None:
            stages[f'img_feat_{fid}'] = x
            stages[f'stereo_feat_{fid}'] = stereo_feat
            stages[f'depth_{fid}'] = depth
            stages[f'lifted_{fid}'] = bev
        bev = custom_resnet3d(sd, 'pre_process_net', bev, pc.pre_layers,
                              pc.pre_stride)[0]
        bev_list.append(bev)
        feat_prev_iv = stereo_feat
    bev = torch.cat(bev_list, dim=1)          # [adj, key] order (:240,266)
    feats = custom_resnet3d(sd, 'img_bev_encoder_backbone', bev,
                            pc.enc_layers, pc.enc_stride)
    x = lss_fpn3d(sd, 'img_bev_encoder_neck', feats)
    if stages is not None:
        stages['bev_cat'] = bev
        stages['encoded'] = x
    return x


def voxel_features(sd, pc, inputs, stages=None):
    """preworld.py:166-169: final_conv (+implicit ReLU) -> [B,X,Y,Z,C]."""
    img_inputs = prepare_inputs(inputs, pc.num_frame, pc.num_frame - 1)
    x = extract_img_feat(sd, pc, img_inputs, stages)
    x = conv_module(sd, 'final_conv', x, 1, 1, norm=False)
    return x.permute(0, 4, 3, 2, 1)


def _attr_mlp(sd, p, x, final_softplus=False):
    x = _linear(sd, p + '.2', F.softplus(_linear(sd, p + '.0', x)))
    return F.softplus(x) if final_softplus else x


def occ_from_density(sd, pc, voxel_feats):
    """preworld.py:173-194."""
    density = _attr_mlp(sd, 'density_mlp', voxel_feats, True)[..., 0]
    semantic = _attr_mlp(sd, 'semantic_mlp', voxel_feats)
    no_empty = density > pc.test_threshold
    occ = torch.full(density.shape, pc.num_classes - 1, dtype=torch.long)
    occ[no_empty] = semantic.argmax(-1)[no_empty]
    geo_occ = torch.full(density.shape, pc.num_classes - 1, dtype=torch.long)
    geo_occ[no_empty] = 0
    return (occ.squeeze(0).numpy().astype(np.uint8),
            geo_occ.squeeze(0).numpy().astype(np.uint8))


def occ_from_head(sd, pc, voxel_feats, return_logits=False):
    """preworld.py:196-221 (Nuscenes)."""
    vf = voxel_feats[0].permute(3, 0, 1, 2).unsqueeze(0)
    logits = occ_head(sd, 'occupancy_head', vf)          # [1,18,X,Y,Z]
    occ_pred = logits.squeeze(0).permute(1, 2, 3, 0).argmax(-1)
    geo_occ = torch.full_like(occ_pred, pc.num_classes - 1)
    geo_occ[occ_pred != 17] = 0
    out = (occ_pred.numpy().astype(np.uint8), geo_occ.numpy().astype(np.uint8))
    return out + (logits,) if return_logits else out


def preworld_simple_test(sd, pc, inputs, stages=None):
    """PreWorld.simple_test, detectors/preworld.py:159-226."""
    vf = voxel_features(sd, pc, inputs, stages)
    if stages is not None:
        stages['voxel_feats'] = vf
    if pc.if_post_finetune:
        occ, geo_occ, logits = occ_from_head(sd, pc, vf, True)
        if stages is not None:
            stages['logits'] = logits
    else:
        occ, geo_occ = occ_from_density(sd, pc, vf)
    return {'semantic_occ': [occ], 'geo_occ': [geo_occ]}


def forecast_step(sd, voxel_feats, ego_states):
    """preworld_temporal_traj.py:329-341: plan_head -> broadcast -> cat ->
    fusion_head -> residual."""
    e = ego_states.reshape(ego_states.shape[0], -1)
    e = F.relu(_linear(sd, 'plan_head.0', e))
    e = F.relu(_linear(sd, 'plan_head.2', e))
    e = _linear(sd, 'plan_head.4', e)
    B, X, Y, Z, C = voxel_feats.shape
    e = e.view(B, 1, 1, 1, C).expand(B, X, Y, Z, C)
    upd = torch.cat([voxel_feats, e], dim=-1)
    res = _linear(sd, 'fusion_head.2',
                  F.softplus(_linear(sd, 'fusion_head.0', upd)))
    return res + voxel_feats


def plan_trajectory(sd, fused_voxel_feats, ego_states):
    """The planning branch of one forecasting step, preworld_temporal_traj.py:454-472
    with DownScaleModule3DCustom (heads/occupancy_head.py:180-200): fused voxel
    features [B,X,Y,Z,C] (reference layout) -> predicted displacement [B, 2]."""
    e = ego_states.reshape(ego_states.shape[0], -1)
    e = F.relu(_linear(sd, 'plan_head.0', e))
    e = F.relu(_linear(sd, 'plan_head.2', e))
    identity = _linear(sd, 'plan_head.4', e)
    x = fused_voxel_feats.permute(0, 4, 1, 2, 3).contiguous()
    for i in (1, 2, 3):
        x = F.conv3d(x, sd[f'downscale.downscale{i}.weight'],
                     sd[f'downscale.downscale{i}.bias'], stride=2)
    scene = F.adaptive_avg_pool3d(x, (1, 1, 1)).flatten(1)
    u = torch.cat([identity, scene], dim=-1)
    for i in (0, 2, 4):
        u = F.softplus(_linear(sd, f'ego_fusion_head.{i}', u))
    fused = identity + _linear(sd, 'ego_fusion_head.6', u)
    return _linear(sd, 'traj_head.2', F.softplus(_linear(sd, 'traj_head.0', fused)))


def preworld4d_simple_test(sd, pc, inputs, temporal_ego_states, stages=None):
    """PreWorld4DTraj.simple_test, preworld_temporal_traj.py:213-371.  Every
    step feeds ``temporal_ego_states[0]`` (:331), as the reference does."""
    vf = voxel_features(sd, pc, inputs, stages)
    res = {}
    first = 1 if pc.if_post_finetune else 2
    fn = occ_from_head if pc.if_post_finetune else occ_from_density
    if stages is not None:
        stages['voxel_feats'] = vf
        if pc.if_post_finetune:
            stages['logits'] = occ_from_head(sd, pc, vf, True)[2]
    occ, geo_occ = fn(sd, pc, vf)
    res['semantic_occ_0s'], res['geo_occ_0s'] = [occ], [geo_occ]
    for k in range(6):
        vf = forecast_step(sd, vf, temporal_ego_states[0])
        occ, geo_occ = fn(sd, pc, vf)
        res[f'semantic_occ_{k + first}s'] = [occ]
        res[f'geo_occ_{k + first}s'] = [geo_occ]
        if stages is not None:
            stages[f'voxel_feats_{k + first}s'] = vf
    return res


# ----------------------------------------------------- attribute proj + render
def attribute_projection(sd, voxel_feats):
    """preworld.py:251-254."""
    density = _attr_mlp(sd, 'density_mlp', voxel_feats, True)[..., 0]
    semantic = _attr_mlp(sd, 'semantic_mlp', voxel_feats)
    color = _attr_mlp(sd, 'color_mlp', voxel_feats)
    return density, semantic, color


class NerfGeometry:
    """NerfHead.__init__ constants, nerf/nerf_head.py:104-163."""

    def __init__(self, point_cloud_range, radius=39, step_size=0.5,
                 alpha_init=1e-6, fast_color_thres=1e-7):
        xyz_min = torch.Tensor(point_cloud_range[:3])
        xyz_max = torch.Tensor(point_cloud_range[3:])
        xyz_range = (xyz_max - xyz_min).float()
        self.bg_len = (xyz_range[0] // 2 - radius) / radius
        self.radius = radius
        self.scene_center = (xyz_min + xyz_max) * 0.5
        self.scene_radius = torch.Tensor([radius, radius, radius])
        self.step_size = step_size
        z_ = xyz_range[2] / xyz_range[0]
        self.xyz_min = torch.Tensor([-1 - self.bg_len, -1 - self.bg_len, -z_])
        self.xyz_max = torch.Tensor([1 + self.bg_len, 1 + self.bg_len, z_])
        self.act_shift = float(torch.FloatTensor(
            [np.log(1 / (1 - alpha_init) - 1)]))
        self.world_len = 200
        self.fast_color_thres = fast_color_thres


def sample_ray(ng, rays_o, rays_d, bda):
    """nerf/nerf_head.py:32-55."""
    rays_o = (rays_o - ng.scene_center) / ng.scene_radius
    rays_d = rays_d / rays_d.norm(dim=-1, keepdim=True)
    N_inner = int(2 / (2 + 2 * ng.bg_len) * ng.world_len / ng.step_size) + 1
    N_outer = N_inner // 15
    b_inner = torch.linspace(0, 2, N_inner + 1)
    b_outer = 2 / torch.linspace(1, 1 / 64, N_outer + 1)
    t = torch.cat([(b_inner[1:] + b_inner[:-1]) * 0.5,
                   (b_outer[1:] + b_outer[:-1]) * 0.5]).to(rays_o)
    ray_pts = rays_o[:, None, :] + rays_d[:, None, :] * t[None, :, None]
    norm = ray_pts.norm(dim=-1, keepdim=True)
    inner_mask = (norm <= 1)
    ray_pts = torch.where(inner_mask, ray_pts,
                          ray_pts / norm * ((1 + ng.bg_len) - ng.bg_len / norm))
    ray_pts = bda.matmul(ray_pts.unsqueeze(-1)).squeeze(-1)
    return ray_pts, inner_mask.squeeze(-1), t


def render_rays(ng, rays, bda, density, semantic, color):
    """One batch element of NerfHead.forward/render_one_scene + the three
    segment sums (nerf/nerf_head.py:165-269,332-353,361-407).  rays [R,16];
    density [X,Y,Z]; semantic [X,Y,Z,17]; color [X,Y,Z,3].  Returns per-kept-ray
    rendered depth, semantic, color, alphainv_last and the ray mask."""
    gt_depth = rays[:, 2].clone()
    gt_depth[gt_depth > 52] = 0
    mask = gt_depth > 0
    rays_o, rays_d = rays[:, 4:7][mask], rays[:, 7:10][mask]
    ray_pts, inner_mask, t = sample_ray(ng, rays_o, rays_d, bda)
    n_ray, n_step = ray_pts.shape[:2]
    ray_id = torch.arange(n_ray).view(-1, 1).expand(n_ray, n_step).flatten()
    m = inner_mask.clone()
    dist_thres = (2 + 2 * ng.bg_len) / ng.world_len * ng.step_size * 0.95
    dist = (ray_pts[:, 1:] - ray_pts[:, :-1]).norm(dim=-1)
    m[:, 1:] |= torch.from_numpy(
        c_ref.cumdist_thres(dist.numpy(), float(dist_thres)))
    ray_pts = ray_pts[m]
    t = t[None].repeat(n_ray, 1)[m]
    ray_id = ray_id[m.flatten()]
    xyz = ray_pts.reshape(1, 1, 1, -1, 3)
    ind_norm = ((xyz - ng.xyz_min) / (ng.xyz_max - ng.xyz_min)).flip((-1,)) \
        * 2 - 1
    dens = F.grid_sample(density[None, None], ind_norm, mode='bilinear',
                         align_corners=True).reshape(-1)
    sem = F.grid_sample(semantic.permute(3, 0, 1, 2)[None], ind_norm,
                        mode='bilinear', align_corners=True)
    ncls = sem.shape[1]
    sem = sem.reshape(ncls, -1).T
    col = F.grid_sample(color.permute(3, 0, 1, 2)[None], ind_norm,
                        mode='bilinear', align_corners=True).reshape(3, -1).T
    alpha = torch.from_numpy(c_ref.raw2alpha(dens.numpy(), ng.act_shift, 0.5))
    keep = alpha > ng.fast_color_thres
    t, ray_id, alpha, sem, col = t[keep], ray_id[keep], alpha[keep], \
        sem[keep], col[keep]
    w, last = c_ref.alpha2weight(alpha.numpy(), ray_id.numpy(), n_ray)
    w, last = torch.from_numpy(w), torch.from_numpy(last)
    keep = w > ng.fast_color_thres
    t, ray_id, w, sem, col = t[keep], ray_id[keep], w[keep], sem[keep], \
        col[keep]
    s = 1 - 1 / (1 + t)
    depth = torch.zeros(n_ray).index_add_(0, ray_id, w * s) + 1e-7
    r_sem = torch.zeros(n_ray, ncls).index_add_(0, ray_id, w[:, None] * sem)
    r_col = torch.zeros(n_ray, 3).index_add_(0, ray_id, w[:, None] * col)
    return dict(render_depth=depth * ng.radius, render_semantic=r_sem,
                render_color=r_col, alphainv_last=last, ray_mask=mask,
                n_samples=int(keep.sum()))


def nerf_compute_loss(res, target_depth, target_semantic, target_color, class_weights,
                      weight_depth=1.0, weight_semantic=1.0, weight_color=1.0,
                      weight_entropy_last=0.01, use_depth_sup=True):
    """NerfHead.compute_loss, nerf/nerf_head.py:271-291 (silog_loss / l1_loss,
    nerf/utils.py:71-87), without the distortion term.  ``res`` holds the MASKED rays'
    render_depth / render_semantic / render_color / alphainv_last."""
    out = {}
    if use_depth_sup:
        d = torch.log(res['render_depth'] + 1e-7) - torch.log(target_depth)
        out['loss_render_depth'] = torch.sqrt((d ** 2).mean() - 0.85 * d.mean() ** 2) * weight_depth
    crit = torch.nn.CrossEntropyLoss(weight=class_weights.type_as(res['render_semantic']),
                                     reduction='mean')
    out['loss_render_semantic'] = crit(res['render_semantic'], target_semantic.long()) * weight_semantic
    out['loss_render_color'] = torch.sum(torch.mean(torch.abs(res['render_color'] - target_color),
                                                    dim=0)) * weight_color
    if weight_entropy_last > 0:
        p = res['alphainv_last'].clamp(1e-6, 1 - 1e-6)
        out['loss_sdf_entropy'] = -(p * torch.log(p) + (1 - p) * torch.log(1 - p)).mean() * weight_entropy_last
    return out
