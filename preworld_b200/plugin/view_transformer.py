"""``LSSViewTransformerBEVStereo``: DepthNet + stereo cost volume + voxel lift.

Mirrors reference necks/view_transformer.py (classes LSSViewTransformer :15,
LSSViewTransformerBEVDepth :702, LSSViewTransformerBEVStereo :807, DepthNet
:473, ASPP :355, Mlp :421, SELayer :440) -- same constructor kwargs, attribute
names read by the detectors (``D``, ``out_channels``, ``downsample``,
``grid_config``, ``cv_frustum``, ``grid_lower_bound``, ``grid_interval``,
``grid_size``, ``get_mlp_input``) and state_dict keys.  The forward launches
only C-ABI kernels:

* DepthNet convs / MLPs / SE gates  -> pw_conv_fwd (+ small element-wise ops)
* 64x grid_sample cost volume       -> pw_cost_volume (one kernel)
* softmax over D                    -> pw_softmax_depth
* get_lidar_coor + voxel_pooling_prepare_v2 + bev_pool_v2 + zero-fill + permute
                                    -> pw_lift_fused
"""
import torch
import torch.nn as nn

from .. import ops
from .base import BaseModule, bn_tuple, pack_conv, pack_linear
from .builder import NECKS


class Mlp(nn.Module):
    def __init__(self, in_features, hidden_features=None, out_features=None):
        super().__init__()
        self.fc1 = nn.Linear(in_features, hidden_features or in_features)
        self.fc2 = nn.Linear(hidden_features or in_features,
                             out_features or in_features)


class SELayer(nn.Module):
    def __init__(self, channels):
        super().__init__()
        self.conv_reduce = nn.Conv2d(channels, channels, 1, bias=True)
        self.conv_expand = nn.Conv2d(channels, channels, 1, bias=True)


class _ASPPModule(nn.Module):
    def __init__(self, inplanes, planes, kernel_size, padding, dilation):
        super().__init__()
        self.atrous_conv = nn.Conv2d(inplanes, planes, kernel_size, stride=1,
                                     padding=padding, dilation=dilation,
                                     bias=False)
        self.bn = nn.BatchNorm2d(planes)


class ASPP(nn.Module):
    def __init__(self, inplanes, mid_channels=256):
        super().__init__()
        dil = [1, 6, 12, 18]
        self.aspp1 = _ASPPModule(inplanes, mid_channels, 1, 0, dil[0])
        self.aspp2 = _ASPPModule(inplanes, mid_channels, 3, dil[1], dil[1])
        self.aspp3 = _ASPPModule(inplanes, mid_channels, 3, dil[2], dil[2])
        self.aspp4 = _ASPPModule(inplanes, mid_channels, 3, dil[3], dil[3])
        self.global_avg_pool = nn.Sequential(
            nn.AdaptiveAvgPool2d((1, 1)),
            nn.Conv2d(inplanes, mid_channels, 1, stride=1, bias=False),
            nn.BatchNorm2d(mid_channels), nn.ReLU())
        self.conv1 = nn.Conv2d(int(mid_channels * 5), inplanes, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(inplanes)
        self.mid_channels = mid_channels


class _DepthBasicBlock(nn.Module):
    """mmdet BasicBlock as instantiated at view_transformer.py:515-517."""

    def __init__(self, inplanes, planes, downsample=None):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 3, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.downsample = downsample


def _pad32(c):
    return (c + 31) // 32 * 32


class DepthNet(nn.Module):
    """Parameter container with the reference DepthNet's keys
    (view_transformer.py:473-544)."""

    def __init__(self, in_channels, mid_channels, context_channels,
                 depth_channels, use_dcn=True, use_aspp=True, with_cp=False,
                 stereo=False, bias=0.0, aspp_mid_channels=-1, D=100):
        super().__init__()
        if use_dcn:
            raise NotImplementedError('DCN is disabled in the PreWorld configs '
                                      '(bevstereo-occ.py:84)')
        if not (stereo and use_aspp):
            raise NotImplementedError('only the stereo + ASPP DepthNet is on '
                                      'the PreWorld path')
        self.reduce_conv = nn.Sequential(
            nn.Conv2d(in_channels, mid_channels, 3, stride=1, padding=1),
            nn.BatchNorm2d(mid_channels), nn.ReLU(inplace=True))
        self.context_conv = nn.Conv2d(mid_channels, context_channels, 1)
        self.bn = nn.BatchNorm1d(27)
        self.depth_mlp = Mlp(27, mid_channels, mid_channels)
        self.depth_se = SELayer(mid_channels)
        self.context_mlp = Mlp(27, mid_channels, mid_channels)
        self.context_se = SELayer(mid_channels)
        cin = mid_channels + depth_channels
        downsample = nn.Conv2d(cin, mid_channels, 1, 1, 0)
        net = []
        for _ in range(2):
            net.extend([nn.Conv2d(depth_channels, depth_channels, 3, stride=2,
                                  padding=1),
                        nn.BatchNorm2d(depth_channels)])
        self.cost_volumn_net = nn.Sequential(*net)
        self.bias = bias
        if aspp_mid_channels < 0:
            aspp_mid_channels = mid_channels
        self.depth_conv = nn.Sequential(
            _DepthBasicBlock(cin, mid_channels, downsample=downsample),
            _DepthBasicBlock(mid_channels, mid_channels),
            _DepthBasicBlock(mid_channels, mid_channels),
            ASPP(mid_channels, aspp_mid_channels),
            nn.Conv2d(mid_channels, depth_channels, 1, 1, 0))
        self.mid_channels = mid_channels
        self.depth_channels = depth_channels
        self.context_channels = context_channels

    def pack(self):
        bn = self.bn
        s = bn.weight.detach() / torch.sqrt(bn.running_var + bn.eps)
        t = bn.bias.detach() - bn.running_mean * s
        P = {}
        P['reduce'] = pack_conv(self.reduce_conv[0], self.reduce_conv[1])
        P['context'] = pack_conv(self.context_conv)
        for k in ('depth', 'context'):
            mlp, se = getattr(self, k + '_mlp'), getattr(self, k + '_se')
            # BatchNorm1d(27) in front of fc1 is folded into fc1
            P[k + '_fc1'] = pack_linear(mlp.fc1, in_scale=s, in_shift=t)
            P[k + '_fc2'] = pack_linear(mlp.fc2)
            P[k + '_red'] = pack_conv(se.conv_reduce)
            P[k + '_exp'] = pack_conv(se.conv_expand)
        # 88- and (mid+88)-channel tensors are carried zero-padded to a multiple
        # of 32 channels (zero weights on the padding) so these convs run on
        # the tensor-core kernels
        dpad = _pad32(self.depth_channels)
        cpad = self.mid_channels + dpad
        P['cvnet'] = [pack_conv(self.cost_volumn_net[i],
                                self.cost_volumn_net[i + 1], cin_pad=dpad,
                                cout_pad=dpad) for i in (0, 2)]
        blocks = []
        for j, b in enumerate(list(self.depth_conv)[:3]):
            kw = dict(cin_pad=cpad) if j == 0 else {}
            blocks.append(dict(
                c1=pack_conv(b.conv1, b.bn1, **kw),
                c2=pack_conv(b.conv2, b.bn2),
                down=pack_conv(b.downsample, **kw)
                if b.downsample is not None else None))
        P['blocks'] = blocks
        aspp = self.depth_conv[3]
        P['aspp'] = [pack_conv(m.atrous_conv, m.bn)
                     for m in (aspp.aspp1, aspp.aspp2, aspp.aspp3, aspp.aspp4)]
        P['aspp_gap'] = pack_conv(aspp.global_avg_pool[1],
                                  aspp.global_avg_pool[2])
        P['aspp_out'] = pack_conv(aspp.conv1, aspp.bn1)
        P['depth_out'] = pack_conv(self.depth_conv[4])
        return P


@NECKS.register_module()
class LSSViewTransformerBEVStereo(BaseModule):

    def __init__(self, grid_config, input_size, downsample=16,
                 in_channels=512, out_channels=64, accelerate=False, sid=False,
                 collapse_z=True, loss_depth_weight=3.0, depthnet_cfg=dict(),
                 init_cfg=None):
        super().__init__(init_cfg)
        if sid or collapse_z:
            raise NotImplementedError(
                'sid / collapse_z are off in the PreWorld configs '
                '(bevstereo-occ.py:76-89)')
        self.grid_config = grid_config
        self.input_size = tuple(input_size)
        self.downsample = downsample
        self.create_grid_infos(**grid_config)
        self.sid = sid
        self.frustum = self.create_frustum(grid_config['depth'], input_size,
                                           downsample)
        self.cv_frustum = self.create_frustum(grid_config['depth'],
                                              input_size, downsample=4)
        self.out_channels = out_channels
        self.in_channels = in_channels
        self.accelerate = accelerate
        self.collapse_z = collapse_z
        self.loss_depth_weight = loss_depth_weight
        self.depth_net = DepthNet(in_channels, in_channels, out_channels,
                                  self.D, **depthnet_cfg)
        self._consts = {}
        # accelerate=True (view_transformer.py:31-33,263-267): constant cameras,
        # the voxel lists are built by the first forward and reused
        self.initial_flag = True
        self._accel_ws = None

    # -- reference-visible geometry attributes ------------------------------
    def create_grid_infos(self, x, y, z, **kwargs):
        """view_transformer.py:66-82."""
        self.grid_lower_bound = torch.Tensor([c[0] for c in [x, y, z]])
        self.grid_interval = torch.Tensor([c[2] for c in [x, y, z]])
        self.grid_size = torch.Tensor([(c[1] - c[0]) / c[2]
                                       for c in [x, y, z]])

    def create_frustum(self, depth_cfg, input_size, downsample):
        """view_transformer.py:84-112 (sid=False).  Returns [D,H,W,3] and sets
        ``self.D`` like the reference."""
        H_in, W_in = input_size
        H_feat, W_feat = H_in // downsample, W_in // downsample
        d = torch.arange(*depth_cfg, dtype=torch.float) \
            .view(-1, 1, 1).expand(-1, H_feat, W_feat)
        self.D = d.shape[0]
        x = torch.linspace(0, W_in - 1, W_feat, dtype=torch.float) \
            .view(1, 1, W_feat).expand(self.D, H_feat, W_feat)
        y = torch.linspace(0, H_in - 1, H_feat, dtype=torch.float) \
            .view(1, H_feat, 1).expand(self.D, H_feat, W_feat)
        return torch.stack((x, y, d), -1)

    # -- training side ---------------------------------------
    def get_depth_loss(self, depth_labels, depth_preds):
        """view_transformer.py:774-789 (+ get_downsampled_gt_depth, :736-771):
        one fused kernel (csrc/losses.cu) instead of ~20 tensor ops;
        differentiable w.r.t. ``depth_preds``."""
        from .. import losses
        return losses.get_depth_loss(depth_labels, depth_preds, self.downsample,
                                     self.grid_config['depth'],
                                     self.loss_depth_weight)

    def get_downsampled_gt_depth(self, gt_depths):
        """view_transformer.py:736-771: [B,N,H,W] -> one-hot [B*N*h*w, D] float
        (all-zero rows = background), from the bin labels the kernel computes."""
        B, N, H, W = gt_depths.shape
        dummy = torch.full((B * N, self.D, H // self.downsample,
                            W // self.downsample), 0.5, device=gt_depths.device)
        _, labels, _ = ops.depth_loss(gt_depths.reshape(B * N, H, W), dummy,
                                      self.downsample, self.grid_config['depth'][0],
                                      self.grid_config['depth'][2], 1.0)
        lab = labels.long()
        onehot = torch.zeros((lab.numel(), self.D), device=lab.device)
        fg = lab >= 0
        onehot[fg, lab[fg]] = 1.0
        return onehot

    def get_mlp_input(self, sensor2ego, ego2global, intrin, post_rot,
                      post_tran, bda):
        """view_transformer.py:713-734 (pure indexing: 27 floats/camera)."""
        B, N, _, _ = sensor2ego.shape
        bda = bda.view(B, 1, 3, 3).repeat(1, N, 1, 1)
        mlp_input = torch.stack([
            intrin[:, :, 0, 0], intrin[:, :, 1, 1], intrin[:, :, 0, 2],
            intrin[:, :, 1, 2], post_rot[:, :, 0, 0], post_rot[:, :, 0, 1],
            post_tran[:, :, 0], post_rot[:, :, 1, 0], post_rot[:, :, 1, 1],
            post_tran[:, :, 1], bda[:, :, 0, 0], bda[:, :, 0, 1],
            bda[:, :, 1, 0], bda[:, :, 1, 1], bda[:, :, 2, 2]], dim=-1)
        sensor2ego = sensor2ego[:, :, :3, :].reshape(B, N, -1)
        return torch.cat([mlp_input, sensor2ego], dim=-1)

    # -- packing --------------------------------------------------------------
    def _build_packs(self):
        return self.depth_net.pack()

    def _frustum_axes(self, frustum, device):
        key = (id(frustum), device)
        c = self._consts.get(key)
        if c is None:
            xs = frustum[0, 0, :, 0].contiguous().to(device)
            ys = frustum[0, :, 0, 1].contiguous().to(device)
            ds = frustum[:, 0, 0, 2].contiguous().to(device)
            c = self._consts[key] = (xs, ys, ds)
        return c

    # -- DepthNet (view_transformer.py:606-638) -------------------------------
    def _se_gates(self, P, mlp_in):
        """sigmoid(conv_expand(relu(conv_reduce(Mlp(bn(mlp_input)))))) of the context and
        the depth branch (view_transformer.py:609-617): both 4-layer chains in ONE
        launch instead of eight 12-row GEMMs."""
        chain = lambda k: [(P[k + '_fc1'], 'relu'), (P[k + '_fc2'], None),
                           (P[k + '_red'], 'relu'), (P[k + '_exp'], 'sigmoid')]
        return ops.dense_chains(mlp_in, [chain('context'), chain('depth')])

    def _cost_volume(self, P, metas, BN, H, W, device, out=None):
        dn = self.depth_net
        prev, curr = metas['cv_feat_list']
        if prev is None:
            # view_transformer.py:619-625: all-zero cost volume
            s = float(metas['downsample']) / metas['cv_downsample']
            cv = torch.zeros((BN, int(H * s), int(W * s),
                              _pad32(dn.depth_channels)),
                             device=device, dtype=torch.float32)
        else:
            xs, ys, ds = self._frustum_axes(metas['frustum'], device)
            cam = ops.cv_camera_params(metas['k2s_sensor'], metas['intrins'],
                                       metas['post_rots'],
                                       metas['post_trans'])
            curr_cl, prev_cl = ops.from_logical(curr), ops.from_logical(prev)
            hf, wf = curr_cl.shape[1:3]
            cv = ops.cost_volume(curr_cl, prev_cl, cam, xs, ys, ds, dn.bias,
                                 (hf * 4, wf * 4),
                                 pad_to=_pad32(dn.depth_channels))
        cv = ops.conv(cv, P['cvnet'][0])
        return ops.conv(cv, P['cvnet'][1], out=out)

    def _depth_net(self, P, x, mlp_input, metas):
        dn = self.depth_net
        BN, H, W, _ = x.shape
        mlp_in = mlp_input.reshape(-1, mlp_input.shape[-1]).contiguous()
        mlp_in = torch.nn.functional.pad(mlp_in, (0, 1))   # 27 -> 28 (cin%4)
        x = ops.conv(x, P['reduce'], 'relu')
        mid = dn.mid_channels
        out = torch.empty((BN, H, W, dn.depth_channels + dn.context_channels),
                          device=x.device, dtype=torch.float32)
        gate_ctx, gate_depth = self._se_gates(P, mlp_in)
        # context branch
        ctx = ops.scale_channels(x, gate_ctx)
        ops.conv(ctx, P['context'], out=out[..., dn.depth_channels:])
        # depth branch: cat([gated x, cost volume]) built in place
        cat = torch.empty((BN, H, W, mid + _pad32(dn.depth_channels)),
                          device=x.device, dtype=torch.float32)
        ops.scale_channels(x, gate_depth, out=cat[..., :mid])
        # cost_volumn_net's last conv writes its (zero-padded) output in place
        self._cost_volume(P, metas, BN, H, W, x.device, out=cat[..., mid:])
        d = cat
        for b in P['blocks']:
            identity = ops.conv(d, b['down']) if b['down'] is not None else d
            y = ops.conv(d, b['c1'], 'relu')
            d = ops.conv(y, b['c2'], 'relu', residual=identity)
        # ASPP (view_transformer.py:355-418; dropout is an eval no-op)
        am = dn.depth_conv[3].mid_channels
        acat = torch.empty((BN, H, W, am * 5), device=x.device,
                           dtype=torch.float32)
        for i, pc in enumerate(P['aspp']):
            ops.conv(d, pc, 'relu', out=acat[..., i * am:(i + 1) * am])
        g = ops.linear(ops.global_avgpool(d), P['aspp_gap'], 'relu')
        ops.broadcast_channels_(acat[..., 4 * am:], g)
        d = ops.conv(acat, P['aspp_out'], 'relu')
        ops.conv(d, P['depth_out'], out=out[..., :dn.depth_channels])
        return out

    # -- forward (view_transformer.py:791-804) --------------------------------
    def depth_stage(self, x, mlp_input, stereo_metas):
        """DepthNet + softmax over D for the cameras in ``x`` [B,n,C,H,W]
        (per-camera work: this is what a camera shard runs locally).
        -> depth [B*n,D,H,W], context features cl [B*n,H,W,C_out] (a channel
        slice of the DepthNet output)."""
        P = self.packs()
        B, N, C, H, W = x.shape
        x_cl = ops.from_logical(x.reshape(B * N, C, H, W))
        feat = self._depth_net(P, x_cl, mlp_input, stereo_metas)
        depth = ops.softmax_depth(feat, self.D)          # [B*N, D, H, W]
        return depth, feat[..., self.D:self.D + self.out_channels]

    def lift_stage(self, depth, tran_feat, sensor2keyego, intrin, post_rot,
                   post_tran, bda, B, N):
        """Voxel lift of ALL cameras (view_transformer.py:114-153,176-261)."""
        dev = depth.device
        cam = ops.lift_camera_params(sensor2keyego, intrin, post_rot,
                                     post_tran)
        xs, ys, ds = self._frustum_axes(self.frustum, dev)
        grid = tuple(int(g) for g in self.grid_size)
        bda9 = bda.reshape(B, 9).contiguous().float()
        if self.accelerate:
            if self.initial_flag:                       # pre_compute(), :263-267
                self._accel_ws = ops.lift_prepare(
                    cam, bda9, xs, ys, ds, self.grid_lower_bound.tolist(),
                    self.grid_interval.tolist(), B, N, grid)
                self.initial_flag = False
            bev = ops.lift_pool(depth, tran_feat, self._accel_ws, B, N, grid)
        else:
            bev = ops.lift_fused(
                depth, tran_feat, cam, bda9, xs, ys, ds,
                self.grid_lower_bound.tolist(), self.grid_interval.tolist(),
                B, N, grid)
        return ops.to_logical(bev)

    def forward(self, input, stereo_metas=None, depth_gt=None):
        (x, sensor2keyego, ego2global, intrin, post_rot, post_tran, bda,
         mlp_input) = input[:8]
        B, N = x.shape[:2]
        depth, tran_feat = self.depth_stage(x, mlp_input, stereo_metas)
        bev = self.lift_stage(depth, tran_feat, sensor2keyego, intrin,
                              post_rot, post_tran, bda, B, N)
        return bev, depth
