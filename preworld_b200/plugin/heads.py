"""``OccHead`` (+ ``DownScaleModule3DCustom``) and ``NerfHead``.

OccHead mirrors reference heads/occupancy_head.py:45-177 for the configuration
the PreWorld configs use (num_level=1, use_deblock=False, soft_weights=True,
SyncBN).  With one level the soft-weight branch is a softmax over ONE channel
(== 1.0 exactly, :142-144) and the same-size trilinear interpolate is the
identity, so the forward is conv3^3+BN+ReLU -> 1x1x1+BN+ReLU -> 1x1x1; the
``voxel_soft_weights`` parameters are kept for checkpoint compatibility.

NerfHead mirrors reference nerf/nerf_head.py:104-163 (constructor, buffers) and
renders with the fused warp-per-ray kernel (pw_render_rays) instead of
sample_ray + cumdist_thres + 3x grid_sample + Raw2Alpha + Alphas2Weights +
segment_coo (:32-55,165-269,332-353).
"""
import ctypes

import numpy as np
import torch
import torch.nn as nn

from .. import ops
from .._lib import RenderDesc
from .base import BaseModule, build_conv_layer, build_norm_layer, pack_conv
from .builder import HEADS


def _is_reversed_cl(t):
    """True if logical [B,C,X,Y,Z] tensor ``t`` is a view of library-order
    memory [B,Z,Y,X,C] (what the B200 detectors pass around)."""
    v = t.permute(0, 4, 3, 2, 1)
    if v.stride(-1) != 1:
        return False
    try:
        ops.cl_ld(v)
        return True
    except ValueError:
        return False


@HEADS.register_module()
class OccHead(BaseModule):
    def __init__(self, in_channels, out_channel, num_level=1,
                 soft_weights=False, conv_cfg=dict(type='Conv3d', bias=False),
                 norm_cfg=dict(type='GN', num_groups=32, requires_grad=True),
                 point_cloud_range=[-40., -40., -1., 40., 40., 5.4],
                 final_occ_size=[200, 200, 16], empty_idx=17,
                 balance_cls_weight=True, with_cp=False, use_deblock=False):
        super().__init__()
        if type(in_channels) is not list:
            in_channels = [in_channels]
        if num_level != 1 or use_deblock or norm_cfg['type'] not in (
                'SyncBN', 'BN3d', 'BN'):
            raise NotImplementedError(
                'OccHead variant not used by the PreWorld configs '
                '(num_level=1, use_deblock=False, SyncBN)')
        self.with_cp = with_cp
        self.use_deblock = use_deblock
        self.in_channels = in_channels
        self.out_channel = out_channel
        self.num_level = num_level
        self.point_cloud_range = torch.tensor(
            np.array(point_cloud_range)).float()
        norm_cfg = dict(norm_cfg, type='BN3d')
        self.occ_convs = nn.ModuleList()
        mid = in_channels[0] // 2
        self.occ_convs.append(nn.Sequential(
            build_conv_layer(conv_cfg, in_channels=in_channels[0],
                             out_channels=mid, kernel_size=3, stride=1,
                             padding=1),
            build_norm_layer(norm_cfg, mid)[1], nn.ReLU(inplace=True)))
        self.occ_pred_conv = nn.Sequential(
            build_conv_layer(conv_cfg, in_channels=mid, out_channels=mid // 2,
                             kernel_size=1, stride=1, padding=0),
            build_norm_layer(norm_cfg, mid // 2)[1], nn.ReLU(inplace=True),
            build_conv_layer(conv_cfg, in_channels=mid // 2,
                             out_channels=out_channel, kernel_size=1,
                             stride=1, padding=0))
        self.soft_weights = soft_weights
        self.num_point_sampling_feat = num_level
        if soft_weights:
            self.voxel_soft_weights = nn.Sequential(
                build_conv_layer(conv_cfg, in_channels=mid,
                                 out_channels=mid // 2, kernel_size=1,
                                 stride=1, padding=0),
                build_norm_layer(norm_cfg, mid // 2)[1],
                nn.ReLU(inplace=True),
                build_conv_layer(conv_cfg, in_channels=mid // 2,
                                 out_channels=self.num_point_sampling_feat,
                                 kernel_size=1, stride=1, padding=0))
        self.empty_idx = empty_idx

    def _build_packs(self):
        oc, pc = self.occ_convs[0], self.occ_pred_conv
        p0, p1 = pack_conv(pc[0], pc[1]), pack_conv(pc[3])
        mid, ncls = pc[0].out_channels, pc[3].out_channels
        # pw_occhead_tail: plain [out, in] matrices + the folded BatchNorm affine
        tail = (pc[0].weight.detach().float().reshape(mid, -1).contiguous(),
                p0.scale, p0.bias,
                pc[3].weight.detach().float().reshape(ncls, mid).contiguous(),
                p1.bias)
        return dict(
            c0=pack_conv(oc[0], oc[1]),
            c0_rev=pack_conv(oc[0], oc[1], spatial_perm=(2, 1, 0)),
            p0=p0, p1=p1, tail=tail)

    def occupancy_pair_cl(self, x_cl, free_idx, geo_value, out=None):
        """cl array [1,Z,Y,X,C] (library order) -> uint8 [2,X,Y,Z]: class argmax
        and geometry grid (preworld.py:196-221), two launches: occ_convs[0] on the
        tensor cores, then the fused 16->8->18->argmax tail (no 8- / 18-channel
        tensors in HBM)."""
        P = self.packs()
        y = ops.conv(x_cl, P['c0_rev'], 'relu')
        return ops.occhead_tail(y, *P['tail'], free_idx, geo_value, out=out)

    def logits_cl(self, x_cl, reversed_order=True):
        """cl array [B,Z,Y,X,C] (library order; weights spatially transposed)
        or [B,X,Y,Z,C] (reference order) -> class logits cl array."""
        P = self.packs()
        y = ops.conv(x_cl, P['c0_rev' if reversed_order else 'c0'], 'relu')
        y = ops.conv(y, P['p0'], 'relu')
        return ops.conv(y, P['p1'])

    def forward(self, voxel_feats, **kwargs):
        assert type(voxel_feats) is list and len(voxel_feats) == self.num_level
        x = voxel_feats[0]                    # logical [B,C,X,Y,Z]
        if _is_reversed_cl(x):
            lg = self.logits_cl(x.permute(0, 4, 3, 2, 1), True)
            out = lg.permute(0, 4, 3, 2, 1)
        else:
            lg = self.logits_cl(ops.from_logical(x), False)
            out = ops.to_logical(lg)
        return {'output_voxels': [out]}


class DownScaleModule3DCustom(BaseModule):
    """heads/occupancy_head.py:180-200: three ``Conv3d(k=2, s=2)`` (C -> 2C -> 4C -> 4C)
    and a global average pool -- the scene descriptor of the planning branch
    (preworld_temporal_traj.py:464-470).  Same parameter keys as the reference
    (``downscale{1,2,3}.{weight,bias}``)."""

    def __init__(self, in_dim):
        super().__init__()
        self.in_dim = in_dim
        self.downscale1 = nn.Conv3d(in_dim, in_dim * 2, 2, stride=2)
        self.downscale2 = nn.Conv3d(in_dim * 2, in_dim * 4, 2, stride=2)
        self.downscale3 = nn.Conv3d(in_dim * 4, in_dim * 4, 2, stride=2)

    def _build_packs(self):
        convs = (self.downscale1, self.downscale2, self.downscale3)
        return dict(ref=[pack_conv(c) for c in convs],
                    # the library's volumes are [Z,Y,X]: spatially transposed kernels
                    rev=[pack_conv(c, spatial_perm=(2, 1, 0)) for c in convs])

    def pooled_cl(self, x_cl, reversed_order=True):
        """cl array [B,Z,Y,X,C] (library order) or [B,X,Y,Z,C] -> [B, 4C]."""
        P = self.packs()['rev' if reversed_order else 'ref']
        for pc in P:
            x_cl = ops.conv(x_cl, pc)
        return ops.global_avgpool(x_cl)

    def forward(self, feats):
        """feats [b, X, Y, Z, C] (the reference's layout) -> [b, 1, 1, 1, 4C]."""
        b = feats.shape[0]
        return self.pooled_cl(feats.contiguous(), False).view(b, 1, 1, 1, -1)


# occ3d-nuscenes class frequencies (nerf_head.py:22-24)
NUSC_CLASS_FREQUENCIES = np.array([
    1163161, 2309034, 188743, 2997643, 20317180, 852476, 243808, 2457947, 497017, 2731022,
    7224789, 214411435, 5565043, 63191967, 76098082, 128860031, 141625221, 2307405309])


@HEADS.register_module()
class NerfHead(nn.Module):
    def __init__(self, point_cloud_range, voxel_size, scene_center=None,
                 radius=39, step_size=0.5, use_depth_sup=True,
                 balance_cls_weight=True, weight_depth=1.0,
                 weight_semantic=1.0, weight_color=1.0,
                 weight_entropy_last=0.01, weight_distortion=0.01,
                 alpha_init=1e-6, fast_color_thres=1e-7):
        super().__init__()
        self.weight_entropy_last = weight_entropy_last
        self.weight_distortion = weight_distortion
        xyz_min = torch.Tensor(point_cloud_range[:3])
        xyz_max = torch.Tensor(point_cloud_range[3:])
        xyz_range = (xyz_max - xyz_min).float()
        self.bg_len = (xyz_range[0] // 2 - radius) / radius
        self.radius = radius
        # `scene_center` kwarg is ignored by the reference too (:134)
        self.register_buffer('scene_center', (xyz_min + xyz_max) * 0.5)
        self.register_buffer('scene_radius',
                             torch.Tensor([radius, radius, radius]))
        self.step_size = step_size
        self.use_depth_sup = use_depth_sup
        z_ = xyz_range[2] / xyz_range[0]
        self.register_buffer('xyz_min', torch.Tensor(
            [-1 - self.bg_len, -1 - self.bg_len, -z_]))
        self.register_buffer('xyz_max', torch.Tensor(
            [1 + self.bg_len, 1 + self.bg_len, z_]))
        self.alpha_init = alpha_init
        self.register_buffer('act_shift', torch.FloatTensor(
            [np.log(1 / (1 - alpha_init) - 1)]))
        self.voxel_size = voxel_size / radius
        self.world_size = torch.Tensor([200, 200, 16]).long()
        self.world_len = self.world_size[0].item()
        self.fast_color_thres = fast_color_thres
        self.weight_depth = weight_depth
        self.weight_semantic = weight_semantic
        self.weight_color = weight_color
        # nerf_head.py:160-163
        self.class_weights = (torch.from_numpy(1 / np.log(NUSC_CLASS_FREQUENCIES[:17] + 0.001))
                              if balance_cls_weight else torch.ones(17) / 17)
        self._t_cache = {}

    def ray_parameters(self, device):
        """t of sample_ray (nerf_head.py:36-44), computed once with the
        reference's own expressions: 391 inner + 26 outer mid-points."""
        t = self._t_cache.get(device)
        if t is None:
            n_inner = int(2 / (2 + 2 * self.bg_len) * self.world_len
                          / self.step_size) + 1
            n_outer = n_inner // 15
            b_inner = torch.linspace(0, 2, n_inner + 1)
            b_outer = 2 / torch.linspace(1, 1 / 64, n_outer + 1)
            t = torch.cat([(b_inner[1:] + b_inner[:-1]) * 0.5,
                           (b_outer[1:] + b_outer[:-1]) * 0.5])
            t = self._t_cache[device] = t.to(device).contiguous()
        return t

    def _desc(self, n_sem, strides):
        d = RenderDesc()
        for name in ('scene_center', 'scene_radius', 'xyz_min', 'xyz_max'):
            v = getattr(self, name).detach().cpu().tolist()
            setattr(d, name, (ctypes.c_float * 3)(*v))
        d.bg_len = float(self.bg_len)
        d.act_shift = float(self.act_shift.detach().cpu()[0])
        d.interval = 0.5            # activate_density(density, interval=0.5)
        d.step_size = float(self.step_size)
        d.fast_color_thres = float(self.fast_color_thres)
        d.radius = float(self.radius)
        d.max_depth = 52.0          # nerf_head.py:382
        d.world_len = int(self.world_len)
        d.gx, d.gy, d.gz = [int(v) for v in self.world_size]
        d.n_sem = n_sem
        d.vs_x, d.vs_y, d.vs_z = strides
        return d

    def render(self, density, semantic, color, rays, bda, library_order=False):
        """One batch element.  Reference layout: density [X,Y,Z], semantic
        [X,Y,Z,17], color [X,Y,Z,3]; with ``library_order`` the voxel order is
        [Z,Y,X] (channel slices of one attribute buffer are fine).
        Returns a dict with per-ray render_depth / render_semantic /
        render_color / alphainv_last and the ray mask (0 < depth <= 52)."""
        gx, gy, gz = [int(v) for v in self.world_size]
        strides = (1, gx, gx * gy) if library_order else (gy * gz, gz, 1)
        if density.dim() == 3:
            density = density[..., None]
        desc = self._desc(semantic.shape[-1], strides)
        t = self.ray_parameters(rays.device)
        d, s, c, last, valid = ops.render_rays(
            desc, rays, t, bda.reshape(9).float(), density, semantic, color)
        return dict(render_depth=d, render_semantic=s, render_color=c,
                    alphainv_last=last, ray_mask=valid)

    def compute_loss(self, results, rays, interval=None):
        """nerf_head.py:271-291 (``interval`` given: compute_loss_temporal, :301-329,
        the same terms under ``_{interval}s`` keys): the renderings of ONE sample
        reduced to loss values by one kernel (nine fp64 sums) + a few scalar
        operations.  Values only -- the fused ray march has no backward pass.  The
        distortion term (flatten_eff_distloss over every sample's weight) needs the
        per-sample weights the fused march never materialises and is not produced."""
        s = ops.render_loss_sums(rays, results['render_depth'], results['render_semantic'],
                                 results['render_color'], results['alphainv_last'],
                                 results['ray_mask'], self.class_weights.to(rays.device))
        n = s[0]
        sfx = '' if interval is None else f'_{int(interval)}s'
        out = {}
        if self.use_depth_sup:
            silog = torch.sqrt(s[2] / n - 0.85 * (s[1] / n) ** 2)       # utils.py:76-78
            out['loss_render_depth' + sfx] = (silog * self.weight_depth).float()
        out['loss_render_semantic' + sfx] = (s[3] / s[4] * self.weight_semantic).float()
        out['loss_render_color' + sfx] = ((s[5] + s[6] + s[7]) / n * self.weight_color).float()
        if self.weight_entropy_last > 0:
            out['loss_sdf_entropy' + sfx] = (-(s[8] / n) * self.weight_entropy_last).float()
        return out

    def forward(self, density, semantic, color, if_pretrain=False,
                if_temporal=False, dataset_type='Nuscenes', rays=None,
                bda=None, interval=0, library_order=False, return_loss=False,
                **kwargs):
        """nerf_head.py:361-420: renders every batch element; returns the per-ray
        renderings (list of dicts, one per sample) or, with ``return_loss=True``, the
        reference's dict of losses averaged over the batch (:410-418)."""
        if dataset_type != 'Nuscenes':
            raise NotImplementedError('only the nuScenes ray layout is on the path')
        renders = [self.render(density[b], semantic[b], color[b], rays[b],
                               bda[b], library_order)
                   for b in range(rays.shape[0])]
        if not return_loss:
            return renders
        losses = {}
        for b, res in enumerate(renders):
            one = self.compute_loss(res, rays[b], interval if if_temporal else None)
            for k, v in one.items():
                losses[k] = losses[k] + v if k in losses else v
        return {k: v / len(renders) for k, v in losses.items()}
