"""Image side: ``ResNet`` backbone and ``CustomFPN`` neck.

``ResNet`` mirrors mmdet 2.24.0 ``mmdet/models/backbones/resnet.py`` (pinned by
the reference's requirements.txt:15; call sites detectors/bevdet.py:10,38,
577-588) -- same constructor kwargs and parameter names (``conv1``, ``bn1``,
``layer{i}.{j}.conv{1,2,3}`` / ``bn{1,2,3}`` / ``downsample.{0,1}``), which are
also torchvision's.  ``CustomFPN`` mirrors reference necks/fpn.py:10-203.
All arithmetic runs in the C-ABI conv kernel with BatchNorm, residual add and
ReLU fused into its epilogue.
"""
import math

import torch
import torch.nn as nn

from .. import ops
from .base import BaseModule, ConvModule, pack_conv
from .builder import BACKBONES, NECKS


class _Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, inplanes, planes, stride, downsample, style):
        super().__init__()
        s1, s2 = (1, stride) if style == 'pytorch' else (stride, 1)
        self.conv1 = nn.Conv2d(inplanes, planes, 1, stride=s1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, stride=s2, padding=1,
                               bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.conv3 = nn.Conv2d(planes, planes * 4, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(planes * 4)
        self.downsample = downsample

    def pack(self):
        p = [pack_conv(self.conv1, self.bn1), pack_conv(self.conv2, self.bn2),
             pack_conv(self.conv3, self.bn3)]
        if self.downsample is not None:
            p.append(pack_conv(self.downsample[0], self.downsample[1]))
        return p

    @staticmethod
    def run(p, x, out=None):
        identity = ops.conv(x, p[3]) if len(p) == 4 else x
        y = ops.conv(x, p[0], 'relu')
        y = ops.conv(y, p[1], 'relu')
        return ops.conv(y, p[2], 'relu', residual=identity, out=out)


class _BasicBlock(nn.Module):
    expansion = 1

    def __init__(self, inplanes, planes, stride, downsample, style):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 3, stride=stride, padding=1,
                               bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.downsample = downsample

    def pack(self):
        p = [pack_conv(self.conv1, self.bn1), pack_conv(self.conv2, self.bn2)]
        if self.downsample is not None:
            p.append(pack_conv(self.downsample[0], self.downsample[1]))
        return p

    @staticmethod
    def run(p, x, out=None):
        identity = ops.conv(x, p[2]) if len(p) == 3 else x
        y = ops.conv(x, p[0], 'relu')
        return ops.conv(y, p[1], 'relu', residual=identity, out=out)


@BACKBONES.register_module()
class ResNet(BaseModule):
    arch_settings = {18: (_BasicBlock, (2, 2, 2, 2)),
                     34: (_BasicBlock, (3, 4, 6, 3)),
                     50: (_Bottleneck, (3, 4, 6, 3)),
                     101: (_Bottleneck, (3, 4, 23, 3)),
                     152: (_Bottleneck, (3, 8, 36, 3))}

    def __init__(self, depth, in_channels=3, stem_channels=None,
                 base_channels=64, num_stages=4, strides=(1, 2, 2, 2),
                 dilations=(1, 1, 1, 1), out_indices=(0, 1, 2, 3),
                 style='pytorch', deep_stem=False, avg_down=False,
                 frozen_stages=-1, conv_cfg=None,
                 norm_cfg=dict(type='BN', requires_grad=True), norm_eval=True,
                 dcn=None, stage_with_dcn=(False, False, False, False),
                 plugins=None, with_cp=False, zero_init_residual=True,
                 pretrained=None, init_cfg=None):
        super().__init__(init_cfg)
        if deep_stem or avg_down or dcn is not None or plugins is not None \
                or tuple(dilations) != (1, 1, 1, 1):
            raise NotImplementedError(
                'ResNet variant not used by the PreWorld configs')
        block, stage_blocks = self.arch_settings[depth]
        self.depth = depth
        self.deep_stem = deep_stem
        self.out_indices = tuple(out_indices)
        stem_channels = stem_channels or base_channels
        self.conv1 = nn.Conv2d(in_channels, stem_channels, 7, 2, 3, bias=False)
        self.norm1_name = 'bn1'
        self.bn1 = nn.BatchNorm2d(stem_channels)
        self.res_layers = []
        inplanes = stem_channels
        for i, nb in enumerate(stage_blocks[:num_stages]):
            planes = base_channels * 2 ** i
            blocks = []
            for b in range(nb):
                stride = strides[i] if b == 0 else 1
                downsample = None
                if b == 0 and (stride != 1
                               or inplanes != planes * block.expansion):
                    downsample = nn.Sequential(
                        nn.Conv2d(inplanes, planes * block.expansion, 1,
                                  stride=stride, bias=False),
                        nn.BatchNorm2d(planes * block.expansion))
                blocks.append(block(inplanes, planes, stride, downsample,
                                    style))
                inplanes = planes * block.expansion
            name = f'layer{i + 1}'
            self.add_module(name, nn.Sequential(*blocks))
            self.res_layers.append(name)

    @property
    def norm1(self):
        return self.bn1

    @property
    def stage0_is_stereo(self):
        """layer1's output is the stereo feature AND the input of layer2
        (what detectors.encode_frames batches over all frames)."""
        return 0 in self.out_indices

    def _build_packs(self):
        return dict(stem=pack_conv(self.conv1, self.bn1),
                    stem_s2d=self._pack_stem_s2d(),
                    layers=[[blk.pack() for blk in getattr(self, n)]
                            for n in self.res_layers])

    def _pack_stem_s2d(self):
        """The 7x7 / stride 2 / pad 3 stem as a 4x4 stride-1 conv over the
        space-to-depth(2) image (ops.nchw_to_s2d): out[Y,X] = sum_{ky,kx}
        W[ky,kx] x[2Y+ky-3, 2X+kx-3]; with ky-3 = 2a+dy (a in -2..1, dy in 0,1)
        this is sum_{a,b} W'[a,b] x'[Y+a, X+b], W'[o,(dy,dx,c),a,b] =
        W[o,c,2a+dy+3,2b+dx+3] (zero where that index leaves the 7x7 window).
        Cin is padded to 32, so the conv runs on the tensor-core kernel."""
        w = self.conv1.weight.detach().float()
        co, ci, kh, kw = w.shape
        if (kh, kw) != (7, 7) or ci > 4 or self.conv1.stride != (2, 2) or \
                self.conv1.padding != (3, 3):
            return None
        w2 = torch.zeros((co, 32, 4, 4), device=w.device)
        for a in range(-2, 2):
            for dy in range(2):
                ky = 2 * a + dy + 3
                if not 0 <= ky < 7:
                    continue
                for b in range(-2, 2):
                    for dx in range(2):
                        kx = 2 * b + dx + 3
                        if 0 <= kx < 7:
                            c0 = (dy * 2 + dx) * 4
                            w2[:, c0:c0 + ci, a + 2, b + 2] = w[:, :, ky, kx]
        from .base import bn_tuple
        return ops.PackedConv(w2, None, bn_tuple(self.bn1), stride=1, padding=2)

    # -- stage-level entry points (bevdet.py:577-588 runs stem + layer1 only
    #    for the stereo reference frame) ------------------------------------
    def run_stem(self, img_nchw):
        """NCHW image batch -> cl array after conv1/bn1/relu/maxpool."""
        return self.run_stem_cl(self.convert_images(img_nchw))

    def stem_input_shape(self, h, w):
        """(H, W, C) of the channels-last stem input for h x w images."""
        p = self.packs()
        if p['stem_s2d'] is not None and h % 2 == 0 and w % 2 == 0:
            return (h // 2, w // 2, 32)
        return (h, w, p['stem'].cin)

    def convert_images(self, img_nchw, out=None):
        """NCHW images -> the stem's channels-last input: space-to-depth(2)
        for the tensor-core stem, plain padded NHWC otherwise."""
        n, c, h, w = img_nchw.shape
        if self.stem_input_shape(h, w)[2] == 32 and \
                self.packs()['stem_s2d'] is not None and h % 2 == 0 and w % 2 == 0:
            return ops.nchw_to_s2d(img_nchw, 32, out=out)
        return ops.nchw_to_nhwc(img_nchw, self.packs()['stem'].cin, out=out)

    def run_stem_cl(self, x_cl):
        """stem input from ``convert_images`` -> conv1/bn1/relu/maxpool."""
        p = self.packs()
        if p['stem_s2d'] is not None and x_cl.shape[-1] == 32 and \
                p['stem'].cin != 32:
            x = ops.conv(x_cl, p['stem_s2d'], 'relu',
                         out_size=tuple(x_cl.shape[1:3]))
        else:
            x = ops.conv(x_cl, p['stem'], 'relu')
        return ops.maxpool3x3s2(x)

    def run_from_layer(self, i0, x):
        """layers i0.. on a cl array -> logical maps for out_indices >= i0."""
        outs = []
        for i in range(i0, len(self.res_layers)):
            x = self.run_layer(i, x)
            if i in self.out_indices:
                outs.append(ops.to_logical(x))
        return tuple(outs)

    def run_layer(self, i, x, out=None):
        """Residual stage i on a cl array; ``out`` receives the last block's
        output (e.g. one frame's rows of a batch other frames are written to)."""
        p = self.packs()
        blk = type(getattr(self, self.res_layers[i])[0])
        bps = p['layers'][i]
        for k, bp in enumerate(bps):
            x = blk.run(bp, x, out=out if k == len(bps) - 1 else None)
        return x

    def forward(self, x):
        """[N,3,H,W] -> tuple of logical [N,C,h,w] feature maps
        (channels_last strides) for ``out_indices``."""
        x = self.run_stem(x)
        outs = []
        for i in range(len(self.res_layers)):
            x = self.run_layer(i, x)
            if i in self.out_indices:
                outs.append(ops.to_logical(x))
        return tuple(outs)


@NECKS.register_module()
class CustomFPN(BaseModule):
    """necks/fpn.py:10-203 for the configuration the path uses: no norm, no
    activation, nearest top-down upsampling to the finer map's size, outputs
    selected by ``out_ids``."""

    def __init__(self, in_channels, out_channels, num_outs, start_level=0,
                 end_level=-1, out_ids=[], add_extra_convs=False,
                 relu_before_extra_convs=False, no_norm_on_lateral=False,
                 conv_cfg=None, norm_cfg=None, act_cfg=None,
                 upsample_cfg=dict(mode='nearest'),
                 init_cfg=dict(type='Xavier', layer='Conv2d',
                               distribution='uniform')):
        super().__init__(init_cfg)
        assert isinstance(in_channels, list)
        if norm_cfg is not None or act_cfg is not None or add_extra_convs \
                or upsample_cfg.get('mode') != 'nearest' \
                or 'scale_factor' in upsample_cfg:
            raise NotImplementedError(
                'CustomFPN variant not used by the PreWorld configs')
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.num_ins = len(in_channels)
        self.num_outs = num_outs
        self.out_ids = list(out_ids)
        self.backbone_end_level = self.num_ins if end_level == -1 \
            else end_level
        self.start_level = start_level
        assert num_outs <= len(self.out_ids) or num_outs == 1
        self.lateral_convs = nn.ModuleList()
        self.fpn_convs = nn.ModuleList()
        for i in range(self.start_level, self.backbone_end_level):
            self.lateral_convs.append(ConvModule(
                in_channels[i], out_channels, 1, act_cfg=None))
            if i in self.out_ids:
                self.fpn_convs.append(ConvModule(
                    out_channels, out_channels, 3, padding=1, act_cfg=None))

    def _build_packs(self):
        return dict(lat=[m.pack() for m in self.lateral_convs],
                    fpn=[m.pack() for m in self.fpn_convs])

    def forward(self, inputs):
        assert len(inputs) == len(self.in_channels)
        p = self.packs()
        lat = [ops.conv(ops.from_logical(inputs[i + self.start_level]), pc)
               for i, pc in enumerate(p['lat'])]
        for i in range(len(lat) - 1, 0, -1):
            ops.upsample_nearest_add_(lat[i - 1], lat[i])
        outs = [ops.to_logical(ops.conv(lat[i], p['fpn'][k]))
                for k, i in enumerate(self.out_ids)]
        return outs[0]


@NECKS.register_module()
class FPN_LSS(BaseModule):
    """necks/lss_fpn.py:13-99 -- the image neck of the shipped Swin config
    (bevstereo-occ.py:68-74): bilinear (align_corners) upsampling of the
    coarser map, concatenation with the finer one, two 3x3 conv + BN + ReLU
    (optionally ``up2``: one more upsampling + 3x3 conv + BN + ReLU + 1x1 conv).
    The upsampled map and the finer map are written straight into the two
    channel slices of the concatenation buffer."""

    def __init__(self, in_channels, out_channels, scale_factor=4,
                 input_feature_index=(0, 2), norm_cfg=dict(type='BN'),
                 extra_upsample=2, lateral=None, use_input_conv=False):
        super().__init__()
        from .base import build_norm_layer
        self.input_feature_index = input_feature_index
        self.extra_upsample = extra_upsample is not None
        self.scale_factor = scale_factor
        self.up = nn.Upsample(scale_factor=scale_factor, mode='bilinear',
                              align_corners=True)
        cf = 2 if self.extra_upsample else 1
        bn = lambda ch: build_norm_layer(norm_cfg, ch, postfix=0)[1]
        self.input_conv = nn.Sequential(
            nn.Conv2d(in_channels, out_channels * cf, 1, padding=0, bias=False),
            bn(out_channels * cf), nn.ReLU(inplace=True)) \
            if use_input_conv else None
        if use_input_conv:
            in_channels = out_channels * cf
        self.conv = nn.Sequential(
            nn.Conv2d(in_channels, out_channels * cf, 3, padding=1, bias=False),
            bn(out_channels * cf), nn.ReLU(inplace=True),
            nn.Conv2d(out_channels * cf, out_channels * cf, 3, padding=1,
                      bias=False),
            bn(out_channels * cf), nn.ReLU(inplace=True))
        if self.extra_upsample:
            self.extra_scale = extra_upsample
            self.up2 = nn.Sequential(
                nn.Upsample(scale_factor=extra_upsample, mode='bilinear',
                            align_corners=True),
                nn.Conv2d(out_channels * cf, out_channels, 3, padding=1,
                          bias=False),
                bn(out_channels), nn.ReLU(inplace=True),
                nn.Conv2d(out_channels, out_channels, 1, padding=0))
        self.lateral = lateral is not None
        if self.lateral:
            self.lateral_conv = nn.Sequential(
                nn.Conv2d(lateral, lateral, 1, padding=0, bias=False),
                bn(lateral), nn.ReLU(inplace=True))

    def _build_packs(self):
        p = dict(conv=[pack_conv(self.conv[0], self.conv[1]),
                       pack_conv(self.conv[3], self.conv[4])])
        if self.input_conv is not None:
            p['input'] = pack_conv(self.input_conv[0], self.input_conv[1])
        if self.extra_upsample:
            p['up2'] = [pack_conv(self.up2[1], self.up2[2]),
                        pack_conv(self.up2[4])]
        if self.lateral:
            p['lateral'] = pack_conv(self.lateral_conv[0], self.lateral_conv[1])
        return p

    @staticmethod
    def _upsample_into(out_slice, x):
        """bilinear, align_corners=True: the trilinear kernel on a depth-1 volume."""
        ops.upsample_trilinear_(out_slice[:, None], x[:, None])

    def forward(self, feats):
        p = self.packs()
        x2 = ops.from_logical(feats[self.input_feature_index[0]])
        x1 = ops.from_logical(feats[self.input_feature_index[1]])
        if self.lateral:
            x2 = ops.conv(x2, p['lateral'], 'relu')
        n, h1, w1, c1 = x1.shape
        oh, ow = int(math.floor(h1 * self.scale_factor)), \
            int(math.floor(w1 * self.scale_factor))
        c2 = x2.shape[-1]
        assert tuple(x2.shape[1:3]) == (oh, ow), (x2.shape, oh, ow)
        cat = torch.empty((n, oh, ow, c2 + c1), device=x1.device,
                          dtype=torch.float32)
        ops.copy_channels_(cat[..., :c2], x2)
        self._upsample_into(cat[..., c2:], x1)
        x = cat
        if self.input_conv is not None:
            x = ops.conv(x, p['input'], 'relu')
        x = ops.conv(ops.conv(x, p['conv'][0], 'relu'), p['conv'][1], 'relu')
        if self.extra_upsample:
            n, h, w, c = x.shape
            up = torch.empty((n, int(math.floor(h * self.extra_scale)),
                              int(math.floor(w * self.extra_scale)), c),
                             device=x.device, dtype=torch.float32)
            self._upsample_into(up, x)
            x = ops.conv(ops.conv(up, p['up2'][0], 'relu'), p['up2'][1])
        return ops.to_logical(x)
