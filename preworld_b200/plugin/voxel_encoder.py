"""3-D occupancy encoder: ``CustomResNet3D`` and ``LSSFPN3D``.

Mirrors reference backbones/resnet.py:88-184 (BasicBlock3D, CustomResNet3D)
and necks/lss_fpn.py:103-148 (LSSFPN3D): same kwargs, same state_dict keys
(``layers.{i}.{j}.conv{1,2}.{conv,bn}``, ``layers.{i}.0.downsample.{conv,bn}``,
``conv.{conv,bn}``).  Volumes are channels-last [B,Z,Y,X,C]; the logical
tensors exchanged with callers are [B,C,Z,Y,X] with channels_last_3d strides.

Fusions relative to the reference graph:
* conv1 (+BN+ReLU) and the 3^3 ``downsample`` shortcut (+BN) of the first
  block of every stage read the same input with the same stride -> ONE launch
  with 2*Cout output channels (ReLU on the first half only);
* BN, the residual add and the final ReLU of a block live in the epilogue of
  conv2;
* LSSFPN3D: the x2 / x4 trilinear up-samplings write straight into channel
  slices of the concatenation buffer read by the 1x1x1 conv.
"""
import torch
import torch.nn as nn

from .. import ops
from .base import BaseModule, ConvModule
from .builder import BACKBONES, NECKS

_BN3D = dict(type='BN3d')
_C3D = dict(type='Conv3d')


class BasicBlock3D(nn.Module):
    def __init__(self, channels_in, channels_out, stride=1, downsample=None):
        super().__init__()
        self.conv1 = ConvModule(channels_in, channels_out, 3, stride=stride,
                                padding=1, bias=False, conv_cfg=_C3D,
                                norm_cfg=_BN3D, act_cfg=dict(type='ReLU'))
        self.conv2 = ConvModule(channels_out, channels_out, 3, stride=1,
                                padding=1, bias=False, conv_cfg=_C3D,
                                norm_cfg=_BN3D, act_cfg=None)
        self.downsample = downsample
        self.channels_out = channels_out

    def pack(self):
        c1, c2 = self.conv1.pack(), self.conv2.pack()
        if self.downsample is None:
            return dict(c1=c1, c2=c2, fused=False)
        ds = self.downsample.pack()
        # one conv with [conv1 | downsample] output channels
        f = ops.PackedConv.__new__(ops.PackedConv)
        f.w = torch.cat([c1.w[:, :c1.cout], ds.w[:, :ds.cout]], 1).contiguous()
        f.scale = torch.cat([c1.scale, ds.scale]).contiguous()
        f.bias = torch.cat([c1.bias, ds.bias]).contiguous()
        f.cin, f.cout, f.k, f.w_ld = c1.cin, c1.cout + ds.cout, c1.k, \
            c1.cout + ds.cout
        f.stride, f.pad, f.dil = c1.stride, c1.pad, c1.dil
        f.wt_hi = f.wt_lo = None
        if c1.wt_hi is not None and ds.wt_hi is not None:
            f.wt_hi = torch.cat([c1.wt_hi, ds.wt_hi], 0).contiguous()
            f.wt_lo = torch.cat([c1.wt_lo, ds.wt_lo], 0).contiguous()
        # x-tap-folded layout: rows are (slab of 32 channels, kx, n) -- the two
        # halves concatenate when each is a whole number of slabs
        f.wf_hi = f.wf_lo = None
        if c1.wf_hi is not None and ds.wf_hi is not None and \
                c1.cout % 32 == 0 and ds.cout % 32 == 0:
            f.wf_hi = torch.cat([c1.wf_hi, ds.wf_hi], 0).contiguous()
            f.wf_lo = torch.cat([c1.wf_lo, ds.wf_lo], 0).contiguous()
        assert f.w_ld % 4 == 0 and c1.cout % 4 == 0
        return dict(c1=f, c2=c2, fused=True, split=c1.cout)

    @staticmethod
    def run(p, x, out=None):
        """``out``: optional destination of the block's result (may be a channel slice of
        a wider cl array -- the caller's concatenation buffer)."""
        if p['fused']:
            y = ops.conv(x, p['c1'], 'relu', act_channels=p['split'])
            s = p['split']
            return ops.conv(y[..., :s], p['c2'], 'relu', residual=y[..., s:], out=out)
        y = ops.conv(x, p['c1'], 'relu')
        return ops.conv(y, p['c2'], 'relu', residual=x, out=out)


@BACKBONES.register_module()
class CustomResNet3D(BaseModule):

    def __init__(self, numC_input, num_layer=[2, 2, 2], num_channels=None,
                 stride=[2, 2, 2], backbone_output_ids=None, with_cp=False):
        super().__init__()
        assert len(num_layer) == len(stride)
        num_channels = [numC_input * 2 ** (i + 1)
                        for i in range(len(num_layer))] \
            if num_channels is None else num_channels
        self.backbone_output_ids = range(len(num_layer)) \
            if backbone_output_ids is None else backbone_output_ids
        layers = []
        curr = numC_input
        for i in range(len(num_layer)):
            layer = [BasicBlock3D(
                curr, num_channels[i], stride=stride[i],
                downsample=ConvModule(curr, num_channels[i], 3,
                                      stride=stride[i], padding=1, bias=False,
                                      conv_cfg=_C3D, norm_cfg=_BN3D,
                                      act_cfg=None))]
            curr = num_channels[i]
            layer.extend([BasicBlock3D(curr, curr)
                          for _ in range(num_layer[i] - 1)])
            layers.append(nn.Sequential(*layer))
        self.layers = nn.Sequential(*layers)
        self.with_cp = with_cp

    def _build_packs(self):
        return [[blk.pack() for blk in layer] for layer in self.layers]

    def forward(self, x, out=None):
        """[B,C,Z,Y,X] (channels_last_3d) -> list of the same.  ``out`` (a cl array, e.g.
        a channel slice of the frame-concatenation buffer) receives the LAST block's
        result in place of a fresh tensor."""
        x = ops.from_logical(x)
        feats = []
        packs = self.packs()
        for lid, layer in enumerate(packs):
            for bi, bp in enumerate(layer):
                last = out is not None and lid == len(packs) - 1 and bi == len(layer) - 1
                x = BasicBlock3D.run(bp, x, out if last else None)
            if lid in self.backbone_output_ids:
                feats.append(ops.to_logical(x))
        return feats


@NECKS.register_module()
class LSSFPN3D(BaseModule):

    def __init__(self, in_channels, out_channels, levels=3, with_cp=False):
        super().__init__()
        if levels != 3:
            raise NotImplementedError('PreWorld configs use levels=3')
        self.levels = levels
        self.conv = ConvModule(in_channels, out_channels, 1, stride=1,
                               padding=0, bias=False, conv_cfg=_C3D,
                               norm_cfg=_BN3D, act_cfg=dict(type='ReLU'))
        self.with_cp = with_cp

    def _build_packs(self):
        return {}

    def _split_packs(self, c8, c16, c32):
        """The 1x1x1 conv is linear and so is the interpolation, and BN's
        per-channel scale commutes with both:
            relu(bn(W [a; up(b); up(c)])) = relu(s*(Wa a) + up(s*Wb b)
                                                 + up(s*Wc c) + t)
        so Wb, Wc run at the COARSE resolutions and the [B,224,16,200,200]
        concatenation (573 MB) of lss_fpn.py:139-147 never exists."""
        P = self.packs()
        key = (c8, c16, c32)
        if key not in P:
            w = self.conv.conv.weight
            bn = self.conv.bn
            assert w.shape[1] == c8 + c16 + c32
            zero = torch.zeros_like(bn.bias)
            scale_only = (bn.weight, zero, zero, bn.running_var, bn.eps)
            full = (bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps)
            with torch.no_grad():
                P[key] = (
                    ops.PackedConv(w[:, :c8], None, full),
                    ops.PackedConv(w[:, c8:c8 + c16], None, scale_only),
                    ops.PackedConv(w[:, c8 + c16:], None, scale_only))
        return P[key]

    def forward(self, feats):
        x8, x16, x32 = [ops.from_logical(f) for f in feats]
        pa, pb, pc = self._split_packs(x8.shape[-1], x16.shape[-1],
                                       x32.shape[-1])
        r = ops.upsample_trilinear2(ops.conv(x16, pb), ops.conv(x32, pc),
                                    tuple(x8.shape[1:4]))
        return ops.to_logical(ops.conv(x8, pa, 'relu', residual=r))
