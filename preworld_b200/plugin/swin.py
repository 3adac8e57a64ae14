"""``SwinTransformer`` -- the image backbone of the configuration the reference ships
(configs/preworld/nuscenes/bevstereo-occ.py:45-67: Swin-B, window 12, 512x1408 input).

Mirrors reference mmdet3d/models/backbones/swin.py:679-976: same constructor kwargs, same
parameter / buffer names (``patch_embed.projection``, ``stages.{i}.blocks.{j}.norm1``,
``.attn.w_msa.{relative_position_bias_table, relative_position_index, qkv, proj}``,
``.ffn.layers.0.0`` / ``.ffn.layers.1``, ``stages.{i}.downsample.{norm, reduction}``,
``norm{i}``), so the released checkpoints load by key.  Forward-only (eval): dropout and
DropPath are identities.

B200 layout: the token sequence [B, H*W, C] of the reference IS the channels-last feature
map [B,H,W,C] of the conv kernels, kept for the whole backbone.  Every linear (qkv, proj,
fc1+GELU, fc2, reduction, the 4x4 patch projection as a 2x2/s2 conv over the
space-to-depth(2) image) runs on the tcgen05 conv kernel with bias, GELU and the residual
add in its epilogue; LayerNorm, the 2x2 merge gather and the shifted-window attention are
the three kernels of csrc/swin.cu.  Padding, roll, window partition / reverse and crop
(swin.py:364-440) are index arithmetic inside the attention kernel.
"""
import math
import warnings
from collections import OrderedDict

import torch
import torch.nn as nn

from .. import ops
from .base import BaseModule, pack_linear
from .builder import BACKBONES


def _to_2tuple(v):
    return tuple(v) if isinstance(v, (tuple, list)) else (v, v)


def _norm(norm_cfg, dims):
    if norm_cfg is None:
        return None
    cfg = dict(norm_cfg)
    if cfg.pop('type') != 'LN':
        raise KeyError('SwinTransformer: only LayerNorm (norm_cfg type "LN") is on the path')
    cfg.pop('requires_grad', None)
    return nn.LayerNorm(dims, **cfg)


def _ln(x, m, out=None):
    return ops.layernorm(x, m.weight, m.bias, m.eps, out=out)


class PatchEmbed(nn.Module):
    """swin.py:76-166 -- Conv2d(k = stride = patch) + LayerNorm."""

    def __init__(self, in_channels, embed_dims, patch_size, norm_cfg):
        super().__init__()
        self.patch_size = _to_2tuple(patch_size)
        self.projection = nn.Conv2d(in_channels, embed_dims, self.patch_size,
                                    stride=self.patch_size)
        self.norm = _norm(norm_cfg, embed_dims)

    def pack(self):
        """The 4x4 / stride 4 projection as a 2x2 / stride 2 conv over the space-to-depth(2)
        image of ops.nchw_to_s2d (channel (dy*2+dx)*4 + c, padded to 32):
        W'[o, (dy,dx,c), a, b] = W[o, c, 2a+dy, 2b+dx]."""
        w = self.projection.weight.detach().float()
        co, ci, kh, kw = w.shape
        if (kh, kw) != (4, 4) or ci > 4:
            raise NotImplementedError('patch_size 4 with <= 4 image channels')
        w2 = torch.zeros((co, 32, 2, 2), device=w.device)
        for dy in range(2):
            for dx in range(2):
                c0 = (dy * 2 + dx) * 4
                w2[:, c0:c0 + ci] = w[:, :, dy::2, dx::2]
        return ops.PackedConv(w2, self.projection.bias, None, stride=2, padding=0)


class PatchMerging(nn.Module):
    """swin.py:169-206 -- nn.Unfold(2, stride 2) + LayerNorm(4C) + Linear(4C, 2C, no bias)."""

    def __init__(self, in_channels, out_channels, stride, norm_cfg):
        super().__init__()
        if stride != 2:
            raise NotImplementedError('PatchMerging stride 2 (every Swin variant)')
        self.in_channels, self.out_channels, self.stride = in_channels, out_channels, stride
        self.sampler = nn.Unfold(kernel_size=stride, dilation=1, padding=0, stride=stride)
        self.norm = _norm(norm_cfg, 4 * in_channels)
        self.reduction = nn.Linear(4 * in_channels, out_channels, bias=False)

    def pack(self):
        """nn.Unfold orders the gathered channels (c, ky, kx); pw_patch_merge_ln gathers
        whole tokens, i.e. (ky, kx, c): permute the norm's affine and the reduction's
        input columns."""
        c = self.in_channels
        dev = self.reduction.weight.device
        perm = (torch.arange(c, device=dev)[None, :] * 4 +
                torch.arange(4, device=dev)[:, None]).reshape(-1)     # new (s, c) <- old c*4+s
        if self.norm is not None:
            g = self.norm.weight.detach().float()[perm].contiguous()
            b = self.norm.bias.detach().float()[perm].contiguous()
            eps = self.norm.eps
        else:
            g = b = None
            eps = 0.
        return dict(gamma=g, beta=b, eps=eps, perm=perm,
                    red=ops.PackedConv(self.reduction.weight.detach()[:, perm],
                                       self.reduction.bias))

    @staticmethod
    def run(p, x):
        if p['gamma'] is None:
            raise NotImplementedError('PatchMerging without a norm layer')
        return ops.conv(ops.patch_merge_ln(x, p['gamma'], p['beta'], p['eps']), p['red'])


class WindowMSA(nn.Module):
    """swin.py:208-313 (parameters; the arithmetic is pw_window_attention)."""

    def __init__(self, embed_dims, num_heads, window_size, qkv_bias=True, qk_scale=None):
        super().__init__()
        self.embed_dims, self.num_heads = embed_dims, num_heads
        self.window_size = window_size                       # (Wh, Ww)
        head_dims = embed_dims // num_heads
        self.scale = qk_scale or head_dims ** -0.5
        wh, ww = window_size
        self.relative_position_bias_table = nn.Parameter(
            torch.zeros((2 * wh - 1) * (2 * ww - 1), num_heads))
        self.register_buffer('relative_position_index', self._index(wh, ww))
        self.qkv = nn.Linear(embed_dims, embed_dims * 3, bias=qkv_bias)
        self.proj = nn.Linear(embed_dims, embed_dims)

    @staticmethod
    def _index(wh, ww):
        """[N,N] table row of (query, key): ((qy-ky) + wh-1) * (2ww-1) + (qx-kx) + ww-1
        (what swin.py:246-252 builds from double_step_seq + flip)."""
        ys, xs = torch.meshgrid(torch.arange(wh), torch.arange(ww), indexing='ij')
        ys, xs = ys.reshape(-1), xs.reshape(-1)
        return ((ys[:, None] - ys[None, :] + wh - 1) * (2 * ww - 1) +
                (xs[:, None] - xs[None, :] + ww - 1)).contiguous()

    def pack(self):
        wh, ww = self.window_size
        if wh != ww or self.embed_dims != self.num_heads * 32:
            raise NotImplementedError('square windows, head dim 32')
        if not torch.equal(self.relative_position_index.cpu(), self._index(wh, ww)):
            raise ValueError('relative_position_index differs from the layout of '
                             'swin.py:246-252 the attention kernel indexes by')
        return dict(qkv=pack_linear(self.qkv), proj=pack_linear(self.proj),
                    qkv_bias=(self.qkv.bias.detach().float().contiguous()
                              if self.qkv.bias is not None else None),
                    table=self.relative_position_bias_table.detach().float().t().contiguous())


class ShiftWindowMSA(nn.Module):
    """swin.py:315-440."""

    def __init__(self, embed_dims, num_heads, window_size, shift_size=0, qkv_bias=True,
                 qk_scale=None):
        super().__init__()
        assert 0 <= shift_size < window_size
        self.window_size, self.shift_size = window_size, shift_size
        self.w_msa = WindowMSA(embed_dims, num_heads, _to_2tuple(window_size), qkv_bias,
                               qk_scale)


class FFN(nn.Module):
    """mmcv 1.6.0 cnn/bricks/transformer.py FFN with num_fcs = 2 (third-party, absent from
    /root/reference; call site swin.py:501-509): layers = Sequential(Sequential(Linear,
    act, Dropout), Linear, Dropout), out = identity + layers(x)."""

    def __init__(self, embed_dims, feedforward_channels, act_cfg):
        super().__init__()
        if act_cfg.get('type') != 'GELU':
            raise KeyError('SwinTransformer: only GELU FFNs are on the path')
        self.layers = nn.Sequential(
            nn.Sequential(nn.Linear(embed_dims, feedforward_channels), nn.GELU(),
                          nn.Dropout(0.)),
            nn.Linear(feedforward_channels, embed_dims), nn.Dropout(0.))


class SwinBlock(nn.Module):
    """swin.py:443-520."""

    def __init__(self, embed_dims, num_heads, feedforward_channels, window_size, shift,
                 qkv_bias, qk_scale, act_cfg, norm_cfg):
        super().__init__()
        self.norm1 = _norm(norm_cfg, embed_dims)
        self.attn = ShiftWindowMSA(embed_dims, num_heads, window_size,
                                   window_size // 2 if shift else 0, qkv_bias, qk_scale)
        self.norm2 = _norm(norm_cfg, embed_dims)
        self.ffn = FFN(embed_dims, feedforward_channels, act_cfg)

    def pack(self):
        p = self.attn.w_msa.pack()
        p.update(fc1=pack_linear(self.ffn.layers[0][0]), fc2=pack_linear(self.ffn.layers[1]))
        return p

    def run(self, p, x, out=None):
        """x [B,H,W,C] -> x + attn(norm1(x)) -> + ffn(norm2(.)) (swin.py:511-520)."""
        msa = self.attn.w_msa
        t = _ln(x, self.norm1)
        qkv = ops.conv(t, p['qkv'])
        ops.window_attention(qkv, p['qkv_bias'], p['table'], msa.num_heads,
                             self.attn.window_size, self.attn.shift_size, msa.scale, out=t)
        del qkv
        x = ops.conv(t, p['proj'], residual=x)
        _ln(x, self.norm2, out=t)
        hidden = ops.conv(t, p['fc1'], 'gelu')
        return ops.conv(hidden, p['fc2'], residual=x, out=out)


class SwinBlockSequence(nn.Module):
    """swin.py:523-611."""

    def __init__(self, embed_dims, num_heads, feedforward_channels, depth, window_size,
                 qkv_bias, qk_scale, downsample, act_cfg, norm_cfg):
        super().__init__()
        self.blocks = nn.ModuleList(
            SwinBlock(embed_dims, num_heads, feedforward_channels, window_size,
                      shift=i % 2 == 1, qkv_bias=qkv_bias, qk_scale=qk_scale,
                      act_cfg=act_cfg, norm_cfg=norm_cfg) for i in range(depth))
        self.downsample = downsample


def swin_convert(ckpt):
    """Key / layout conversion of an official Swin checkpoint (what swin.py:25-73 does
    for ``pretrain_style='official'``): attn.* -> attn.w_msa.*, mlp.fc{1,2} ->
    ffn.layers.{0.0,1}, layers -> stages, patch_embed.proj -> projection, and the
    downsample tensors from the official (x0,x1,x2,x3) token order to nn.Unfold's
    channel order."""
    out = OrderedDict()
    for k, v in ckpt.items():
        if k.startswith('head'):
            continue
        if k.startswith('layers'):
            if 'attn.' in k:
                k = k.replace('attn.', 'attn.w_msa.')
            elif 'mlp.fc1.' in k:
                k = k.replace('mlp.fc1.', 'ffn.layers.0.0.')
            elif 'mlp.fc2.' in k:
                k = k.replace('mlp.fc2.', 'ffn.layers.1.')
            elif 'mlp.' in k:
                k = k.replace('mlp.', 'ffn.')
            elif 'downsample' in k and ('reduction.' in k or 'norm.' in k):
                c4 = v.shape[-1]
                lead = v.shape[:-1]
                v = v.reshape(*lead, 4, c4 // 4)[..., [0, 2, 1, 3], :] \
                    .transpose(-1, -2).reshape(*lead, c4)
            k = k.replace('layers', 'stages', 1)
        elif k.startswith('patch_embed') and 'proj' in k:
            k = k.replace('proj', 'projection')
        out[k] = v
    return out


@BACKBONES.register_module()
class SwinTransformer(BaseModule):
    def __init__(self, pretrain_img_size=224, in_channels=3, embed_dims=96, patch_size=4,
                 window_size=7, mlp_ratio=4, depths=(2, 2, 6, 2), num_heads=(3, 6, 12, 24),
                 strides=(4, 2, 2, 2), out_indices=(0, 1, 2, 3), qkv_bias=True,
                 qk_scale=None, patch_norm=True, drop_rate=0., attn_drop_rate=0.,
                 drop_path_rate=0.1, use_abs_pos_embed=False, act_cfg=dict(type='GELU'),
                 norm_cfg=dict(type='LN'), pretrain_style='official', pretrained=None,
                 init_cfg=None, with_cp=True, return_stereo_feat=False,
                 output_missing_index_as_none=False, frozen_stages=-1):
        super().__init__(init_cfg)
        pretrain_img_size = _to_2tuple(pretrain_img_size)
        assert pretrain_style in ('official', 'mmcls')
        if not (isinstance(pretrained, str) or pretrained is None):
            raise TypeError('pretrained must be a str or None')
        assert strides[0] == patch_size, 'Use non-overlapping patch embed.'
        self.out_indices = tuple(out_indices)
        self.use_abs_pos_embed = use_abs_pos_embed
        self.pretrain_style, self.pretrained = pretrain_style, pretrained
        self.frozen_stages = frozen_stages
        self.return_stereo_feat = return_stereo_feat
        self.output_missing_index_as_none = output_missing_index_as_none
        self.patch_embed = PatchEmbed(in_channels, embed_dims, patch_size,
                                      norm_cfg if patch_norm else None)
        if use_abs_pos_embed:
            n = (pretrain_img_size[0] // patch_size) * (pretrain_img_size[1] // patch_size)
            self.absolute_pos_embed = nn.Parameter(torch.zeros((1, n, embed_dims)))
        self.drop_after_pos = nn.Dropout(p=drop_rate)
        self.stages = nn.ModuleList()
        c = embed_dims
        for i, depth in enumerate(depths):
            down = PatchMerging(c, 2 * c, strides[i + 1], norm_cfg if patch_norm else None) \
                if i < len(depths) - 1 else None
            self.stages.append(SwinBlockSequence(
                c, num_heads[i], mlp_ratio * c, depth, window_size, qkv_bias, qk_scale, down,
                act_cfg, norm_cfg))
            if down is not None:
                c = down.out_channels
        self.num_features = [int(embed_dims * 2 ** i) for i in range(len(depths))]
        for i in self.out_indices:
            self.add_module(f'norm{i}', _norm(norm_cfg, self.num_features[i]))

    # -- swin.py:861-925 ------------------------------------------------------------
    def init_weights(self):
        if self.pretrained is None:
            for m in self.modules():
                if isinstance(m, nn.Linear):
                    nn.init.trunc_normal_(m.weight, std=.02)
                    if m.bias is not None:
                        nn.init.zeros_(m.bias)
                elif isinstance(m, nn.LayerNorm):
                    nn.init.ones_(m.weight)
                    nn.init.zeros_(m.bias)
                elif isinstance(m, WindowMSA):
                    nn.init.trunc_normal_(m.relative_position_bias_table, std=.02)
            if self.use_abs_pos_embed:
                nn.init.trunc_normal_(self.absolute_pos_embed, std=.02)
            return
        ckpt = torch.load(self.pretrained, map_location='cpu')
        sd = ckpt.get('state_dict', ckpt.get('model', ckpt))
        if self.pretrain_style == 'official':
            sd = swin_convert(sd)
        if next(iter(sd)).startswith('module.'):
            sd = OrderedDict((k[7:], v) for k, v in sd.items())
        own = self.state_dict()
        for k in [k for k in sd if 'relative_position_bias_table' in k]:
            l1, h1 = sd[k].shape
            l2, h2 = own[k].shape
            if h1 != h2:
                warnings.warn(f'Error in loading {k}, pass')
                del sd[k]
            elif l1 != l2:
                s1, s2 = int(l1 ** 0.5), int(l2 ** 0.5)
                t = nn.functional.interpolate(sd[k].permute(1, 0).reshape(1, h1, s1, s1),
                                              size=(s2, s2), mode='bicubic')
                sd[k] = t.view(h2, l2).permute(1, 0).contiguous()
        self.load_state_dict(sd, False)

    def _build_packs(self):
        return dict(embed=self.patch_embed.pack(),
                    blocks=[[blk.pack() for blk in st.blocks] for st in self.stages],
                    down=[st.downsample.pack() if st.downsample is not None else None
                          for st in self.stages])

    # -- stage-level entry points, the same surface as plugin/image.py ResNet (the
    #    detectors run stage 0 alone for the stereo reference frame, bevdet.py:589-604) --
    @property
    def stage0_is_stereo(self):
        return self.return_stereo_feat and 0 not in self.out_indices

    def stem_input_shape(self, h, w):
        if h % 4 or w % 4:
            raise NotImplementedError('image sides must be multiples of the patch size')
        return (h // 2, w // 2, 32)

    def convert_images(self, img_nchw, out=None):
        self.stem_input_shape(*img_nchw.shape[2:])
        return ops.nchw_to_s2d(img_nchw, 32, out=out)

    def run_stem_cl(self, x_cl):
        """space-to-depth(2) image -> patch projection + norm (+ position embedding):
        the token map [n, H/4, W/4, C]."""
        x = ops.conv(x_cl, self.packs()['embed'])
        if self.patch_embed.norm is not None:
            x = _ln(x, self.patch_embed.norm, out=x)
        if self.use_abs_pos_embed:
            n, h, w, c = x.shape
            x = x + self.absolute_pos_embed.view(1, h, w, c)
        return x

    def run_stem(self, img_nchw):
        return self.run_stem_cl(self.convert_images(img_nchw))

    def run_layer(self, i, x, out=None):
        """The blocks of stage i on its input token map (already merged for i > 0)."""
        blocks = self.stages[i].blocks
        bps = self.packs()['blocks'][i]
        for k, (blk, bp) in enumerate(zip(blocks, bps)):
            x = blk.run(bp, x, out=out if k == len(bps) - 1 else None)
        return x

    def _out(self, i, x):
        return ops.to_logical(_ln(x, getattr(self, f'norm{i}')))

    def run_from_layer(self, i0, x):
        """x = block output of stage i0-1 -> logical maps of out_indices >= i0."""
        outs = []
        p = self.packs()
        for i in range(i0, len(self.stages)):
            x = PatchMerging.run(p['down'][i - 1], x)
            x = self.run_layer(i, x)
            if i in self.out_indices:
                outs.append(self._out(i, x))
            elif self.output_missing_index_as_none:
                outs.append(None)
        return tuple(outs)

    def forward(self, x):
        """[N,3,H,W] -> list of logical [N,C,h,w] maps (channels_last strides): the
        un-normed stage-0 map first when ``return_stereo_feat`` (swin.py:940-944), then
        norm{i}(stage i) for ``out_indices``."""
        x = self.run_layer(0, self.run_stem(x))
        outs = []
        if self.return_stereo_feat:
            outs.append(ops.to_logical(x))
        if 0 in self.out_indices:
            outs.append(self._out(0, x))
        elif self.output_missing_index_as_none:
            outs.append(None)
        outs.extend(self.run_from_layer(1, x))
        return outs
