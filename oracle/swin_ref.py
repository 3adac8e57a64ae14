"""TEST INFRASTRUCTURE ONLY -- never imported by preworld_b200 or by bench.py's timed path.

Plain-torch fp32 restatement of the reference's Swin image side, evaluated FROM THE
PARAMETERS of a module with the reference's state_dict keys (ours or the reference's own):

  backbone_forward   SwinTransformer.forward          backbones/swin.py:927-970
  block_forward      SwinBlock.forward                 swin.py:511-520
  shifted_window_msa ShiftWindowMSA / WindowMSA        swin.py:262-300, 364-440
  patch_merging      PatchMerging.forward              swin.py:185-206
  neck_forward       FPN_LSS.forward                   necks/lss_fpn.py:83-99
  FFN                mmcv 1.6.0 cnn/bricks/transformer.py (third-party, absent): identity +
                     Linear-GELU-Linear -- the un-pinned assumption, shared with swin_shim

Pinned: tests/test_oracle.py runs it on the seeded tiny configuration and compares with
tests/golden/tiny_swin.npz, which oracle/make_swin_golden.py produced by executing the
reference's own swin.py / lss_fpn.py files (oracle/swin_shim.py).
"""
import zlib

import torch
import torch.nn.functional as F

# the tiny configuration of the golden fixture (window 6 on 16x48 / 8x24 / 4x12 / 2x6 token
# grids: stages pad to the window, odd blocks shift, the last grid is smaller than a
# window); head dim 32 as in Swin-B
TINY_SWIN = dict(pretrain_img_size=224, patch_size=4, window_size=6, mlp_ratio=4, embed_dims=32,
                 depths=[2, 2, 2, 2], num_heads=[1, 2, 4, 8], strides=(4, 2, 2, 2),
                 out_indices=(2, 3), qkv_bias=True, qk_scale=None, patch_norm=True,
                 drop_rate=0., attn_drop_rate=0., drop_path_rate=0.1, use_abs_pos_embed=False,
                 return_stereo_feat=True, act_cfg=dict(type='GELU'),
                 norm_cfg=dict(type='LN', requires_grad=True), pretrain_style='official',
                 output_missing_index_as_none=False)
TINY_NECK = dict(in_channels=128 + 256, out_channels=64, extra_upsample=None,
                 input_feature_index=(0, 1), scale_factor=2)
TINY_INPUT = (2, 3, 64, 192)


def seeded_init_(module, seed):
    """Key-addressed deterministic init shared by the reference modules and ours (same
    state_dict keys): every floating tensor of the state dict from a generator seeded by its
    key; norm scales around 1, variances positive."""
    sd = module.state_dict()
    for k, v in sd.items():
        if not v.is_floating_point():
            continue
        g = torch.Generator().manual_seed((zlib.crc32(k.encode()) + 7919 * seed) % (2 ** 31))
        if k.endswith('running_var'):
            t = torch.rand(v.shape, generator=g) + 0.5
        elif k.endswith('running_mean'):
            t = torch.randn(v.shape, generator=g) * 0.1
        elif v.dim() == 1 and k.endswith('.weight'):                   # norm scales
            t = torch.rand(v.shape, generator=g) + 0.5
        elif v.dim() == 1:                                             # biases
            t = torch.randn(v.shape, generator=g) * 0.1
        elif 'relative_position_bias_table' in k:
            t = torch.randn(v.shape, generator=g) * 0.5
        else:                                                          # linear / conv weights
            t = torch.randn(v.shape, generator=g) * (1.0 / v[0].numel()) ** 0.5
        v.copy_(t.to(v.device))
    return module


def tiny_input(device='cpu'):
    g = torch.Generator().manual_seed(3)
    return torch.randn(*TINY_INPUT, generator=g).to(device)


def _region_labels(hp, wp, ws, shift, device):
    """img_mask of swin.py:381-391: 3 x 3 regions cut at -ws and -shift."""
    lab = torch.zeros((hp, wp), device=device)
    cuts = lambda n: ((0, n - ws), (n - ws, n - shift), (n - shift, n))
    cnt = 0
    for h0, h1 in cuts(hp):
        for w0, w1 in cuts(wp):
            lab[h0:h1, w0:w1] = cnt
            cnt += 1
    return lab


def _windows(t, ws):
    """[B,Hp,Wp,C] -> [B*nW, ws*ws, C]"""
    b, hp, wp, c = t.shape
    return t.view(b, hp // ws, ws, wp // ws, ws, c).permute(0, 1, 3, 2, 4, 5) \
        .reshape(-1, ws * ws, c)


def shifted_window_msa(x, attn):
    """x [B,H,W,C] (already normed) -> attention output [B,H,W,C]; ``attn`` has
    window_size, shift_size and w_msa.{qkv, proj, relative_position_bias_table,
    relative_position_index, num_heads, scale}."""
    msa = attn.w_msa
    ws, shift, nh = attn.window_size, attn.shift_size, msa.num_heads
    b, h, w, c = x.shape
    pb, pr = (ws - h % ws) % ws, (ws - w % ws) % ws
    x = F.pad(x, (0, 0, 0, pr, 0, pb))
    hp, wp = h + pb, w + pr
    mask = None
    if shift > 0:
        x = torch.roll(x, shifts=(-shift, -shift), dims=(1, 2))
        lab = _windows(_region_labels(hp, wp, ws, shift, x.device)[None, :, :, None], ws)[..., 0]
        diff = lab[:, None, :] - lab[:, :, None]                       # [nW,N,N]
        mask = torch.where(diff != 0, torch.full_like(diff, -100.0), torch.zeros_like(diff))
    win = _windows(x, ws)
    nwb, n, _ = win.shape
    qkv = F.linear(win, msa.qkv.weight, msa.qkv.bias).reshape(nwb, n, 3, nh, c // nh) \
        .permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0] * msa.scale, qkv[1], qkv[2]
    a = q @ k.transpose(-2, -1)
    bias = msa.relative_position_bias_table[msa.relative_position_index.view(-1)] \
        .view(n, n, nh).permute(2, 0, 1)
    a = a + bias[None]
    if mask is not None:
        nw = mask.shape[0]
        a = (a.view(nwb // nw, nw, nh, n, n) + mask[None, :, None]).view(-1, nh, n, n)
    a = a.softmax(dim=-1)
    o = (a @ v).transpose(1, 2).reshape(nwb, n, c)
    o = F.linear(o, msa.proj.weight, msa.proj.bias)
    o = o.view(b, hp // ws, wp // ws, ws, ws, c).permute(0, 1, 3, 2, 4, 5).reshape(b, hp, wp, c)
    if shift > 0:
        o = torch.roll(o, shifts=(shift, shift), dims=(1, 2))
    return o[:, :h, :w].contiguous()


def _ln(x, m):
    return F.layer_norm(x, (x.shape[-1],), m.weight, m.bias, m.eps)


def block_forward(x, blk):
    x = x + shifted_window_msa(_ln(x, blk.norm1), blk.attn)
    fc1, fc2 = blk.ffn.layers[0][0], blk.ffn.layers[1]
    t = F.linear(F.gelu(F.linear(_ln(x, blk.norm2), fc1.weight, fc1.bias)), fc2.weight, fc2.bias)
    return x + t


def patch_merging(x, down):
    """x [B,H,W,C] -> [B,ceil(H/2),ceil(W/2),2C]: nn.Unfold(2, stride 2) channel order
    (c, ky, kx), LayerNorm(4C), Linear without bias."""
    b, h, w, c = x.shape
    t = x.permute(0, 3, 1, 2)
    t = F.pad(t, (0, w % 2, 0, h % 2))
    t = F.unfold(t, kernel_size=2, stride=2).transpose(1, 2)           # [B, L/4, 4C]
    t = F.linear(_ln(t, down.norm), down.reduction.weight, down.reduction.bias)
    return t.view(b, (h + 1) // 2, (w + 1) // 2, -1)


def backbone_forward(bb, img):
    """-> list of [N,C,h,w] maps in the order SwinTransformer.forward returns them."""
    pe = bb.patch_embed
    x = F.conv2d(img, pe.projection.weight, pe.projection.bias, stride=pe.projection.stride)
    x = x.permute(0, 2, 3, 1)
    if pe.norm is not None:
        x = _ln(x, pe.norm)
    outs = []
    for i, stage in enumerate(bb.stages):
        for blk in stage.blocks:
            x = block_forward(x, blk)
        if i == 0 and bb.return_stereo_feat:
            outs.append(x.permute(0, 3, 1, 2).contiguous())
        if i in bb.out_indices:
            outs.append(_ln(x, getattr(bb, f'norm{i}')).permute(0, 3, 1, 2).contiguous())
        if stage.downsample is not None:
            x = patch_merging(x, stage.downsample)
    return outs


def _conv_bn_relu(x, conv, bn, padding):
    x = F.conv2d(x, conv.weight, None, padding=padding)
    x = F.batch_norm(x, bn.running_mean, bn.running_var, bn.weight, bn.bias, False, 0., bn.eps)
    return F.relu(x)


def neck_forward(neck, feats):
    """FPN_LSS.forward (lss_fpn.py:83-99) incl. its optional lateral / input_conv / up2
    branches (the shipped configuration uses none of them)."""
    x2, x1 = feats[neck.input_feature_index[0]], feats[neck.input_feature_index[1]]
    if getattr(neck, 'lateral', False):
        x2 = _conv_bn_relu(x2, neck.lateral_conv[0], neck.lateral_conv[1], 0)
    x1 = F.interpolate(x1, scale_factor=neck.up.scale_factor, mode='bilinear',
                       align_corners=True)
    x = torch.cat([x2, x1], dim=1)
    if getattr(neck, 'input_conv', None) is not None:
        x = _conv_bn_relu(x, neck.input_conv[0], neck.input_conv[1], 0)
    x = _conv_bn_relu(x, neck.conv[0], neck.conv[1], 1)
    x = _conv_bn_relu(x, neck.conv[3], neck.conv[4], 1)
    if getattr(neck, 'extra_upsample', False):
        x = F.interpolate(x, scale_factor=neck.up2[0].scale_factor, mode='bilinear',
                          align_corners=True)
        x = _conv_bn_relu(x, neck.up2[1], neck.up2[2], 1)
        x = F.conv2d(x, neck.up2[4].weight, neck.up2[4].bias)
    return x


# ---- parameter trees from a state dict (so oracle/torch_ref.py, which is state-dict
#      driven, can evaluate the Swin image side without any module class) ----------------
class _NS:
    def __init__(self, **kw):
        self.__dict__.update(kw)


def _lin(sd, p, eps=None):
    ns = _NS(weight=sd[p + '.weight'], bias=sd.get(p + '.bias'))
    if eps is not None:
        ns.eps = eps
    return ns


def backbone_from_state_dict(sd, prefix, cfg, ln_eps=1e-5):
    """``cfg`` = the SwinTransformer config dict; -> the attribute tree the functions above
    read (patch_embed / stages[i].blocks[j] / stages[i].downsample / norm{i})."""
    p = prefix + '.' if prefix else ''
    ws = cfg['window_size']
    pe = _NS(projection=_NS(weight=sd[p + 'patch_embed.projection.weight'],
                            bias=sd[p + 'patch_embed.projection.bias'],
                            stride=(cfg['patch_size'],) * 2),
             norm=_lin(sd, p + 'patch_embed.norm', ln_eps) if cfg.get('patch_norm', True) else None)
    stages = []
    c = cfg['embed_dims']
    for i, depth in enumerate(cfg['depths']):
        nh = cfg['num_heads'][i]
        blocks = []
        for j in range(depth):
            q = f'{p}stages.{i}.blocks.{j}.'
            msa = _NS(qkv=_lin(sd, q + 'attn.w_msa.qkv'), proj=_lin(sd, q + 'attn.w_msa.proj'),
                      relative_position_bias_table=sd[q + 'attn.w_msa.relative_position_bias_table'],
                      relative_position_index=sd[q + 'attn.w_msa.relative_position_index'],
                      num_heads=nh, scale=cfg.get('qk_scale') or (c // nh) ** -0.5)
            blocks.append(_NS(
                norm1=_lin(sd, q + 'norm1', ln_eps), norm2=_lin(sd, q + 'norm2', ln_eps),
                attn=_NS(window_size=ws, shift_size=ws // 2 if j % 2 else 0, w_msa=msa),
                ffn=_NS(layers=[[_lin(sd, q + 'ffn.layers.0.0')], _lin(sd, q + 'ffn.layers.1')])))
        down = None
        if i < len(cfg['depths']) - 1:
            q = f'{p}stages.{i}.downsample.'
            down = _NS(norm=_lin(sd, q + 'norm', ln_eps), reduction=_lin(sd, q + 'reduction'))
            c *= 2
        stages.append(_NS(blocks=blocks, downsample=down))
    bb = _NS(patch_embed=pe, stages=stages, out_indices=tuple(cfg['out_indices']),
             return_stereo_feat=cfg.get('return_stereo_feat', False))
    for i in bb.out_indices:
        setattr(bb, f'norm{i}', _lin(sd, f'{p}norm{i}', ln_eps))
    return bb


def neck_from_state_dict(sd, prefix, cfg):
    p = prefix + '.' if prefix else ''
    bn = lambda q: _NS(weight=sd[q + '.weight'], bias=sd[q + '.bias'],
                       running_mean=sd[q + '.running_mean'], running_var=sd[q + '.running_var'],
                       eps=1e-5)
    conv = [_NS(weight=sd[p + 'conv.0.weight']), bn(p + 'conv.1'), None,
            _NS(weight=sd[p + 'conv.3.weight']), bn(p + 'conv.4'), None]
    return _NS(conv=conv, input_feature_index=tuple(cfg['input_feature_index']),
               up=_NS(scale_factor=cfg['scale_factor']))


def stage0_forward(bb, img):
    """BEVStereo4D.extract_stereo_ref_feat, Swin branch (bevdet.py:589-604): patch embedding +
    stage 0, un-normed, [N,C,h,w]."""
    pe = bb.patch_embed
    x = F.conv2d(img, pe.projection.weight, pe.projection.bias, stride=pe.projection.stride)
    x = x.permute(0, 2, 3, 1)
    if pe.norm is not None:
        x = _ln(x, pe.norm)
    for blk in bb.stages[0].blocks:
        x = block_forward(x, blk)
    return x.permute(0, 3, 1, 2).contiguous()
