"""Torch-tensor front-end of the C-ABI kernels (include/preworld_b200.h).

torch is used here only for device memory and streams: every function
extracts raw device pointers, leading dimensions and the current stream and
calls the C library.  There is no CPU or eager-PyTorch fallback -- CPU tensors
raise.

Layout: feature maps are handled as *channels-last arrays*
``[N, H, W, C]`` / ``[B, Z, Y, X, C]`` ("cl" tensors: last dim stride 1, all
other dims dense over a pixel pitch ``ld >= C``).  ``to_logical`` /
``from_logical`` convert (zero-copy) to and from the reference's logical
``[N,C,H,W]`` / ``[B,C,Z,Y,X]`` tensors with channels_last strides.
"""
import ctypes
import os

import torch

from . import _lib
from ._lib import ConvDesc, RenderDesc, check

ACT = {None: 0, 'none': 0, 'relu': 1, 'softplus': 2, 'sigmoid': 3, 'gelu': 4}


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _require_cuda(*ts):
    """Every operand on ONE CUDA device, and that device current: the kernels are
    launched on the current device's current stream (`_stream`), so a tensor of
    another device would be a wrong-device launch, not an error, without this."""
    dev = None
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError(
                'preworld_b200 ops run on CUDA tensors only (no CPU fallback)')
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise RuntimeError(f'operands on different devices: {dev} and {t.device}')
    if dev is not None and dev.index != torch.cuda.current_device():
        raise RuntimeError(
            f'tensors live on {dev} but the current device is cuda:'
            f'{torch.cuda.current_device()}: wrap the call in torch.cuda.device({dev.index})')


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def cl_ld(t):
    """Pixel pitch of a channels-last array [..., C]; validates density."""
    assert t.dtype == torch.float32 and t.stride(-1) == 1, (t.dtype, t.stride())
    ld = exp = None
    for i in range(t.dim() - 2, -1, -1):
        if t.shape[i] == 1:
            continue                      # stride of a size-1 dim is arbitrary
        if ld is None:
            ld = exp = t.stride(i)
        elif t.stride(i) != exp:
            raise ValueError(f'not a dense channels-last array: shape '
                             f'{tuple(t.shape)} strides {t.stride()}')
        exp *= t.shape[i]
    return t.shape[-1] if ld is None else ld


def to_logical(t_cl):
    """[N,H,W,C] -> [N,C,H,W] view (channels_last strides); 5-D alike."""
    d = t_cl.dim()
    return t_cl.permute(0, d - 1, *range(1, d - 1))


def from_logical(t):
    """[N,C,H,W] (channels_last strides) -> [N,H,W,C] view.  A tensor in the
    default NCHW-contiguous format is rejected: callers convert images with
    ``nchw_to_nhwc`` at the model boundary."""
    d = t.dim()
    v = t.permute(0, *range(2, d), 1)
    if v.stride(-1) != 1:
        raise ValueError('expected a channels-last feature map '
                         f'(strides {t.stride()})')
    return v


# --------------------------------------------------------------------- conv
def split_tf32(w):
    """fp32 -> (hi, lo) tf32 pair of the 3xTF32 tensor-core kernels, both
    ROUNDED to nearest (ties away): hi = rna_tf32(w), lo = rna_tf32(w - hi).
    The tensor core truncates its operands to tf32, so un-rounded parts would
    carry a one-sided error (see csrc/tc_ptx.cuh:split2_rn)."""
    def rna(t):
        return ((t.contiguous().view(torch.int32) + 0x1000) & -8192) \
            .view(torch.float32)
    w = w.contiguous().float()
    hi = rna(w)
    return hi.contiguous(), rna(w - hi).contiguous()


class PackedConv:
    """Weights of one conv/linear in the kernel's layout: w [K, w_ld] with
    K = taps*cin_pad (tap-major), plus the folded per-channel affine."""
    __slots__ = ('w', 'scale', 'bias', 'cin', 'cout', 'k', 'stride', 'pad',
                 'dil', 'w_ld', 'wt_hi', 'wt_lo', 'wf_hi', 'wf_lo', 'src',
                 'collapsed')

    def __init__(self, weight, bias=None, bn=None, stride=1, padding=0,
                 dilation=1, in_scale=None, in_shift=None, eps=None,
                 spatial_perm=None, cin_pad=None, cout_pad=None):
        """weight [Cout,Cin,*k] (or [Cout,Cin] for Linear).  bn = (gamma,
        beta, mean, var, eps) folded as y = conv*s + (beta - mean*s + b*s).
        in_scale/in_shift fold a per-input-channel affine applied BEFORE a
        1x1 conv / linear (BatchNorm1d in front of DepthNet's Mlp).
        spatial_perm re-orders the kernel's spatial axes (used to run the
        reference's [X,Y,Z]-ordered OccHead on [Z,Y,X] volumes).
        cin_pad / cout_pad widen the packed conv with ZERO weights (zero
        scale-free outputs): a 344- or 88-channel tensor padded to a multiple
        of 32 runs on the tensor-core kernels; padded outputs are exact 0."""
        w = weight.detach().float()
        if w.dim() == 2:
            w = w[:, :, None, None, None]
        elif w.dim() == 4:
            w = w[:, :, None]
        assert w.dim() == 5
        # what collapse_axes() rebuilds a narrower conv from (references, no copies)
        self.src = None if (spatial_perm is not None or in_scale is not None) else \
            (w, bias, bn, len(tuple(weight.shape)), cin_pad, cout_pad)
        self.collapsed = {}
        if spatial_perm is not None:
            w = w.permute(0, 1, *[2 + p for p in spatial_perm])
        if cin_pad or cout_pad:
            co, ci = w.shape[:2]
            wp = torch.zeros((cout_pad or co, cin_pad or ci, *w.shape[2:]),
                             device=w.device)
            wp[:co, :ci] = w
            w = wp
            if cout_pad and cout_pad > co:
                padv = lambda v, fill: None if v is None else torch.cat(
                    [v.detach().float(), torch.full((cout_pad - co,), fill,
                                                    device=v.device)])
                bias = padv(bias, 0.)
                if bn is not None:
                    g_, b_, m_, v_, e_ = bn
                    bn = (padv(g_, 1.), padv(b_, 0.), padv(m_, 0.),
                          padv(v_, 1.), e_)
        cout, cin = w.shape[:2]
        k = tuple(w.shape[2:])
        b = bias.detach().float() if bias is not None else None
        if in_scale is not None:
            assert k == (1, 1, 1)
            shift_term = (w[:, :, 0, 0, 0] * in_shift[None, :]).sum(1)
            b = shift_term if b is None else b + shift_term
            w = w * in_scale[None, :, None, None, None]
        cin_pad = (cin + 3) // 4 * 4
        w_ld = (cout + 3) // 4 * 4
        wk = torch.zeros(*k, cin_pad, w_ld, device=w.device)
        wk[..., :cin, :cout] = w.permute(2, 3, 4, 1, 0)
        self.w = wk.reshape(-1, w_ld).contiguous()
        self.wt_hi = self.wt_lo = self.wf_hi = self.wf_lo = None
        if cin % 32 == 0:
            self.set_umma_weights(w.permute(0, 2, 3, 4, 1).reshape(cout, -1))
            if 2 <= k[2] <= 4 and cout <= FOLD_MAX_COUT:
                self.set_fold_weights(w)
        if bn is not None:
            gamma, beta, mean, var, eps_ = bn
            s = gamma.detach().float() / torch.sqrt(var.detach().float() + eps_)
            sh = beta.detach().float() - mean.detach().float() * s
            if b is not None:
                sh = sh + b * s
            self.scale, self.bias = s.contiguous(), sh.contiguous()
        else:
            self.scale = None
            self.bias = b.contiguous() if b is not None else None
        self.cin, self.cout, self.k, self.w_ld = cin_pad, cout, k, w_ld

        def t3(v):
            if isinstance(v, int):
                v = (v,) * 3
            v = tuple(v)
            return (1,) * (3 - len(v)) + v if len(v) < 3 else v
        self.stride, self.pad, self.dil = t3(stride), t3(padding), t3(dilation)
        # a 2-D conv has kd == 1: depth stride/pad/dilation are neutral
        if k[0] == 1 and len(tuple(weight.shape)) <= 4:
            self.stride = (1,) + self.stride[1:]
            self.pad = (0,) + self.pad[1:]
            self.dil = (1,) + self.dil[1:]
        if spatial_perm is not None:
            self.stride = tuple(self.stride[p] for p in spatial_perm)
            self.pad = tuple(self.pad[p] for p in spatial_perm)
            self.dil = tuple(self.dil[p] for p in spatial_perm)


    def collapse_axes(self, spatial):
        """A dilated 'same' conv whose dilation is at least the input extent along an axis
        only ever sees zero padding through its off-centre taps there (ASPP's dilation-18
        branch on a 16-row map, view_transformer.py:355-418): the conv with that axis
        reduced to its centre tap gives the same sums with a third of the taps and a halo
        that fits the tensor-core kernel's shared memory.  -> the narrower PackedConv for an
        input of spatial extent ``spatial`` (d, h, w), or self."""
        if self.src is None:
            return self
        axes = tuple(i for i in range(3)
                     if self.k[i] > 1 and self.k[i] % 2 == 1 and self.stride[i] == 1
                     and self.pad[i] == self.dil[i] * (self.k[i] // 2)
                     and self.dil[i] >= spatial[i])
        if not axes:
            return self
        if axes not in self.collapsed:
            w, bias, bn, wdim, cin_pad, cout_pad = self.src
            pad, dil = list(self.pad), list(self.dil)
            for i in axes:
                c = self.k[i] // 2
                w = w.narrow(2 + i, c, 1)
                pad[i], dil[i] = 0, 1
            pc = PackedConv(w.contiguous(), bias, bn, stride=self.stride, padding=tuple(pad),
                            dilation=tuple(dil), cin_pad=cin_pad, cout_pad=cout_pad)
            self.collapsed[axes] = pc
        return self.collapsed[axes]


    def set_umma_weights(self, wt):
        """wt [cout, K] (K-major, tap-major then cin): pre-split for the
        3xTF32 tensor-core path (pw_conv_umma_fwd)."""
        self.wt_hi, self.wt_lo = split_tf32(wt)


    def set_fold_weights(self, w5):
        """w5 [cout, cin, kd, kh, kw]: weights of the x-tap-folded kernel
        (pw_conv_fold_fwd, layout in include/preworld_b200.h): rows
        (slab, kx, n), columns (kz, ky, ci), pre-split hi/lo."""
        cout, cin, kd, kh, kw = w5.shape
        fold_n = 16 if cout <= 16 else 32            # == pw_conv_fold_n(cout)
        slabs = -(-cout // fold_n)
        wp = torch.zeros((slabs * fold_n, cin, kd, kh, kw), device=w5.device)
        wp[:cout] = w5
        wf = wp.view(slabs, fold_n, cin, kd, kh, kw).permute(0, 5, 1, 3, 4, 2) \
            .reshape(slabs * kw * fold_n, kd * kh * cin).contiguous()
        self.wf_hi, self.wf_lo = split_tf32(wf)


# The tensor-core path is used whenever the layer qualifies; set to False to
# force the fp32 SIMT kernel (tests compare the two).
USE_UMMA = True
# ... and among the tensor-core kernels the halo-resident / TMEM-operand one
# (conv_halo.cu) wherever its plan fits; False falls back to conv_umma.cu.
USE_HALO = True
# ... or its x-tap-folded launch (pw_conv_fold_fwd) for stride-1 3-tap convs with
# at most FOLD_AUTO_COUT output channels.  Measured inside the step (persistent
# CTAs): 32->16 k333 (OccHead) 276 -> 207 us; 32->32 k333 with residual 289 -> 324 us
# (its smaller CTAs pay the epilogue more often), wider layers re-split the halo
# once per 32-channel slab and lose more -- hence 16.  PW_HALO_FOLD=0 switches it
# off, PW_HALO_FOLD=1 folds every layer the library can (experiments).
USE_FOLD = os.environ.get('PW_HALO_FOLD') != '0'
FOLD_AUTO_COUT = 128 if os.environ.get('PW_HALO_FOLD') == '1' else 16
FOLD_MAX_COUT = 128          # fold weights are packed up to this width


def conv(x, pc, act=None, residual=None, out=None, act_channels=0,
         out_size=None):
    """x: cl array [N,H,W,C] or [B,Z,Y,X,C] (C >= pc.cin, extra channels
    ignored only if equal to the packed cin).  Returns / fills a cl array with
    pc.cout channels; ``out`` may be a channel slice of a wider cl array."""
    _require_cuda(x, residual, out)
    spatial = x.shape[1:-1]
    sp = (1,) * (3 - len(spatial)) + tuple(spatial)
    n = x.shape[0]
    if out_size is None and max(pc.dil) > 1:
        pc = pc.collapse_axes(sp)
    assert x.shape[-1] == pc.cin, (x.shape, pc.cin)
    o = tuple((sp[i] + 2 * pc.pad[i] - pc.dil[i] * (pc.k[i] - 1) - 1)
              // pc.stride[i] + 1 for i in range(3))
    if out_size is not None:
        # fewer outputs than the symmetric-padding formula gives == less padding
        # at the END of each axis (the kernels take the output extent as given)
        os_ = (1,) * (3 - len(out_size)) + tuple(out_size)
        assert all(0 < a <= b for a, b in zip(os_, o)), (os_, o)
        o = os_
    if out is None:
        out = torch.empty((n, *o[3 - len(spatial):], pc.cout), device=x.device,
                          dtype=torch.float32)
    else:
        assert out.shape[0] == n and out.shape[-1] == pc.cout and \
            tuple(out.shape[1:-1]) == o[3 - len(spatial):], (out.shape, o)
    d = ConvDesc(n=n, d=sp[0], h=sp[1], w=sp[2], cin=pc.cin, in_ld=cl_ld(x),
                 od=o[0], oh=o[1], ow=o[2], cout=pc.cout, out_ld=cl_ld(out),
                 res_ld=cl_ld(residual) if residual is not None else 0,
                 w_ld=pc.w_ld, kd=pc.k[0], kh=pc.k[1], kw=pc.k[2],
                 sd=pc.stride[0], sh=pc.stride[1], sw=pc.stride[2],
                 pd=pc.pad[0], ph=pc.pad[1], pw=pc.pad[2],
                 dd=pc.dil[0], dh=pc.dil[1], dw=pc.dil[2], act=ACT[act],
                 act_channels=act_channels)
    if residual is not None:
        assert residual.shape == out.shape
    L = _lib.lib()
    if USE_UMMA and USE_HALO and USE_FOLD and pc.wf_hi is not None and \
            pc.cout <= FOLD_AUTO_COUT and \
            L.pw_conv_fold_supported(ctypes.byref(d)):
        check(L.pw_conv_fold_fwd(ctypes.byref(d), _ptr(x), _ptr(pc.wf_hi),
                                 _ptr(pc.wf_lo), _ptr(pc.scale), _ptr(pc.bias),
                                 _ptr(residual), _ptr(out), _stream()),
              'pw_conv_fold_fwd')
        return out
    if USE_UMMA and USE_HALO and pc.wt_hi is not None and \
            L.pw_conv_halo_supported(ctypes.byref(d)):
        check(L.pw_conv_halo_fwd(ctypes.byref(d), _ptr(x), _ptr(pc.wt_hi),
                                 _ptr(pc.wt_lo), _ptr(pc.scale), _ptr(pc.bias),
                                 _ptr(residual), _ptr(out), _stream()),
              'pw_conv_halo_fwd')
        return out
    if USE_UMMA and pc.wt_hi is not None and \
            L.pw_conv_umma_supported(ctypes.byref(d)):
        check(L.pw_conv_umma_fwd(ctypes.byref(d), _ptr(x), _ptr(pc.wt_hi),
                                 _ptr(pc.wt_lo), _ptr(pc.scale), _ptr(pc.bias),
                                 _ptr(residual), _ptr(out), _stream()),
              'pw_conv_umma_fwd')
        return out
    check(_lib.lib().pw_conv_fwd(ctypes.byref(d), _ptr(x), _ptr(pc.w),
                                 _ptr(pc.scale), _ptr(pc.bias),
                                 _ptr(residual), _ptr(out), _stream()),
          'pw_conv_fwd')
    return out


def linear(x2d, pc, act=None, residual=None, out=None):
    """x2d [rows, C] -> [rows, cout] through the conv kernel (1x1)."""
    y = conv(x2d[:, None, :], pc, act,
             residual[:, None, :] if residual is not None else None,
             out[:, None, :] if out is not None else None)
    return y[:, 0, :]


# ------------------------------------------------------- fused per-voxel MLP
class PackedMlp2:
    """Weights of ``pw_mlp2``: y = act2(W2 . act1(W1 . x + b1) + b2) [+ res].
    w1 [H, 32], w2 [n2, H] (torch Linear layout).  H is padded to a multiple of
    32 and n2 to a multiple of 4 with zero weights (padded outputs are
    act2(0 + 0))."""
    __slots__ = ('w1_hi', 'w1_lo', 'b1', 'w2_hi', 'w2_lo', 'b2', 'c1', 'hidden',
                 'n2', 'act1', 'act2', 'act2_channels')

    def __init__(self, w1, b1, w2, b2, act1='softplus', act2=None,
                 act2_channels=0):
        w1, w2 = w1.detach().float(), w2.detach().float()
        h, c1 = w1.shape
        n2 = w2.shape[0]
        assert w2.shape[1] == h and c1 == 32
        hp, n2r = -(-h // 32) * 32, -(-n2 // 4) * 4
        n2p = 16 if n2r <= 16 else 32
        dev = w1.device
        w1p = torch.zeros((hp, c1), device=dev)
        w1p[:h] = w1
        w2p = torch.zeros((n2p, hp), device=dev)
        w2p[:n2, :h] = w2
        self.w1_hi, self.w1_lo = split_tf32(w1p)
        self.w2_hi, self.w2_lo = split_tf32(w2p)
        b1p = torch.zeros(hp, device=dev)
        if b1 is not None:
            b1p[:h] = b1.detach().float()
        # a padded hidden unit is act1(0): its W2 column is zero, so it adds nothing
        self.b1 = b1p.contiguous()
        b2p = torch.zeros(n2r, device=dev)
        if b2 is not None:
            b2p[:n2] = b2.detach().float()
        self.b2 = b2p.contiguous()
        self.c1, self.hidden, self.n2 = c1, hp, n2r
        self.act1, self.act2, self.act2_channels = act1, act2, act2_channels


def mlp2(rows, pm, bias1=None, residual=None, out=None):
    """rows [M, 32] (row pitch = rows.stride(0)) -> [M, pm.n2] through ONE fused
    tensor-core launch.  ``bias1`` overrides pm.b1 (per-sample bias)."""
    _require_cuda(rows, residual, out, bias1)
    m, c1 = rows.shape
    assert c1 == pm.c1 and rows.stride(1) == 1 and rows.dtype == torch.float32
    if out is None:
        out = torch.empty((m, pm.n2), device=rows.device, dtype=torch.float32)
    assert out.shape == (m, pm.n2) and out.stride(1) == 1
    b1 = pm.b1
    if bias1 is not None:
        b1 = bias1.contiguous().float()
        assert b1.numel() == pm.hidden
    if residual is not None:
        assert residual.shape == out.shape and residual.stride(1) == 1
    check(_lib.lib().pw_mlp2(
        _ptr(rows), rows.stride(0), m, c1, _ptr(pm.w1_hi), _ptr(pm.w1_lo),
        _ptr(b1), pm.hidden, ACT[pm.act1], _ptr(pm.w2_hi), _ptr(pm.w2_lo),
        _ptr(pm.b2), pm.n2, ACT[pm.act2], pm.act2_channels, _ptr(residual),
        residual.stride(0) if residual is not None else 0, _ptr(out),
        out.stride(0), _stream()), 'pw_mlp2')
    return out


def occhead_tail(feat_cl, w0, s0, b0, w1, b1, free_idx, geo_value, out=None,
                 want_logits=False):
    """feat_cl [1,Z,Y,X,16] (output of occ_convs[0]) -> uint8 [2,X,Y,Z]
    (class argmax, geometry grid) [+ logits cl array [1,Z,Y,X,ncls]]."""
    _require_cuda(feat_cl, w0, w1)
    _, gz, gy, gx, cin = feat_cl.shape
    ncls, mid = w1.shape
    if out is None:
        out = torch.empty((2, gx, gy, gz), device=feat_cl.device, dtype=torch.uint8)
    assert out.shape == (2, gx, gy, gz) and out.is_contiguous() and \
        out.dtype == torch.uint8
    logits = torch.empty((1, gz, gy, gx, ncls), device=feat_cl.device,
                         dtype=torch.float32) if want_logits else None
    check(_lib.lib().pw_occhead_tail(
        _ptr(feat_cl), cl_ld(feat_cl), cin, _ptr(w0), _ptr(s0), _ptr(b0), mid,
        _ptr(w1), _ptr(b1), ncls, _ptr(logits), ncls, _ptr(out[0]), _ptr(out[1]),
        int(free_idx), int(geo_value), gx, gy, gz, _stream()), 'pw_occhead_tail')
    return (out, logits) if want_logits else out


# ---------------------------------------------------------------- image side
def nchw_to_nhwc(x, c_pad=None, out=None):
    """x [n,c,h,w]: each image dense NCHW; images may be strided (a frame
    slice of the loader's camera-major batch).  ``out``: dense [n,h,w,c_pad]
    destination (e.g. a batch slice of a larger buffer)."""
    _require_cuda(x)
    n, c, h, w = x.shape
    if x.stride()[1:] != (h * w, w, 1):
        raise ValueError(f'images must be dense CHW planes, got {x.stride()}')
    c_pad = c_pad or c
    if out is None:
        y = torch.empty((n, h, w, c_pad), device=x.device, dtype=torch.float32)
    else:
        assert out.shape == (n, h, w, c_pad) and out.is_contiguous()
        y = out
    img_stride = x.stride(0) if n > 1 else c * h * w
    check(_lib.lib().pw_nchw_to_nhwc_pad(_ptr(x), img_stride, _ptr(y), n, c,
                                         h, w, c_pad, _stream()),
          'pw_nchw_to_nhwc_pad')
    return y


def nchw_to_s2d(x, c_pad=32, out=None):
    """x [n,c<=4,h,w] (dense CHW planes, images may be strided) ->
    space-to-depth(2) channels-last [n,h/2,w/2,c_pad]: channel (dy*2+dx)*4+c."""
    _require_cuda(x)
    n, c, h, w = x.shape
    if x.stride()[1:] != (h * w, w, 1):
        raise ValueError(f'images must be dense CHW planes, got {x.stride()}')
    if out is None:
        out = torch.empty((n, h // 2, w // 2, c_pad), device=x.device,
                          dtype=torch.float32)
    else:
        assert out.shape == (n, h // 2, w // 2, c_pad) and out.is_contiguous()
    img_stride = x.stride(0) if n > 1 else c * h * w
    check(_lib.lib().pw_nchw_to_s2d_nhwc(_ptr(x), img_stride, _ptr(out), n, c,
                                         h, w, c_pad, _stream()),
          'pw_nchw_to_s2d_nhwc')
    return out


def nhwc_to_nchw(x_cl):
    """cl array [N,...,C] (may be a channel slice) -> contiguous [N,C,...]."""
    _require_cuda(x_cl)
    n, c = x_cl.shape[0], x_cl.shape[-1]
    spatial = tuple(x_cl.shape[1:-1])
    pixels = 1
    for s in spatial:
        pixels *= s
    y = torch.empty((n, c, *spatial), device=x_cl.device, dtype=torch.float32)
    check(_lib.lib().pw_nhwc_to_nchw(_ptr(x_cl), cl_ld(x_cl), _ptr(y), n, c,
                                     pixels, _stream()), 'pw_nhwc_to_nchw')
    return y


def maxpool3x3s2(x):
    n, h, w, c = x.shape
    assert cl_ld(x) == c
    oh, ow = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    y = torch.empty((n, oh, ow, c), device=x.device, dtype=torch.float32)
    check(_lib.lib().pw_maxpool3x3s2(_ptr(x), _ptr(y), n, h, w, c, oh, ow,
                                     _stream()), 'pw_maxpool3x3s2')
    return y


def upsample_nearest_add_(y, x):
    """y += nearest-upsample(x) to y's size (in place)."""
    n, oh, ow, c = y.shape
    assert cl_ld(y) == c and cl_ld(x) == c and x.shape[-1] == c
    check(_lib.lib().pw_upsample_nearest_add(_ptr(x), _ptr(y), n, x.shape[1],
                                             x.shape[2], oh, ow, c, _stream()),
          'pw_upsample_nearest_add')
    return y


def scale_channels(x, gate, out=None):
    """x [N,H,W,C] * gate [N,C]."""
    n, c = x.shape[0], x.shape[-1]
    pixels = x[0, ..., 0].numel()
    if out is None:
        out = torch.empty(x.shape, device=x.device, dtype=torch.float32)
    gate = gate.contiguous()
    check(_lib.lib().pw_scale_channels(_ptr(x), cl_ld(x), _ptr(gate),
                                       _ptr(out), cl_ld(out), n, pixels, c,
                                       _stream()), 'pw_scale_channels')
    return out


def dense_chains(x2d, chains):
    """x2d [rows, C]; ``chains``: list of lists of (PackedConv, act) -- every chain a
    stack of dense layers applied to the same rows.  ONE launch (pw_dense_chains);
    returns a list with one [rows, cout_last] tensor per chain."""
    _require_cuda(x2d)
    assert x2d.dim() == 2 and x2d.stride(1) == 1 and 1 <= len(chains) <= 4
    rows = x2d.shape[0]
    arr = (_lib.DenseChain * len(chains))()
    outs = []
    for ci, layers in enumerate(chains):
        assert 1 <= len(layers) <= 4
        for li, (pc, act) in enumerate(layers):
            assert pc.k == (1, 1, 1)
            L = arr[ci].layer[li]
            L.w = pc.w.data_ptr()
            L.scale = pc.scale.data_ptr() if pc.scale is not None else None
            L.bias = pc.bias.data_ptr() if pc.bias is not None else None
            L.cin, L.cout, L.w_ld, L.act = pc.cin, pc.cout, pc.w_ld, ACT[act]
        out = torch.empty((rows, layers[-1][0].cout), device=x2d.device, dtype=torch.float32)
        arr[ci].n_layers, arr[ci].out, arr[ci].out_ld = len(layers), out.data_ptr(), out.stride(0)
        outs.append(out)
    check(_lib.lib().pw_dense_chains(arr, len(chains), _ptr(x2d), x2d.stride(0), rows,
                                     _stream()), 'pw_dense_chains')
    return outs


def global_avgpool(x):
    n, c = x.shape[0], x.shape[-1]
    pixels = x[0, ..., 0].numel()
    y = torch.empty((n, c), device=x.device, dtype=torch.float32)
    check(_lib.lib().pw_global_avgpool(_ptr(x), cl_ld(x), _ptr(y), n, pixels,
                                       c, _stream()), 'pw_global_avgpool')
    return y


def broadcast_channels_(out, v):
    """out [N,H,W,C] (may be a slice) = v [N,C] broadcast over pixels."""
    n, c = out.shape[0], out.shape[-1]
    pixels = out[0, ..., 0].numel()
    v = v.contiguous()
    check(_lib.lib().pw_broadcast_channels(_ptr(v), _ptr(out), cl_ld(out), n,
                                           pixels, c, _stream()),
          'pw_broadcast_channels')
    return out


def softmax_depth(logits_cl, d):
    """logits_cl [N,H,W,>=d] -> planar probabilities [N,d,H,W] (contiguous,
    the reference's layout)."""
    n, h, w = logits_cl.shape[:3]
    y = torch.empty((n, d, h, w), device=logits_cl.device, dtype=torch.float32)
    check(_lib.lib().pw_softmax_depth(_ptr(logits_cl), cl_ld(logits_cl), None,
                                      _ptr(y), n, h * w, d, _stream()),
          'pw_softmax_depth')
    return y


def cost_volume(curr, prev, cam, xs, ys, ds, bias, img_hw, pad_to=None):
    """-> [n,h,w,d] (or [n,h,w,pad_to] with zero padding channels, so that a
    following conv sees a 32-multiple Cin)."""
    n, h, w, c = curr.shape
    assert cl_ld(curr) == c and cl_ld(prev) == c and prev.shape == curr.shape
    d = ds.numel()
    ld = pad_to or d
    out = torch.empty((n, h, w, ld), device=curr.device, dtype=torch.float32)
    check(_lib.lib().pw_cost_volume(_ptr(curr), _ptr(prev), _ptr(cam), _ptr(xs),
                                    _ptr(ys), _ptr(ds), _ptr(out), ld, n, h, w,
                                    c, d, float(bias), int(img_hw[0]),
                                    int(img_hw[1]), _stream()),
          'pw_cost_volume')
    return out


# ----------------------------------------------------------------------- lift
def bev_pool_v2_(depth, feat, ranks_depth, ranks_feat, ranks_bev,
                 interval_starts, interval_lengths, out):
    """In-place drop-in of bev_pool_v2_ext.bev_pool_v2_forward (note the
    reference's pybind argument order is lengths-before-starts,
    src/bev_pool.cpp:30-57; this wrapper takes keywords-by-name order)."""
    _require_cuda(depth, feat, out)
    c = feat.shape[-1]
    check(_lib.lib().pw_bev_pool_v2(
        c, int(interval_starts.numel()), _ptr(depth), _ptr(feat),
        _ptr(ranks_depth), _ptr(ranks_feat), _ptr(ranks_bev),
        _ptr(interval_starts), _ptr(interval_lengths), _ptr(out), _stream()),
        'pw_bev_pool_v2')
    return out


def bev_pool_v2_grad_(out_grad, depth, feat, ranks_depth, ranks_feat, ranks_bev,
                      interval_starts, interval_lengths, depth_grad, feat_grad):
    """In-place drop-in of bev_pool_v2_ext.bev_pool_v2_backward: lists sorted by
    ranks_feat, intervals = runs of equal ranks_feat, zero-initialised grads
    (bev_pool.py:43-83)."""
    _require_cuda(out_grad, depth, feat, depth_grad, feat_grad)
    c = feat.shape[-1]
    check(_lib.lib().pw_bev_pool_v2_grad(
        c, int(interval_starts.numel()), _ptr(out_grad), _ptr(depth), _ptr(feat),
        _ptr(ranks_depth), _ptr(ranks_feat), _ptr(ranks_bev),
        _ptr(interval_starts), _ptr(interval_lengths), _ptr(depth_grad),
        _ptr(feat_grad), _stream()), 'pw_bev_pool_v2_grad')
    return depth_grad, feat_grad


def _cam_params(fn_name, floats, pose44, intrin, post_rot, post_tran):
    _require_cuda(pose44, intrin, post_rot, post_tran)
    n = pose44.numel() // 16
    args = [t.contiguous().float() for t in (pose44, intrin, post_rot,
                                             post_tran)]
    cam = torch.empty((n, floats), device=pose44.device, dtype=torch.float32)
    check(getattr(_lib.lib(), fn_name)(n, *[_ptr(a) for a in args], _ptr(cam),
                                       _stream()), fn_name)
    return cam


def lift_camera_params(sensor2ego, intrin, post_rot, post_tran):
    """[..,4,4],[..,3,3],[..,3,3],[..,3] -> cam [n, 24] (device)."""
    return _cam_params('pw_lift_camera_params', 24, sensor2ego, intrin,
                       post_rot, post_tran)


def cv_camera_params(k2s_sensor, intrin, post_rot, post_tran):
    return _cam_params('pw_cv_camera_params', 48, k2s_sensor, intrin, post_rot,
                       post_tran)


def _f3(v):
    return (ctypes.c_float * 3)(*[float(x) for x in v])


def lift_ranks(cam, bda, xs, ys, ds, lower, interval, b, n, grid):
    d, h, w = ds.numel(), ys.numel(), xs.numel()
    rank = torch.empty(b * n * d * h * w, device=cam.device, dtype=torch.int32)
    check(_lib.lib().pw_lift_ranks(_ptr(cam), _ptr(bda), _ptr(xs), _ptr(ys),
                                   _ptr(ds), _f3(lower), _f3(interval), b, n,
                                   d, h, w, grid[0], grid[1], grid[2],
                                   _ptr(rank), _stream()), 'pw_lift_ranks')
    return rank


_ws_cache = {}


def lift_fused(depth, feat_cl, cam, bda, xs, ys, ds, lower, interval, b, n,
               grid, out=None):
    """depth [b*n,D,H,W] planar probabilities; feat_cl [b*n,H,W,C] (may be a
    channel slice).  Returns the pooled volume as a cl array [b,Z,Y,X,C]."""
    _require_cuda(depth, feat_cl)
    d, h, w = ds.numel(), ys.numel(), xs.numel()
    c = feat_cl.shape[-1]
    gx, gy, gz = grid
    assert depth.is_contiguous() and depth.shape == (b * n, d, h, w)
    if out is None:
        out = torch.empty((b, gz, gy, gx, c), device=depth.device,
                          dtype=torch.float32)
    nbytes = int(_lib.lib().pw_lift_workspace_bytes(b, n, d, h, w, gx, gy, gz))
    key = (depth.device, torch.cuda.current_stream().cuda_stream)
    ws = _ws_cache.get(key)
    if ws is None or ws.numel() < nbytes:
        # control words must start zeroed; the kernel re-zeroes them on exit
        ws = torch.zeros(nbytes, device=depth.device, dtype=torch.uint8)
        _ws_cache[key] = ws
    check(_lib.lib().pw_lift_fused(
        _ptr(depth), _ptr(feat_cl), cl_ld(feat_cl), _ptr(cam), _ptr(bda),
        _ptr(xs), _ptr(ys), _ptr(ds), _f3(lower), _f3(interval), b, n, d, h, w,
        c, gx, gy, gz, _ptr(out), _ptr(ws), _stream()), 'pw_lift_fused')
    return out


def lift_prepare(cam, bda, xs, ys, ds, lower, interval, b, n, grid):
    """Build the voxel point lists for constant cameras (accelerate=True);
    returns the workspace tensor to hand to ``lift_pool``."""
    _require_cuda(cam, bda)
    d, h, w = ds.numel(), ys.numel(), xs.numel()
    gx, gy, gz = grid
    nbytes = int(_lib.lib().pw_lift_workspace_bytes(b, n, d, h, w, gx, gy, gz))
    ws = torch.zeros(nbytes, device=cam.device, dtype=torch.uint8)
    check(_lib.lib().pw_lift_prepare(
        _ptr(cam), _ptr(bda), _ptr(xs), _ptr(ys), _ptr(ds), _f3(lower),
        _f3(interval), b, n, d, h, w, gx, gy, gz, _ptr(ws), _stream()),
        'pw_lift_prepare')
    return ws


def lift_pool(depth, feat_cl, ws, b, n, grid, out=None):
    """Pool with the lists of ``lift_prepare``: one pass over the output."""
    _require_cuda(depth, feat_cl, ws)
    _, d, h, w = depth.shape
    c = feat_cl.shape[-1]
    gx, gy, gz = grid
    assert depth.is_contiguous() and depth.shape[0] == b * n
    if out is None:
        out = torch.empty((b, gz, gy, gx, c), device=depth.device,
                          dtype=torch.float32)
    check(_lib.lib().pw_lift_pool(
        _ptr(depth), _ptr(feat_cl), cl_ld(feat_cl), b, n, d, h, w, c, gx, gy,
        gz, _ptr(out), _ptr(ws), _stream()), 'pw_lift_pool')
    return out


# ---------------------------------------------------------------- Swin backbone
def layernorm(x, gamma, beta, eps=1e-5, out=None):
    """nn.LayerNorm over the channels of a cl array [..., C] (x / out may be channel
    slices of wider arrays)."""
    _require_cuda(x, gamma, beta, out)
    c = x.shape[-1]
    if out is None:
        out = torch.empty(x.shape, device=x.device, dtype=torch.float32)
    assert out.shape == x.shape and gamma.numel() == c and beta.numel() == c
    rows = x[..., 0].numel()
    check(_lib.lib().pw_layernorm(_ptr(x), cl_ld(x), _ptr(gamma), _ptr(beta), float(eps),
                                  _ptr(out), cl_ld(out), rows, c, _stream()), 'pw_layernorm')
    return out


def patch_merge_ln(x, gamma_kkc, beta_kkc, eps=1e-5):
    """x [B,H,W,C] -> LayerNorm of the 2x2-gathered tokens [B,ceil(H/2),ceil(W/2),4C],
    channels ordered (ky, kx, c) (gamma / beta given in that order)."""
    _require_cuda(x, gamma_kkc, beta_kkc)
    b, h, w, c = x.shape
    assert gamma_kkc.numel() == 4 * c and beta_kkc.numel() == 4 * c
    out = torch.empty((b, (h + 1) // 2, (w + 1) // 2, 4 * c), device=x.device,
                      dtype=torch.float32)
    check(_lib.lib().pw_patch_merge_ln(_ptr(x), cl_ld(x), b, h, w, c, _ptr(gamma_kkc),
                                       _ptr(beta_kkc), float(eps), _ptr(out), 4 * c,
                                       _stream()), 'pw_patch_merge_ln')
    return out


def window_attention(qkv, qkv_bias, table_t, heads, ws, shift, scale, out=None):
    """qkv [B,H,W,3C] (q|k|v, head-major, head dim 32) -> shifted-window attention output
    [B,H,W,C]; table_t [heads,(2ws-1)^2]."""
    _require_cuda(qkv, qkv_bias, table_t, out)
    b, h, w, c3 = qkv.shape
    c = c3 // 3
    assert c * 3 == c3 and c == heads * 32, 'head dim 32 (every Swin variant)'
    assert table_t.shape == (heads, (2 * ws - 1) ** 2) and table_t.is_contiguous()
    if out is None:
        out = torch.empty((b, h, w, c), device=qkv.device, dtype=torch.float32)
    check(_lib.lib().pw_window_attention(_ptr(qkv), cl_ld(qkv), _ptr(qkv_bias), _ptr(table_t),
                                         _ptr(out), cl_ld(out), b, h, w, c, heads, ws, shift,
                                         float(scale), _stream()), 'pw_window_attention')
    return out


# ------------------------------------------------------------------ 3-D side
def upsample_trilinear_(out, x):
    """out [B,OZ,OY,OX,C] (may be a channel slice) = trilinear(x), align_corners."""
    b, z, y, xx, c = x.shape
    check(_lib.lib().pw_upsample_trilinear(
        _ptr(x), cl_ld(x), _ptr(out), cl_ld(out), b, z, y, xx, c, out.shape[1],
        out.shape[2], out.shape[3], _stream()), 'pw_upsample_trilinear')
    return out


def upsample_trilinear2(x1, x2, out_spatial):
    """up(x1) + up(x2) -> [B,OZ,OY,OX,C] (trilinear, align_corners)."""
    b, z1, y1, w1, c = x1.shape
    _, z2, y2, w2, c2 = x2.shape
    assert c == c2 and x2.shape[0] == b
    oz, oy, ox = out_spatial
    out = torch.empty((b, oz, oy, ox, c), device=x1.device, dtype=torch.float32)
    check(_lib.lib().pw_upsample_trilinear2(
        _ptr(x1), cl_ld(x1), z1, y1, w1, _ptr(x2), cl_ld(x2), z2, y2, w2,
        _ptr(out), c, b, c, oz, oy, ox, _stream()), 'pw_upsample_trilinear2')
    return out


def copy_channels_(out, x):
    c = x.shape[-1]
    pixels = x[..., 0].numel()
    check(_lib.lib().pw_copy_channels(_ptr(x), cl_ld(x), _ptr(out), cl_ld(out),
                                      pixels, c, _stream()),
          'pw_copy_channels')
    return out


def argmax_zyx_to_xyz(logits_cl):
    """logits [1,Z,Y,X,ncls] -> uint8 [X,Y,Z] (torch.argmax semantics)."""
    _, gz, gy, gx, ncls = logits_cl.shape
    occ = torch.empty((gx, gy, gz), device=logits_cl.device, dtype=torch.uint8)
    check(_lib.lib().pw_argmax_zyx_to_xyz(_ptr(logits_cl), cl_ld(logits_cl),
                                          ncls, _ptr(occ), gx, gy, gz,
                                          _stream()), 'pw_argmax_zyx_to_xyz')
    return occ


def argmax_geo_zyx_to_xyz(logits_cl, free_idx, geo_value):
    """logits [1,Z,Y,X,ncls] -> uint8 [2,X,Y,Z]: [0] the argmax grid, [1] the
    geometry grid (0 where class != free_idx else geo_value) -- one buffer, so
    one device->host transfer brings both."""
    _, gz, gy, gx, ncls = logits_cl.shape
    both = torch.empty((2, gx, gy, gz), device=logits_cl.device, dtype=torch.uint8)
    check(_lib.lib().pw_argmax_geo_zyx_to_xyz(
        _ptr(logits_cl), cl_ld(logits_cl), ncls, int(free_idx), int(geo_value),
        _ptr(both[0]), _ptr(both[1]), gx, gy, gz, _stream()),
        'pw_argmax_geo_zyx_to_xyz')
    return both


def copy_rows_(dst, src, stream=None):
    """dst [R, ...] (rows contiguous, device) <- src [R, ...] whose rows are
    contiguous but strided (pinned host or device), as ONE asynchronous
    cudaMemcpy2D on ``stream`` (default: the current stream)."""
    assert dst.is_cuda and dst.shape == src.shape and dst.dtype == src.dtype
    assert dst[0].is_contiguous() and src[0].is_contiguous()
    rows = dst.shape[0]
    row_bytes = dst[0].numel() * dst.element_size()
    pitch = lambda t: (t.stride(0) if rows > 1 else t[0].numel()) * t.element_size()
    st = stream.cuda_stream if stream is not None else \
        torch.cuda.current_stream(dst.device).cuda_stream
    check(_lib.lib().pw_copy_rows(_ptr(dst), pitch(dst), _ptr(src), pitch(src),
                                  row_bytes, rows, ctypes.c_void_p(st)),
          'pw_copy_rows')
    return dst


def density_occ_zyx_to_xyz(density_cl, semantic_cl, thr, empty_idx):
    """density [1,Z,Y,X,1+] (channel 0 used), semantic [1,Z,Y,X,ncls]."""
    _, gz, gy, gx, ncls = semantic_cl.shape
    occ = torch.empty((gx, gy, gz), device=semantic_cl.device,
                      dtype=torch.uint8)
    geo = torch.empty_like(occ)
    check(_lib.lib().pw_density_occ_zyx_to_xyz(
        _ptr(density_cl), cl_ld(density_cl), _ptr(semantic_cl),
        cl_ld(semantic_cl), ncls, float(thr), int(empty_idx), _ptr(occ),
        _ptr(geo), gx, gy, gz, _stream()), 'pw_density_occ_zyx_to_xyz')
    return occ, geo


def zyx_to_xyz(x_cl):
    """[B,Z,Y,X,C] -> contiguous [B,X,Y,Z,C] (reference voxel_feats order)."""
    b, gz, gy, gx, c = x_cl.shape
    assert cl_ld(x_cl) == c
    y = torch.empty((b, gx, gy, gz, c), device=x_cl.device,
                    dtype=torch.float32)
    check(_lib.lib().pw_zyx_to_xyz(_ptr(x_cl), _ptr(y), b, gz, gy, gx, c,
                                   _stream()), 'pw_zyx_to_xyz')
    return y


# -------------------------------------------------------------------- render
def raw2alpha(density, shift, interval):
    _require_cuda(density)
    density = density.contiguous()
    exp_d = torch.empty_like(density)
    alpha = torch.empty_like(density)
    check(_lib.lib().pw_raw2alpha(_ptr(density), float(shift), float(interval),
                                  density.numel(), _ptr(exp_d), _ptr(alpha),
                                  _stream()), 'pw_raw2alpha')
    return exp_d, alpha


def alpha2weight(alpha, ray_id, n_rays):
    _require_cuda(alpha, ray_id)
    alpha = alpha.contiguous()
    ray_id = ray_id.contiguous().long()
    n = alpha.numel()
    dev = alpha.device
    weight = torch.empty(n, device=dev)
    T = torch.empty(n, device=dev)
    last = torch.empty(n_rays, device=dev)
    i_s = torch.empty(n_rays, device=dev, dtype=torch.int64)
    i_e = torch.empty(n_rays, device=dev, dtype=torch.int64)
    check(_lib.lib().pw_alpha2weight(_ptr(alpha), _ptr(ray_id), n, n_rays,
                                     _ptr(weight), _ptr(T), _ptr(last),
                                     _ptr(i_s), _ptr(i_e), _stream()),
          'pw_alpha2weight')
    return weight, T, last, i_s, i_e


def raw2alpha_backward(exp_d, grad_back, interval):
    _require_cuda(exp_d, grad_back)
    exp_d, grad_back = exp_d.contiguous(), grad_back.contiguous()
    grad = torch.empty_like(exp_d)
    check(_lib.lib().pw_raw2alpha_backward(_ptr(exp_d), _ptr(grad_back),
                                           float(interval), exp_d.numel(),
                                           _ptr(grad), _stream()),
          'pw_raw2alpha_backward')
    return grad


def alpha2weight_backward(alpha, weight, T, alphainv_last, i_start, i_end,
                          grad_weights, grad_last):
    _require_cuda(alpha, weight, T, grad_weights)
    grad = torch.zeros_like(alpha)
    check(_lib.lib().pw_alpha2weight_backward(
        _ptr(alpha.contiguous()), _ptr(weight.contiguous()),
        _ptr(T.contiguous()), _ptr(alphainv_last.contiguous()),
        _ptr(i_start.contiguous()), _ptr(i_end.contiguous()),
        int(alphainv_last.numel()), _ptr(grad_weights.contiguous()),
        _ptr(grad_last.contiguous()), _ptr(grad), _stream()),
        'pw_alpha2weight_backward')
    return grad


def cumdist_thres(dist, thres):
    _require_cuda(dist)
    dist = dist.contiguous()
    n_rays, n_pts = dist.shape
    mask = torch.empty((n_rays, n_pts), device=dist.device, dtype=torch.bool)
    check(_lib.lib().pw_cumdist_thres(_ptr(dist), float(thres), n_rays, n_pts,
                                      _ptr(mask), _stream()),
          'pw_cumdist_thres')
    return mask


def render_rays(desc, rays, t_vals, bda, density_cl, semantic_cl, color_cl):
    """rays [R,16]; *_cl are cl arrays [Z,Y,X,c] (channel slices allowed).
    Returns (depth [R], semantic [R,n_sem], color [R,3], alphainv_last [R],
    valid [R] bool)."""
    _require_cuda(rays, density_cl, semantic_cl, color_cl)
    rays = rays.contiguous()
    r = rays.shape[0]
    dev = rays.device
    ns = desc.n_sem
    o_d = torch.empty(r, device=dev)
    o_s = torch.empty((r, ns), device=dev)
    o_c = torch.empty((r, 3), device=dev)
    o_l = torch.empty(r, device=dev)
    o_v = torch.empty(r, device=dev, dtype=torch.bool)
    bda = bda.contiguous()
    check(_lib.lib().pw_render_rays(
        ctypes.byref(desc), _ptr(rays), r, _ptr(t_vals), t_vals.numel(),
        _ptr(bda), _ptr(density_cl), cl_ld(density_cl), _ptr(semantic_cl),
        cl_ld(semantic_cl), _ptr(color_cl), cl_ld(color_cl), _ptr(o_d),
        _ptr(o_s), _ptr(o_c), _ptr(o_l), _ptr(o_v), _stream()),
        'pw_render_rays')
    return o_d, o_s, o_c, o_l, o_v


def render_loss_sums(rays, depth, sem, col, last, valid, class_weights):
    """The nine fp64 sums NerfHead.compute_loss needs (pw_render_loss_sums)."""
    _require_cuda(rays, depth, sem, col, last, valid, class_weights)
    rays = rays.contiguous()
    sums = torch.empty(9, device=rays.device, dtype=torch.float64)
    cw = class_weights.float().contiguous()
    assert cw.numel() == sem.shape[-1] and valid.dtype in (torch.bool, torch.uint8)
    check(_lib.lib().pw_render_loss_sums(
        _ptr(rays), rays.shape[0], sem.shape[-1], _ptr(depth.contiguous()),
        _ptr(sem.contiguous()), _ptr(col.contiguous()), _ptr(last.contiguous()),
        _ptr(valid.contiguous()), _ptr(cw), _ptr(sums), _stream()), 'pw_render_loss_sums')
    return sums


# ------------------------------------------------------------------ losses
def voxel_loss_stats(rows, target, camera_mask, class_weights, empty_idx,
                     ignore_index=255):
    """rows [V, C] fp32 logits (row pitch = rows.stride(0)), target uint8 [V],
    camera_mask uint8/bool [V] or None, class_weights fp32 [C].
    -> (stats fp64 [8+3C], losses fp32 [3] = ce, sem_scal, geo_scal) -- one pass
    over the logits (pw_voxel_loss_stats, csrc/losses.cu)."""
    _require_cuda(rows, target, camera_mask, class_weights)
    v, c = rows.shape
    assert rows.dtype == torch.float32 and rows.stride(1) == 1
    assert target.dtype == torch.uint8 and target.numel() == v and target.is_contiguous()
    if camera_mask is not None:
        camera_mask = camera_mask.reshape(-1).to(torch.uint8).contiguous()
        assert camera_mask.numel() == v
    cw = class_weights.float().contiguous()
    assert cw.numel() == c
    L = _lib.lib()
    stats = torch.empty(L.pw_voxel_loss_stats_size(c), device=rows.device,
                        dtype=torch.float64)
    losses = torch.empty(3, device=rows.device, dtype=torch.float32)
    check(L.pw_voxel_loss_stats(_ptr(rows), rows.stride(0), _ptr(target),
                                _ptr(camera_mask), v, c, int(ignore_index),
                                int(empty_idx), _ptr(cw), _ptr(stats),
                                _ptr(losses), _stream()), 'pw_voxel_loss_stats')
    return stats, losses


def voxel_loss_grad(rows, target, camera_mask, class_weights, empty_idx, stats,
                    w_ce, w_sem, w_geo, ignore_index=255):
    """d(w_ce*ce + w_sem*sem + w_geo*geo)/d rows, fp32 [V, C] contiguous."""
    _require_cuda(rows, target, camera_mask, class_weights, stats)
    v, c = rows.shape
    if camera_mask is not None:
        camera_mask = camera_mask.reshape(-1).to(torch.uint8).contiguous()
    cw = class_weights.float().contiguous()
    grad = torch.empty((v, c), device=rows.device, dtype=torch.float32)
    check(_lib.lib().pw_voxel_loss_grad(
        _ptr(rows), rows.stride(0), _ptr(target), _ptr(camera_mask), v, c,
        int(ignore_index), int(empty_idx), _ptr(cw), _ptr(stats), float(w_ce),
        float(w_sem), float(w_geo), _ptr(grad), c, _stream()),
        'pw_voxel_loss_grad')
    return grad


def _depth_strides(depth_preds):
    """element strides (image, bin, y, x) of a logical [BN, D, h, w] tensor"""
    assert depth_preds.dim() == 4 and depth_preds.dtype == torch.float32
    return tuple(int(v) for v in depth_preds.stride())


def depth_loss(gt_depth, depth_preds, downsample, depth_min, depth_step, weight):
    """gt_depth [BN,H,W] fp32, depth_preds logical [BN,D,h,w] (any strides).
    -> (loss fp32 [1], labels int32 [BN*h*w], sums fp64 [2])  (pw_depth_loss)."""
    _require_cuda(gt_depth, depth_preds)
    bn, H, W = gt_depth.shape
    gt_depth = gt_depth.float().contiguous()
    D = depth_preds.shape[1]
    assert depth_preds.shape == (bn, D, H // downsample, W // downsample)
    si, sd, sy, sx = _depth_strides(depth_preds)
    cells = bn * (H // downsample) * (W // downsample)
    labels = torch.empty(cells, device=gt_depth.device, dtype=torch.int32)
    sums = torch.empty(2, device=gt_depth.device, dtype=torch.float64)
    loss = torch.empty(1, device=gt_depth.device, dtype=torch.float32)
    check(_lib.lib().pw_depth_loss(_ptr(gt_depth), bn, H, W, int(downsample),
                                   _ptr(depth_preds), si, sd, sy, sx, D,
                                   float(depth_min), float(depth_step), float(weight),
                                   _ptr(labels), _ptr(sums), _ptr(loss), _stream()),
          'pw_depth_loss')
    return loss, labels, sums


def depth_loss_grad(labels, depth_preds, sums, weight):
    """-> d loss / d depth_preds as a logical [BN, D, h, w] tensor."""
    bn, D, h, w = depth_preds.shape
    si, sd, sy, sx = _depth_strides(depth_preds)
    grad = torch.empty((bn * h * w, D), device=depth_preds.device, dtype=torch.float32)
    check(_lib.lib().pw_depth_loss_grad(_ptr(labels), bn, h, w, _ptr(depth_preds),
                                        si, sd, sy, sx, D, _ptr(sums), float(weight),
                                        _ptr(grad), _stream()), 'pw_depth_loss_grad')
    return grad.view(bn, h, w, D).permute(0, 3, 1, 2)


def lovasz_softmax_rows(rows, is_logits, target, camera_mask, ignore_label,
                        want_grad=True):
    """rows [V, C] fp32 probabilities (or logits), target uint8 [V], camera_mask
    uint8 [V] or None.  -> (loss fp32 [1], d loss / d probabilities [V, C] or
    None)  (pw_lovasz_softmax: keys, segmented sort, per-class scan)."""
    _require_cuda(rows, target, camera_mask)
    v, c = rows.shape
    assert rows.dtype == torch.float32 and rows.stride(1) == 1
    assert target.dtype == torch.uint8 and target.numel() == v
    L = _lib.lib()
    nbytes = L.pw_lovasz_workspace_bytes(v, c)
    if nbytes < 0:
        raise ValueError(f'lovasz_softmax: unsupported size {v} x {c}')
    ws = torch.empty(nbytes, device=rows.device, dtype=torch.uint8)
    loss = torch.empty(1, device=rows.device, dtype=torch.float32)
    gp = torch.empty((v, c), device=rows.device, dtype=torch.float32) \
        if want_grad else None
    check(L.pw_lovasz_softmax(_ptr(rows), rows.stride(0), int(bool(is_logits)),
                              _ptr(target), _ptr(camera_mask), v, c,
                              int(ignore_label), _ptr(ws), nbytes, _ptr(loss),
                              _ptr(gp), _stream()), 'pw_lovasz_softmax')
    return loss, gp


def softmax_backward(rows_logits, grad_probas):
    """d L / d logits [V, C] from d L / d softmax(logits)."""
    v, c = rows_logits.shape
    gl = torch.empty((v, c), device=rows_logits.device, dtype=torch.float32)
    check(_lib.lib().pw_softmax_backward(_ptr(rows_logits), rows_logits.stride(0),
                                         _ptr(grad_probas), v, c, _ptr(gl),
                                         _stream()), 'pw_softmax_backward')
    return gl


def focal_loss(rows, target, camera_mask, class_weights, radial, depth,
               gamma, alpha, loss_weight, ignore_index=255):
    """CustomFocalLoss on rows [V, C] logits (voxel order ..., H, W, D with
    ``depth`` = D and ``radial`` the fp32 [H*W] centre-distance map or None).
    -> (loss fp32 [1], sums fp64 [2] = weighted sum, kept count)."""
    _require_cuda(rows, target, camera_mask, class_weights, radial)
    v, c = rows.shape
    assert rows.dtype == torch.float32 and rows.stride(1) == 1
    cw = class_weights.float().contiguous()
    assert cw.numel() == c and target.dtype == torch.uint8 and target.numel() == v
    sums = torch.empty(2, device=rows.device, dtype=torch.float64)
    loss = torch.empty(1, device=rows.device, dtype=torch.float32)
    hw = radial.numel() if radial is not None else 0
    check(_lib.lib().pw_focal_loss(_ptr(rows), rows.stride(0), _ptr(target),
                                   _ptr(camera_mask), v, c, int(ignore_index), _ptr(cw),
                                   _ptr(radial), hw, int(depth), float(gamma),
                                   float(alpha), float(loss_weight), _ptr(sums),
                                   _ptr(loss), _stream()), 'pw_focal_loss')
    return loss, sums


def focal_loss_grad(rows, target, camera_mask, class_weights, radial, depth,
                    gamma, alpha, loss_weight, sums, ignore_index=255):
    v, c = rows.shape
    cw = class_weights.float().contiguous()
    grad = torch.empty((v, c), device=rows.device, dtype=torch.float32)
    hw = radial.numel() if radial is not None else 0
    check(_lib.lib().pw_focal_loss_grad(_ptr(rows), rows.stride(0), _ptr(target),
                                        _ptr(camera_mask), v, c, int(ignore_index),
                                        _ptr(cw), _ptr(radial), hw, int(depth),
                                        float(gamma), float(alpha), float(loss_weight),
                                        _ptr(sums), _ptr(grad), _stream()),
          'pw_focal_loss_grad')
    return grad
