"""TEST INFRASTRUCTURE ONLY: numpy/ctypes front-end of oracle/oracle_ref.c.

Builds ``oracle/_build/liboracle_ref.so`` with gcc on first use (also built by
``__graft_entry__.build()``).  Never imported by ``preworld_b200``.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, 'oracle_ref.c')
_SO = os.path.join(_HERE, '_build', 'liboracle_ref.so')
_lib = None

_f32p = np.ctypeslib.ndpointer(np.float32, flags='C_CONTIGUOUS')
_i32p = np.ctypeslib.ndpointer(np.int32, flags='C_CONTIGUOUS')
_i64p = np.ctypeslib.ndpointer(np.int64, flags='C_CONTIGUOUS')
_u8p = np.ctypeslib.ndpointer(np.uint8, flags='C_CONTIGUOUS')
_int = ctypes.c_int
_flt = ctypes.c_float


def build(force=False):
    if force or not os.path.exists(_SO) or \
            os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        subprocess.check_call(
            ['gcc', '-O2', '-ffp-contract=off', '-fPIC', '-shared', '-o', _SO,
             _SRC, '-lm'])
    return _SO


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(build())
        L.pw_ref_bev_pool_v2_fwd.argtypes = [
            _int, _int, _f32p, _f32p, _i32p, _i32p, _i32p, _i32p, _i32p, _f32p]
        L.pw_ref_bev_pool_v2_bwd.argtypes = [
            _int, _int, _f32p, _f32p, _f32p, _i32p, _i32p, _i32p, _i32p, _i32p,
            _f32p, _f32p]
        L.pw_ref_lift_camera_params.argtypes = [
            _int, _f32p, _f32p, _f32p, _f32p, _f32p]
        L.pw_ref_lift_ranks.argtypes = [
            _int, _int, _int, _int, _int, _f32p, _f32p, _f32p, _f32p, _f32p,
            _f32p, _f32p, _int, _int, _int, _i32p]
        L.pw_ref_raw2alpha.argtypes = [_int, _f32p, _flt, _flt, _f32p]
        L.pw_ref_alpha2weight.argtypes = [
            _int, _int, _f32p, _i64p, _f32p, _f32p, _f32p, _i64p, _i64p]
        L.pw_ref_cumdist_thres.argtypes = [_int, _int, _f32p, _flt, _u8p]
        L.pw_ref_raw2alpha_bwd.argtypes = [_int, _f32p, _f32p, _flt, _f32p]
        L.pw_ref_alpha2weight_bwd.argtypes = [
            _int, _f32p, _f32p, _f32p, _f32p, _i64p, _i64p, _f32p, _f32p,
            _f32p]
        for f in ('pw_ref_bev_pool_v2_fwd', 'pw_ref_bev_pool_v2_bwd',
                  'pw_ref_raw2alpha_bwd', 'pw_ref_alpha2weight_bwd',
                  'pw_ref_lift_camera_params', 'pw_ref_lift_ranks',
                  'pw_ref_raw2alpha', 'pw_ref_alpha2weight',
                  'pw_ref_cumdist_thres'):
            getattr(L, f).restype = None
        _lib = L
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def bev_pool_v2_fwd(depth, feat, ranks_depth, ranks_feat, ranks_bev,
                    interval_starts, interval_lengths, n_voxels):
    """depth: flat [B*N*D*H*W]; feat: [B*N*H*W, C]; returns [n_voxels, C]."""
    feat = _f32(feat)
    c = feat.shape[-1]
    out = np.zeros((n_voxels, c), np.float32)
    lib().pw_ref_bev_pool_v2_fwd(
        c, len(interval_starts), _f32(depth).ravel(), feat.reshape(-1),
        _i32(ranks_depth), _i32(ranks_feat), _i32(ranks_bev),
        _i32(interval_starts), _i32(interval_lengths), out.reshape(-1))
    return out


def bev_pool_v2_bwd(out_grad, depth, feat, ranks_depth, ranks_feat, ranks_bev,
                    interval_starts, interval_lengths):
    """Inputs already sorted by ranks_feat (bev_pool.py:47-57).  Returns
    (depth_grad flat, feat_grad [*, C])."""
    feat = _f32(feat)
    c = feat.shape[-1]
    depth = _f32(depth).ravel()
    dg = np.zeros_like(depth)
    fg = np.zeros_like(feat.reshape(-1))
    lib().pw_ref_bev_pool_v2_bwd(
        c, len(interval_starts), _f32(out_grad).reshape(-1), depth,
        feat.reshape(-1), _i32(ranks_depth), _i32(ranks_feat), _i32(ranks_bev),
        _i32(interval_starts), _i32(interval_lengths), dg, fg)
    return dg, fg.reshape(feat.shape)


def lift_camera_params(sensor2ego, intrin, post_rot, post_tran):
    s = _f32(sensor2ego).reshape(-1, 16)
    n = s.shape[0]
    cam = np.zeros((n, 24), np.float32)
    lib().pw_ref_lift_camera_params(
        n, s.reshape(-1), _f32(intrin).reshape(-1), _f32(post_rot).reshape(-1),
        _f32(post_tran).reshape(-1), cam.reshape(-1))
    return cam


def lift_ranks(B, N, xs, ys, ds, cam, bda, lower, interval, grid_size):
    D, H, W = len(ds), len(ys), len(xs)
    rank = np.empty(B * N * D * H * W, np.int32)
    gx, gy, gz = (int(g) for g in grid_size)
    lib().pw_ref_lift_ranks(
        B, N, D, H, W, _f32(xs), _f32(ys), _f32(ds), _f32(cam).reshape(-1),
        _f32(bda).reshape(-1), _f32(lower), _f32(interval), gx, gy, gz, rank)
    return rank


def raw2alpha(density, shift, interval):
    d = _f32(density).ravel()
    a = np.empty_like(d)
    lib().pw_ref_raw2alpha(len(d), d, shift, interval, a)
    return a


def alpha2weight(alpha, ray_id, n_rays, full=False):
    a = _f32(alpha).ravel()
    rid = np.ascontiguousarray(ray_id, dtype=np.int64)
    w = np.empty_like(a)
    T = np.empty_like(a)
    last = np.empty(n_rays, np.float32)
    i_s = np.empty(n_rays, np.int64)
    i_e = np.empty(n_rays, np.int64)
    lib().pw_ref_alpha2weight(len(a), n_rays, a, rid, w, T, last, i_s, i_e)
    if full:
        return w, T, last, i_s, i_e
    return w, last


def raw2alpha_bwd(exp_d, grad_back, interval):
    e = _f32(exp_d).ravel()
    g = np.empty_like(e)
    lib().pw_ref_raw2alpha_bwd(len(e), e, _f32(grad_back).ravel(), interval, g)
    return g


def alpha2weight_bwd(alpha, weight, T, alphainv_last, i_start, i_end,
                     grad_weights, grad_last):
    a = _f32(alpha).ravel()
    g = np.zeros_like(a)
    lib().pw_ref_alpha2weight_bwd(
        len(alphainv_last), a, _f32(weight).ravel(), _f32(T).ravel(),
        _f32(alphainv_last).ravel(),
        np.ascontiguousarray(i_start, dtype=np.int64),
        np.ascontiguousarray(i_end, dtype=np.int64),
        _f32(grad_weights).ravel(), _f32(grad_last).ravel(), g)
    return g


def cumdist_thres(dist, thres):
    d = _f32(dist)
    m = np.empty(d.shape, np.uint8)
    lib().pw_ref_cumdist_thres(d.shape[0], d.shape[1], d.reshape(-1), thres,
                               m.reshape(-1))
    return m.astype(bool)
