"""TEST / BASELINE INFRASTRUCTURE ONLY -- ctypes binding of the reference's own
bev_pool_v2 CUDA source compiled UNMODIFIED for sm_100a (oracle/build_ref.sh ->
oracle/_ref/libbev_pool_v2_ref.so; host entry points of
mmdet3d/ops/bev_pool_v2/src/bev_pool_cuda.cu:125-139, declared at
src/bev_pool.cpp:7-14).  Never imported by ``preworld_b200``.

The reference launches on the legacy default stream (bev_pool_cuda.cu:127) --
callers keep torch on its default stream, which is the same stream."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, '_ref', 'libbev_pool_v2_ref.so')
_FWD = '_Z11bev_pool_v2iiPKfS0_PKiS2_S2_S2_S2_Pf'
_BWD = '_Z16bev_pool_v2_gradiiPKfS0_S0_PKiS2_S2_S2_S2_PfS3_'
_lib = None


def available():
    return os.path.exists(LIB)


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(LIB)
        p = ctypes.c_void_p
        getattr(_lib, _FWD).argtypes = [ctypes.c_int, ctypes.c_int] + [p] * 8
        getattr(_lib, _FWD).restype = None
        getattr(_lib, _BWD).argtypes = [ctypes.c_int, ctypes.c_int] + [p] * 11
        getattr(_lib, _BWD).restype = None
    return _lib


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def bev_pool_v2_forward(depth, feat, out, ranks_depth, ranks_feat, ranks_bev,
                        interval_lengths, interval_starts):
    """Same argument order as the pybind entry (src/bev_pool.cpp:30-57:
    lengths before starts); ``out`` zero-initialised by the caller
    (bev_pool.py:27).  CUDA tensors, fp32 / int32, contiguous."""
    c = feat.shape[-1]
    getattr(lib(), _FWD)(c, interval_lengths.numel(), _p(depth), _p(feat),
                         _p(ranks_depth), _p(ranks_feat), _p(ranks_bev),
                         _p(interval_starts), _p(interval_lengths), _p(out))


def bev_pool_v2_backward(out_grad, depth_grad, feat_grad, depth, feat,
                         ranks_depth, ranks_feat, ranks_bev, interval_lengths,
                         interval_starts):
    """src/bev_pool.cpp:74-111."""
    c = feat.shape[-1]
    getattr(lib(), _BWD)(c, interval_lengths.numel(), _p(out_grad), _p(depth),
                         _p(feat), _p(ranks_depth), _p(ranks_feat),
                         _p(ranks_bev), _p(interval_starts),
                         _p(interval_lengths), _p(depth_grad), _p(feat_grad))
