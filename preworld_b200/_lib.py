"""ctypes binding of the C-ABI library (include/preworld_b200.h).

The library is the product: there is NO fallback.  If it is missing or fails
to load, importing an op raises; if a kernel launch reports an error the op
raises ``RuntimeError`` with the CUDA error code.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# PW_LIB selects another build of the SAME library (the instrumented `make debug` one)
LIB_PATH = os.environ.get('PW_LIB') or os.path.join(_HERE, 'lib', 'libpreworld_b200.so')

c_int = ctypes.c_int
c_ll = ctypes.c_longlong
c_f = ctypes.c_float
c_p = ctypes.c_void_p


class ConvDesc(ctypes.Structure):
    """struct pw_conv_desc"""
    _fields_ = [(n, c_int) for n in (
        'n', 'd', 'h', 'w', 'cin', 'in_ld',
        'od', 'oh', 'ow', 'cout', 'out_ld', 'res_ld', 'w_ld',
        'kd', 'kh', 'kw', 'sd', 'sh', 'sw', 'pd', 'ph', 'pw',
        'dd', 'dh', 'dw', 'act', 'act_channels')]


class RenderDesc(ctypes.Structure):
    """struct pw_render_desc"""
    _fields_ = [('scene_center', c_f * 3), ('scene_radius', c_f * 3),
                ('xyz_min', c_f * 3), ('xyz_max', c_f * 3),
                ('bg_len', c_f), ('act_shift', c_f), ('interval', c_f),
                ('step_size', c_f), ('fast_color_thres', c_f),
                ('radius', c_f), ('max_depth', c_f),
                ('world_len', c_int), ('gx', c_int), ('gy', c_int),
                ('gz', c_int), ('n_sem', c_int),
                ('vs_x', c_ll), ('vs_y', c_ll), ('vs_z', c_ll)]


class DenseLayer(ctypes.Structure):
    """struct pw_dense_layer"""
    _fields_ = [('w', c_p), ('scale', c_p), ('bias', c_p), ('cin', c_int),
                ('cout', c_int), ('w_ld', c_int), ('act', c_int)]


class DenseChain(ctypes.Structure):
    """struct pw_dense_chain"""
    _fields_ = [('layer', DenseLayer * 4), ('n_layers', c_int), ('out', c_p),
                ('out_ld', c_int)]


# name -> argtypes (restype is int unless noted); mirrors preworld_b200.h
SIGNATURES = {
    'pw_abi_version': [],
    'pw_launch_count': [],
    'pw_conv_fwd': [ctypes.POINTER(ConvDesc), c_p, c_p, c_p, c_p, c_p, c_p,
                    c_p],
    'pw_conv_umma_supported': [ctypes.POINTER(ConvDesc)],
    'pw_conv_umma_fwd': [ctypes.POINTER(ConvDesc), c_p, c_p, c_p, c_p, c_p,
                         c_p, c_p, c_p],
    'pw_conv_halo_supported': [ctypes.POINTER(ConvDesc)],
    'pw_conv_halo_fwd': [ctypes.POINTER(ConvDesc), c_p, c_p, c_p, c_p, c_p,
                         c_p, c_p, c_p],
    'pw_conv_fold_n': [c_int],
    'pw_conv_fold_supported': [ctypes.POINTER(ConvDesc)],
    'pw_conv_fold_fwd': [ctypes.POINTER(ConvDesc), c_p, c_p, c_p, c_p, c_p,
                         c_p, c_p, c_p],
    'pw_mlp2_supported': [c_int, c_int, c_int],
    'pw_mlp2': [c_p, c_int, c_ll, c_int, c_p, c_p, c_p, c_int, c_int, c_p, c_p,
                c_p, c_int, c_int, c_int, c_p, c_int, c_p, c_int, c_p],
    'pw_occhead_tail': [c_p, c_int, c_int, c_p, c_p, c_p, c_int, c_p, c_p, c_int,
                        c_p, c_int, c_p, c_p, c_int, c_int, c_int, c_int, c_int,
                        c_p],
    'pw_nchw_to_nhwc_pad': [c_p, c_ll, c_p, c_int, c_int, c_int, c_int, c_int,
                            c_p],
    'pw_nchw_to_s2d_nhwc': [c_p, c_ll, c_p, c_int, c_int, c_int, c_int, c_int,
                            c_p],
    'pw_nhwc_to_nchw': [c_p, c_int, c_p, c_int, c_int, c_ll, c_p],
    'pw_maxpool3x3s2': [c_p, c_p, c_int, c_int, c_int, c_int, c_int, c_int,
                        c_p],
    'pw_upsample_nearest_add': [c_p, c_p, c_int, c_int, c_int, c_int, c_int,
                                c_int, c_p],
    'pw_scale_channels': [c_p, c_int, c_p, c_p, c_int, c_int, c_ll, c_int,
                          c_p],
    'pw_global_avgpool': [c_p, c_int, c_p, c_int, c_ll, c_int, c_p],
    'pw_broadcast_channels': [c_p, c_p, c_int, c_int, c_ll, c_int, c_p],
    'pw_softmax_depth': [c_p, c_int, c_p, c_p, c_int, c_ll, c_int, c_p],
    'pw_cost_volume': [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_int, c_int, c_int,
                       c_int, c_int, c_int, c_f, c_int, c_int, c_p],
    'pw_bev_pool_v2': [c_int, c_int, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p,
                       c_p],
    'pw_bev_pool_v2_grad': [c_int, c_int, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p,
                            c_p, c_p, c_p],
    'pw_raw2alpha_backward': [c_p, c_p, c_f, c_ll, c_p, c_p],
    'pw_alpha2weight_backward': [c_p, c_p, c_p, c_p, c_p, c_p, c_int, c_p, c_p,
                                 c_p, c_p],
    'pw_lift_camera_params': [c_int, c_p, c_p, c_p, c_p, c_p, c_p],
    'pw_cv_camera_params': [c_int, c_p, c_p, c_p, c_p, c_p, c_p],
    'pw_lift_workspace_bytes': [c_int] * 8,
    'pw_lift_fused': [c_p, c_p, c_int, c_p, c_p, c_p, c_p, c_p,
                      ctypes.POINTER(c_f), ctypes.POINTER(c_f),
                      c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                      c_int, c_p, c_p, c_p],
    'pw_lift_prepare': [c_p, c_p, c_p, c_p, c_p, ctypes.POINTER(c_f),
                        ctypes.POINTER(c_f), c_int, c_int, c_int, c_int, c_int,
                        c_int, c_int, c_int, c_p, c_p],
    'pw_lift_pool': [c_p, c_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                     c_int, c_int, c_int, c_p, c_p, c_p],
    'pw_lift_ranks': [c_p, c_p, c_p, c_p, c_p, ctypes.POINTER(c_f),
                      ctypes.POINTER(c_f), c_int, c_int, c_int, c_int, c_int,
                      c_int, c_int, c_int, c_p, c_p],
    'pw_upsample_trilinear': [c_p, c_int, c_p, c_int, c_int, c_int, c_int,
                              c_int, c_int, c_int, c_int, c_int, c_p],
    'pw_upsample_trilinear2': [c_p, c_int, c_int, c_int, c_int, c_p, c_int, c_int,
                               c_int, c_int, c_p, c_int, c_int, c_int, c_int,
                               c_int, c_int, c_p],
    'pw_copy_channels': [c_p, c_int, c_p, c_int, c_ll, c_int, c_p],
    'pw_argmax_zyx_to_xyz': [c_p, c_int, c_int, c_p, c_int, c_int, c_int, c_p],
    'pw_argmax_geo_zyx_to_xyz': [c_p, c_int, c_int, c_int, c_int, c_p, c_p,
                                 c_int, c_int, c_int, c_p],
    'pw_copy_rows': [c_p, c_ll, c_p, c_ll, c_ll, c_ll, c_p],
    'pw_resample_rows_u8': [c_p, c_ll, c_int, c_int, c_int, c_int, c_p, c_p, c_p, c_int,
                            c_int, c_p, c_int, c_p],
    'pw_resample_view_norm': [c_p, c_int, c_int, c_int, c_int, c_p, c_p, c_p, c_int, c_int,
                              c_int, c_int, c_int, c_p, c_p, c_p, c_int, c_int, c_p],
    'pw_layernorm': [c_p, c_int, c_p, c_p, c_f, c_p, c_int, c_ll, c_int, c_p],
    'pw_patch_merge_ln': [c_p, c_int, c_int, c_int, c_int, c_int, c_p, c_p, c_f,
                          c_p, c_int, c_p],
    'pw_window_attention': [c_p, c_int, c_p, c_p, c_p, c_int, c_int, c_int, c_int,
                            c_int, c_int, c_int, c_int, c_f, c_p],
    'pw_density_occ_zyx_to_xyz': [c_p, c_int, c_p, c_int, c_int, c_f, c_int,
                                  c_p, c_p, c_int, c_int, c_int, c_p],
    'pw_zyx_to_xyz': [c_p, c_p, c_int, c_int, c_int, c_int, c_int, c_p],
    'pw_voxel_loss_stats_size': [c_int],
    'pw_voxel_loss_stats': [c_p, c_int, c_p, c_p, c_ll, c_int, c_int, c_int, c_p,
                            c_p, c_p, c_p],
    'pw_voxel_loss_grad': [c_p, c_int, c_p, c_p, c_ll, c_int, c_int, c_int, c_p,
                           c_p, c_f, c_f, c_f, c_p, c_int, c_p],
    'pw_depth_loss': [c_p, c_int, c_int, c_int, c_int, c_p, c_ll, c_ll, c_ll, c_ll,
                      c_int, c_f, c_f, c_f, c_p, c_p, c_p, c_p],
    'pw_depth_loss_grad': [c_p, c_int, c_int, c_int, c_p, c_ll, c_ll, c_ll, c_ll,
                           c_int, c_p, c_f, c_p, c_p],
    'pw_lovasz_workspace_bytes': [c_ll, c_int],
    'pw_lovasz_softmax': [c_p, c_int, c_int, c_p, c_p, c_ll, c_int, c_int, c_p, c_ll,
                          c_p, c_p, c_p],
    'pw_softmax_backward': [c_p, c_int, c_p, c_ll, c_int, c_p, c_p],
    'pw_focal_loss': [c_p, c_int, c_p, c_p, c_ll, c_int, c_int, c_p, c_p, c_int, c_int,
                      c_f, c_f, c_f, c_p, c_p, c_p],
    'pw_focal_loss_grad': [c_p, c_int, c_p, c_p, c_ll, c_int, c_int, c_p, c_p, c_int,
                           c_int, c_f, c_f, c_f, c_p, c_p, c_p],
    'pw_pts2ray': [c_p, c_p, c_p, c_p, c_p, c_p, c_ll, c_p, c_p],
    'pw_occ_confusion': [c_p, c_p, c_p, c_ll, c_int, c_int, c_p, c_p, c_p],
    'pw_raw2alpha': [c_p, c_f, c_f, c_ll, c_p, c_p, c_p],
    'pw_alpha2weight': [c_p, c_p, c_ll, c_int, c_p, c_p, c_p, c_p, c_p, c_p],
    'pw_cumdist_thres': [c_p, c_f, c_int, c_int, c_p, c_p],
    'pw_render_rays': [ctypes.POINTER(RenderDesc), c_p, c_int, c_p, c_int, c_p,
                       c_p, c_int, c_p, c_int, c_p, c_int, c_p, c_p, c_p, c_p,
                       c_p, c_p],
    'pw_dense_chains': [ctypes.POINTER(DenseChain), c_int, c_p, c_int, c_int, c_p],
    'pw_render_loss_sums': [c_p, c_int, c_int, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p],
}
_LONGLONG_RET = {'pw_launch_count', 'pw_lift_workspace_bytes',
                 'pw_lovasz_workspace_bytes'}

_lib = None


def lib():
    """Load (once) and return the ctypes handle; raises if the CUDA library
    has not been built -- there is no CPU path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f'{LIB_PATH} not found: build it with '
                '`python -c "import __graft_entry__ as g; g.build()"` '
                '(preworld_b200 has no CPU fallback)')
        L = ctypes.CDLL(LIB_PATH)
        for name, argtypes in SIGNATURES.items():
            fn = getattr(L, name)          # AttributeError == ABI mismatch
            fn.argtypes = argtypes
            fn.restype = c_ll if name in _LONGLONG_RET else c_int
        if L.pw_abi_version() != 1:
            raise RuntimeError('preworld_b200 ABI version mismatch')
        _lib = L
    return _lib


def check(code, what):
    if code != 0:
        raise RuntimeError(
            f'{what} failed: '
            + ('invalid argument (contract violation)' if code < 0
               else f'cudaError {code}'))


def launch_count():
    return int(lib().pw_launch_count())
