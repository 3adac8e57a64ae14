"""The C-ABI library loads and exports every symbol include/*.h declares
(no compute calls: runs without a GPU)."""
import ctypes
import os
import re

from preworld_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, 'include', 'preworld_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(pw_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    names = _declared()
    assert len(names) >= 25
    L = ctypes.CDLL(_lib.LIB_PATH)
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing


def test_bindings_cover_the_header():
    assert sorted(_lib.SIGNATURES) == _declared()


def test_abi_version_and_launch_counter():
    L = _lib.lib()
    assert L.pw_abi_version() == 1
    assert _lib.launch_count() >= 0


def test_struct_sizes_match_header():
    # pw_conv_desc: 27 ints; pw_render_desc: 19 floats + 5 ints + 3 int64
    assert ctypes.sizeof(_lib.ConvDesc) == 27 * 4
    assert ctypes.sizeof(_lib.RenderDesc) == 19 * 4 + 5 * 4 + 3 * 8


def test_invalid_arguments_are_rejected_without_a_gpu():
    L = _lib.lib()
    d = _lib.ConvDesc()            # all zeros: n == 0
    assert L.pw_conv_fwd(ctypes.byref(d), None, None, None, None, None, None,
                         None) == -1
    assert L.pw_bev_pool_v2(0, 0, None, None, None, None, None, None, None,
                            None, None) == -1
