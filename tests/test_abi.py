"""CPU: the C-ABI library builds in-tree, loads, and exports every symbol include/ctunet_b200.h
declares (no compute calls without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ctunet_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ctu_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib_path():
    from ctunet_b200 import _lib
    if not os.path.isfile(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    return _lib.LIB_PATH


def test_header_and_binding_table_agree():
    from ctunet_b200 import _lib
    assert _declared() == sorted(_lib.SIGNATURES)


def test_library_exports_every_declared_symbol(lib_path):
    lib = ctypes.CDLL(lib_path)
    names = _declared()
    assert len(names) >= 40
    for n in names:
        assert hasattr(lib, n), n


def test_version_and_error_string(lib_path):
    from ctunet_b200 import _lib
    lib = _lib.load()
    assert lib.ctu_version() >= 100
    assert isinstance(_lib.last_error(), str)
    assert lib.ctu_has_tensor_path() in (0, 1)
    # size helpers are pure host arithmetic
    ch = _lib.int_array([7, 7])
    assert lib.ctu_conv_wpack_floats(3, 1, 2, ch) == 1 * 2 * 1 * 64
    assert lib.ctu_conv_wpack_floats(7, 3, 1, _lib.int_array([28])) == 1 * 4 * 27 * 64
    assert lib.ctu_conv_wpack_dgrad_floats(7, 3, 28) == 4 * 1 * 27 * 64
    assert lib.ctu_convt_wpack_floats(112, 2, _lib.int_array([56, 56])) == 14 * 14 * 8 * 64


def test_sass_is_sm100a(lib_path):
    """The shipped library carries sm_100a code only."""
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", lib_path], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    archs = set(re.findall(r"sm_(\d+a?)", out.stdout))
    assert archs == {"100a"}, archs


def test_kernel_selection_is_host_logic_and_needs_no_gpu():
    """Coverage predicates / geometry of the three tensor-path convolution kernels answer on the CPU (no launch):
    resident-weights kernel for h, w multiples of 16 above 16x16, weight-streaming kernel on small grids and for the
    wide 5^3 layers, tap-stationary weight gradient on 8x8 planes."""
    from ctunet_b200 import _lib
    from ctunet_b200.engine import tc_variant
    lib = _lib.load()
    assert tc_variant(3, [7], 7, 4, 128, 128, 128) == 1          # level 0 of UNetSP
    assert tc_variant(3, [14, 14, 1], 64, 4, 64, 64, 64) == 1    # fused up stage: concatenated sources
    assert tc_variant(3, [56], 56, 4, 16, 16, 16) == 2           # small grid: streamed weights preferred
    assert tc_variant(5, [64], 64, 4, 16, 16, 16) == 2           # 5^3 weights do not fit shared memory
    assert tc_variant(5, [128], 128, 4, 8, 8, 8) == 2            # 8 x 8 planes (M = 64 tiles)
    assert tc_variant(3, [7], 7, 1, 12, 12, 12) == 0             # neither: CUDA cores
    assert tc_variant(1, [14], 3, 4, 128, 128, 128) == 0         # 1^3 is the head kernel's job
    assert lib.ctu_conv_wide_wimg_bytes(5, 64, 64, 4, 16, 16, 16) == 64 * 64 * 125 * 2
    assert lib.ctu_conv_wide_wgrad_supported(5, 128, 128, 8, 8, 8) == 1
    assert lib.ctu_conv_wide_wgrad_supported(5, 136, 128, 8, 8, 8) == 0     # more than 128 input channels
    assert lib.ctu_conv_tc_wgrad_supported(3, 1, _lib.int_array([7]), 7, 8, 8, 8) == 0
    assert lib.ctu_upfuse_workspace_floats(28, 7, 3) > 0
