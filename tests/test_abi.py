"""CPU-side checks of the C ABI: the library builds/loads, exports every symbol the header declares,
and validates arguments before touching CUDA (no compute calls here -- there is no GPU in CI)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "nndepth_b200.h")


@pytest.fixture(scope="module")
def lib():
    from nndepth_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build_library()
    return _lib.load()


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nnd_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_entry_points():
    syms = declared_symbols()
    for name in ["nnd_corr1d_build", "nnd_corr1d_lookup", "nnd_corr1d_lookup_indices", "nnd_groupcorr_build",
                 "nnd_group_lookup", "nnd_geo_transpose_pool", "nnd_soft_argmin", "nnd_agcl_offset", "nnd_agcl_iter",
                 "nnd_avgpool_pairs", "nnd_last_error_string", "nnd_abi_version", "nnd_row_pitch"]:
        assert name in syms


def test_library_exports_every_declared_symbol(lib):
    from nndepth_b200 import _lib
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes prototype"


def test_abi_version_and_row_pitch(lib):
    assert lib.nnd_abi_version() == 1
    assert [lib.nnd_row_pitch(w) for w in (1, 4, 19, 39, 156, 160)] == [4, 4, 20, 40, 156, 160]
    assert lib.nnd_row_pitch(0) == 0


def test_invalid_arguments_are_reported_not_crashed(lib):
    from nndepth_b200 import _lib
    # null pointers
    assert lib.nnd_soft_argmin(None, 1, 1, 1, 1, None, None) == 1
    assert b"null" in lib.nnd_last_error_string()
    # a 1-wide level: linear_sampler would divide by zero (raft_stereo/utils.py:16)
    dummy = ctypes.c_void_p(16)
    lv = (ctypes.c_void_p * 1)(16)
    st = lib.nnd_corr1d_lookup(lv, _lib.int_array([1]), _lib.int_array([4]), dummy, 1, 1, 1, 1, 4, dummy, None)
    assert st == 1 and b"width" in lib.nnd_last_error_string()
    # channel count not divisible by the 4 AGCL groups
    st = lib.nnd_agcl_iter(dummy, dummy, dummy, 1, 30, 8, 8, 0, dummy, None)
    assert st == 1 and b"divisible" in lib.nnd_last_error_string()
    # group build reading past the channel count (the reference raises IndexError there)
    st = lib.nnd_groupcorr_build(dummy, dummy, 1, 32, 2, 8, 8, 8, 8, 1.0, 1, lv, _lib.int_array([8]), None)
    assert st == 1 and b"exceeds C" in lib.nnd_last_error_string()
    # unknown precision
    st = lib.nnd_corr1d_build(dummy, dummy, 1, 8, 1, 8, 8, 1, 7, lv, _lib.int_array([8]), None)
    assert st != 0
    # interleaved IGEV pyramids are an 8-group layout
    st = lib.nnd_gev_interleave_pool(dummy, 0, 16, 1, 4, 16, 2, 8, 1, lv, None)
    assert st == 1 and b"8 groups" in lib.nnd_last_error_string()
    st = lib.nnd_gev_lookup(lv, lv, dummy, 1, 8, 12, 2, 8, 1, 4, dummy, None)
    assert st == 1 and b"multiple of 8" in lib.nnd_last_error_string()
    # channels-last AGCL needs C % 16 == 0
    st = lib.nnd_agcl_offset_nhwc(dummy, dummy, dummy, dummy, 1, 24, 4, 4, 0, dummy, None)
    assert st == 1 and b"16" in lib.nnd_last_error_string()
    # upsampling rates are 2, 4, 8
    st = lib.nnd_convex_upsample(dummy, dummy, None, 1, 4, 4, 3, 1.0, 0, dummy, None)
    assert st == 1 and b"rate" in lib.nnd_last_error_string()
    # fused lookup + 1x1 convolution is the 4-level, radius-4 configuration
    st = lib.nnd_corr1d_lookup_conv1x1(lv, _lib.int_array([8]), _lib.int_array([8]), dummy, 1, 1, 8, 1, 4, dummy, None, 16, 1, 0, 0,
                                       dummy, None)
    assert st == 1 and b"4-level" in lib.nnd_last_error_string()
    # GRU glue: channel counts in quads
    st = lib.nnd_gru_gate_r(dummy, dummy, dummy, 8, 6, dummy, dummy, 40, None)
    assert st == 1 and b"ch % 4" in lib.nnd_last_error_string()


def test_python_layer_refuses_cpu_tensors():
    """No CPU fallback: the mirror classes must raise on CPU tensors instead of computing anything."""
    import torch
    import nndepth_b200 as nb
    f = torch.zeros(1, 8, 2, 8)
    with pytest.raises(RuntimeError, match="no CPU path"):
        nb.CorrBlock1D(f, f)
    with pytest.raises(RuntimeError, match="no CPU path"):
        nb.AGCL(f, f)
    with pytest.raises(RuntimeError, match="no CPU path"):
        nb.soft_argmin(torch.zeros(1, 4, 2, 2))


def test_product_package_does_not_import_the_oracle():
    """The oracle is test infrastructure; nothing under nndepth_b200/ may reference it."""
    pkg = os.path.join(ROOT, "nndepth_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), fn


TORCH_OPS = ["abi_version", "corr1d_build", "groupcorr_build", "avgpool_pairs", "corr1d_lookup", "corr1d_lookup_conv1x1",
             "corr1d_lookup_backward", "pyramid_unpool_", "corr1d_lookup_indices", "group_lookup", "geo_transpose_pool",
             "gev_interleave_pool", "gev_lookup", "soft_argmin", "gev_squeeze_soft_argmin", "nchw_to_nhwc", "agcl_offset",
             "agcl_iter", "convex_upsample", "agcl_warp", "agcl_offset_backward", "agcl_iter_backward", "volume_grad", "corr1d_skew", "corr1d_lookup_conv1x1_skewed", "corr1d_lookup_skewed", "corr1d_build_nhwc_f16"]


def test_torch_extension_registers_every_operator_and_refuses_cpu_tensors():
    """The thin PyTorch C++ extension over the C ABI (csrc/torch_ext.cpp): every operator the mirror classes call is
    registered under torch.ops.nndepth_b200, reports the C ABI's version, and raises (no compute, no fallback) on CPU
    tensors, wrong dtypes and non-contiguous tensors."""
    import torch
    from nndepth_b200 import _lib
    ops = _lib.ops()
    assert ops.abi_version() == _lib.load().nnd_abi_version() == 1
    for name in TORCH_OPS:
        assert hasattr(torch.ops.nndepth_b200, name), name
    with pytest.raises(RuntimeError, match="no CPU path"):
        ops.soft_argmin(torch.zeros(1, 4, 2, 2))
    with pytest.raises(RuntimeError, match="no CPU path"):
        ops.corr1d_lookup(torch.zeros(64), 8, torch.zeros(1, 1, 1, 8), 1, 4)
    with pytest.raises(RuntimeError, match="no CPU path"):
        ops.agcl_iter(torch.zeros(1, 16, 2, 2), torch.zeros(1, 16, 2, 2), torch.zeros(1, 2, 2, 2), False, False, None)
    # the mirror classes go through the operator library, not ctypes
    import inspect
    import nndepth_b200.corr as corr
    import nndepth_b200.igev as igev
    import nndepth_b200.agcl as agcl
    import nndepth_b200.upsample as upsample
    for mod in (corr, igev, agcl, upsample):
        src = inspect.getsource(mod)
        assert "_lib.ops()" in src and "_lib.load()" not in src, mod.__name__
