"""Bindings of the C ABI declared in ``include/nndepth_b200.h`` (``libnndepth_b200.so``).

* ``ops()`` -- the product binding: ``libnndepth_b200_torch.so``, a thin PyTorch C++ extension (``csrc/torch_ext.cpp``,
  ``TORCH_LIBRARY(nndepth_b200, ...)``) whose operators check device / dtype / contiguity / shapes, take the current
  CUDA stream and call the C ABI.  The mirror classes of this package (``corr.py``, ``igev.py``, ``agcl.py``,
  ``upsample.py``) call ``torch.ops.nndepth_b200.*`` through it.
* ``load()`` -- a ``ctypes`` view of the same C ABI, used by ``tests/test_abi.py`` (symbol / signature checks), by the
  timing tools, and by the model shell's out-of-scope glue kernels (ConvGRU gates, one-channel convolutions).

There is deliberately no CPU or PyTorch fallback: if a shared library is missing, or a tensor is not a contiguous
CUDA tensor of the expected dtype, the call raises.  PyTorch is used for device memory and the current stream only.
"""
import ctypes
import os
import subprocess
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libnndepth_b200.so")
TORCH_EXT_PATH = os.path.join(_HERE, "libnndepth_b200_torch.so")
CSRC_DIR = os.path.join(_HERE, "csrc")

NND_OK = 0
NND_MAX_LEVELS = 8
PREC_FP32 = 0
PREC_TF32 = 1

# every symbol of include/nndepth_b200.h: name -> (restype, argtypes)
_P = ctypes.c_void_p
_I = ctypes.c_int
SIGNATURES = {
    "nnd_abi_version": (_I, []),
    "nnd_last_error_string": (ctypes.c_char_p, []),
    "nnd_row_pitch": (_I, [_I]),
    "nnd_corr1d_build": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P]),
    "nnd_corr1d_build_nhwc_f16": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _P, _P, _P]),
    "nnd_groupcorr_build": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _I, ctypes.c_float, _I, _P, _P, _P]),
    "nnd_avgpool_pairs": (_I, [_P, _I, _I, _P, _I, ctypes.c_int64, _P]),
    "nnd_corr1d_lookup": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P]),
    "nnd_corr1d_lookup_conv1x1": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _I, _I, _I, _I, _P, _P]),
    "nnd_corr1d_lookup_backward": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P]),
    "nnd_avgpool_pairs_backward": (_I, [_P, _I, _I, _P, _I, ctypes.c_int64, _P]),
    "nnd_corr1d_skew": (_I, [_P, _P, _P, _I, _I, _I, _I, _P, _I, _P]),
    "nnd_corr1d_lookup_conv1x1_skewed": (_I, [_P, _P, _I, _P, _I, _I, _I, _I, _I, _P, _P, _I, _I, _I, _P, _P]),
    "nnd_corr1d_lookup_skewed": (_I, [_P, _P, _I, _P, _I, _I, _I, _I, _I, _P, _P]),
    "nnd_volume_grad": (_I, [_P, _I, _P, _I, _I, _I, _I, _I, _I, _I, ctypes.c_float, _I, _P, _P]),
    "nnd_corr1d_lookup_indices": (_I, [_P, _P, _I, _I, _I, _I, _I, _P, _P, _P]),
    "nnd_group_lookup": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P, _P]),
    "nnd_geo_transpose_pool": (_I, [_P, _I, _I, _I, _I, _I, _I, _P, _P, _P]),
    "nnd_soft_argmin": (_I, [_P, _I, _I, _I, _I, _P, _P]),
    "nnd_gev_squeeze_soft_argmin": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _P]),
    "nnd_agcl_offset": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P]),
    "nnd_agcl_iter": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _P, _P]),
    "nnd_flow_conv7x7_relu": (_I, [_P, _P, _P, _I, _I, _I, _I, _P, _I, _P]),
    "nnd_nhwc_cat_f16": (_I, [_P, _I, _I, _P, _I, _I, ctypes.c_longlong, _P, _P]),
    "nnd_flow_head_tail": (_I, [_P, _I, _P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P]),
    "nnd_gru_stage_f16": (_I, [_P, _I, _I, _I, ctypes.c_longlong, _P, _I, _I, _P]),
    "nnd_gru_gate_r_f16": (_I, [_P, _P, _P, ctypes.c_longlong, _I, _P, _P, _I, _P]),
    "nnd_gru_gate_h_f16": (_I, [_P, _P, _P, ctypes.c_longlong, _I, _P, _P, _I, _P, _P]),
    "nnd_gru_stage": (_I, [_P, _I, _I, _I, ctypes.c_longlong, _P, _I, _I, _P]),
    "nnd_gru_gate_r": (_I, [_P, _P, _P, ctypes.c_longlong, _I, _P, _P, _I, _P]),
    "nnd_gru_gate_h": (_I, [_P, _P, _P, ctypes.c_longlong, _I, _P, _P, _I, _P]),
    "nnd_convex_upsample": (_I, [_P, _P, _P, _I, _I, _I, _I, ctypes.c_float, _I, _P, _P]),
    "nnd_gev_interleave_pool": (_I, [_P, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P]),
    "nnd_gev_lookup": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P, _P]),
    "nnd_nchw_to_nhwc": (_I, [_P, _I, _I, _I, _I, _P, _P]),
    "nnd_agcl_offset_nhwc": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P]),
    "nnd_agcl_iter_nhwc": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _P]),
    "nnd_agcl_warp_nhwc": (_I, [_P, _P, _I, _I, _I, _I, _P, _P]),
    "nnd_agcl_offset_backward_nhwc": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P]),
    "nnd_agcl_iter_backward_nhwc": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P]),
}

_lock = threading.Lock()
_lib = None
_launches = 0


class NNDepthError(RuntimeError):
    """A non-zero status from the C ABI (message = ``nnd_last_error_string()``)."""


def build_library(verbose=False):
    """Compile ``csrc/*.cu`` into ``libnndepth_b200.so`` for sm_100a (nvcc cross-compiles without a GPU)."""
    torch_dir = os.path.dirname(torch.__file__)
    cmd = ["make", "-C", CSRC_DIR, "-j", str(min(8, os.cpu_count() or 1)), f"TORCH_DIR={torch_dir}",
           f"CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}"]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout)
    if res.returncode != 0:
        raise RuntimeError("building libnndepth_b200.so failed:\n" + res.stdout[-4000:])
    return LIB_PATH


def load():
    """Load the shared library (once) and attach the prototypes.  Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                f"or `make -C {CSRC_DIR}`.  nndepth_b200 has no CPU / PyTorch fallback."
            )
        lib = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the library lacks a declared symbol
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = lib
    return _lib


class _Ops:
    """``torch.ops.nndepth_b200`` with a launch counter in front (``launch_count()``)."""

    def __getattr__(self, name):
        fn = getattr(torch.ops.nndepth_b200, name)

        def call(*args):
            global _launches
            try:
                out = fn(*args)
            except RuntimeError as e:
                if " failed (status " in str(e):        # a non-zero status of the C ABI (TORCH_CHECK in torch_ext.cpp)
                    raise NNDepthError(str(e).split("\n")[0]) from None
                raise
            _launches += 1
            return out

        setattr(self, name, call)
        return call


_ops = None


def ops():
    """The PyTorch operator library over the C ABI (loaded once).  Raises if it is not built."""
    global _ops
    if _ops is None:
        with _lock:
            if _ops is None:
                if not os.path.exists(TORCH_EXT_PATH):
                    raise ImportError(
                        f"{TORCH_EXT_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                        f"or `make -C {CSRC_DIR}`.  nndepth_b200 has no CPU / PyTorch fallback."
                    )
                torch.ops.load_library(TORCH_EXT_PATH)
                if torch.ops.nndepth_b200.abi_version() != 1:
                    raise ImportError("libnndepth_b200_torch.so was built against another C ABI version")
                _ops = _Ops()
    return _ops


def last_error():
    msg = load().nnd_last_error_string()
    return msg.decode("utf-8", "replace") if msg else ""


def check(status, what):
    """Raise on a non-zero status; count successful C-ABI compute calls (one kernel launch each)."""
    global _launches
    if status != NND_OK:
        raise NNDepthError(f"{what} failed (status {status}): {last_error()}")
    _launches += 1


def launch_count():
    """Number of C-ABI compute calls issued by this process so far (each enqueues >= 1 kernel)."""
    return _launches


def require_cuda_f32(t, name):
    """The C ABI takes raw device pointers: reject everything that is not a dense fp32 CUDA tensor."""
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor, got {type(t).__name__}")
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: nndepth_b200 has no CPU path (got device {t.device})")
    if t.dtype != torch.float32:
        raise RuntimeError(f"{name} must be float32, got {t.dtype}")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous (shape {tuple(t.shape)}, strides {t.stride()})")
    return t


def as_cuda_f32(t, name, allow_grad=False):
    """Reference call sites hand over whatever the encoder produced: make it dense fp32 on its device."""
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor, got {type(t).__name__}")
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: nndepth_b200 has no CPU path (got device {t.device})")
    if torch.is_grad_enabled() and t.requires_grad and not allow_grad:
        raise RuntimeError(
            f"{name} requires grad: the B200 correlation path is inference-only (wrap the call in torch.no_grad())"
        )
    return t.detach().float().contiguous()


def stream_ptr(t):
    return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def ptr(t):
    return ctypes.c_void_p(t.data_ptr())


def ptr_array(tensors):
    arr = (ctypes.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])
    return arr


def int_array(values):
    return (ctypes.c_int * len(values))(*[int(v) for v in values])


def row_pitch(width):
    return (int(width) + 3) & ~3
