"""Import shim for the UNMODIFIED reference at /root/reference (build container only).

The reference pulls timm / matplotlib / h5py at package-import time; none of them is on the
correlation hot path, so absent modules are stubbed in ``sys.modules`` before the import.
Used only by ``make_goldens.py`` -- nothing under ``tests/`` proper, ``bench.py`` or the package
imports this file (``/root/reference`` does not exist on the GPU box).
"""
import importlib
import sys
import types

import torch

REFERENCE_ROOT = "/root/reference"


def _stub(name, **attrs):
    try:
        importlib.import_module(name)
        return
    except Exception:
        pass
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules[name] = mod
    parent, _, leaf = name.rpartition(".")
    if parent and parent in sys.modules:
        setattr(sys.modules[parent], leaf, mod)


class _DropPath(torch.nn.Module):
    def __init__(self, p=0.0):
        super().__init__()

    def forward(self, x):
        return x


def _no_timm(*a, **k):
    raise RuntimeError("timm backbone unavailable offline")


def install():
    _stub("timm")
    _stub("timm.models")
    _stub("timm.models.layers", trunc_normal_=torch.nn.init.trunc_normal_, DropPath=_DropPath)
    _stub("timm.models.mobilenetv3", tf_mobilenetv3_large_100=_no_timm)
    _stub("matplotlib")
    _stub("matplotlib.cm")
    _stub("matplotlib.pyplot")
    _stub("h5py")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
