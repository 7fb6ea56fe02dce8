"""Import shim for the UNMODIFIED reference (test infrastructure -- never imported by ``nndepth_b200``).

The reference is a pure-Python package without ``setup.py``; ``oracle/vendor_reference.py`` (run by
``__graft_entry__.build()`` in the build container, where ``/root/reference`` exists) copies its ``nndepth``
package and the shipped KITTI sample pair, byte for byte, into the git-ignored ``oracle/_ref/`` so that they travel
to the GPU box with the snapshot.  This module puts that copy (or ``/root/reference`` itself when it is there)
on ``sys.path`` and stubs the three third-party modules the reference imports at package-import time but never
touches on the correlation path (``timm``, ``matplotlib``, ``h5py``).  A module is only
stubbed when its real import fails.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s reference / baseline legs may import this file.
"""
import importlib
import os
import sys
import types

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
VENDORED_ROOT = os.path.join(_HERE, "_ref")
SOURCE_ROOT = "/root/reference"
KITTI_PAIR = ("samples/kitti-stereo-2015/training/image_2/000000_10.png",
              "samples/kitti-stereo-2015/training/image_3/000000_10.png")


def reference_root():
    """Directory that holds the reference's ``nndepth`` package: the vendored copy first, else the source tree."""
    for root in (VENDORED_ROOT, SOURCE_ROOT):
        if os.path.isfile(os.path.join(root, "nndepth", "__init__.py")):
            return root
    return None


def available():
    return reference_root() is not None


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
    """Make ``import nndepth...`` resolve to the unmodified reference.  Raises if no copy of it is present."""
    root = reference_root()
    if root is None:
        raise ImportError(
            f"the reference is neither vendored under {VENDORED_ROOT} nor present at {SOURCE_ROOT}: run "
            "`python -c 'import __graft_entry__ as g; g.build()'` in the build container first"
        )
    _stub("timm")
    _stub("timm.models")
    _stub("timm.models.layers", trunc_normal_=torch.nn.init.trunc_normal_, DropPath=_DropPath)
    _stub("timm.models.mobilenetv3", tf_mobilenetv3_large_100=_no_timm)
    _stub("matplotlib")
    _stub("matplotlib.cm")
    _stub("matplotlib.pyplot")
    _stub("h5py")
    if root not in sys.path:
        sys.path.insert(0, root)
    return root


def kitti_sample_pair():
    """The shipped KITTI-2015 pair as ``(1,3,375,1242)`` fp32 tensors normalised like the reference's dataloader
    (``(x - 127.5) / 127.5``, RGB; ``nndepth/data/dataloaders/disparity/kitti2015_disparity.py:32``)."""
    import cv2
    import numpy as np
    root = reference_root()
    if root is None:
        raise ImportError("reference samples are not available (see install())")
    out = []
    for rel in KITTI_PAIR:
        path = os.path.join(root, rel)
        img = cv2.imread(path, cv2.IMREAD_COLOR)
        if img is None:
            raise FileNotFoundError(path)
        img = cv2.cvtColor(img, cv2.COLOR_BGR2RGB).astype(np.float32)
        out.append(torch.from_numpy((img - 127.5) / 127.5).permute(2, 0, 1)[None].contiguous())
    return out[0], out[1]
