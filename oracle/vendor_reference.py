"""Recipe that stages the UNMODIFIED reference under the git-ignored ``oracle/_ref/`` (build container only).

The reference is pure Python (no ``setup.py`` / ``pyproject.toml``: ``pip install --target`` has nothing to install,
and there is nothing to compile), so "building" it is a byte-for-byte copy of

    /root/reference/nndepth/                                  -> oracle/_ref/nndepth/
    /root/reference/samples/kitti-stereo-2015/.../000000_10.png (image_2, image_3)  -> oracle/_ref/samples/...
    /root/reference/LICENSE                                   -> oracle/_ref/LICENSE

``oracle/_ref/`` is listed in ``.gitignore`` (never enters the history) but not in ``.gpurunignore``, so the copy
travels to the GPU box, where ``/root/reference`` does not exist.  There the tests, ``smoke()`` and ``bench.py``'s
reference arm import it through ``oracle/ref_shim.py`` as the checker / CPU baseline.  The product
(``nndepth_b200``) never imports it.  On a box without ``/root/reference`` this recipe is a no-op.
"""
import filecmp
import os
import shutil

from . import ref_shim


def vendor(verbose=False):
    """Copy the reference into ``oracle/_ref``; returns the vendored root, or None when there is no source tree."""
    src = ref_shim.SOURCE_ROOT
    dst = ref_shim.VENDORED_ROOT
    if not os.path.isfile(os.path.join(src, "nndepth", "__init__.py")):
        return dst if os.path.isdir(dst) else None
    os.makedirs(dst, exist_ok=True)
    pkg_dst = os.path.join(dst, "nndepth")
    if os.path.isdir(pkg_dst):
        shutil.rmtree(pkg_dst)
    shutil.copytree(os.path.join(src, "nndepth"), pkg_dst,
                    ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    for rel in ref_shim.KITTI_PAIR + ("LICENSE",):
        s, d = os.path.join(src, rel), os.path.join(dst, rel)
        if os.path.isfile(s):
            os.makedirs(os.path.dirname(d), exist_ok=True)
            shutil.copyfile(s, d)
    if verbose:
        print(f"vendored {src} -> {dst}")
    return dst


def verify():
    """True when every vendored ``.py`` file is byte-identical to its source (build container only)."""
    src = os.path.join(ref_shim.SOURCE_ROOT, "nndepth")
    dst = os.path.join(ref_shim.VENDORED_ROOT, "nndepth")
    if not (os.path.isdir(src) and os.path.isdir(dst)):
        return None
    for root, _, files in os.walk(src):
        for f in files:
            if not f.endswith(".py"):
                continue
            a = os.path.join(root, f)
            b = os.path.join(dst, os.path.relpath(a, src))
            if not (os.path.isfile(b) and filecmp.cmp(a, b, shallow=False)):
                return False
    return True


if __name__ == "__main__":
    print(vendor(verbose=True), "identical" if verify() else "DIFFERS")
