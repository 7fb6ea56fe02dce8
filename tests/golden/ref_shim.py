"""Kept for ``make_goldens.py``: the import shim of the unmodified reference now lives in ``oracle/ref_shim.py``."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle.ref_shim import install  # noqa: E402,F401

REFERENCE_ROOT = "/root/reference"
