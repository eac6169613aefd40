"""Import the UNMODIFIED reference (`/root/reference`) as a live oracle -- build container only.

TEST INFRASTRUCTURE ONLY.  The reference tree does not exist on the GPU box; callers must check
``available()`` first.  ``utils/Forward_model.py:5`` imports ``utils.functions`` which imports
``skimage.restoration.unwrap_phase`` (absent here): a stub module is registered first, exactly as
SURVEY.md section 8(c) describes.  Nothing is copied out of the reference tree.
"""
from __future__ import annotations

import os
import sys
import types

REF_ROOT = os.environ.get("ASM_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "utils", "Angular_Spectrum_Method.py"))


def load():
    """Returns (ASM, Holo_Generator, Back_prop) from the reference checkout."""
    if not available():
        raise RuntimeError(f"reference checkout not found at {REF_ROOT}")
    sys.dont_write_bytecode = True                       # the reference tree is read-only
    if "skimage" not in sys.modules:
        sk = types.ModuleType("skimage")
        skr = types.ModuleType("skimage.restoration")

        def _no_unwrap(*_a, **_k):
            raise RuntimeError("skimage is not installed; unwrap=True is out of scope")

        skr.unwrap_phase = _no_unwrap
        sk.restoration = skr
        sys.modules["skimage"] = sk
        sys.modules["skimage.restoration"] = skr
    # the reference's package is called ``utils``; make sure ours/anyone else's is not shadowing it
    for k in [k for k in sys.modules if k == "utils" or k.startswith("utils.")]:
        del sys.modules[k]
    sys.path.insert(0, REF_ROOT)
    try:
        from utils.Angular_Spectrum_Method import ASM          # noqa
        from utils.Forward_model import Holo_Generator, Back_prop  # noqa
    finally:
        sys.path.remove(REF_ROOT)
    return ASM, Holo_Generator, Back_prop
