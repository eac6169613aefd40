"""Import the UNMODIFIED reference as a live oracle / timed baseline.

TEST / BENCH INFRASTRUCTURE ONLY.  Two places are searched, in this order:
  1. ``$ASM_REFERENCE_ROOT`` or ``/root/reference``   (the read-only checkout; build container only)
  2. ``<repo>/baseline/_ref``                          (git-ignored staging directory; travels to the GPU box)
``stage()`` (called by ``__graft_entry__.build()`` in the build container) copies the handful of reference files
that the ASM path and the training-step example import into ``baseline/_ref`` so that ``bench.py --impl reference``
and ``examples/train_step.py`` can run the reference's own code on the GPU box.  ``baseline/_ref`` is listed in
``.gitignore``: reference sources never enter this repository's history.

``utils/Forward_model.py:5`` imports ``utils.functions`` which imports ``skimage.restoration.unwrap_phase`` (absent
in this image): a stub module is registered first, exactly as SURVEY.md section 8(c) describes.
"""
from __future__ import annotations

import os
import shutil
import sys
import types

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHECKOUT = os.environ.get("ASM_REFERENCE_ROOT", "/root/reference")
STAGED = os.path.join(REPO, "baseline", "_ref")
# what the hot path (and the config-5 training-step example) import from the reference
FILES = ["utils/Angular_Spectrum_Method.py", "utils/Forward_model.py", "utils/functions.py", "net.py", "function.py",
         "LICENSE"]


def _has(root: str) -> bool:
    return os.path.isfile(os.path.join(root, "utils", "Angular_Spectrum_Method.py"))


def root() -> str | None:
    for r in (CHECKOUT, STAGED):
        if _has(r):
            return r
    return None


def available() -> bool:
    return root() is not None


def stage() -> bool:
    """Copy FILES from the checkout into baseline/_ref (no-op when the checkout is absent).  Returns True if staged."""
    if not _has(CHECKOUT):
        return _has(STAGED)
    for f in FILES:
        src = os.path.join(CHECKOUT, f)
        if not os.path.isfile(src):
            continue
        dst = os.path.join(STAGED, f)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
    return _has(STAGED)


def _stub_skimage() -> None:
    try:
        import skimage.restoration  # noqa: F401
        return
    except Exception:
        pass
    sk = types.ModuleType("skimage")
    skr = types.ModuleType("skimage.restoration")

    def _no_unwrap(*_a, **_k):
        raise RuntimeError("skimage is not installed; unwrap=True is out of scope")

    skr.unwrap_phase = _no_unwrap
    sk.restoration = skr
    sys.modules["skimage"] = sk
    sys.modules["skimage.restoration"] = skr


def _import_from(ref_root: str, names):
    sys.dont_write_bytecode = True                       # the checkout is read-only
    _stub_skimage()
    # the reference's package is called ``utils``; make sure nobody else's is shadowing it
    for k in [k for k in sys.modules if k == "utils" or k.startswith("utils.") or k in ("net", "function")]:
        del sys.modules[k]
    sys.path.insert(0, ref_root)
    try:
        out = []
        for mod, attr in names:
            m = __import__(mod, fromlist=[attr])
            out.append(getattr(m, attr))
    finally:
        sys.path.remove(ref_root)
    return out


def load():
    """Returns (ASM, Holo_Generator, Back_prop) of the reference."""
    r = root()
    if r is None:
        raise RuntimeError(f"reference not found at {CHECKOUT} nor staged at {STAGED}")
    return tuple(_import_from(r, [("utils.Angular_Spectrum_Method", "ASM"), ("utils.Forward_model", "Holo_Generator"),
                                  ("utils.Forward_model", "Back_prop")]))


def load_net():
    """Returns the reference's ``net`` module (AdaIN network: ``net.Net``, ``net.Distance_G``, ``net.vgg``, ...)."""
    r = root()
    if r is None:
        raise RuntimeError(f"reference not found at {CHECKOUT} nor staged at {STAGED}")
    _stub_skimage()
    for k in [k for k in sys.modules if k in ("net", "function")]:
        del sys.modules[k]
    sys.dont_write_bytecode = True
    sys.path.insert(0, r)
    try:
        import net as ref_net  # noqa
    finally:
        sys.path.remove(r)
    return ref_net
