"""Locate an unmodified reference checkout and stitch it under this package's `maskrcnn_benchmark`
(INTEGRATION.md section 1): `modeling/ structures/ config/ utils/ engine/ solver/ data/` then resolve from the
reference tree while `_C` and `layers` stay this repository's.

Search order: $MRB_REFERENCE_ROOT, <repo>/baseline/_ref (the git-ignored mirror that travels to the GPU box,
baseline/sync_ref.py), /root/reference.  The image lacks yacs / apex / pycocotools and ships a torch without
torch._six; `activate(shims=True)` puts the small stand-ins of tests/_shims on sys.path for them."""
import os
import sys

_REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def find_reference_root():
    for cand in (os.environ.get("MRB_REFERENCE_ROOT"), os.path.join(_REPO, "baseline", "_ref"), "/root/reference"):
        if cand and os.path.isdir(os.path.join(cand, "maskrcnn_benchmark", "modeling")):
            return cand
    return None


def config_path(name, root=None):
    root = root or find_reference_root()
    return os.path.join(root, "configs", name)


def activate(root=None, shims=True):
    """-> reference root (str) or None when no checkout is available."""
    root = root or find_reference_root()
    if root is None:
        return None
    if shims:
        sh = os.path.join(_REPO, "tests", "_shims")
        if sh not in sys.path:
            sys.path.insert(0, sh)
        import mrb_test_compat  # noqa: F401  (torch._six, numpy aliases)
    import maskrcnn_benchmark
    ref_pkg = os.path.join(root, "maskrcnn_benchmark")
    if ref_pkg not in maskrcnn_benchmark.__path__:
        maskrcnn_benchmark.__path__.append(ref_pkg)      # our directory stays first: `_C`, `layers` come from here
    return root
