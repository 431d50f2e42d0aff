"""Build oracle/_ref/mrb_ref_C*.so from the reference's own csrc/cpu sources.

Recipe (committed); outputs go ONLY to oracle/_ref/ (git-ignored, travels to
the GPU box with gpurun).  Needs /root/reference (or $MRB_REFERENCE); on a box
without it this script is a no-op and the prebuilt .so is used.
"""
import os
import sys


def build(verbose=False):
    here = os.path.dirname(os.path.abspath(__file__))
    ref = os.environ.get("MRB_REFERENCE", "/root/reference")
    csrc = os.path.join(ref, "maskrcnn_benchmark", "csrc")
    out = os.path.join(here, "_ref")
    if not os.path.isdir(csrc):
        return None
    os.makedirs(out, exist_ok=True)
    from torch.utils.cpp_extension import load
    mod = load(
        name="mrb_ref_C",
        sources=[os.path.join(here, "ref_shim.cpp")],
        extra_include_paths=[csrc],
        extra_cflags=[
            "-O2", "-w"],
        build_directory=out,
        verbose=verbose,
    )
    return mod


if __name__ == "__main__":
    m = build(verbose=True)
    print("built" if m is not None else "reference tree absent: nothing built", file=sys.stderr)
