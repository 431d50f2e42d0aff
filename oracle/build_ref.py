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


def build_cuda(verbose=False):
    """oracle/_ref/cuda/mrb_ref_cuda.so: the reference's OWN CUDA sources (csrc/cuda/*.cu: ROIAlign, ROIPool, NMS,
    SigmoidFocalLoss, deform_conv, deform_psroi_pooling) compiled where they lie for sm_100a through the compat shims of
    oracle/ref_cuda + oracle/thc_compat (THC was removed from PyTorch; no kernel code is touched).  GPU-side checker and
    "reference CUDA kernels on this box" baseline.  ~2 minutes once (ninja caches); no-op without the reference tree."""
    here = os.path.dirname(os.path.abspath(__file__))
    ref = os.environ.get("MRB_REFERENCE", "/root/reference")
    csrc = os.path.join(ref, "maskrcnn_benchmark", "csrc")
    out = os.path.join(here, "_ref", "cuda")
    if not os.path.isdir(os.path.join(csrc, "cuda")):
        return None
    os.makedirs(out, exist_ok=True)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0")      # no GPU in the authoring container: do not probe
    from torch.utils.cpp_extension import load
    srcs = [os.path.join(here, "ref_cuda", f) for f in ("roi_nms_focal.cu", "dcn_kernels.cu", "dcn_host.cu", "dpool_kernels.cu",
                                                        "dpool_host.cu", "bind.cpp")]
    load(name="mrb_ref_cuda", sources=srcs,
         extra_include_paths=[csrc, os.path.join(here, "thc_compat"), os.path.join(here, "ref_cuda")],
         extra_cflags=["-O2", "-w"], extra_cuda_cflags=["-O2", "-w", "-gencode", "arch=compute_100a,code=sm_100a"],
         build_directory=out, verbose=verbose, is_python_module=False)
    return os.path.join(out, "mrb_ref_cuda.so")


if __name__ == "__main__":
    m = build(verbose=True)
    print("built" if m is not None else "reference tree absent: nothing built", file=sys.stderr)
    print(build_cuda(verbose=True), file=sys.stderr)
