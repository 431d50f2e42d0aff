"""Mirror the UNMODIFIED reference python package and its configs into baseline/_ref/ (git-ignored, travels to the
GPU box with the gpurun snapshot exactly like the built .so files and oracle/_ref).

    baseline/_ref/maskrcnn_benchmark/{modeling,structures,config,utils,engine,solver,data,layers}/**/*.py
    baseline/_ref/configs/**/*.yaml

Nothing is edited and nothing from here enters the git history (see .gitignore).  The GPU box has no
/root/reference; tests and bench.py's `--model reference` arm find this mirror through mrb_b200.refenv.
csrc/ (the C++/CUDA sources) is NOT mirrored: the product replaces it, and the checker build of its CPU files is
oracle/_ref."""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")


def sync(src="/root/reference", verbose=False):
    pkg = os.path.join(src, "maskrcnn_benchmark")
    if not os.path.isdir(pkg):
        return None
    n = 0
    for sub, exts in (("maskrcnn_benchmark", (".py",)), ("configs", (".yaml",))):
        for root, dirs, files in os.walk(os.path.join(src, sub)):
            dirs[:] = [d for d in dirs if d not in ("csrc", "__pycache__")]
            rel = os.path.relpath(root, src)
            for f in files:
                if not f.endswith(exts):
                    continue
                d = os.path.join(DST, rel)
                os.makedirs(d, exist_ok=True)
                s, t = os.path.join(root, f), os.path.join(d, f)
                if not os.path.exists(t) or os.path.getmtime(t) < os.path.getmtime(s) or os.path.getsize(t) != os.path.getsize(s):
                    shutil.copy2(s, t)
                n += 1
    if verbose:
        print("baseline/_ref: %d files mirrored from %s" % (n, src))
    return DST


if __name__ == "__main__":
    print(sync(sys.argv[1] if len(sys.argv) > 1 else "/root/reference", verbose=True))
