"""TEST INFRASTRUCTURE ONLY -- recipe that stages the UNMODIFIED reference for the CPU baseline.

Copies the Python package of the reference (`/root/reference/projects`, *.py only: the hot path lives in
projects/mvsdetection/models/ray_marching.py) to `oracle/_ref/projects`.  `oracle/_ref/` is git-ignored -- the
reference's sources never enter this repository's history -- but it is NOT gpurun-ignored, so the copy travels to the
GPU box with the snapshot, where `bench.py`'s CPU legs time the reference's own functions through oracle/ref_shim.py
(`cpu_baseline.kind == "reference"`).  Run by `__graft_entry__.build()` wherever /root/reference is mounted; a no-op
elsewhere (the GPU box only uses the staged copy).
"""
import os
import shutil

_HERE = os.path.dirname(os.path.abspath(__file__))
SOURCE = os.environ.get("CNRMA_REFERENCE_SOURCE", "/root/reference")
TARGET = os.path.join(_HERE, "_ref")


def stage(verbose=False):
    src = os.path.join(SOURCE, "projects")
    if not os.path.isdir(os.path.join(src, "mvsdetection")):
        return os.path.isdir(os.path.join(TARGET, "projects", "mvsdetection"))
    dst = os.path.join(TARGET, "projects")
    n = 0
    for root, _dirs, files in os.walk(src):
        rel = os.path.relpath(root, src)
        for name in files:
            if not name.endswith(".py"):
                continue
            out_dir = os.path.join(dst, rel)
            os.makedirs(out_dir, exist_ok=True)
            s, d = os.path.join(root, name), os.path.join(out_dir, name)
            if not os.path.exists(d) or os.path.getmtime(d) < os.path.getmtime(s) or os.path.getsize(d) != os.path.getsize(s):
                shutil.copy2(s, d)
                n += 1
    if verbose:
        print(f"oracle/_ref: staged {n} file(s) of the reference's projects/ package")
    return True


if __name__ == "__main__":
    stage(verbose=True)
