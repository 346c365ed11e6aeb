"""Builds libsurfel_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

No torch headers are involved: the library is plain CUDA + a C ABI (include/surfel_rasterizer.h).
preprocess_fwd.cu is compiled with -fmad=false so that the integer-valued outputs (radii, tile
rects, sort keys) are bit-identical to the CPU oracle ("Parity").
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIB_DIR, "libsurfel_b200.so")
OBJ_DIR = os.path.join(HERE, "build")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--use_fast_math=false"]
SOURCES = {
    "api.cu": [],
    "profile.cu": [],
    "preprocess_fwd.cu": ["-fmad=false"],
    "preprocess_bwd.cu": [],
    "binning.cu": [],
    "radix_sort.cu": [],
    "bucket_sort.cu": [],
    "render_fwd.cu": [],
    "render_bwd.cu": [],
    "postprocess.cu": [],
    "loss.cu": [],
    "optim.cu": [],
    "ply_pack.cu": [],
}


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    os.makedirs(LIB_DIR, exist_ok=True)
    os.makedirs(OBJ_DIR, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h", ".inc"))]
    headers.append(os.path.join(HERE, "..", "include", "surfel_rasterizer.h"))
    objs, procs = [], []
    for src, extra in SOURCES.items():
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _newer(o, [s] + headers):
            cmd = [_nvcc()] + ARCH + [f for f in COMMON if f != "--use_fast_math=false"] + extra + \
                  (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, pr in procs:
        out, _ = pr.communicate()
        if pr.returncode != 0 or verbose:
            sys.stderr.write(f"--- nvcc {src} ---\n{out}\n")
        failed |= pr.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    if force or procs or not os.path.exists(LIB):
        # link next to the target and rename: a reader (a process loading the library, a snapshot of the tree)
        # never sees a half-written file
        tmp = LIB + ".link"
        cmd = [_nvcc()] + ARCH + ["-shared", "-o", tmp] + objs + ["-lcudart"]
        subprocess.check_call(cmd)
        os.replace(tmp, LIB)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
