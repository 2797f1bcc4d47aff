"""
Builds ``libnemo_b200.so`` (hand-written sm_100a CUDA behind a C ABI, include/nemo_b200.h) in-tree with nvcc.
There is no GPU requirement for building: nvcc cross-compiles.

    python -m nemo_b200._build [--force]
"""

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJDIR = os.path.join(CSRC, "build")
LIBPATH = os.path.join(HERE, "libnemo_b200.so")
SOURCES = ["api.cu", "fft_kernels.cu", "fft_bluestein.cu", "fft_bluestein_cols.cu", "sfft_rows.cu", "sfft_cols_f64.cu", "sfft_cols_f32.cu", "sfft_cols_tma.cu", "filter_build.cu", "noise.cu", "noise_onebin.cu", "noise_smart.cu", "detect.cu", "realspace.cu", "xmatch.cu", "paint.cu", "shapes.cu"]
# IEEE arithmetic everywhere (no fast-math); the CUDA runtime is linked statically so the library does not
# depend on which libcudart the host process (PyTorch) happens to load
NVCC_FLAGS = ["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC"]


if os.environ.get("NEMO_B200_NVCC_EXTRA"):
    NVCC_FLAGS += os.environ["NEMO_B200_NVCC_EXTRA"].split()      # experiment knobs, e.g. -DOC_THREADS=1024


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _deps_mtime():
    newest = 0.0
    for root in (CSRC, os.path.join(HERE, "..", "include")):
        for name in os.listdir(root):
            if name.endswith((".cuh", ".h")):
                newest = max(newest, os.path.getmtime(os.path.join(root, name)))
    return newest


def _compile_one(src, force):
    obj = os.path.join(OBJDIR, src.replace(".cu", ".o"))
    srcPath = os.path.join(CSRC, src)
    if not force and os.path.exists(obj) and os.path.getmtime(obj) >= max(os.path.getmtime(srcPath), _deps_mtime()):
        return obj, None
    cmd = [_nvcc()] + NVCC_FLAGS + ["-c", srcPath, "-o", obj]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        return obj, "nvcc failed for %s:\n%s\n%s" % (src, res.stdout, res.stderr)
    return obj, None


def build(force=False, verbose=True):
    os.makedirs(OBJDIR, exist_ok=True)
    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 4)) as ex:
        results = list(ex.map(lambda s: _compile_one(s, force), SOURCES))
    errors = [e for _, e in results if e]
    if errors:
        raise RuntimeError("\n".join(errors))
    objs = [o for o, _ in results]
    if force or not os.path.exists(LIBPATH) or any(os.path.getmtime(o) > os.path.getmtime(LIBPATH) for o in objs):
        cmd = [_nvcc(), "-shared", "-o", LIBPATH] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (res.stdout, res.stderr))
        if verbose:
            print("built", LIBPATH)
    elif verbose:
        print("up to date:", LIBPATH)
    return LIBPATH


if __name__ == "__main__":
    build(force="--force" in sys.argv)
