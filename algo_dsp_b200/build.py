"""Builds libalgodsp_cuda.so (hand-written sm_100a kernels + C ABI) in-tree with nvcc.

    python -m algo_dsp_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU; the .so is git-ignored but travels with the repo snapshot.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "libalgodsp_cuda.so")
SOURCES = ["api.cu", "fftconv_f64.cu", "fftconv_f32.cu", "fdl.cu", "staging.cu", "siggen.cu", "diag.cu", "post.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-ffp-contract=off", "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    return "nvcc"


def _deps():
    return [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "algodsp_cuda.h")]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(src: str, verbose: bool, defines=(), tag: str = "") -> str:
    obj = os.path.join(BUILD, src.replace(".cu", tag + ".o"))
    cmd = [_nvcc(), *NVCC_FLAGS, *[f"-D{d}" for d in defines], "-c", os.path.join(CSRC, src), "-o", obj]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    if verbose:
        with open(obj + ".ptxas.log", "w") as f:
            f.write(r.stderr)
    return obj


def build(force: bool = False, verbose: bool = False, defines=(), out: str | None = None) -> str:
    """defines/out build an experimental variant (e.g. defines=["ADSP_COLS_CTA_THREADS=128"],
    out="libvariant.so") that tools/ can select with ADSP_LIB_PATH; the default build takes neither."""
    os.makedirs(BUILD, exist_ok=True)
    deps = _deps()
    lib = os.path.join(HERE, out) if out else LIB
    if not force and not _stale(lib, deps):
        return lib
    tag = ("_" + os.path.splitext(out)[0]) if out else ""
    with cf.ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(lambda s: _compile(s, verbose, defines, tag), SOURCES))
    cmd = [_nvcc(), "-shared", "-o", lib, *objs, "-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return lib


if __name__ == "__main__":
    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    outs = [a.split("=", 1)[1] for a in sys.argv[1:] if a.startswith("--out=")]
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, defines=defs, out=outs[0] if outs else None))
