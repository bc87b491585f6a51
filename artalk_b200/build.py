"""In-tree build of libartalk_b200.so for sm_100a (nvcc cross-compiles without a GPU).

``python -m artalk_b200.build`` or ``__graft_entry__.build()``. Objects are cached by source mtime under
artalk_b200/csrc/_build/ ; the shared library is written to artalk_b200/lib/libartalk_b200.so (git-ignored,
travels to the GPU box with the gpurun snapshot).
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_build")
LIB_DIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIB_DIR, "libartalk_b200.so")
SOURCES = ["norms.cu", "gemm_simt.cu", "gemm_tc.cu", "skinny.cu", "split.cu", "attention.cu", "attention_tc.cu", "posconv_tc.cu", "conv0_fold.cu", "bits.cu", "flame.cu", "flame_tc.cu", "postproc.cu", "mesh.cu", "frontend.cu",
           "engine.cu", "api.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
         "-Xcompiler", "-Wall", "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _headers_mtime() -> float:
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs.append(os.path.join(HERE, "..", "include", "artalk_b200.h"))
    return max(os.path.getmtime(h) for h in hs)


def _compile(src: str, verbose: bool) -> str:
    obj = os.path.join(OBJ, src.replace(".cu", ".o"))
    sp = os.path.join(CSRC, src)
    if os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(sp), _headers_mtime()):
        return obj
    cmd = [NVCC] + FLAGS + ["-c", sp, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log = os.path.join(OBJ, src + ".log")
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s" % (src, r.stderr[-6000:]))
    if verbose:
        sys.stderr.write("compiled %s\n" % src)
    return obj


def build(verbose: bool = False, force: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(LIB_DIR, exist_ok=True)
    if force:
        for f in os.listdir(OBJ):
            if f.endswith(".o"):
                os.remove(os.path.join(OBJ, f))
    with cf.ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        objs = list(ex.map(lambda s: _compile(s, verbose), SOURCES))
    if (not os.path.exists(LIB)) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs):
        # shared cudart: torch already loads libcudart.so.12 into the process, and the shipped library then carries no copy of
        # the runtime (the static runtime embeds entry-point names this repo never calls). The driver entry point for TMA
        # descriptors is fetched at run time with cudaGetDriverEntryPoint, so the library also loads on a box without libcuda
        cudart = "static" if os.environ.get("ARTALK_STATIC_CUDART") == "1" else "shared"      # developer switch
        cmd = [NVCC, "-shared", "--cudart", cudart, "-o", LIB] + objs
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stderr[-4000:])
    return LIB


if __name__ == "__main__":
    print(build(verbose=True, force="--force" in sys.argv))
