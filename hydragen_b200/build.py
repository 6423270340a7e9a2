"""Build recipe: nvcc -> hydragen_b200/_C/libhydragen_b200.so (sm_100a only, in-tree).

The library has a plain C ABI (include/hydragen_b200.h) and links only the CUDA runtime
(statically); the driver API entry point needed for TMA descriptors is fetched at run time
through cudaGetDriverEntryPoint, so there is no link-time dependency on libcuda and the
library cross-compiles and dlopens on a box without a GPU.
"""

from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
OUT_DIR = os.path.join(PKG_DIR, "_C")
LIB_PATH = os.path.join(OUT_DIR, "libhydragen_b200.so")
SOURCES = ["api.cu", "combine.cu", "rowwise.cu", "prefix_sm100.cu", "prefix_unit_sm100.cu", "prefix_unit_sm100_causal.cu", "allreduce.cu", "oproj_allreduce.cu", "rope.cu"]
HEADERS = [os.path.join(CSRC, "common.cuh"), os.path.join(CSRC, "prefix_sched.h"), os.path.join(CSRC, "oproj_sched.h"), os.path.join(CSRC, "sm100_ptx.cuh"), os.path.join(ROOT, "include", "hydragen_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


# development only: e.g. HG_EXTRA_NVCC_FLAGS="-DHG_PREFIX_TRACE" builds a separate, instrumented library
EXTRA_FLAGS = os.environ.get("HG_EXTRA_NVCC_FLAGS", "").split()
if EXTRA_FLAGS:
    OUT_DIR = os.path.join(PKG_DIR, "_C_dev_" + "".join(ch if ch.isalnum() else "_" for ch in "".join(EXTRA_FLAGS)).strip("_"))
    LIB_PATH = os.path.join(OUT_DIR, "libhydragen_b200.so")


def _nvcc() -> str:
    cand = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise RuntimeError("nvcc not found (set NVCC=...)")
    return cand


def _stamp() -> str:
    h = hashlib.sha256()
    for path in [os.path.join(CSRC, s) for s in SOURCES] + HEADERS + [os.path.abspath(__file__)]:
        with open(path, "rb") as f:
            h.update(f.read())
    h.update(" ".join(EXTRA_FLAGS).encode())
    return h.hexdigest()


def needs_build() -> bool:
    stamp_file = os.path.join(OUT_DIR, "stamp")
    if not os.path.exists(LIB_PATH) or not os.path.exists(stamp_file):
        return True
    with open(stamp_file) as f:
        return f.read().strip() != _stamp()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a and link the shared library.  Returns its path."""
    if not force and not needs_build():
        return LIB_PATH
    os.makedirs(OUT_DIR, exist_ok=True)
    nvcc = _nvcc()
    objs = []

    def compile_one(src: str) -> str:
        obj = os.path.join(OUT_DIR, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, *EXTRA_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log = os.path.join(OUT_DIR, src + ".log")
        with open(log, "w") as f:
            f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [nvcc, "-shared", "-o", LIB_PATH, *objs, "-cudart", "static", "-Xcompiler", "-fPIC"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(os.path.join(OUT_DIR, "stamp"), "w") as f:
        f.write(_stamp())
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
