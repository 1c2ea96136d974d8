"""Build libkanconv.so (hand-written sm_100a CUDA behind the C ABI of include/kanconv.h) in-tree with nvcc.

    python -m kanconv_b200.build        # or:  python __graft_entry__.py build

The shared library is built next to its sources (csrc/libkanconv.so) so that it travels with the repo snapshot to the
GPU box; it is git-ignored.  nvcc cross-compiles for sm_100a without a GPU.
"""
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
ROOT = os.path.dirname(HERE)
INCLUDE = os.path.join(ROOT, "include")
LIB = os.path.join(CSRC, "libkanconv.so")
SOURCES = ["kc_api.cu", "kc_simt.cu", "kc_dw.cu", "kc_norm.cu", "kc_norm_cluster.cu", "kc_pool.cu", "kc_tc.cu", "kc_tc_wgrad.cu"]
HEADERS = ["kc_common.cuh", "kc_umma.cuh", "kc_tc_basis.cuh", "kc_norm_common.cuh"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "--expt-relaxed-constexpr",
              "-Xcompiler", "-fPIC", "-I" + INCLUDE, "-diag-suppress", "177"]
# KANCONV_DEBUG=1: also build the micro-benchmarks (kc_debug.cu) and the in-kernel timeline trace (kc_debug_* exports used by
# tools/mma_rate.py, bulk_bench.py, trace_*.py).  The product library carries neither.
DEBUG = os.environ.get("KANCONV_DEBUG") == "1"
if DEBUG:
    SOURCES = SOURCES + ["kc_debug.cu"]
    NVCC_FLAGS = NVCC_FLAGS + ["-DKANCONV_DEBUG"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build libkanconv.so")


def have_nvcc():
    try:
        _nvcc()
        return True
    except RuntimeError:
        return False


HASHFILE = LIB + ".srchash"


def source_hash():
    """sha256 over the CUDA sources, headers and compiler flags the library is built from (mtimes do not survive the copy
    to a GPU box, contents do)."""
    import hashlib
    h = hashlib.sha256(" ".join(NVCC_FLAGS[:NVCC_FLAGS.index("-Xcompiler")]).encode())
    for p in sorted([os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.join(INCLUDE, "kanconv.h")]):
        h.update(os.path.basename(p).encode())
        with open(p, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def _stale():
    """True when the library is missing or was built from different sources than the ones in the tree."""
    if not os.path.exists(LIB) or not os.path.exists(HASHFILE):
        return True
    with open(HASHFILE) as fh:
        return fh.read().strip() != source_hash()


def build(force=False, verbose=True):
    """Compile every CUDA source for sm_100a and link csrc/libkanconv.so.  Returns the library path.  Concurrent callers
    (one process per GPU under torchrun) are serialised with a file lock; the losers find an up-to-date library."""
    if not force and not _stale():
        return LIB
    import fcntl
    with open(os.path.join(CSRC, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        if not force and not _stale():
            return LIB
        return _build_locked(verbose)


def _build_locked(verbose):
    nvcc = _nvcc()
    objdir = os.path.join(CSRC, "build")
    os.makedirs(objdir, exist_ok=True)

    def compile_one(src):
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [nvcc] + NVCC_FLAGS + ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    tmp = LIB + ".tmp"
    cmd = [nvcc, "-shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a", "-o", tmp] + objs
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    os.replace(tmp, LIB)
    with open(HASHFILE, "w") as fh:
        fh.write(source_hash() + "\n")
    if verbose:
        print("built", LIB, file=sys.stderr)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
