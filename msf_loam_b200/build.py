"""Builds libmsfl.so (hand-written sm_100a CUDA + the C ABI of include/msfl.h) in-tree with nvcc.

Every translation unit of csrc/ is compiled to its own object (in parallel, re-used while neither the source nor any
header changed) and the objects are linked into msf_loam_b200/libmsfl.so.  No CPU fallback exists: a failed build raises.
"""
from __future__ import annotations

import concurrent.futures as cf
import glob
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libmsfl.so")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-O3,-Wall", "--expt-relaxed-constexpr",
]
# per-file extra flags (none yet; a place for e.g. --fmad=false on a bit-exact fp32 unit)
FILE_FLAGS: dict[str, list[str]] = {}


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def headers():
    return sorted(glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h")) + [
        os.path.join(HERE, "..", "include", "msfl.h")])


def stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in sources() + headers())


def _extra():
    return os.environ.get("MSFL_NVCC_EXTRA", "").split()  # development: e.g. -DMSFL_LM_TIMING


def _obj_path(src, extra):
    tag = hashlib.sha1(" ".join(NVCC_FLAGS + extra + FILE_FLAGS.get(os.path.basename(src), [])).encode()).hexdigest()[:8]
    return os.path.join(OBJ, os.path.basename(src)[:-3] + "." + tag + ".o")


def _compile(nvcc, src, obj, extra, verbose):
    tmp = f"{obj}.tmp.{os.getpid()}"
    cmd = [nvcc] + NVCC_FLAGS + extra + FILE_FLAGS.get(os.path.basename(src), []) + (
        ["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", tmp, src]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        if os.path.exists(tmp):
            os.unlink(tmp)
        return src, r.returncode, r.stdout + r.stderr
    os.replace(tmp, obj)
    return src, 0, r.stderr


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not stale():
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    extra = _extra()
    os.makedirs(OBJ, exist_ok=True)
    hdr_t = max(os.path.getmtime(h) for h in headers())
    jobs, objs = [], []
    for src in sources():
        obj = _obj_path(src, extra)
        objs.append(obj)
        if force or verbose or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr_t):
            jobs.append((src, obj))
    log = []
    with cf.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        for src, rc, out in ex.map(lambda j: _compile(nvcc, j[0], j[1], extra, verbose), jobs):
            if rc != 0:
                sys.stderr.write(out)
                raise RuntimeError(f"nvcc failed on {os.path.basename(src)}")
            log.append(out)
    # link into a private file and rename: several ranks of one torchrun may find the library stale at once, and a
    # reader must never see a half-written .so
    tmp = f"{LIB}.tmp.{os.getpid()}"
    r = subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", tmp] + objs + ["-ldl"],
                       capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        if os.path.exists(tmp):
            os.unlink(tmp)
        raise RuntimeError("nvcc failed linking libmsfl.so")
    os.replace(tmp, os.environ.get("MSFL_LIB_OUT", LIB))  # development: build a variant next to the tree's library
    if verbose:
        sys.stderr.write("".join(log))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
