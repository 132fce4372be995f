"""In-tree build of the CUDA library (sm_100a) with plain nvcc — no JIT cache, so the .so
travels with the repo snapshot.  `python -m epipolar_transformers_b200.build [--force]`."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libepipolar_b200.so")
SOURCES = ["epi_abi.cu", "epi_aux.cu", "epi_fusion_warp.cu", "epi_fusion_tile.cu", "epi_fusion_pipe.cu", "epi_fusion_bwd.cu", "epi_stage.cu", "epi_peaks.cu", "epi_zgemm.cu", "epi_umma_selftest.cu"]
HEADERS = ["epi_common.cuh", "epi_kernels.cuh", "epi_umma.cuh", os.path.join("..", "..", "include", "epipolar_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build libepipolar_b200.so")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False, timers: bool = False) -> str:
    """timers=True builds libepipolar_b200_timers.so with the in-kernel clock64 phase timers (developer tool)."""
    if not timers and not force and not needs_build():
        return LIB
    nvcc = _nvcc()
    objs = []
    procs = []
    lib = LIB.replace(".so", "_timers.so") if timers else LIB
    for s in SOURCES:
        obj = os.path.join(CSRC, s.replace(".cu", ".timers.o" if timers else ".o"))
        cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, s), "-o", obj]
        if timers:
            cmd[1:1] = ["-DEPI_PIPE_TIMERS", "-DEPI_TILE_TIMERS"]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for s, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out)
        if p.returncode != 0:
            raise RuntimeError("nvcc failed on %s" % s)
    subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", lib, *objs, "-lcudart"])
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, timers="--timers" in sys.argv))
