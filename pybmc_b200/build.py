"""Build libbmc_b200.so (sm_100a only) in-tree with nvcc.

    python -m pybmc_b200.build [--force] [--verbose]

The shared object lands next to the sources (``pybmc_b200/csrc/libbmc_b200.so``) so it
travels with the repository snapshot to the GPU box; nothing is cached elsewhere.
"""
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
LIB = os.path.join(CSRC, "libbmc_b200.so")
PROBE_LIB = os.path.join(CSRC, "bench", "libbmc_probe.so")     # benchmark tooling, separate from the product ABI
UNITS = ["linalg.cu", "gibbs.cu", "simplex.cu", "predict.cu", "literal.cu"]
HEADERS = ["common.h", "rng.cuh", "fp64_tables.cuh", "gibbs_kernels.cuh", "linalg_kernels.cuh", "predict_kernels.cuh", "predict_tc_kernels.cuh",
           "literal_kernels.cuh", "tma.cuh", "select_logic.h", os.path.join("..", "..", "include", "bmc_b200.h")]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
         "--expt-relaxed-constexpr"]


def _nvcc():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: libbmc_b200.so cannot be built")
    return nvcc


def _units():
    return [u for u in UNITS if os.path.exists(os.path.join(CSRC, u))]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in _units() + HEADERS]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build_probe_library(force=False):
    """libbmc_probe.so: the pipe-peak probes bench.py divides by (pybmc_b200/csrc/bench/probe.h)."""
    src = os.path.join(CSRC, "bench", "probe.cu")
    if not force and os.path.exists(PROBE_LIB) and os.path.getmtime(PROBE_LIB) >= max(
            os.path.getmtime(src), os.path.getmtime(os.path.join(CSRC, "bench", "probe.h"))):
        return PROBE_LIB
    cmd = [_nvcc(), *ARCH, *FLAGS, "-shared", src, "-o", PROBE_LIB, "-cudart", "static"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed on probe.cu:\n{r.stdout}\n{r.stderr}")
    return PROBE_LIB


def build_library(force=False, verbose=False):
    """Compile every translation unit for sm_100a and link the shared library."""
    if not force and not _stale():
        return LIB
    nvcc = _nvcc()
    objs = []

    def compile_one(unit):
        obj = os.path.join(CSRC, unit.replace(".cu", ".o"))
        cmd = [nvcc, *ARCH, *FLAGS, "-c", os.path.join(CSRC, unit), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {unit}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:
        objs = list(pool.map(compile_one, _units()))
    cmd = [nvcc, *ARCH, "-shared", "-o", LIB, *objs, "-cudart", "static"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    for obj in objs:            # only the shared object needs to travel with the snapshot
        try:
            os.remove(obj)
        except OSError:
            pass
    return LIB


if __name__ == "__main__":
    path = build_library(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
    print(build_probe_library(force="--force" in sys.argv))
