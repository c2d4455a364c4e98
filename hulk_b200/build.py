"""In-tree build of the CUDA library (sm_100a only).  `python -m hulk_b200.build [--force]`."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libhulk_b200.so")
CLI = os.path.join(HERE, "bin", "hulk")

CUDA_SOURCES = ["api.cu", "smash.cu"]
HOST_SOURCES = ["host_io.cpp", "ingest.cpp", "sketch_json.cpp", "group.cpp", "pack.cpp"]
DEPS = CUDA_SOURCES + HOST_SOURCES + ["hd_math.h", "pgzip.h", "ptx_util.cuh", "k1_minimizer.cuh", "k1_scan.h", "k1_scan2.h", "k1_long.h", "k1_long.cuh", "k1_minhash.cuh", "k2_countmin.cuh",
                                      "k3_cws.cuh", "go_rng_cooked.inc", "../../include/hulk_b200.h"]
CLI_SOURCES = ["cli/hulk_main.cpp"]

NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",   # B200 only; no other targets, no PTX-JIT fallback
    "-lineinfo",
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-Wall",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA extension cannot be built (there is no CPU fallback)")


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build_lib(force: bool = False, verbose: bool = False) -> str:
    deps = [os.path.join(CSRC, d) for d in DEPS]
    if force or _stale(LIB, deps):
        cmd = [_nvcc()] + NVCC_FLAGS + ["-shared", "-o", LIB] + [os.path.join(CSRC, s) for s in CUDA_SOURCES + HOST_SOURCES] + \
              ["-lz", "-lpthread"]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), file=sys.stderr)
        subprocess.run(cmd, check=True, cwd=CSRC)
    return LIB


def build_cli(force: bool = False, verbose: bool = False) -> str:
    srcs = [os.path.join(CSRC, s) for s in CLI_SOURCES]
    if not all(os.path.exists(s) for s in srcs):
        return ""
    if force or _stale(CLI, srcs + [LIB]):
        os.makedirs(os.path.dirname(CLI), exist_ok=True)
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else (shutil.which("g++") or "g++")
        cmd = [cxx, "-O2", "-std=c++17", "-Wall", "-pthread", "-o", CLI] + srcs + \
              ["-I", os.path.join(HERE, "..", "include"), "-L", HERE, "-lhulk_b200", "-lz",
               "-Wl,-rpath,$ORIGIN/.."]
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        subprocess.run(cmd, check=True)
    return CLI


def build_all(force: bool = False, verbose: bool = False):
    lib = build_lib(force, verbose)
    cli = build_cli(force, verbose)
    return lib, cli


if __name__ == "__main__":
    print(build_all(force="--force" in sys.argv, verbose="-v" in sys.argv))
