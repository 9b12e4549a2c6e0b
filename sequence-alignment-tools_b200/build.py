"""Builds the in-tree native artefacts of the package.

  lib/libpm_b200.so   the CUDA library behind include/pm_b200.h (sm_100a only)
  bin/primer_match    the drop-in CLI (C++ host code over the C ABI)

nvcc cross-compiles without a GPU, so this also is the "does it build" check.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "lib", "libpm_b200.so")
CLI = os.path.join(HERE, "bin", "primer_match")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-pthread"]


def _newer(target: str, sources) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def lib_sources():
    srcs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh", "pm_pcr.hpp", "pm_reader.hpp"))]
    srcs.append(os.path.join(ROOT, "include", "pm_b200.h"))
    return srcs


def build_lib(force: bool = False, verbose: bool = False) -> str:
    srcs = lib_sources()
    if not force and _newer(LIB, srcs):
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    cmd = [_nvcc(), *NVCC_FLAGS, "-shared", "-o", LIB, os.path.join(CSRC, "pm_b200.cu")]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    subprocess.check_call(cmd)
    return LIB


def build_cli(force: bool = False, tool: str = "primer_match") -> str | None:
    """bin/primer_match or bin/pcr_match: the drop-in command line tools (C++ over the C ABI)."""
    src = os.path.join(CSRC, tool + "_main.cc")
    if not os.path.exists(src):
        return None
    exe = os.path.join(os.path.dirname(CLI), tool)
    deps = [src, os.path.join(CSRC, "pm_cli_common.hpp"), os.path.join(ROOT, "include", "pm_b200.h"),
            os.path.join(ROOT, "include", "pm_pattern_match.hpp")]
    deps = [d for d in deps if os.path.exists(d)]
    if not force and _newer(exe, deps + [LIB]):
        return exe
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "include"), "-I", CSRC, "-o", exe, src,
                           "-L", os.path.dirname(LIB), "-lpm_b200", "-Wl,-rpath,$ORIGIN/../lib"])
    return exe


def build(force: bool = False, verbose: bool = False):
    build_lib(force, verbose)
    build_cli(force)
    build_cli(force, "pcr_match")


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)
