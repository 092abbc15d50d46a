"""Build recipe for libzenyth_b200.so (the C-ABI CUDA library) — sm_100a only, in-tree.

    python -m zenyth_b200.build [--verbose]

The .so lands in zenyth_b200/lib/ (git-ignored, but it travels with the repo snapshot to the GPU
box).  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB_DIR = os.path.join(PKG, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libzenyth_b200.so")
SOURCES = ["zn_ctx.cu", "zn_scene.cu", "zn_mesh.cu", "zn_synth.cu", "zn_hostmath.cpp", "zn_io.cpp", "zn_hierarchy.cpp", "zn_exchange.cpp"]

NVCC_FLAGS = [
    "-std=c++17", "-O3",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    # The kernels spell every float op as an explicit _rn intrinsic; -fmad=false is the belt to
    # that pair of braces (no contraction anywhere in device code).
    "-fmad=false",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math,-Wall",
    "-shared",
]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(PKG, "..", "include", "zenyth_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(verbose: bool = False, force: bool = False) -> str:
    if not force and not needs_build():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    cmd = [nvcc_path(), *NVCC_FLAGS]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += ["-o", LIB_PATH, *SOURCES]
    proc = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if verbose or proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed building libzenyth_b200.so")
    return LIB_PATH


HOST_SRC = os.path.join(PKG, "host", "sandbox_headless.cpp")
HOST_EXE = os.path.join(PKG, "host", "sandbox_headless")


def build_host(force: bool = False) -> str:
    """The headless C++ Sandbox over include/zenyth_b200.hpp (plain g++, no nvcc)."""
    build()
    inc = os.path.join(PKG, "..", "include")
    deps = [HOST_SRC, os.path.join(inc, "zenyth_b200.hpp"), os.path.join(inc, "zenyth_b200.h"), LIB_PATH]
    if not force and os.path.exists(HOST_EXE) and all(os.path.getmtime(d) <= os.path.getmtime(HOST_EXE) for d in deps):
        return HOST_EXE
    cxx = os.environ.get("CXX") or shutil.which("g++") or "g++"
    cmd = [cxx, "-std=c++17", "-O2", "-Wall", "-I", inc, HOST_SRC, "-L", LIB_DIR, "-lzenyth_b200",
           "-Wl,-rpath,$ORIGIN/../lib", "-o", HOST_EXE]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
        raise RuntimeError("g++ failed building sandbox_headless")
    return HOST_EXE


CPP_TEST_SRC = os.path.join(PKG, "..", "tests", "cpp", "exchange_ranks.cpp")
CPP_TEST_EXE = os.path.join(PKG, "host", "exchange_ranks")


def build_cpp_tests(force: bool = False) -> str:
    """tests/cpp/exchange_ranks.cpp: the C++-only multi-process exchange test (plain g++, no nvcc, no CUDA headers)."""
    build()
    inc = os.path.join(PKG, "..", "include")
    deps = [CPP_TEST_SRC, os.path.join(inc, "zenyth_b200.hpp"), os.path.join(inc, "zenyth_b200.h"), LIB_PATH]
    if not force and os.path.exists(CPP_TEST_EXE) and all(os.path.getmtime(d) <= os.path.getmtime(CPP_TEST_EXE) for d in deps):
        return CPP_TEST_EXE
    cxx = os.environ.get("CXX") or shutil.which("g++") or "g++"
    cmd = [cxx, "-std=c++17", "-O2", "-Wall", "-I", inc, CPP_TEST_SRC, "-L", LIB_DIR, "-lzenyth_b200",
           "-Wl,-rpath,$ORIGIN/../lib", "-o", CPP_TEST_EXE]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
        raise RuntimeError("g++ failed building exchange_ranks")
    return CPP_TEST_EXE


if __name__ == "__main__":
    print(build(verbose="--verbose" in sys.argv, force=True))
    print(build_host(force=True))
    print(build_cpp_tests(force=True))
