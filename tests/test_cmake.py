"""The CMake project (CMakeLists.txt) a reference-side build consumes must stay in step with the
Python build recipe: same sources, same architecture, same floating-point flags — and configure."""
import os
import shutil
import subprocess

import pytest

from conftest import ROOT


def test_cmake_lists_the_sources_and_flags_of_the_python_recipe():
    from zenyth_b200 import build as zb
    text = open(os.path.join(ROOT, "CMakeLists.txt")).read()
    for src in zb.SOURCES:
        assert f"zn_csrc_placeholder/{src}".replace("zn_csrc_placeholder", "${ZN_CSRC}") in text, src
    assert "CMAKE_CUDA_ARCHITECTURES 100a" in text
    for flag in ("-fmad=false", "-ffp-contract=off", "-lineinfo"):
        assert flag in text and any(flag in f for f in zb.NVCC_FLAGS), flag
    assert "zenyth_b200::zenyth_b200" in text          # the target INTEGRATION.md §1 links


def test_cmake_configures(tmp_path):
    cmake = shutil.which("cmake")
    if cmake is None or shutil.which("nvcc") is None:
        pytest.skip("cmake / nvcc not installed")
    gen = ["-G", "Ninja"] if shutil.which("ninja") else []
    r = subprocess.run([cmake, "-S", ROOT, "-B", str(tmp_path), *gen], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert (tmp_path / "CMakeCache.txt").exists()
