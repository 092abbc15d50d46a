"""The drop-in boundary, checked without a GPU: the shared library loads, exports exactly the
symbols include/zenyth_b200.h declares, has no torch / Python / oracle dependency, its host-side
camera maths equals the oracle bit for bit, and it fails loudly (never falls back) with no device.
"""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT, bits

HEADER = os.path.join(ROOT, "include", "zenyth_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(zn_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(zn_lib):
    from zenyth_b200 import _capi
    names = declared_symbols()
    assert len(names) >= 40
    for name in names:
        assert hasattr(zn_lib, name), f"{name} is declared in include/zenyth_b200.h but not exported"
    assert sorted(_capi.PROTOTYPES) == names, "zenyth_b200/_capi.py PROTOTYPES and the header disagree"
    assert zn_lib.zn_abi_version() == 2


def test_library_is_self_contained(zn_lib):
    from zenyth_b200 import _capi
    out = subprocess.run(["ldd", _capi.LIB_PATH], capture_output=True, text=True).stdout
    for forbidden in ("torch", "python", "zo_oracle", "c10"):
        assert forbidden not in out.lower(), f"libzenyth_b200.so links {forbidden}:\n{out}"
    syms = subprocess.run(["nm", "-D", "--defined-only", _capi.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\bT (zn_[a-z0-9_]+)\b", syms))
    assert exported == set(declared_symbols()), exported ^ set(declared_symbols())


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "zenyth_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in text.lower() and "zo_" not in text, f"{f} refers to the oracle"


def test_host_camera_maths_equals_the_oracle_bit_for_bit(zn_lib, oracle):
    import zenyth_b200 as zn
    rng = np.random.default_rng(21)
    for _ in range(200):
        eye, center = rng.uniform(-100, 100, 3), rng.uniform(-100, 100, 3)
        up = (0.0, 1.0, 0.0)
        assert np.array_equal(bits(zn.mat4.look_at(eye, center, up)), bits(oracle.look_at(eye, center, up)))
        fov, aspect, near, far = rng.uniform(0.3, 2.5), rng.uniform(0.5, 2.5), rng.uniform(0.01, 1.0), rng.uniform(10, 10000)
        p = zn.mat4.perspective(fov, aspect, near, far)
        assert np.array_equal(bits(p), bits(oracle.perspective(fov, aspect, near, far)))
        assert np.array_equal(bits(zn.mat4.perspective_reverse_z(fov, aspect, near)), bits(oracle.perspective_reverse_z(fov, aspect, near)))
        o = (-aspect, aspect, -1.0, 1.0, near, far)
        assert np.array_equal(bits(zn.mat4.orthographic(*o)), bits(oracle.orthographic(*o)))
        v = zn.mat4.look_at(eye, center, up)
        vp = zn.mat4.mul(p, v)
        assert np.array_equal(bits(vp), bits(oracle.mat_mul(p, v)))
        assert np.array_equal(bits(zn.mat4.frustum_planes(vp)), bits(oracle.frustum_planes(vp)))
    a, b = rng.normal(size=16).astype(np.float32), rng.normal(size=16).astype(np.float32)
    assert np.array_equal(bits(zn.mat4.mul(a, b)), bits(oracle.mat_mul(a, b)))


def has_gpu(zn_lib) -> bool:
    n = C.c_int(0)
    zn_lib.zn_device_count(C.byref(n))
    return n.value > 0


def test_no_device_is_an_error_not_a_fallback(zn_lib):
    if has_gpu(zn_lib):
        pytest.skip("a CUDA device is present")
    import zenyth_b200 as zn
    from zenyth_b200 import _capi
    h = C.c_void_p()
    assert zn_lib.zn_ctx_create(0, C.byref(h)) == _capi.ZN_ERR_NO_DEVICE
    assert h.value is None
    assert b"no CPU fallback" in zn_lib.zn_last_error()
    with pytest.raises(zn.ZenythError) as e:
        zn.GpuContext(0)
    assert e.value.code == _capi.ZN_ERR_NO_DEVICE


def test_null_arguments_are_rejected(zn_lib):
    from zenyth_b200 import _capi
    assert zn_lib.zn_ctx_create(0, None) == _capi.ZN_ERR_INVALID
    assert zn_lib.zn_scene_update(None) == _capi.ZN_ERR_INVALID
    assert zn_lib.zn_scene_create(None, None, None) == _capi.ZN_ERR_INVALID
    assert zn_lib.zn_mesh_transform(None) == _capi.ZN_ERR_INVALID
    assert zn_lib.zn_scene_destroy(None) == _capi.ZN_OK and zn_lib.zn_ctx_destroy(None) == _capi.ZN_OK
    assert b"NULL" in zn_lib.zn_last_error() or b"null" in zn_lib.zn_last_error().lower()


def test_missing_library_is_a_loud_import_error(monkeypatch, tmp_path):
    from zenyth_b200 import _capi
    monkeypatch.setattr(_capi, "_lib", None)
    monkeypatch.setattr(_capi, "LIB_PATH", str(tmp_path / "libzenyth_b200.so"))
    with pytest.raises(ImportError, match="no CPU or Python fallback"):
        _capi.load()


def test_cpp_host_wrapper_compiles_and_reports_no_gpu(zn_lib):
    """include/zenyth_b200.hpp (the engine-idiom C++ layer) builds with a plain g++ and its
    GpuContext throws std::runtime_error — the reference's error convention
    (Core/src/Application.cpp:28-30) — when there is no device."""
    from zenyth_b200 import build as zb
    exe = zb.build_host()
    if has_gpu(zn_lib):
        pytest.skip("a CUDA device is present")
    r = subprocess.run([exe, "--expect-no-gpu"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "no CPU fallback" in r.stdout
