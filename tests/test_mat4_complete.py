"""contrib/mat4_complete.cpp gives the reference's stubbed zenyth::math::mat4
(Core/src/math/matrix.cpp:5-121) real bodies, written against the reference's own headers.  Built
with the reference's vector.cpp (oracle/Makefile `mat4`) it must agree bit for bit with the oracle
— and therefore with the GPU path.  Skipped where neither /root/reference nor a prebuilt
oracle/_ref/libmat4_complete.so exists."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import ROOT, bits

LIB = os.path.join(ROOT, "oracle", "_ref", "libmat4_complete.so")
F = np.float32


@pytest.fixture(scope="module")
def mc(oracle):
    if not os.path.exists(LIB):
        pytest.skip("oracle/_ref/libmat4_complete.so not built (no /root/reference here)")
    lib = C.CDLL(LIB)
    lib.mc_determinant.restype = C.c_float
    return lib


def p(a):
    a = np.ascontiguousarray(a, np.float32)
    return a.ctypes.data_as(C.c_void_p), a


def call(lib, name, *args, n=16):
    out = np.zeros(n, F)
    keep = []
    cargs = []
    for a in args:
        if isinstance(a, (float, int, np.floating)):
            cargs.append(C.c_float(float(a)))
        else:
            ptr, arr = p(a)
            keep.append(arr)
            cargs.append(ptr)
    getattr(lib, name)(*cargs, out.ctypes.data_as(C.c_void_p))
    return out


def test_default_and_diagonal(mc, oracle):
    assert np.array_equal(call(mc, "mc_identity"), oracle.identity())
    assert np.array_equal(call(mc, "mc_diagonal", 2.5), oracle.diagonal(2.5))     # the stub ignores its argument (matrix.cpp:6)


def test_every_method_matches_the_oracle_bit_for_bit(mc, oracle):
    rng = np.random.default_rng(31)
    for _ in range(200):
        t, r, s = rng.uniform(-100, 100, 3).astype(F), rng.uniform(-3.14, 3.14, 3).astype(F), rng.uniform(0.5, 2, 3).astype(F)
        a = oracle.local_matrix(t, r, s)
        b = (rng.normal(size=16) * 3).astype(F)
        v = rng.normal(size=4).astype(F)
        pairs = [
            (call(mc, "mc_from_euler", r[0], r[1], r[2]), oracle.from_euler(r[0], r[1], r[2])),
            (call(mc, "mc_from_axis_angle", t, r[0]), oracle.from_axis_angle(t, r[0])),
            (call(mc, "mc_from_scale", s), oracle.from_scale(s)),
            (call(mc, "mc_from_translation", t), oracle.from_translation(t)),
            (call(mc, "mc_from_basis", t, r, s), oracle.from_basis(t, r, s)),
            (call(mc, "mc_add", a, b), oracle.mat_add(a, b)),
            (call(mc, "mc_sub", a, b), oracle.mat_sub(a, b)),
            (call(mc, "mc_scale", a, 1.7), oracle.mat_scale(a, 1.7)),
            (call(mc, "mc_div", a, 1.7), oracle.mat_div(a, 1.7)),
            (call(mc, "mc_mul", a, b), oracle.mat_mul(a, b)),
            (call(mc, "mc_mul_assign", a, b), oracle.mat_mul(a, b)),
            (call(mc, "mc_mul_vec4", a, v, n=4), oracle.mat_mul_vec4(a, v)),
            (call(mc, "mc_transpose", b), oracle.transpose(b)),
            (call(mc, "mc_inverse", a), oracle.inverse(a)),
            (call(mc, "mc_inverse", b), oracle.inverse(b)),
            (call(mc, "mc_inverse_transpose", a), oracle.inverse_transpose(a)),
            (call(mc, "mc_transform_point", a, t, n=3), oracle.transform_point(a, t)),
            (call(mc, "mc_transform_normal", a, t, n=3), oracle.transform_normal(a, t)),
            (call(mc, "mc_look_at", t, r, F([0, 1, 0])), oracle.look_at(t, r, (0, 1, 0))),
            (call(mc, "mc_perspective", 1.0, 1.6, 0.1, 900.0), oracle.perspective(1.0, 1.6, 0.1, 900.0)),
            (call(mc, "mc_orthographic", -1, 2, -3, 4, 0.1, 50), oracle.orthographic(-1, 2, -3, 4, 0.1, 50)),
            (call(mc, "mc_perspective_reverse_z", 1.0, 1.6, 0.1), oracle.perspective_reverse_z(1.0, 1.6, 0.1)),
        ]
        for k, (got, want) in enumerate(pairs):
            assert np.array_equal(bits(got), bits(want)), k
        ptr, arr = p(b)
        assert bits(F(mc.mc_determinant(ptr))) == bits(oracle.determinant(b))


def test_trs_chain_through_the_completed_class_equals_the_scene_oracle(mc, oracle):
    """world = parent * (T*R*S) built only from completed mat4 methods == the oracle's scene path."""
    from oracle.binding import BASE_SEED, camera, h8_geo_level_sizes
    sc = oracle.synth_scene(BASE_SEED + 3, h8_geo_level_sizes(3, 3, 4), fanout=4, center_range=0.2)
    view, proj = camera(oracle)
    ref = oracle.scene_update(sc, view, proj)
    vp = call(mc, "mc_mul", proj, view)
    world = np.zeros_like(ref["world"])
    for i in range(sc["position"].shape[0]):
        trs = call(mc, "mc_mul", call(mc, "mc_mul", call(mc, "mc_from_translation", sc["position"][i]),
                                      call(mc, "mc_from_euler", *sc["rotation"][i])), call(mc, "mc_from_scale", sc["scale"][i]))
        par = sc["parent"][i]
        world[i] = trs if par == 0xFFFFFFFF else call(mc, "mc_mul", world[par], trs)
        assert np.all(world[i] == ref["world"][i]), i
        assert np.all(call(mc, "mc_mul", vp, world[i]) == ref["mvp"][i]), i
