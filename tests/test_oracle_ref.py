"""The restated oracle types against the reference's OWN code.

oracle/_ref/libzo_oracle_ref.so is the same oracle source compiled with -DZO_REFERENCE_TYPES, i.e.
on zenyth::math::vec3 / vec4 / mat4 from /root/reference/Core/{include,src}/math, built in place
behind oracle/shim/pch.hpp (oracle/Makefile `ref`).  Bit-identical results from both builds pin
the scalar restatement in oracle/zo_vec.hpp to the reference's real SSE arithmetic
(Core/src/math/vector.cpp:41-128) — the only arithmetic the reference implements.
Skipped when neither /root/reference nor a prebuilt oracle/_ref is present.
"""
import numpy as np
import pytest

from conftest import bits

F = np.float32


def rnd(rng, n):
    # wide dynamic range, including negatives, zeros and denormal-ish values
    v = rng.normal(size=n) * 10.0 ** rng.integers(-6, 6, n)
    v[rng.random(n) < 0.05] = 0.0
    return v.astype(F)


def test_backends_identify_themselves(oracle, oracle_ref):
    assert oracle.lib.zo_backend() == 0 and oracle_ref.lib.zo_backend() == 1


def test_vec3_vec4_ops_bit_identical(oracle, oracle_ref):
    rng = np.random.default_rng(11)
    for _ in range(300):
        a3, b3 = rnd(rng, 3), rnd(rng, 3)
        a4, b4 = rnd(rng, 4), rnd(rng, 4)
        s = rnd(rng, 1)
        if s[0] == 0:
            s[0] = F(3.0)
        for op in (0, 1, 2, 5):     # + - cross neg
            assert np.array_equal(bits(oracle.vec3_op(op, a3, b3)), bits(oracle_ref.vec3_op(op, a3, b3))), op
        for op in (3, 4):           # * / by scalar
            assert np.array_equal(bits(oracle.vec3_op(op, a3, s)), bits(oracle_ref.vec3_op(op, a3, s))), op
            assert np.array_equal(bits(oracle.vec4_op(op, a4, s)), bits(oracle_ref.vec4_op(op, a4, s))), op
        for op in (0, 1, 5):
            assert np.array_equal(bits(oracle.vec4_op(op, a4, b4)), bits(oracle_ref.vec4_op(op, a4, b4))), op
        assert bits(oracle.vec3_dot(a3, b3)) == bits(oracle_ref.vec3_dot(a3, b3))
        assert bits(oracle.vec4_dot(a4, b4)) == bits(oracle_ref.vec4_dot(a4, b4))
        assert bits(oracle.vec3_length(a3)) == bits(oracle_ref.vec3_length(a3))
        assert bits(oracle.vec4_length(a4)) == bits(oracle_ref.vec4_length(a4))


def test_reference_dot_really_sums_pairs(oracle_ref):
    # DPPS order (p0+p1)+(p2+p3), observed on the reference's own vec4::dot (vector.cpp:125)
    a = F([1.0e8, 1.0, -1.0e8, 1.0])
    assert oracle_ref.vec4_dot(a, F([1, 1, 1, 1])) == F(0.0)


def test_completed_mat4_bit_identical(oracle, oracle_ref):
    rng = np.random.default_rng(12)
    for _ in range(100):
        t, r, s = rng.uniform(-100, 100, 3), rng.uniform(-3.14, 3.14, 3), rng.uniform(0.5, 2, 3)
        a = oracle.local_matrix(t, r, s)
        assert np.array_equal(bits(a), bits(oracle_ref.local_matrix(t, r, s)))
        b = rnd(rng, 16)
        v = rnd(rng, 4)
        for fn, args in [("from_euler", tuple(r)), ("from_axis_angle", (t, float(r[0]))), ("from_scale", (s,)), ("from_translation", (t,)),
                         ("from_basis", (t, r, s)), ("mat_mul", (a, b)), ("mat_add", (a, b)), ("mat_sub", (a, b)), ("mat_scale", (a, 1.7)),
                         ("mat_div", (a, 1.7)), ("mat_mul_vec4", (a, v)), ("transpose", (b,)), ("inverse", (a,)), ("inverse", (b,)),
                         ("inverse_transpose", (a,)), ("transform_point", (a, t)), ("transform_normal", (a, t)),
                         ("look_at", (t, r, (0, 1, 0))), ("perspective", (1.0, 1.6, 0.1, 900.0)), ("orthographic", (-1, 2, -3, 4, 0.1, 50)),
                         ("perspective_reverse_z", (1.0, 1.6, 0.1))]:
            x, y = getattr(oracle, fn)(*args), getattr(oracle_ref, fn)(*args)
            assert np.array_equal(bits(x), bits(y)), fn
        assert bits(oracle.determinant(b)) == bits(oracle_ref.determinant(b))


def test_scene_and_mesh_bit_identical(oracle, oracle_ref):
    from oracle.binding import BASE_SEED, camera, h8_geo_level_sizes
    sc = oracle.synth_scene(BASE_SEED + 3, h8_geo_level_sizes(7, 4), fanout=8, center_range=0.25)
    sc_ref = oracle_ref.synth_scene(BASE_SEED + 3, h8_geo_level_sizes(7, 4), fanout=8, center_range=0.25)
    for k in ("position", "rotation", "scale", "aabb_center", "aabb_half", "parent"):
        assert np.array_equal(sc[k].view(np.uint32), sc_ref[k].view(np.uint32)), k
    for eye in [(0, 0, 0), (0, 0, 1500)]:
        view, proj = camera(oracle, eye=eye)
        vr, pr = camera(oracle_ref, eye=eye)
        assert np.array_equal(bits(view), bits(vr)) and np.array_equal(bits(proj), bits(pr))
        a = oracle.scene_update(sc, view, proj, threads=3)
        b = oracle_ref.scene_update(sc, view, proj, threads=1)
        assert np.array_equal(bits(a["world"]), bits(b["world"]))
        assert np.array_equal(bits(a["mvp"]), bits(b["mvp"]))
        assert np.array_equal(a["visible"], b["visible"])
    pos, nrm = oracle.synth_mesh(BASE_SEED + 4, 6 * 50)
    first = (np.arange(7) * 50).astype(np.uint32)
    pa, na = oracle.mesh_transform(first, a["world"][:6], pos, nrm)
    pb, nb = oracle_ref.mesh_transform(first, a["world"][:6], pos, nrm)
    assert np.array_equal(bits(pa), bits(pb)) and np.array_equal(bits(na), bits(nb))
