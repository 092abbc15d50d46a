"""Pins for the CPU oracle.

The reference ships no tests, fixtures or golden vectors for this path and its mat4 bodies are
stubs (PARITY UNPINNED by the reference), so the oracle is pinned here by
(1) closed-form known answers, (2) algebraic properties, (3) an independent float64 shadow
implementation (oracle/zo_f64.cpp), and in test_oracle_ref.py by the reference's own vec3/vec4 code.
"""
import math

import numpy as np
import pytest

from conftest import rel_err

F = np.float32
I4 = np.eye(4, dtype=np.float32)


def M(a):
    """16 column-major floats -> 4x4 row/col matrix."""
    return np.asarray(a, np.float64).reshape(4, 4).T


def rand_affine(oracle, rng):
    return oracle.local_matrix(rng.uniform(-5, 5, 3), rng.uniform(-3.1, 3.1, 3), rng.uniform(0.5, 2.0, 3))


# ---- storage / trivial factories --------------------------------------------------------------
def test_default_mat4_is_zero_like_the_reference_prints(oracle):
    # Sandbox/src/main.cpp:10-17 prints mat4{}[0,0] == 0: the only number the reference pins.
    assert np.array_equal(oracle.diagonal(0.0), np.zeros(16, F))


def test_identity_translation_scale_exact(oracle):
    assert np.array_equal(M(oracle.identity()), np.eye(4))
    t = M(oracle.from_translation((1.5, -2.0, 3.25)))
    assert np.array_equal(t[:3, 3], [1.5, -2.0, 3.25]) and np.array_equal(t[:3, :3], np.eye(3)) and t[3, 3] == 1
    # column-major storage: translation lives in m_data[12..14]  (matrix.hpp:7-8)
    assert np.array_equal(oracle.from_translation((1.5, -2.0, 3.25))[12:15], F([1.5, -2.0, 3.25]))
    s = M(oracle.from_scale((2, 3, 4)))
    assert np.array_equal(s, np.diag([2.0, 3.0, 4.0, 1.0]))


def test_from_basis_columns(oracle):
    b = M(oracle.from_basis((1, 2, 3), (4, 5, 6), (7, 8, 9)))
    assert np.array_equal(b[:3, 0], [1, 2, 3]) and np.array_equal(b[:3, 1], [4, 5, 6]) and np.array_equal(b[:3, 2], [7, 8, 9])
    assert np.array_equal(b[3], [0, 0, 0, 1])


# ---- rotations --------------------------------------------------------------------------------
def test_euler_zero_is_exact_identity(oracle):
    assert np.array_equal(M(oracle.from_euler(0.0, 0.0, 0.0)), np.eye(4))


@pytest.mark.parametrize("angles,expect", [
    ((math.pi / 2, 0, 0), [[1, 0, 0], [0, 0, -1], [0, 1, 0]]),    # pitch: about X
    ((0, math.pi / 2, 0), [[0, 0, 1], [0, 1, 0], [-1, 0, 0]]),    # yaw:   about Y
    ((0, 0, math.pi / 2), [[0, -1, 0], [1, 0, 0], [0, 0, 1]]),    # roll:  about Z
    ((-math.pi / 2, 0, 0), [[1, 0, 0], [0, 0, 1], [0, -1, 0]]),
    ((math.pi, 0, 0), [[1, 0, 0], [0, -1, 0], [0, 0, -1]]),
])
def test_euler_quarter_turns(oracle, angles, expect):
    r = M(oracle.from_euler(*angles))[:3, :3]
    assert np.abs(r - np.asarray(expect, np.float64)).max() < 2e-7   # float32(pi/2) is not pi/2


def test_euler_order_is_yaw_pitch_roll(oracle):
    p, y, r = 0.3, -0.7, 1.1
    composed = oracle.mat_mul(oracle.mat_mul(oracle.from_euler(0, y, 0), oracle.from_euler(p, 0, 0)), oracle.from_euler(0, 0, r))
    assert np.abs(M(composed) - M(oracle.from_euler(p, y, r))).max() < 3e-7


def test_euler_is_right_handed(oracle):
    # positive yaw about +Y takes +Z towards +X (right-handed, matrix.hpp:7)
    v = oracle.transform_normal(oracle.from_euler(0, math.pi / 2, 0), (0, 0, 1))
    assert np.abs(v - F([1, 0, 0])).max() < 2e-7


def test_axis_angle_matches_euler_on_the_axes(oracle):
    a = 0.83
    assert np.abs(M(oracle.from_axis_angle((2, 0, 0), a)) - M(oracle.from_euler(a, 0, 0))).max() < 3e-7
    assert np.abs(M(oracle.from_axis_angle((0, 3, 0), a)) - M(oracle.from_euler(0, a, 0))).max() < 3e-7
    assert np.abs(M(oracle.from_axis_angle((0, 0, 0.5), a)) - M(oracle.from_euler(0, 0, a))).max() < 3e-7


def test_rotations_are_orthonormal(oracle):
    rng = np.random.default_rng(1)
    for _ in range(50):
        r = M(oracle.from_euler(*rng.uniform(-3.14, 3.14, 3)))[:3, :3]
        assert np.abs(r @ r.T - np.eye(3)).max() < 5e-7
        assert abs(np.linalg.det(r) - 1.0) < 5e-7
        q = M(oracle.from_axis_angle(rng.normal(size=3), rng.uniform(-3, 3)))[:3, :3]
        assert np.abs(q @ q.T - np.eye(3)).max() < 1e-6


def test_euler_against_float64_shadow(oracle):
    rng = np.random.default_rng(2)
    for _ in range(200):
        p, y, r = (float(F(v)) for v in rng.uniform(-3.14159, 3.14159, 3))
        assert np.abs(oracle.from_euler(p, y, r).astype(np.float64) - oracle.from_euler64(p, y, r)).max() < 3e-7


# ---- deterministic sincos ---------------------------------------------------------------------
def ulp_error(got, ref64):
    ulp = np.spacing(np.abs(ref64).astype(np.float32)).astype(np.float64)
    return np.abs(got.astype(np.float64) - ref64) / ulp


@pytest.mark.parametrize("span", [math.pi, 100.0, 1.0e4, 1.0e5])
def test_sincos_within_two_ulp_of_float64(oracle, span):
    rng = np.random.default_rng(3)
    x = rng.uniform(-span, span, 400_000).astype(np.float32)
    s, c = oracle.sincos(x)
    s64, c64 = oracle.sincos64(x)
    assert ulp_error(s, s64).max() <= 2.0
    assert ulp_error(c, c64).max() <= 2.0
    assert np.abs(s - s64).max() < 1.2e-7 and np.abs(c - c64).max() < 1.2e-7


def test_sincos_agrees_with_libm_float32(oracle):
    x = np.linspace(-math.pi, math.pi, 200_001).astype(np.float32)
    s, c = oracle.sincos(x)
    assert np.abs(s - np.sin(x)).max() <= 2.4e-7 and np.abs(c - np.cos(x)).max() <= 2.4e-7


def test_sincos_special_values_and_symmetry(oracle):
    s, c = oracle.sincos(F([0.0]))
    assert s[0] == 0.0 and c[0] == 1.0
    x = np.random.default_rng(4).uniform(-50, 50, 10_000).astype(np.float32)
    s1, c1 = oracle.sincos(x)
    s2, c2 = oracle.sincos(-x)
    assert np.array_equal(s1, -s2) and np.array_equal(c1, c2)
    assert np.abs(s1 * s1 + c1 * c1 - 1).max() < 4e-7


# ---- arithmetic -------------------------------------------------------------------------------
def test_mat_mul_matches_float64_and_column_vector_convention(oracle):
    rng = np.random.default_rng(5)
    a, b = rng.normal(size=16).astype(F), rng.normal(size=16).astype(F)
    assert np.abs(M(oracle.mat_mul(a, b)) - M(a) @ M(b)).max() < 2e-6
    v = rng.normal(size=4).astype(F)
    assert np.abs(oracle.mat_mul_vec4(a, v) - M(a) @ v.astype(np.float64)).max() < 2e-6
    # (A*B)*v == A*(B*v) up to rounding
    lhs = oracle.mat_mul_vec4(oracle.mat_mul(a, b), v)
    rhs = oracle.mat_mul_vec4(a, oracle.mat_mul_vec4(b, v))
    assert np.abs(lhs - rhs).max() < 5e-6


def test_mat_mul_operation_order_is_left_to_right_over_columns(oracle):
    # out = ((c0*x + c1*y) + c2*z) + c3*w in float32, no FMA: a cancellation case tells orders apart
    a = np.zeros(16, F)
    a[0], a[4], a[8], a[12] = 1.0, 1.0, 1.0, 1.0          # row 0 = (1,1,1,1)
    v = F([1.0e8, 1.0, -1.0e8, 1.0])
    got = oracle.mat_mul_vec4(a, v)[0]
    expect = F(F(F(F(1.0e8) + F(1.0)) + F(-1.0e8)) + F(1.0))   # = 1.0 (the first +1 is absorbed)
    assert got == expect == F(1.0)


def test_add_sub_scale_div_transpose(oracle):
    rng = np.random.default_rng(6)
    a, b = rng.normal(size=16).astype(F), rng.normal(size=16).astype(F)
    assert np.array_equal(oracle.mat_add(a, b), a + b) and np.array_equal(oracle.mat_sub(a, b), a - b)
    assert np.array_equal(oracle.mat_scale(a, 3.0), a * F(3.0)) and np.array_equal(oracle.mat_div(a, 3.0), a / F(3.0))
    assert np.array_equal(M(oracle.transpose(a)), M(a).T)
    assert np.array_equal(oracle.transpose(oracle.transpose(a)), a)


def test_determinant_and_inverse(oracle):
    rng = np.random.default_rng(7)
    for _ in range(50):
        a = rand_affine(oracle, rng)
        assert abs(float(oracle.determinant(a)) - np.linalg.det(M(a))) < 1e-5 * max(1.0, abs(np.linalg.det(M(a))))
        inv = oracle.inverse(a)
        assert np.abs(M(oracle.mat_mul(a, inv)) - np.eye(4)).max() < 2e-5
        assert np.abs(M(inv) - np.linalg.inv(M(a))).max() < 2e-5
        assert np.array_equal(oracle.inverse_transpose(a), oracle.transpose(inv))
    g = rng.normal(size=16).astype(F)      # a general (projective) matrix too
    assert np.abs(M(oracle.inverse(g)) @ M(g) - np.eye(4)).max() < 1e-4
    b = rng.normal(size=16).astype(F)
    assert abs(float(oracle.determinant(oracle.mat_mul(g, b))) - float(oracle.determinant(g)) * float(oracle.determinant(b))) < 1e-4 * (
        1 + abs(float(oracle.determinant(g)) * float(oracle.determinant(b))))


def test_inverse_of_singular_is_zero(oracle):
    assert np.array_equal(oracle.inverse(oracle.from_scale((1, 0, 1))), np.zeros(16, F))
    assert float(oracle.determinant(oracle.identity())) == 1.0


def test_transform_point_applies_translation_normal_ignores_it(oracle):
    m = oracle.mat_mul(oracle.from_translation((10, 20, 30)), oracle.from_scale((2, 2, 2)))
    assert np.array_equal(oracle.transform_point(m, (1, 1, 1)), F([12, 22, 32]))     # matrix.hpp:45-46
    assert np.array_equal(oracle.transform_normal(m, (1, 1, 1)), F([2, 2, 2]))      # matrix.hpp:47-48


# ---- camera matrices ----------------------------------------------------------------------------
def project(m, p):
    clip = M(m) @ np.asarray([*p, 1.0])
    return clip[:3] / clip[3]


def test_perspective_known_answer(oracle):
    # fov 90 deg, aspect 1, near 1, far 2: m00 = m11 = 1/tan(45 deg), m22 = 2/(1-2), m23 = 1*2/(1-2), m32 = -1
    p = M(oracle.perspective(math.pi / 2, 1.0, 1.0, 2.0))
    assert abs(p[0, 0] - 1) < 2e-7 and abs(p[1, 1] - 1) < 2e-7
    assert p[2, 2] == -2.0 and p[2, 3] == -2.0 and p[3, 2] == -1.0 and p[3, 3] == 0.0
    assert np.count_nonzero(p) == 5
    # right-handed, depth 0..1: the camera looks down -Z, near -> 0, far -> 1
    assert abs(project(oracle.perspective(math.pi / 2, 1.0, 1.0, 2.0), (0, 0, -1))[2]) < 1e-7
    assert abs(project(oracle.perspective(math.pi / 2, 1.0, 1.0, 2.0), (0, 0, -2))[2] - 1) < 1e-7
    ndc = project(oracle.perspective(math.pi / 2, 2.0, 1.0, 2.0), (2, 1, -1))
    assert abs(ndc[0] - 1) < 1e-6 and abs(ndc[1] - 1) < 1e-6   # aspect 2: x = 2 at z = -1 is the right edge


def test_perspective_reverse_z_known_answer(oracle):
    p = oracle.perspective_reverse_z(math.pi / 2, 1.0, 0.5)
    assert abs(project(p, (0, 0, -0.5))[2] - 1) < 1e-7        # near -> 1
    assert project(p, (0, 0, -1.0e6))[2] < 1e-6               # infinity -> 0
    assert M(p)[3, 2] == -1.0 and M(p)[2, 3] == 0.5 and M(p)[2, 2] == 0.0


def test_orthographic_known_answer(oracle):
    o = oracle.orthographic(-2, 4, -1, 3, 0.5, 10.5)
    assert np.abs(project(o, (-2, -1, -0.5)) - [-1, -1, 0]).max() < 1e-6
    assert np.abs(project(o, (4, 3, -10.5)) - [1, 1, 1]).max() < 1e-6


@pytest.mark.parametrize("target,up", [((1, 0, 0), (0, 1, 0)), ((-1, 0, 0), (0, 1, 0)), ((0, 0, -1), (0, 1, 0)),
                                       ((0, 0, 1), (0, 1, 0)), ((0, 1, 0), (0, 0, 1)), ((0, -1, 0), (0, 0, 1))])
def test_look_at_down_each_axis(oracle, target, up):
    eye = np.asarray([3.0, -2.0, 5.0])
    v = oracle.look_at(eye, eye + np.asarray(target), up)
    # the target direction maps to -Z (camera looks down -Z), the eye to the origin
    assert np.abs(oracle.transform_normal(v, target) - F([0, 0, -1])).max() < 1e-6
    assert np.abs(oracle.transform_point(v, eye)).max() < 1e-5
    r = M(v)[:3, :3]
    assert np.abs(r @ r.T - np.eye(3)).max() < 1e-6 and abs(np.linalg.det(r) - 1) < 1e-6


def test_look_at_identity_case_is_exact(oracle):
    v = oracle.look_at((0, 0, 0), (0, 0, -1), (0, 1, 0))
    assert np.array_equal(M(v), np.eye(4))


# ---- scene pieces -------------------------------------------------------------------------------
def test_local_matrix_equals_t_times_r_times_s(oracle):
    rng = np.random.default_rng(8)
    for _ in range(50):
        t, r, s = rng.uniform(-100, 100, 3), rng.uniform(-3.14, 3.14, 3), rng.uniform(0.5, 2, 3)
        trs = oracle.mat_mul(oracle.mat_mul(oracle.from_translation(t), oracle.from_euler(*r)), oracle.from_scale(s))
        assert np.all(oracle.local_matrix(t, r, s) == trs)       # numerically identical (sign of zero aside)
        assert np.abs(oracle.local_matrix(t, r, s) - oracle.local_matrix64(t, r, s)).max() < 2e-5


def test_frustum_planes_of_identity_are_the_clip_cube(oracle):
    planes = oracle.frustum_planes(oracle.identity())
    expect = [[1, 0, 0, 1], [-1, 0, 0, 1], [0, 1, 0, 1], [0, -1, 0, 1], [0, 0, 1, 0], [0, 0, -1, 1]]
    assert np.array_equal(planes, F(expect))    # -1<=x<=1, -1<=y<=1, 0<=z<=1 (D3D depth)


def test_aabb_culling_inside_outside_straddle_and_inclusive_boundary(oracle):
    vp = oracle.identity()
    assert oracle.aabb_visible(vp, (0, 0, 0.5), (0.1, 0.1, 0.1))
    assert not oracle.aabb_visible(vp, (2, 0, 0.5), (0.5, 0.5, 0.1))
    assert oracle.aabb_visible(vp, (1.25, 0, 0.5), (0.5, 0.5, 0.1))          # straddles x = 1
    assert oracle.aabb_visible(vp, (1.5, 0, 0.5), (0.5, 0.1, 0.1))           # touches x = 1 exactly: inclusive
    assert not oracle.aabb_visible(vp, (1.5, 0, 0.5), (0.4999999, 0.1, 0.1))
    assert not oracle.aabb_visible(vp, (0, 0, -0.2), (0.1, 0.1, 0.1))        # behind the near plane (z < 0)


def test_world_aabb_under_rotation_and_scale(oracle):
    m = oracle.mat_mul(oracle.from_translation((5, 0, 0)), oracle.mat_mul(oracle.from_euler(0, 0, math.pi / 2), oracle.from_scale((2, 1, 1))))
    c, e = oracle.world_aabb(m, (1, 0, 0), (1, 2, 3))
    assert np.abs(c - F([5, 2, 0])).max() < 1e-6          # (1,0,0) -> scaled (2,0,0) -> rolled to (0,2,0) -> +5 x
    assert np.abs(e - F([2, 2, 3])).max() < 1e-6          # extents (2,2,3) with x and y swapped by the quarter roll


def test_vec_dot_sums_pairs_like_dpps(oracle):
    # (p0+p1)+(p2+p3): Core/src/math/vector.cpp:88,125
    a = F([1.0e8, 1.0, -1.0e8, 1.0])
    ones = F([1, 1, 1, 1])
    assert oracle.vec4_dot(a, ones) == F(F(F(1.0e8) + F(1.0)) + F(F(-1.0e8) + F(1.0)))    # = 0, not 1 or 2
    assert oracle.vec3_dot((1.0e8, 1.0, -1.0e8), (1, 1, 1)) == F(F(F(1.0e8) + F(1.0)) + F(F(-1.0e8) + F(0.0)))


# ---- whole scene ----------------------------------------------------------------------------------
def small_tree(oracle, center_range=0.3):
    from oracle.binding import BASE_SEED, h8_geo_level_sizes
    return oracle.synth_scene(BASE_SEED + 3, h8_geo_level_sizes(3, 4, 5), fanout=5, center_range=center_range)


def test_scene_world_is_parent_times_local(oracle):
    from oracle.binding import camera
    sc = small_tree(oracle)
    view, proj = camera(oracle)
    out = oracle.scene_update(sc, view, proj)
    vp = oracle.mat_mul(proj, view)
    for i in np.random.default_rng(9).integers(0, sc["position"].shape[0], 60):
        local = oracle.local_matrix(sc["position"][i], sc["rotation"][i], sc["scale"][i])
        p = sc["parent"][i]
        expect = local if p == 0xFFFFFFFF else oracle.mat_mul(out["world"][p], local)
        assert np.array_equal(out["world"][i], expect)
        assert np.array_equal(out["mvp"][i], oracle.mat_mul(vp, expect))
        assert bool(out["flags"][i]) == oracle.aabb_visible(vp, *oracle.world_aabb(expect, sc["aabb_center"][i], sc["aabb_half"][i]))
    assert np.array_equal(out["visible"], np.flatnonzero(out["flags"]).astype(np.uint32))


def test_scene_threads_do_not_change_results(oracle):
    from oracle.binding import camera
    sc = small_tree(oracle)
    view, proj = camera(oracle)
    a, b = oracle.scene_update(sc, view, proj, threads=1), oracle.scene_update(sc, view, proj, threads=7)
    assert np.array_equal(a["world"].view(np.uint32), b["world"].view(np.uint32)) and np.array_equal(a["visible"], b["visible"])


@pytest.mark.parametrize("eye", [(0, 0, 0), (0, 0, 1500), (0, 0, 3000)])
def test_scene_against_float64_shadow(oracle, eye):
    from oracle.binding import BASE_SEED, camera, h8_geo_level_sizes
    sc = oracle.synth_scene(BASE_SEED + 3, h8_geo_level_sizes(7, 5), fanout=8, center_range=0.25)
    view, proj = camera(oracle, eye=eye)
    f32 = oracle.scene_update(sc, view, proj, threads=4)
    f64 = oracle.scene_update64(sc, view, proj, eps=1e-5)
    assert rel_err(f32["world"], f64["world"]) < 1e-5 and rel_err(f32["mvp"], f64["mvp"]) < 1e-5
    differ = f32["flags"] != f64["visible"]
    assert not np.any(differ & (f64["boundary"] == 0))      # lists agree outside the documented boundary set
    assert f64["boundary"].sum() < 0.01 * differ.size


def test_mesh_against_float64_shadow_and_normal_properties(oracle):
    from oracle.binding import BASE_SEED
    rng = np.random.default_rng(10)
    inst, verts = 5, 64
    pos, nrm = oracle.synth_mesh(BASE_SEED + 4, inst * verts)
    assert np.abs(np.linalg.norm(nrm.astype(np.float64), axis=1) - 1).max() < 2e-7
    models = np.stack([rand_affine(oracle, rng) for _ in range(inst)])
    first = (np.arange(inst + 1) * verts).astype(np.uint32)
    p32, n32 = oracle.mesh_transform(first, models, pos, nrm)
    p64, n64 = oracle.mesh_transform64(first, models, pos, nrm)
    assert rel_err(p32, p64) < 1e-5 and rel_err(n32, n64) < 1e-5
    assert np.abs(np.linalg.norm(n32.astype(np.float64), axis=1) - 1).max() < 3e-7
    # a normal stays perpendicular to transformed tangents under non-uniform scale
    for k in range(inst):
        m3 = M(models[k])[:3, :3]
        for v in range(k * verts, k * verts + 8):
            n = nrm[v].astype(np.float64)
            t = np.cross(n, [0.3, -0.5, 0.8])
            assert abs(np.dot(n32[v].astype(np.float64), m3 @ t)) < 1e-5 * np.linalg.norm(m3 @ t)
