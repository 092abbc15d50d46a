"""External pins for the [chosen] conventions — authorities that share no author
with the oracle.

The reference declares these functions (Core/include/math/matrix.hpp:23-24,40-43,52-55) over stub
bodies (Core/src/math/matrix.cpp:23-29,79-93,107-121), so it cannot arbitrate what they compute.
Every other pin in tests/ (KATs, float64 shadow, numpy restatement) was written by the oracle's own
author from the same conventions text.  This file checks the oracle against:

  * scipy.spatial.transform.Rotation — from_euler (intrinsic 'YXZ' = Ry(yaw)*Rx(pitch)*Rz(roll))
    and from_axis_angle (rotation vectors);
  * the published DirectXMath right-handed formulas the D3D12 target implies
    (Core/CMakeLists.txt:14 links d3d12): XMMatrixPerspectiveFovRH, XMMatrixOrthographicOffCenterRH,
    XMMatrixLookAtRH.  DirectXMath uses ROW vectors (v' = v*M, row-major storage); the reference
    uses COLUMN vectors (matrix.hpp:50), so every DirectXMath matrix is transposed before comparing;
  * numpy.linalg (LAPACK) in float64 for inverse / determinant / transpose / products;
  * the D3D clip-volume definition (-w <= x <= w, -w <= y <= w, 0 <= z <= w) for the frustum planes,
    and the eight-corner definition of "a box is outside a plane" for the AABB test.

Tolerance 1e-6 absolute on O(1) entries unless stated.
"""
import math

import numpy as np
import pytest
from scipy.spatial.transform import Rotation

TOL = 1e-6


def M(a):
    """16 column-major floats (mat4::data(), matrix.hpp:57-61) -> 4x4 [row, col] float64."""
    return np.asarray(a, np.float64).reshape(4, 4).T


# ---- DirectXMath, as published (DirectXMathMatrix.inl), row-vector convention -------------------
def xm_perspective_fov_rh(fov_y, aspect, near_z, far_z):
    height = math.cos(0.5 * fov_y) / math.sin(0.5 * fov_y)
    width = height / aspect
    f_range = far_z / (near_z - far_z)
    m = np.zeros((4, 4))
    m[0, 0] = width
    m[1, 1] = height
    m[2, 2] = f_range
    m[2, 3] = -1.0
    m[3, 2] = f_range * near_z
    return m


def xm_orthographic_off_center_rh(left, right, bottom, top, near_z, far_z):
    rw = 1.0 / (right - left)
    rh = 1.0 / (top - bottom)
    f_range = 1.0 / (near_z - far_z)
    m = np.zeros((4, 4))
    m[0, 0] = rw + rw
    m[1, 1] = rh + rh
    m[2, 2] = f_range
    m[3] = [-(left + right) * rw, -(top + bottom) * rh, f_range * near_z, 1.0]
    return m


def xm_look_to_lh(eye, direction, up):
    r2 = direction / np.linalg.norm(direction)
    r0 = np.cross(up, r2)
    r0 /= np.linalg.norm(r0)
    r1 = np.cross(r2, r0)
    neg = -eye
    m = np.zeros((4, 4))
    m[0] = [*r0, r0 @ neg]
    m[1] = [*r1, r1 @ neg]
    m[2] = [*r2, r2 @ neg]
    m[3] = [0, 0, 0, 1]
    return m.T          # XMMatrixLookToLH ends with XMMatrixTranspose


def xm_look_at_rh(eye, focus, up):
    return xm_look_to_lh(eye, eye - focus, up)


def to_column_convention(xm):
    return xm.T


# ---- rotations vs scipy -------------------------------------------------------------------------
def test_from_euler_is_scipy_intrinsic_YXZ(oracle):
    rng = np.random.default_rng(11)
    worst = 0.0
    for _ in range(500):
        pitch, yaw, roll = (float(np.float32(v)) for v in rng.uniform(-math.pi, math.pi, 3))
        got = M(oracle.from_euler(pitch, yaw, roll))
        want = Rotation.from_euler("YXZ", [yaw, pitch, roll]).as_matrix()
        worst = max(worst, np.abs(got[:3, :3] - want).max())
        assert np.array_equal(got[3], [0, 0, 0, 1]) and np.array_equal(got[:3, 3], [0, 0, 0])
    assert worst <= TOL, worst


def test_from_axis_angle_is_scipy_rotvec(oracle):
    rng = np.random.default_rng(12)
    worst = 0.0
    for _ in range(300):
        axis = rng.normal(size=3).astype(np.float32)
        angle = float(np.float32(rng.uniform(-math.pi, math.pi)))
        got = M(oracle.from_axis_angle(axis, angle))[:3, :3]
        unit = axis.astype(np.float64) / np.linalg.norm(axis.astype(np.float64))
        worst = max(worst, np.abs(got - Rotation.from_rotvec(unit * angle).as_matrix()).max())
    assert worst <= TOL, worst


def test_local_trs_is_translate_after_rotate_after_scale(oracle):
    """C5 against scipy: a point goes through S, then R (YXZ), then T."""
    rng = np.random.default_rng(13)
    for _ in range(200):
        t = rng.uniform(-5, 5, 3).astype(np.float32)
        r = rng.uniform(-3.1, 3.1, 3).astype(np.float32)       # (pitch, yaw, roll)
        s = rng.uniform(0.5, 2.0, 3).astype(np.float32)
        p = rng.uniform(-1, 1, 3).astype(np.float32)
        got = oracle.transform_point(oracle.local_matrix(t, r, s), p).astype(np.float64)
        rot = Rotation.from_euler("YXZ", [float(r[1]), float(r[0]), float(r[2])])
        want = rot.apply(p.astype(np.float64) * s) + t
        assert np.abs(got - want).max() <= 5e-6


# ---- projections / view vs DirectXMath ----------------------------------------------------------
@pytest.mark.parametrize("fov_deg,aspect,near_z,far_z", [
    (60.0, 1280.0 / 720.0, 0.1, 5000.0),      # the bench camera (Sandbox/include/SandboxApp.hpp:6 viewport)
    (90.0, 1.0, 1.0, 2.0),
    (45.0, 4.0 / 3.0, 0.5, 100.0),
    (120.0, 2.39, 0.01, 10.0),
])
def test_perspective_is_XMMatrixPerspectiveFovRH(oracle, fov_deg, aspect, near_z, far_z):
    fov = float(np.float32(math.radians(fov_deg)))
    got = M(oracle.perspective(fov, aspect, near_z, far_z))
    want = to_column_convention(xm_perspective_fov_rh(fov, float(np.float32(aspect)), float(np.float32(near_z)), float(np.float32(far_z))))
    assert np.abs(got - want).max() <= TOL * max(1.0, np.abs(want).max())
    assert got[3, 2] == -1.0 and got[3, 3] == 0.0          # w_clip = -z_view: right-handed


@pytest.mark.parametrize("fov_deg,aspect,near_z", [(60.0, 16.0 / 9.0, 0.1), (75.0, 1.0, 1.0), (30.0, 2.0, 0.25)])
def test_perspective_reverse_z_is_the_swapped_plane_limit_of_XMMatrixPerspectiveFovRH(oracle, fov_deg, aspect, near_z):
    """Reverse-Z with an infinite far plane = the standard D3D projection with near and far swapped,
    far -> infinity (depth 1 at the near plane, 0 at infinity)."""
    fov = float(np.float32(math.radians(fov_deg)))
    got = M(oracle.perspective_reverse_z(fov, aspect, near_z))
    n = float(np.float32(near_z))
    want = to_column_convention(xm_perspective_fov_rh(fov, float(np.float32(aspect)), 1e30, n))
    assert np.abs(got - want).max() <= TOL
    for z_view, depth in ((-n, 1.0), (-1e9, 0.0)):
        clip = got @ np.array([0.0, 0.0, z_view, 1.0])
        assert abs(clip[2] / clip[3] - depth) <= 1e-6


@pytest.mark.parametrize("box", [(-1, 1, -1, 1, 0, 1), (-10, 30, -5, 2, 0.1, 100), (0, 1280, 0, 720, -1, 1)])
def test_orthographic_is_XMMatrixOrthographicOffCenterRH(oracle, box):
    got = M(oracle.orthographic(*box))
    want = to_column_convention(xm_orthographic_off_center_rh(*(float(np.float32(v)) for v in box)))
    assert np.abs(got - want).max() <= TOL * max(1.0, np.abs(want).max())


def test_look_at_is_XMMatrixLookAtRH(oracle):
    rng = np.random.default_rng(14)
    worst = 0.0
    for _ in range(300):
        eye = rng.uniform(-50, 50, 3).astype(np.float32)
        focus = (eye + rng.normal(size=3) * rng.uniform(0.5, 20)).astype(np.float32)
        up = np.array([0, 1, 0], np.float32) if rng.random() < 0.5 else rng.normal(size=3).astype(np.float32)
        f = (focus - eye).astype(np.float64)
        if np.linalg.norm(np.cross(f / np.linalg.norm(f), up / np.linalg.norm(up))) < 0.2:
            continue        # nearly parallel: ill-conditioned for any implementation
        got = M(oracle.look_at(eye, focus, up))
        want = to_column_convention(xm_look_at_rh(eye.astype(np.float64), focus.astype(np.float64), up.astype(np.float64)))
        worst = max(worst, np.abs(got - want).max() / max(1.0, np.abs(want).max()))
    assert worst <= 2e-6, worst


def test_bench_camera_chain_is_directxmath(oracle):
    """The 'mid' camera of the bench (eye (0,0,1500) looking down -Z, 60 deg, 16:9, 0.1..5000):
    P*V as the GPU path consumes it, against XMMatrixLookAtRH * XMMatrixPerspectiveFovRH."""
    from oracle.binding import camera
    view, proj = camera(oracle)
    got = M(oracle.mat_mul(proj, view))
    fov = float(np.float32(60.0) * np.float32(np.pi) / np.float32(180.0))
    xm = xm_look_at_rh(np.array([0, 0, 1500.0]), np.array([0, 0, 1499.0]), np.array([0, 1.0, 0])) @ \
        xm_perspective_fov_rh(fov, float(np.float32(1280.0 / 720.0)), float(np.float32(0.1)), 5000.0)
    want = to_column_convention(xm)       # row vectors: v*V*P  ==  column vectors: P*V*v
    assert np.abs(got - want).max() / np.abs(want).max() <= TOL


# ---- linear algebra vs LAPACK -------------------------------------------------------------------
def _random_matrices(oracle, rng, n):
    out = []
    for _ in range(n):
        m = oracle.local_matrix(rng.uniform(-5, 5, 3), rng.uniform(-3.1, 3.1, 3), rng.uniform(0.5, 2.0, 3))
        out.append(m)
        g = rng.uniform(-2, 2, 16).astype(np.float32)      # a general (projective) matrix too
        if abs(np.linalg.det(M(g))) > 0.5:
            out.append(g)
    return out


def test_inverse_determinant_transpose_are_lapack(oracle):
    rng = np.random.default_rng(15)
    for m in _random_matrices(oracle, rng, 200):
        a = M(m)
        inv = np.linalg.inv(a)
        scale = max(1.0, np.abs(inv).max())
        assert np.abs(M(oracle.inverse(m)) - inv).max() <= 5e-6 * scale
        assert np.abs(M(oracle.inverse_transpose(m)) - inv.T).max() <= 5e-6 * scale
        assert abs(float(oracle.determinant(m)) - np.linalg.det(a)) <= 5e-6 * max(1.0, abs(np.linalg.det(a)))
        assert np.array_equal(M(oracle.transpose(m)), a.T)


def test_affine_inverse_within_1e6(oracle):
    """Well-conditioned rigid transforms: the judge's 1e-6 bar."""
    rng = np.random.default_rng(16)
    for _ in range(200):
        m = oracle.local_matrix(rng.uniform(-1, 1, 3), rng.uniform(-3.1, 3.1, 3), (1.0, 1.0, 1.0))
        assert np.abs(M(oracle.inverse(m)) - np.linalg.inv(M(m))).max() <= TOL


def test_products_are_numpy_matmul(oracle):
    rng = np.random.default_rng(17)
    for _ in range(200):
        a = rng.uniform(-2, 2, 16).astype(np.float32)
        b = rng.uniform(-2, 2, 16).astype(np.float32)
        v = rng.uniform(-2, 2, 4).astype(np.float32)
        assert np.abs(M(oracle.mat_mul(a, b)) - M(a) @ M(b)).max() <= 4e-6            # column vectors: (a*b)*v = a*(b*v)
        assert np.abs(oracle.mat_mul_vec4(a, v).astype(np.float64) - M(a) @ v.astype(np.float64)).max() <= 4e-6
        p = v[:3]
        assert np.abs(oracle.transform_point(a, p) - (M(a) @ np.append(p, 1.0))[:3]).max() <= 4e-6      # w = 1 (matrix.hpp:45-46)
        assert np.abs(oracle.transform_normal(a, p) - (M(a) @ np.append(p, 0.0))[:3]).max() <= 4e-6     # w = 0 (matrix.hpp:47-48)


# ---- frustum planes / AABB test vs their definitions ---------------------------------------------
def test_frustum_planes_describe_the_d3d_clip_volume(oracle):
    """A point is inside all six planes  <=>  its clip coordinates satisfy the D3D clip volume."""
    from oracle.binding import camera
    rng = np.random.default_rng(18)
    for eye in ((0.0, 0.0, 1500.0), (0.0, 0.0, 0.0), (300.0, -200.0, 900.0)):
        view, proj = camera(oracle, eye=eye)
        vp = oracle.mat_mul(proj, view)
        planes = oracle.frustum_planes(vp).astype(np.float64)
        pts = rng.uniform(-3000, 3000, (20000, 3))
        clip = (M(vp) @ np.c_[pts, np.ones(len(pts))].T).T
        x, y, z, w = clip.T
        margins = np.stack([w + x, w - x, w + y, w - y, z, w - z], axis=1)
        by_planes = np.c_[pts, np.ones(len(pts))] @ planes.T
        safe = np.abs(margins).min(axis=1) > 1e-9 * (1 + np.abs(clip).max(axis=1))     # both sides are float64: only exact ties are excluded
        assert np.array_equal((by_planes >= 0).all(axis=1)[safe], (margins >= 0).all(axis=1)[safe])
        assert safe.mean() > 0.95
        assert 0.005 < (margins >= 0).all(axis=1).mean() < 0.9      # the sample straddles the frustum


def test_aabb_test_is_the_eight_corner_definition(oracle):
    """C8: a box is culled iff, for some plane, all 8 corners of its WORLD-space AABB are outside."""
    from oracle.binding import camera
    rng = np.random.default_rng(19)
    view, proj = camera(oracle, eye=(0.0, 0.0, 40.0))
    vp = oracle.mat_mul(proj, view)
    planes = oracle.frustum_planes(vp).astype(np.float64)
    signs = np.array([[sx, sy, sz] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)], np.float64)
    checked = 0
    for _ in range(3000):
        m = oracle.local_matrix(rng.uniform(-60, 60, 3), rng.uniform(-3.1, 3.1, 3), rng.uniform(0.5, 2.0, 3))
        c = rng.uniform(-1, 1, 3).astype(np.float32)
        e = rng.uniform(0.1, 3.0, 3).astype(np.float32)
        wc, we = oracle.world_aabb(m, c, e)
        # the world AABB encloses the transformed local box (Arvo): check against the 8 transformed corners
        corners_local = c.astype(np.float64) + signs * e.astype(np.float64)
        corners_world = (M(m) @ np.c_[corners_local, np.ones(8)].T).T[:, :3]
        assert np.all(corners_world.min(axis=0) >= wc - we - 1e-4) and np.all(corners_world.max(axis=0) <= wc + we + 1e-4)
        assert np.abs((corners_world.max(axis=0) + corners_world.min(axis=0)) / 2 - wc).max() <= 1e-4
        assert np.abs((corners_world.max(axis=0) - corners_world.min(axis=0)) / 2 - we).max() <= 1e-4
        box = wc.astype(np.float64) + signs * we.astype(np.float64)
        d = np.c_[box, np.ones(8)] @ planes.T                  # [corner, plane]
        best = d.max(axis=0)                                   # the most-inside corner per plane
        scale = np.abs(planes[:, :3]) @ (np.abs(wc) + we).astype(np.float64) + np.abs(planes[:, 3])
        if (np.abs(best) < 1e-4 * scale).any():
            continue                                           # within float32 rounding of a plane
        checked += 1
        assert oracle.aabb_visible(vp, wc, we) == bool((best >= 0).all())
    assert checked > 2500
