"""A third, independent statement of the path: plain numpy float64 written from the CONVENTIONS
text, not from the oracle's code.  The reference pins nothing (no tests, a
stub mat4), so the oracle is pinned from several sides: closed-form KATs, the C++ float64 shadow,
the reference's own vector code — and this file, which shares no line with any of them."""
import numpy as np
import pytest

from conftest import rel_err
from oracle.binding import BASE_SEED, h8_geo_level_sizes


def rot_x(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[1, 0, 0], [0, c, -s], [0, s, c]])


def rot_y(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])


def rot_z(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]])


def local_matrix(t, r, s):
    """C4/C5: local = T * R * S with R = Ry(yaw) * Rx(pitch) * Rz(roll), rotation = (pitch, yaw, roll)."""
    m = np.eye(4)
    m[:3, :3] = rot_y(r[1]) @ rot_x(r[0]) @ rot_z(r[2]) @ np.diag(s)
    m[:3, 3] = t
    return m


def look_at(eye, center, up):
    """C6: f = norm(center - eye), s = norm(f x up), u = s x f, rows (s, u, -f); right-handed."""
    f = (center - eye) / np.linalg.norm(center - eye)
    s = np.cross(f, up)
    s /= np.linalg.norm(s)
    u = np.cross(s, f)
    m = np.eye(4)
    m[0, :3], m[1, :3], m[2, :3] = s, u, -f
    m[:3, 3] = -m[:3, :3] @ eye
    return m


def perspective(fov_y, aspect, n, f):
    """C6: right-handed, clip depth 0..1."""
    t = 1.0 / np.tan(fov_y / 2)
    m = np.zeros((4, 4))
    m[0, 0], m[1, 1], m[2, 2], m[2, 3], m[3, 2] = t / aspect, t, f / (n - f), n * f / (n - f), -1.0
    return m


def frustum_planes(vp):
    """C7: Gribb-Hartmann from the rows of P*V, D3D depth: r3+-r0, r3+-r1, r2, r3-r2."""
    r = vp
    return np.stack([r[3] + r[0], r[3] - r[0], r[3] + r[1], r[3] - r[1], r[2], r[3] - r[2]])


def scene_numpy(sc, view, proj):
    n = sc["position"].shape[0]
    world = np.empty((n, 4, 4))
    parent = sc["parent"]
    for i in range(n):          # level order: a parent always comes first
        local = local_matrix(sc["position"][i].astype(np.float64), sc["rotation"][i].astype(np.float64), sc["scale"][i].astype(np.float64))
        world[i] = local if parent[i] == 0xFFFFFFFF else world[parent[i]] @ local
    vp = proj @ view
    mvp = vp @ world
    planes = frustum_planes(vp)
    c = np.einsum("nij,nj->ni", world, np.concatenate([sc["aabb_center"].astype(np.float64), np.ones((n, 1))], axis=1))   # C8: c' = M (c, 1)
    e = np.einsum("nij,nj->ni", np.abs(world[:, :3, :3]), sc["aabb_half"].astype(np.float64))                              # e' = |M3x3| e
    margin = c @ planes.T + e @ np.abs(planes[:, :3]).T                                                                   # plane . (c',1) + |n| . e'
    scale = np.abs(c[:, :3]) @ np.abs(planes[:, :3]).T + e @ np.abs(planes[:, :3]).T + np.abs(planes[:, 3])
    return world, mvp, margin, scale


def col_major(m):
    """[n,4,4] row/column matrices -> the float[16] column-major layout of mat4::data() (C1)."""
    return np.ascontiguousarray(np.swapaxes(m, -1, -2)).reshape(m.shape[0], 16)


@pytest.mark.parametrize("eye", [(0.0, 0.0, 0.0), (0.0, 0.0, 1500.0), (900.0, -200.0, 1200.0)])
def test_oracle_scene_matches_a_numpy_float64_statement_of_the_conventions(oracle, eye):
    sc = oracle.synth_scene(BASE_SEED + 3, h8_geo_level_sizes(7, 4), fanout=8, center_range=0.25)
    eye64 = np.asarray(eye, np.float64)
    fov = float(np.float32(60.0) * np.float32(np.pi) / np.float32(180.0))
    view64 = look_at(eye64, eye64 + np.array([0.0, 0.0, -1.0]), np.array([0.0, 1.0, 0.0]))
    proj64 = perspective(fov, 1280.0 / 720.0, 0.1, 5000.0)
    view = oracle.look_at(np.asarray(eye, np.float32), np.asarray(eye, np.float32) + np.asarray([0, 0, -1], np.float32), (0.0, 1.0, 0.0))
    proj = oracle.perspective(np.float32(fov), 1280.0 / 720.0, 0.1, 5000.0)
    assert np.abs(col_major(view64[None])[0] - view.astype(np.float64)).max() < 1e-6 * max(1.0, np.abs(eye64).max())
    assert rel_err(proj, col_major(proj64[None])[0]) < 1e-6
    got = oracle.scene_update(sc, view, proj, threads=4)
    world, mvp, margin, scale = scene_numpy(sc, view64, proj64)
    assert rel_err(got["world"], col_major(world)) < 1e-5          # the north star's tolerance
    assert rel_err(got["mvp"], col_major(mvp)) < 1e-5
    visible64 = np.all(margin >= 0.0, axis=1)
    boundary = np.any(np.abs(margin) <= 1e-5 * scale, axis=1)      # the documented plane-boundary set
    differ = got["flags"].astype(bool) != visible64
    assert not np.any(differ & ~boundary)
    assert boundary.sum() < 0.01 * boundary.size
    assert np.array_equal(got["visible"], np.flatnonzero(got["flags"]).astype(np.uint32))


def test_oracle_mesh_matches_numpy_float64(oracle):
    rng = np.random.default_rng(31)
    inst, verts = 6, 40
    pos, nrm = oracle.synth_mesh(BASE_SEED + 4, inst * verts)
    models = np.stack([local_matrix(rng.uniform(-5, 5, 3), rng.uniform(-3, 3, 3), rng.uniform(0.5, 2.0, 3) * rng.choice([-1, 1], 3))
                       for _ in range(inst)])
    first = (np.arange(inst + 1) * verts).astype(np.uint32)
    p32, n32 = oracle.mesh_transform(first, col_major(models).astype(np.float32), pos, nrm)
    for k in range(inst):
        m = col_major(models[k:k + 1]).astype(np.float32).astype(np.float64).reshape(4, 4).T    # the matrix the oracle actually saw
        sl = slice(k * verts, (k + 1) * verts)
        p = pos[sl].astype(np.float64) @ m[:3, :3].T + m[:3, 3]                                   # transform_point: w = 1
        nn = nrm[sl].astype(np.float64) @ np.linalg.inv(m[:3, :3])                                # (M^-1)^T n, row-vector form (C9)
        nn /= np.linalg.norm(nn, axis=1, keepdims=True)
        assert rel_err(p32[sl], p) < 1e-5 and rel_err(n32[sl], nn) < 1e-5
