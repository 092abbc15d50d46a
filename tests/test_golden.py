"""Committed golden vectors (tests/golden/*.npz, made by tests/golden/make_golden.py with the
oracle built on the reference's own zenyth::math types).  CPU: the restated oracle reproduces them
bit for bit.  GPU (-m gpu): the CUDA path reproduces them — visible lists exactly, matrices and
vertices equal value for value (tolerance gate 1e-5 relative)."""
import os

import numpy as np
import pytest

from conftest import bits, rel_err, same_values

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-5


def load(name):
    return dict(np.load(os.path.join(G, name)))


def test_oracle_reproduces_scene_goldens(oracle):
    g = load("scene_h3.npz")
    sc = {k: g[k] for k in ("position", "rotation", "scale", "aabb_center", "aabb_half", "parent", "level_offsets")}
    for name in ("inside", "mid", "wide"):
        r = oracle.scene_update(sc, g[f"{name}_view"], g[f"{name}_proj"])
        assert np.array_equal(r["visible"], g[f"{name}_visible"]), name
        assert np.array_equal(bits(r["mvp"]), bits(g[f"{name}_mvp"])), name
    assert np.array_equal(bits(r["world"]), bits(g["world"]))
    counts = [g[f"{n}_visible"].size for n in ("inside", "mid", "wide")]
    assert counts[0] < counts[1] <= counts[2] and counts[0] > 0   # the three cameras really differ
    f = load("scene_flat300.npz")
    fl = {k: f[k] for k in ("position", "rotation", "scale", "aabb_center", "aabb_half")}
    r = oracle.scene_update(fl, f["view"], f["proj"])
    assert np.array_equal(r["visible"], f["visible"]) and np.array_equal(bits(r["world"]), bits(f["world"]))


def test_oracle_reproduces_mesh_and_mat4_goldens(oracle):
    m = load("mesh_ragged.npz")
    p, n = oracle.mesh_transform(m["first"], m["models"], m["pos"], m["nrm"])
    assert np.array_equal(bits(p), bits(m["pos_out"])) and np.array_equal(bits(n), bits(m["nrm_out"]))
    k = load("mat4_kat.npz")
    assert np.array_equal(bits(oracle.perspective(np.float32(np.pi / 2), 1.0, 1.0, 2.0)), bits(k["perspective"]))
    assert np.array_equal(bits(oracle.look_at((3, -2, 5), (0, 0, 0), (0, 1, 0))), bits(k["look_at"]))
    assert np.array_equal(bits(oracle.from_euler(0.3, -0.7, 1.1)), bits(k["euler"]))


def test_product_host_maths_reproduces_mat4_goldens(zn_lib):
    import zenyth_b200 as zn
    k = load("mat4_kat.npz")
    assert np.array_equal(bits(zn.mat4.perspective(np.float32(np.pi / 2), 1.0, 1.0, 2.0)), bits(k["perspective"]))
    v = zn.mat4.look_at((3, -2, 5), (0, 0, 0), (0, 1, 0))
    assert np.array_equal(bits(v), bits(k["look_at"]))
    assert np.array_equal(bits(zn.mat4.frustum_planes(zn.mat4.mul(zn.mat4.perspective(1.0, 1.6, 0.1, 900.0), v))), bits(k["planes"]))


@pytest.mark.gpu
def test_gpu_reproduces_scene_goldens(gpu_ctx):
    import zenyth_b200 as zn
    g = load("scene_h3.npz")
    sc = zn.Scene(gpu_ctx, g["position"].shape[0], g["level_offsets"])
    sc.Load(g)
    for name in ("inside", "mid", "wide"):
        sc.SetCamera(zn.Camera(view=g[f"{name}_view"], proj=g[f"{name}_proj"]))
        sc.Update()
        assert np.array_equal(sc.Visible(), g[f"{name}_visible"]), name
        mvp = sc.Mvp()
        assert rel_err(mvp, g[f"{name}_mvp"]) <= TOL and same_values(mvp, g[f"{name}_mvp"]), name
    world = sc.World()
    assert rel_err(world, g["world"]) <= TOL and same_values(world, g["world"])
    sc.Close()
    f = load("scene_flat300.npz")
    fl = zn.Scene(gpu_ctx, 300)
    fl.Load(f)
    fl.SetCamera(zn.Camera(view=f["view"], proj=f["proj"]))
    fl.Update()
    assert np.array_equal(fl.Visible(), f["visible"]) and same_values(fl.World(), f["world"]) and same_values(fl.Mvp(), f["mvp"])
    fl.Close()


@pytest.mark.gpu
def test_gpu_reproduces_mesh_goldens(gpu_ctx):
    import zenyth_b200 as zn
    m = load("mesh_ragged.npz")
    mesh = zn.Mesh(gpu_ctx, m["pos"].shape[0], m["first"].size - 1, m["first"])
    mesh.SetVertices(m["pos"], m["nrm"])
    mesh.SetModels(m["models"])
    mesh.Transform()
    p, n = mesh.Read()
    mesh.Close()
    assert rel_err(p, m["pos_out"]) <= TOL and rel_err(n, m["nrm_out"]) <= TOL
    assert same_values(p, m["pos_out"]) and same_values(n, m["nrm_out"])
