"""-m gpu: BASELINE.json's full-size configurations.  The CPU oracle finishes a 16.7M-entity frame
in a fraction of a second on the GPU box's host cores, so configs[2] is compared entity for entity;
the 64Mi-vertex mesh batch (configs[3]) is compared chunk by chunk; on top come the
size-independent properties the domain offers (ascending unique visible list, count == popcount,
idempotence, children follow a moved root)."""
import subprocess

import numpy as np
import pytest

from conftest import rel_err, same_values
from oracle.binding import BASE_SEED, camera, h8_geo_level_sizes

pytestmark = pytest.mark.gpu


def test_h8geo_16m_entity_for_entity(oracle, gpu_ctx):
    import zenyth_b200 as zn
    sizes = h8_geo_level_sizes(7, 8)
    E = sum(sizes)
    assert E == 16_777_215
    arrays = oracle.synth_scene(BASE_SEED + 3, sizes, fanout=8, center_range=0.0)
    view, proj = camera(oracle, eye=(0, 0, 1500))
    ref = oracle.scene_update(arrays, view, proj, threads=oracle.hardware_threads())

    sc = zn.Scene(gpu_ctx, E, arrays["level_offsets"])
    sc.FillSynthetic(BASE_SEED + 3, 8, 0.0)              # device-side generator, same bits as the host's
    sc.SetCamera(zn.Camera(view=view, proj=proj))
    sc.Update()
    vis = sc.Visible()
    # properties
    assert vis.size == sc.VisibleCount() and np.all(np.diff(vis.astype(np.int64)) > 0) and vis[-1] < E
    # parity, exact
    assert np.array_equal(vis, ref["visible"])
    step = 1 << 21
    for first in range(0, E, step):
        n = min(step, E - first)
        w, m = sc.World(first, n), sc.Mvp(first, n)
        assert rel_err(w, ref["world"][first:first + n]) <= 1e-5 and same_values(w, ref["world"][first:first + n]), first
        assert rel_err(m, ref["mvp"][first:first + n]) <= 1e-5 and same_values(m, ref["mvp"][first:first + n]), first
    # idempotence
    sc.Update()
    assert np.array_equal(sc.Visible(), vis)
    # moving a root by d moves its whole subtree's world translation by d (to rounding), nothing else
    d = np.float32([128.0, -64.0, 32.0])
    root = 3
    moved = arrays["position"][root:root + 1] + d
    sc.SetTransforms(position=moved, first=root, count=1)
    sc.Update()
    last = sc.World(E - 4096, 4096)          # tail of the deepest level: descendants of the last root (6)
    assert same_values(last, ref["world"][E - 4096:])
    lvl7 = int(arrays["level_offsets"][7])
    per_root = sizes[7] // 7
    sub = sc.World(lvl7 + root * per_root, 1024)[:, 12:15]
    want = ref["world"][lvl7 + root * per_root: lvl7 + root * per_root + 1024, 12:15] + d
    assert np.abs(sub - want).max() <= 1e-3    # positions ~1e3: one float32 ulp is 6e-5
    sc.Close()


def _compare_matrices(sc, ref, E, what=("world", "mvp"), step=1 << 21):
    for first in range(0, E, step):
        n = min(step, E - first)
        if "world" in what:
            w = sc.World(first, n)
            assert rel_err(w, ref["world"][first:first + n]) <= 1e-5 and same_values(w, ref["world"][first:first + n]), first
        if "mvp" in what:
            m = sc.Mvp(first, n)
            assert rel_err(m, ref["mvp"][first:first + n]) <= 1e-5 and same_values(m, ref["mvp"][first:first + n]), first


def test_h8geo_16m_mixed_visibility_cameras(oracle, gpu_ctx):
    """The 16.7M-entity scene seen from INSIDE (f ~ 0.21: mixed masks in most chunks, the general
    compaction path) and from far away (wide), entity for entity; the 'mid' camera above is
    all-or-nothing per root subtree.  One through the fused update, one through a cull pass."""
    import zenyth_b200 as zn
    sizes = h8_geo_level_sizes(7, 8)
    E = sum(sizes)
    arrays = oracle.synth_scene(BASE_SEED + 3, sizes, fanout=8, center_range=0.0)
    sc = zn.Scene(gpu_ctx, E, arrays["level_offsets"])
    sc.FillSynthetic(BASE_SEED + 3, 8, 0.0)
    fractions = {}
    for name, eye, fused in (("inside", (0, 0, 0), True), ("wide", (0, 0, 3000), False), ("oblique", (400, -300, 900), False)):
        view, proj = camera(oracle, eye=eye)
        ref = oracle.scene_update(arrays, view, proj, threads=oracle.hardware_threads())
        cam = zn.Camera(view=view, proj=proj)
        if fused:
            sc.SetCamera(cam)
            sc.Update()
        else:
            sc.Cull(cam)
        vis = sc.Visible()
        assert np.array_equal(vis, ref["visible"]), name
        _compare_matrices(sc, ref, E, what=("world", "mvp") if fused else ("mvp",))
        fractions[name] = vis.size / E
        # the general path was really taken: plenty of 8192-entity chunks are neither empty nor full
        flags = np.zeros((E + 8191) // 8192 * 8192, np.uint8)
        flags[vis] = 1
        per_chunk = flags.reshape(-1, 8192).sum(axis=1)
        if name == "inside":
            assert np.count_nonzero((per_chunk > 0) & (per_chunk < 8192)) > 200
    assert 0.05 < fractions["inside"] < 0.5 < fractions["wide"]
    sc.Close()


def test_flat_1m_entity_for_entity(oracle, gpu_ctx):
    """BASELINE configs[1]: 1,048,576 flat entities, three cameras, exact."""
    import zenyth_b200 as zn
    E = 1_048_576
    arrays = oracle.synth_scene(BASE_SEED + 3, [E], fanout=1, center_range=0.0)
    sc = zn.Scene(gpu_ctx, E)
    sc.FillSynthetic(BASE_SEED + 3, 1, 0.0)
    seen = []
    for eye in ((0, 0, 0), (0, 0, 1500), (0, 0, 3000)):
        view, proj = camera(oracle, eye=eye)
        ref = oracle.scene_update(dict(arrays, parent=None, level_offsets=None), view, proj, threads=oracle.hardware_threads())
        sc.SetCamera(zn.Camera(view=view, proj=proj))
        sc.Update()
        vis = sc.Visible()
        assert np.array_equal(vis, ref["visible"]) and vis.size == sc.VisibleCount()
        _compare_matrices(sc, ref, E)
        seen.append(vis.size)
    assert seen[0] < seen[1] < seen[2]          # per-entity (incoherent) visibility: a flat scene has no subtrees
    sc.Close()


def test_h8geo_32m_scene_with_index_base(oracle, gpu_ctx):
    """One scene of BASELINE configs[4]: 33,554,430 entities (14 roots), numbered from a non-zero
    global base as the sharded bench does (scene 3 of 8 would start at 3 * E; any base shows the add)."""
    import zenyth_b200 as zn
    sizes = h8_geo_level_sizes(14, 8)
    E = sum(sizes)
    assert E == 33_554_430
    base = 3 * E
    arrays = oracle.synth_scene(BASE_SEED + 5, sizes, fanout=8, center_range=0.0)
    view, proj = camera(oracle, eye=(0, 0, 1500))
    ref = oracle.scene_update(arrays, view, proj, threads=oracle.hardware_threads())
    sc = zn.Scene(gpu_ctx, E, arrays["level_offsets"])
    sc.FillSynthetic(BASE_SEED + 5, 8, 0.0)
    sc.SetCamera(zn.Camera(view=view, proj=proj))
    sc.SetIndexBase(base)
    sc.Update()
    vis = sc.Visible()
    assert vis.size == ref["visible"].size and np.array_equal(vis.astype(np.int64), ref["visible"].astype(np.int64) + base)
    _compare_matrices(sc, ref, E)
    sc.Close()


def test_h8chain_16m_entity_for_entity(oracle, gpu_ctx):
    """SURVEY §8d stress shape: 8 levels x 2,097,152 entities, fan-out 1 (a distinct parent per child)."""
    import zenyth_b200 as zn
    sizes = [2_097_152] * 8
    E = sum(sizes)
    arrays = oracle.synth_scene(BASE_SEED + 3, sizes, fanout=1, center_range=0.0)
    view, proj = camera(oracle, eye=(0, 0, 1500))
    ref = oracle.scene_update(arrays, view, proj, threads=oracle.hardware_threads())
    sc = zn.Scene(gpu_ctx, E, arrays["level_offsets"])
    sc.FillSynthetic(BASE_SEED + 3, 1, 0.0)
    sc.SetCamera(zn.Camera(view=view, proj=proj))
    sc.Update()
    assert np.array_equal(sc.Visible(), ref["visible"])
    _compare_matrices(sc, ref, E)
    sc.Close()


def test_mesh_64m_vertices_chunked(oracle, gpu_ctx):
    import zenyth_b200 as zn
    I, verts = 65536, 1024
    V = I * verts
    flat = oracle.synth_scene(BASE_SEED + 2, [I], fanout=1)
    flat["parent"] = None
    flat["level_offsets"] = None
    view, proj = camera(oracle)
    models = oracle.scene_update(flat, view, proj, threads=oracle.hardware_threads(), want_mvp=False)["world"]
    mesh = zn.Mesh(gpu_ctx, V, I)
    mesh.FillSynthetic(BASE_SEED + 4)
    mesh.SetModels(models)
    mesh.Transform()
    pos, nrm = oracle.synth_mesh(BASE_SEED + 4, V)
    chunk_inst = 8192
    for i0 in range(0, I, chunk_inst * 2):   # every other 8M-vertex chunk: 32M vertices compared exactly
        v0, n = i0 * verts, chunk_inst * verts
        first = (np.arange(chunk_inst + 1, dtype=np.uint32) * verts).astype(np.uint32)
        rp, rn = oracle.mesh_transform(first, models[i0:i0 + chunk_inst], pos[v0:v0 + n], nrm[v0:v0 + n], threads=oracle.hardware_threads())
        gp, gn = mesh.Read(v0, n)
        assert same_values(gp, rp) and same_values(gn, rn), i0
    gp, gn = mesh.Read(V - 4096, 4096)
    assert np.abs(np.linalg.norm(gn.astype(np.float64), axis=1) - 1.0).max() < 3e-7
    mesh.Close()


def test_cpp_headless_sandbox_runs_the_reference_loop_shape(gpu_ctx):
    """config 1 stand-in: 4,681-entity 5-level tree through the C++ wrapper."""
    import json

    from zenyth_b200 import build as zb
    exe = zb.build_host()
    r = subprocess.run([exe, "200"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    out = json.loads(r.stdout.strip().splitlines()[-1])
    assert out["entities"] == 4681 and 0 < out["visible"] <= 4681 and out["ms_per_frame"] > 0
    assert out["file_round_trip"] is True             # Scene::SaveFile / LoadFile / FrameHost of the C++ layer


def test_reference_application_loop_drives_the_gpu_scene(gpu_ctx):
    """the reference's own Application::Run (Core/src/Application.cpp:26-57) and Timer,
    compiled in place with a windowless Window, run the config-1 stand-in scene on the GPU through
    the IRenderer seam.  Skipped where oracle/_ref/sandbox_reference_loop was not built."""
    import json
    import os

    from conftest import ROOT
    exe = os.path.join(ROOT, "oracle", "_ref", "sandbox_reference_loop")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/sandbox_reference_loop not built (needs /root/reference at build time)")
    env = dict(os.environ, ZENYTH_HEADLESS_FRAMES="150")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    out = json.loads(r.stdout.strip().splitlines()[-1])
    assert out["entities"] == 4681 and out["frames"] == 150 and out["resizes"] == 1
    assert 0 < out["visible"] <= 4681
    assert abs(out["aspect"] - 2.0) < 1e-3                # Application::OnEvent forwarded the 1600x800 resize to IRenderer::Resize
    assert out["split_equals_fused"] is True              # OnUpdate -> UpdateWorld, OnRender -> Cull: the same list as the fused Update
