"""-m gpu: the CUDA mesh vertex/normal transform, through the C ABI, against the CPU oracle."""
import numpy as np
import pytest

from conftest import rel_err, same_values
from oracle.binding import BASE_SEED, camera

pytestmark = pytest.mark.gpu
TOL = 1e-5


def models_from_scene(oracle, n):
    arrays = oracle.synth_scene(BASE_SEED + 2, [n], fanout=1)
    arrays["parent"] = None
    view, proj = camera(oracle)
    return oracle.scene_update(arrays, view, proj)["world"]


@pytest.mark.parametrize("instances,verts", [(1, 4), (3, 1024), (5, 2048), (7, 36), (64, 1000)])
def test_uniform_instances(oracle, gpu_ctx, instances, verts):
    import zenyth_b200 as zn
    V = instances * verts
    pos, nrm = oracle.synth_mesh(BASE_SEED + 4, V)
    models = models_from_scene(oracle, instances)
    first = (np.arange(instances + 1, dtype=np.uint32) * verts).astype(np.uint32)
    ref_p, ref_n = oracle.mesh_transform(first, models, pos, nrm)
    m = zn.Mesh(gpu_ctx, V, instances)
    m.SetVertices(pos, nrm)
    m.SetModels(models)
    m.Transform()
    got_p, got_n = m.Read()
    m.Close()
    assert rel_err(got_p, ref_p) <= TOL and rel_err(got_n, ref_n) <= TOL
    assert same_values(got_p, ref_p)
    assert same_values(got_n, ref_n)


def test_ragged_instance_ranges_and_f64(oracle, gpu_ctx):
    import zenyth_b200 as zn
    counts = np.array([1, 2, 3, 1023, 1024, 1025, 5, 4000, 0, 17], np.uint32)
    first = np.concatenate([[0], np.cumsum(counts)]).astype(np.uint32)
    V = int(first[-1])
    pos, nrm = oracle.synth_mesh(BASE_SEED + 4, V)
    models = models_from_scene(oracle, counts.size)
    ref_p, ref_n = oracle.mesh_transform(first, models, pos, nrm)
    m = zn.Mesh(gpu_ctx, V, counts.size, first)
    m.SetVertices(pos, nrm)
    m.SetModels(models)
    m.Transform()
    got_p, got_n = m.Read()
    m.Close()
    assert same_values(got_p, ref_p) and same_values(got_n, ref_n)
    p64, n64 = oracle.mesh_transform64(first, models, pos, nrm)
    assert rel_err(got_p, p64) <= TOL and rel_err(got_n, n64) <= TOL
    assert np.allclose(np.linalg.norm(got_n.astype(np.float64), axis=1), 1.0, atol=1e-6)


def test_models_bound_to_scene_world(oracle, gpu_ctx):
    import zenyth_b200 as zn
    I, verts = 16, 1024
    arrays = oracle.synth_scene(BASE_SEED + 2, [I], fanout=1)
    arrays["parent"] = None
    arrays["level_offsets"] = None
    view, proj = camera(oracle)
    world = oracle.scene_update(arrays, view, proj)["world"]
    sc = zn.Scene(gpu_ctx, I)
    sc.Load(arrays)
    sc.SetCamera(zn.Camera(view=view, proj=proj))
    sc.Update()
    m = zn.Mesh(gpu_ctx, I * verts, I)
    m.FillSynthetic(BASE_SEED + 4)
    pos, nrm = oracle.synth_mesh(BASE_SEED + 4, I * verts)
    gp, gn = m.Inputs()
    assert np.array_equal(gp.view(np.uint32), pos.view(np.uint32)) and np.array_equal(gn.view(np.uint32), nrm.view(np.uint32))
    m.BindSceneModels(sc, 0)
    m.Transform()
    got_p, got_n = m.Read()
    first = (np.arange(I + 1, dtype=np.uint32) * verts).astype(np.uint32)
    ref_p, ref_n = oracle.mesh_transform(first, world, pos, nrm)
    assert same_values(got_p, ref_p) and same_values(got_n, ref_n)
    m.Close()
    sc.Close()


def test_mesh_file_round_trip(oracle, gpu_ctx, tmp_path):
    """N4: the ZNMESH01 dump — a file written with numpy loads into a batch that matches the oracle,
    saving it reproduces the file byte for byte, and a batch bound to a scene saves its world matrices."""
    import zenyth_b200 as zn
    counts = np.array([3, 1024, 0, 700, 2049], np.uint32)          # ragged, one empty instance
    first = np.concatenate([[0], np.cumsum(counts)]).astype(np.uint32)
    V, I = int(first[-1]), counts.size
    pos, nrm = oracle.synth_mesh(BASE_SEED + 21, V)
    models = models_from_scene(oracle, I)
    path = tmp_path / "batch.znmesh"
    with open(path, "wb") as f:
        f.write(b"ZNMESH01")
        f.write(np.asarray([V, I], np.uint32).tobytes())
        f.write(first.tobytes())
        f.write(pos.astype(np.float32).tobytes())
        f.write(nrm.astype(np.float32).tobytes())
        f.write(models.astype(np.float32).reshape(I, 16).tobytes())
    m = zn.Mesh.LoadFile(gpu_ctx, path)
    assert (m.vertex_count, m.instance_count) == (V, I)
    m.Transform()
    got_p, got_n = m.Read()
    ref_p, ref_n = oracle.mesh_transform(first, models, pos, nrm)
    assert same_values(got_p, ref_p) and same_values(got_n, ref_n)
    out = tmp_path / "again.znmesh"
    m.SaveFile(out)
    assert out.read_bytes() == path.read_bytes()
    # bound to a scene: the scene's world matrices are what gets written
    arrays = oracle.synth_scene(BASE_SEED + 2, [I], fanout=1)
    arrays["parent"] = None
    view, proj = camera(oracle)
    sc = zn.Scene(gpu_ctx, I)
    sc.Load(arrays)
    sc.SetCamera(zn.Camera(view=view, proj=proj))
    sc.Update()
    m.BindSceneModels(sc, 0)
    bound = tmp_path / "bound.znmesh"
    m.SaveFile(bound)
    a, b = bound.read_bytes(), path.read_bytes()                # models_from_scene builds the same scene
    assert a[:-64 * I] == b[:-64 * I]
    assert np.array_equal(np.frombuffer(a[-64 * I:], np.float32), np.frombuffer(b[-64 * I:], np.float32))   # -0.0 == 0.0
    m.Close()
    sc.Close()
    for name, payload in (("bad.znmesh", b"NOTAMESH" + bytes(64)), ("short.znmesh", path.read_bytes()[:-20])):
        p = tmp_path / name
        p.write_bytes(payload)
        with pytest.raises(zn.ZenythError):
            zn.Mesh.LoadFile(gpu_ctx, p)
    with pytest.raises(zn.ZenythError):
        zn.Mesh.LoadFile(gpu_ctx, tmp_path / "missing.znmesh")


@pytest.mark.parametrize("verts", [1024, 40], ids=["one-instance-per-tile", "many-instances-per-tile"])
def test_model_matrices_that_are_not_plain_affine(oracle, gpu_ctx, verts):
    """The kernels fold the general 4x4 inverse for affine models (bottom row exactly 0,0,0,1, moderate finite
    entries) — results must stay equal to the oracle's GENERAL inverse bit for bit, and every model that does
    not qualify (projective bottom row, huge entries whose products overflow, an infinity, a NaN, w != 1) must
    take the general routine and propagate exactly what the oracle propagates."""
    import zenyth_b200 as zn
    rng = np.random.default_rng(31)
    base = models_from_scene(oracle, 24).copy()
    models = base.copy()
    models[1] = rng.uniform(-2, 2, 16).astype(np.float32)                 # general projective matrix
    models[2, 3] = 0.25                                                    # bottom row (0.25, 0, 0, 1)
    models[3, 15] = 2.0                                                    # w = 2
    models[4, 12:15] = np.float32([3e19, -2e19, 1e19])                     # huge translation: s-terms the general routine forms overflow
    models[5, 0] = np.float32(np.inf)
    models[6, 13] = np.float32(np.nan)
    models[7, :] = 0.0; models[7, 15] = 1.0                                # singular affine: zero matrix by C9
    models[8, 0:3] = np.float32([1e-30, 0, 0])                             # nearly singular, tiny determinant
    models[9, 12:15] = np.float32([1e17, 1e17, -1e17])                     # large but still on the folded path
    I = models.shape[0]
    V = I * verts
    pos, nrm = oracle.synth_mesh(BASE_SEED + 6, V)
    first = (np.arange(I + 1, dtype=np.uint32) * verts).astype(np.uint32)
    ref_p, ref_n = oracle.mesh_transform(first, models, pos, nrm)
    m = zn.Mesh(gpu_ctx, V, I)
    m.SetVertices(pos, nrm)
    m.SetModels(models)
    m.Transform()
    got_p, got_n = m.Read()
    m.Close()
    for k in range(I):
        sl = slice(k * verts, (k + 1) * verts)
        assert np.array_equal(got_p[sl], ref_p[sl], equal_nan=True), f"positions of instance {k}"
        assert np.array_equal(got_n[sl], ref_n[sl], equal_nan=True), f"normals of instance {k}"
    assert np.isnan(ref_n[6 * verts:7 * verts]).any() or np.isnan(ref_p[6 * verts:7 * verts]).any()   # the NaN really propagates
