"""-m gpu: meshes made of many small instances — tiles that span several instances take the
per-vertex instance lookup path of mesh_tile_kernel."""
import numpy as np
import pytest

from conftest import same_values
from oracle.binding import BASE_SEED, camera

pytestmark = pytest.mark.gpu


def models_for(oracle, n):
    arrays = oracle.synth_scene(BASE_SEED + 2, [n], fanout=1)
    arrays["parent"] = None
    arrays["level_offsets"] = None
    view, proj = camera(oracle)
    return oracle.scene_update(arrays, view, proj, want_mvp=False)["world"]


def run(oracle, gpu_ctx, counts):
    import zenyth_b200 as zn
    counts = np.asarray(counts, np.uint32)
    first = np.concatenate([[0], np.cumsum(counts)]).astype(np.uint32)
    V = int(first[-1])
    pos, nrm = oracle.synth_mesh(BASE_SEED + 4, V)
    models = models_for(oracle, counts.size)
    ref_p, ref_n = oracle.mesh_transform(first, models, pos, nrm, threads=oracle.hardware_threads())
    mesh = zn.Mesh(gpu_ctx, V, counts.size, first)
    mesh.SetVertices(pos, nrm)
    mesh.SetModels(models)
    mesh.Transform()
    p, n = mesh.Read()
    mesh.Close()
    assert same_values(p, ref_p) and same_values(n, ref_n)


def test_cubes_36_vertices_each(oracle, gpu_ctx):
    run(oracle, gpu_ctx, [36] * 5000)


def test_triangles_and_unaligned_tiny_instances(oracle, gpu_ctx):
    rng = np.random.default_rng(3)
    run(oracle, gpu_ctx, rng.integers(0, 9, 20000))          # 0..8 vertices, many empty, nothing 4-aligned


def test_more_instances_per_tile_than_the_shared_window(oracle, gpu_ctx):
    run(oracle, gpu_ctx, [1] * 3000 + [2000] + [1] * 777)     # > 512 instances inside one 1024-vertex tile


def test_mixed_large_and_small(oracle, gpu_ctx):
    rng = np.random.default_rng(4)
    counts = np.concatenate([rng.integers(1, 50, 500), [5000, 1024, 1023, 1025], rng.integers(100, 3000, 40), [3]])
    run(oracle, gpu_ctx, counts)


def test_vertex_count_not_a_multiple_of_four(oracle, gpu_ctx):
    for tail in (1, 2, 3):
        run(oracle, gpu_ctx, [1024, 512 + tail])


@pytest.mark.parametrize("instances,verts", [(20000, 1), (9000, 2), (7001, 3), (5000, 4), (3000, 8), (5000, 36), (300, 100),
                                             (33, 1000), (9, 1025), (5, 3000), (1024, 1), (1025, 1)])
def test_equal_ranges_without_a_table_equal_the_explicit_table_and_the_oracle(oracle, gpu_ctx, instances, verts):
    """A mesh created WITHOUT instance_first_vertex has equal ranges: its kernel derives every boundary arithmetically
    (no table read).  Same results as the same ranges passed explicitly, and as the oracle."""
    import zenyth_b200 as zn
    V = instances * verts
    first = (np.arange(instances + 1, dtype=np.uint64) * verts).astype(np.uint32)
    pos, nrm = oracle.synth_mesh(BASE_SEED + 5, V)
    models = models_for(oracle, instances)
    ref_p, ref_n = oracle.mesh_transform(first, models, pos, nrm, threads=oracle.hardware_threads())
    got = []
    for table in (None, first):
        mesh = zn.Mesh(gpu_ctx, V, instances, table)
        mesh.SetVertices(pos, nrm)
        mesh.SetModels(models)
        mesh.Transform()
        got.append(mesh.Read())
        mesh.Close()
    for p, n in got:
        assert same_values(p, ref_p) and same_values(n, ref_n)
    assert np.array_equal(got[0][0].view(np.uint32), got[1][0].view(np.uint32))
    assert np.array_equal(got[0][1].view(np.uint32), got[1][1].view(np.uint32))
