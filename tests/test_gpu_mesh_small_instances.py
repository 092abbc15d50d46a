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
