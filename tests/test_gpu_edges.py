"""-m gpu: edge cases of the CUDA scene / mesh path against the oracle: deep and tiny hierarchies,
alternate projections, empty and full visible sets, partial uploads across level boundaries,
degenerate scales, large angles and non-finite inputs."""
import numpy as np
import pytest

from conftest import same_values
from oracle.binding import BASE_SEED, camera, h8_geo_level_sizes

pytestmark = pytest.mark.gpu


def gpu_frame(gpu_ctx, arrays, view, proj):
    import zenyth_b200 as zn
    sc = zn.Scene(gpu_ctx, arrays["position"].shape[0], arrays.get("level_offsets"))
    sc.Load(arrays)
    sc.SetCamera(zn.Camera(view=view, proj=proj))
    sc.Update()
    out = (sc.World(), sc.Mvp(), sc.Visible())
    sc.Close()
    return out


def check(oracle, gpu_ctx, arrays, view, proj):
    ref = oracle.scene_update(arrays, view, proj)
    w, m, v = gpu_frame(gpu_ctx, arrays, view, proj)
    assert same_values(w, ref["world"]) and same_values(m, ref["mvp"])
    assert np.array_equal(v, ref["visible"])
    return ref


def test_forty_levels_of_three(oracle, gpu_ctx):
    arrays = oracle.synth_scene(BASE_SEED + 7, [3] * 40, fanout=1, center_range=0.2)
    view, proj = camera(oracle, eye=(0, 0, 2500))
    check(oracle, gpu_ctx, arrays, view, proj)


def test_single_entity_levels_and_odd_sizes(oracle, gpu_ctx):
    arrays = oracle.synth_scene(BASE_SEED + 8, [1, 1, 5, 257, 1, 511, 2], fanout=3, center_range=0.2)
    view, proj = camera(oracle, eye=(0, 0, 2500))
    check(oracle, gpu_ctx, arrays, view, proj)


@pytest.mark.parametrize("kind", ["orthographic", "reverse_z", "narrow"])
def test_alternate_projections(oracle, gpu_ctx, kind):
    arrays = oracle.synth_scene(BASE_SEED + 9, [7, 56, 448, 3584], fanout=8, center_range=0.2)
    view = oracle.look_at((200, 300, 1800), arrays["position"][0] if kind == "narrow" else (0, 0, 0), (0, 1, 0))
    proj = {"orthographic": oracle.orthographic(-600, 600, -400, 400, 1.0, 4000.0),
            "reverse_z": oracle.perspective_reverse_z(1.0, 16 / 9, 0.5),
            "narrow": oracle.perspective(0.2, 16 / 9, 10.0, 3000.0)}[kind]
    ref = check(oracle, gpu_ctx, arrays, view, proj)
    assert 0 < ref["visible"].size < arrays["position"].shape[0]


def test_nothing_visible_and_everything_visible(oracle, gpu_ctx):
    arrays = oracle.synth_scene(BASE_SEED + 9, [7, 56, 448], fanout=8)
    away = oracle.look_at((0, 0, 5000), (0, 0, 10000), (0, 1, 0))         # looking away from the scene
    _, proj = camera(oracle)
    ref = check(oracle, gpu_ctx, arrays, away, proj)
    assert ref["visible"].size == 0
    wide = oracle.orthographic(-1e5, 1e5, -1e5, 1e5, -1e5, 1e5)
    ref = check(oracle, gpu_ctx, arrays, oracle.identity(), wide)
    assert ref["visible"].size == arrays["position"].shape[0]


def test_partial_uploads_across_level_boundaries(oracle, gpu_ctx):
    import zenyth_b200 as zn
    arrays = oracle.synth_scene(BASE_SEED + 10, [300, 700, 1500], fanout=3, center_range=0.2)
    view, proj = camera(oracle)
    E = 2500
    sc = zn.Scene(gpu_ctx, E, arrays["level_offsets"])
    sc.Load(arrays)
    sc.SetCamera(zn.Camera(view=view, proj=proj))
    sc.Update()
    rng = np.random.default_rng(5)
    for first, count in [(290, 20), (0, 1), (999, 2), (2499, 1), (250, 1800), (292, 16), (990, 17)]:   # ranges straddling level edges; <= 16 entities travel as kernel parameters
        arrays["position"][first:first + count] += rng.uniform(-5, 5, (count, 3)).astype(np.float32)
        arrays["rotation"][first:first + count] = rng.uniform(-3, 3, (count, 3)).astype(np.float32)
        sc.SetTransforms(arrays["position"][first:first + count], arrays["rotation"][first:first + count], None, first=first)
        arrays["aabb_half"][first:first + count] *= np.float32(1.5)
        sc.SetBounds(None, arrays["aabb_half"][first:first + count], first=first)
        sc.Update()
        ref = oracle.scene_update(arrays, view, proj)
        assert same_values(sc.World(), ref["world"]) and np.array_equal(sc.Visible(), ref["visible"]), (first, count)
    back = sc.Inputs()
    assert np.array_equal(back["position"], arrays["position"]) and np.array_equal(back["parent"], arrays["parent"])
    sc.Close()


def test_handful_of_entities_with_padded_records_and_single_arrays(oracle, gpu_ctx):
    """<= 16 entities take the one-launch path (values as kernel parameters): 16-byte records (the reference's vec3),
    one array at a time, across a level edge."""
    import zenyth_b200 as zn
    arrays = oracle.synth_scene(BASE_SEED + 12, [300, 700, 1500], fanout=3, center_range=0.2)
    view, proj = camera(oracle)
    sc = zn.Scene(gpu_ctx, 2500, arrays["level_offsets"])
    sc.Load(arrays)
    sc.SetCamera(zn.Camera(view=view, proj=proj))
    rng = np.random.default_rng(8)

    def pad(a):
        out = np.full((a.shape[0], 4), np.float32(7.0), np.float32)     # the w lane is never read
        out[:, :3] = a
        return out
    for key, first, count in [("scale", 296, 9), ("rotation", 0, 16), ("position", 995, 12), ("scale", 2499, 1)]:
        lo, hi = {"scale": (0.5, 2.0), "rotation": (-3.0, 3.0), "position": (-50.0, 50.0)}[key]
        arrays[key][first:first + count] = rng.uniform(lo, hi, (count, 3)).astype(np.float32)
        sc.SetTransforms(**{key: pad(arrays[key][first:first + count])}, first=first, count=count)
        sc.Update()
        ref = oracle.scene_update(arrays, view, proj)
        assert same_values(sc.World(), ref["world"]) and np.array_equal(sc.Visible(), ref["visible"]), (key, first, count)
    back = sc.Inputs()
    for key in ("position", "rotation", "scale"):
        assert np.array_equal(back[key], arrays[key]), key
    sc.Close()


def test_large_angles_zero_and_negative_scale(oracle, gpu_ctx):
    arrays = oracle.synth_scene(BASE_SEED + 11, [64, 512], fanout=8, center_range=0.2)
    rng = np.random.default_rng(6)
    arrays["rotation"][:] = rng.uniform(-1.0e4, 1.0e4, arrays["rotation"].shape).astype(np.float32)
    arrays["scale"][5] = (0.0, 1.0, 1.0)
    arrays["scale"][70] = (-1.5, 0.7, 2.0)
    arrays["scale"][71] = (0.0, 0.0, 0.0)
    view, proj = camera(oracle)
    check(oracle, gpu_ctx, arrays, view, proj)


def test_non_finite_inputs_stay_local(oracle, gpu_ctx):
    """A NaN / Inf entity (and its subtree) is invisible on both sides; everything else is untouched."""
    arrays = oracle.synth_scene(BASE_SEED + 12, [8, 64, 512], fanout=8)
    arrays["position"][3, 1] = np.nan
    arrays["scale"][20, 0] = np.inf
    arrays["rotation"][100, 2] = np.nan
    view, proj = camera(oracle)
    ref = oracle.scene_update(arrays, view, proj)
    w, m, v = gpu_frame(gpu_ctx, arrays, view, proj)
    assert np.array_equal(v, ref["visible"])
    assert 3 not in v and 20 not in v and 100 not in v
    finite = np.isfinite(ref["world"]).all(axis=1)
    assert np.array_equal(np.isfinite(w).all(axis=1), finite)
    assert same_values(w[finite], ref["world"][finite]) and same_values(m[finite], ref["mvp"][finite])


def test_mesh_singular_and_mirrored_models(oracle, gpu_ctx):
    import zenyth_b200 as zn
    models = np.stack([oracle.local_matrix((1, 2, 3), (0.3, 0.2, 0.1), s) for s in [(1, 1, 1), (0, 1, 1), (-1, 2, 0.5), (0, 0, 0), (3, 3, 3)]])
    verts = 100
    pos, nrm = oracle.synth_mesh(BASE_SEED + 13, 5 * verts)
    first = (np.arange(6) * verts).astype(np.uint32)
    ref_p, ref_n = oracle.mesh_transform(first, models, pos, nrm)
    mesh = zn.Mesh(gpu_ctx, 5 * verts, 5)
    mesh.SetVertices(pos, nrm)
    mesh.SetModels(models)
    mesh.Transform()
    p, n = mesh.Read()
    mesh.Close()
    assert same_values(p, ref_p) and same_values(n, ref_n)
    assert np.all(n[verts:2 * verts] == 0) and np.all(n[3 * verts:4 * verts] == 0)   # singular model: zero normal matrix, left as is


def test_many_frames_no_races(oracle, gpu_ctx):
    """compute-sanitizer is closed on this pool, so hammer the asynchronous machinery instead:
    300 back-to-back frames (PDL-chained level launches, ticketed persistent CTAs, TMA pipelines,
    parity-buffered compaction) with a camera that changes every frame, every 10th frame checked
    against the oracle in full and every frame's count checked."""
    import zenyth_b200 as zn
    sizes = [7, 56, 448, 3584, 28672, 229376]
    E = sum(sizes)
    arrays = oracle.synth_scene(BASE_SEED + 15, sizes, fanout=8, center_range=0.25)
    sc = zn.Scene(gpu_ctx, E, arrays["level_offsets"])
    sc.Load(arrays)
    _, proj = camera(oracle)
    views = [oracle.look_at((300.0 * np.sin(0.05 * f), 100.0, 1500.0 - 4.0 * f), (0, 0, 0), (0, 1, 0)) for f in range(300)]
    counts = []
    for f, view in enumerate(views):
        sc.SetCamera(zn.Camera(view=view, proj=proj))
        sc.Update()
        if f % 10 == 0:
            ref = oracle.scene_update(arrays, view, proj, threads=oracle.hardware_threads())
            assert np.array_equal(sc.Visible(), ref["visible"]), f
            assert same_values(sc.Mvp(), ref["mvp"]) and same_values(sc.World(), ref["world"]), f
        counts.append(sc.VisibleCount())
    # counts of unchecked frames: recompute a sample on the CPU
    for f in (3, 77, 151, 299):
        assert counts[f] == oracle.scene_update(arrays, views[f], proj, threads=oracle.hardware_threads(), want_mvp=False)["visible"].size, f
    sc.Close()


def test_tile_tickets_wrap_around_2_to_the_32(oracle, gpu_ctx):
    """The persistent kernels' ticket counters are 32-bit and only ever grow (level 7 of the 16M scene
    wraps every ~74k frames).  Start them just below 2^32 and run across the wrap: update, cull and the
    record expansion must keep producing the oracle's results on every frame."""
    import torch
    import zenyth_b200 as zn
    sizes = h8_geo_level_sizes(7, 5)                    # 4,681 entities: 40 tiles, ~100 tickets per frame and level
    arrays = oracle.synth_scene(BASE_SEED + 3, sizes, fanout=8, center_range=0.25)
    E = sum(sizes)
    view, proj = camera(oracle)
    ref = oracle.scene_update(arrays, view, proj)
    sc = zn.Scene(gpu_ctx, E, arrays["level_offsets"])
    sc.Load(arrays)
    cam = zn.Camera(view=view, proj=proj)
    sc.SetCamera(cam)
    words = sc.DeviceBuffers().visibility_record_words
    lists = torch.zeros(2 * E, dtype=torch.int32, device="cuda")
    counts = torch.zeros(2, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    sc.SetTicketOrigin(0xFFFFFFFF - 150)                # every counter crosses 2^32 within the first frames
    for frame in range(40):
        if frame % 2:
            sc.UpdateWorld()
            sc.Cull(cam)
        else:
            sc.Update()
        rec = sc.DeviceBuffers().visibility_record
        sc.SetExpandMode(frame % 3 == 2)                    # both expansion strategies
        sc.ExpandVisible(rec, 1, words, lists.data_ptr(), E, counts.data_ptr(), index_bases=[9])
        assert np.array_equal(sc.Visible(), ref["visible"]), frame
        assert counts[0].item() == ref["visible"].size
        assert np.array_equal(lists[:ref["visible"].size].cpu().numpy().astype(np.uint32), ref["visible"] + 9), frame
    assert same_values(sc.World(), ref["world"]) and same_values(sc.Mvp(), ref["mvp"])
    sc.Close()
