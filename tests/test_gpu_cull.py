"""-m gpu: Scene::Update(dt) and Scene::Cull(camera) as separate calls (zn_scene_update_world /
zn_scene_cull), the shape the reference's frame loop has room for — OnUpdate at
Core/src/Application.cpp:47, OnRender between BeginFrame/EndFrame at :49-51.

Parity bar: a cull pass over existing world matrices equals the fused zn_scene_update with the same
camera BIT FOR BIT (MVPs, list, count, visibility record), and both equal the CPU oracle."""
import numpy as np
import pytest

from conftest import same_values
from oracle.binding import BASE_SEED, camera, h8_geo_level_sizes

pytestmark = pytest.mark.gpu

CAMERAS = [(0.0, 0.0, 0.0), (0.0, 0.0, 1500.0), (0.0, 0.0, 3000.0), (400.0, -300.0, 900.0)]


class _Dev:
    def __init__(self, ptr, n, typestr="<u4"):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (int(ptr), False), "version": 3}


def read_record(sc):
    import torch
    b = sc.DeviceBuffers()
    sc.ctx.Sync()
    return torch.as_tensor(_Dev(b.visibility_record, b.visibility_record_words, "<i4"), device="cuda").cpu().numpy().copy()


def ragged_scene(oracle):
    """Level sizes that are not multiples of the 256-entity tile: every level ends in a partial tile."""
    sizes = [7, 61, 500, 4001, 30011]
    return oracle.synth_scene(BASE_SEED + 9, sizes, fanout=8, center_range=0.25)


def test_update_world_writes_world_only(oracle, gpu_ctx):
    import zenyth_b200 as zn
    arrays = ragged_scene(oracle)
    view, proj = camera(oracle)
    ref = oracle.scene_update(arrays, view, proj, threads=4)
    E = arrays["position"].shape[0]
    sc = zn.Scene(gpu_ctx, E, arrays["level_offsets"])
    sc.Load(arrays)
    sc.Update()                                   # no camera set: the world-only form
    assert same_values(sc.World(), ref["world"])
    assert sc.VisibleCount() == 0                 # nothing has culled yet
    sc.SetCamera(zn.Camera(view=view, proj=proj))
    sc.Update()                                   # fused
    mvp, vis = sc.Mvp(), sc.Visible()
    assert same_values(mvp, ref["mvp"]) and np.array_equal(vis, ref["visible"])
    # move a root, world-only: world follows, MVP / list stay as the last cull left them
    moved = arrays["position"][:1] + np.float32([3.0, 0.0, 0.0])
    sc.SetTransforms(position=moved, first=0, count=1)
    sc.UpdateWorld()
    arrays2 = dict(arrays, position=arrays["position"].copy())
    arrays2["position"][0] = moved[0]
    ref2 = oracle.scene_update(arrays2, view, proj, threads=4)
    assert same_values(sc.World(), ref2["world"])
    assert same_values(sc.Mvp(), mvp) and np.array_equal(sc.Visible(), vis)
    sc.ClearCamera()
    sc.Update()
    assert same_values(sc.Mvp(), mvp)
    sc.Close()


@pytest.mark.parametrize("shape", ["ragged-hierarchy", "flat-ragged", "h8geo-6-levels"])
def test_cull_equals_fused_update_and_oracle(oracle, gpu_ctx, shape):
    import zenyth_b200 as zn
    if shape == "ragged-hierarchy":
        arrays = ragged_scene(oracle)
    elif shape == "flat-ragged":
        arrays = oracle.synth_scene(BASE_SEED + 2, [70001], fanout=1, center_range=0.5)
        arrays["parent"] = None
        arrays["level_offsets"] = None
    else:
        arrays = oracle.synth_scene(BASE_SEED + 3, h8_geo_level_sizes(7, 6), fanout=8, center_range=0.25)
    E = arrays["position"].shape[0]
    fused = zn.Scene(gpu_ctx, E, arrays.get("level_offsets"))
    fused.Load(arrays)
    split = zn.Scene(gpu_ctx, E, arrays.get("level_offsets"))
    split.Load(arrays)
    split.UpdateWorld()
    seen = set()
    for eye in CAMERAS:
        view, proj = camera(oracle, eye=eye)
        cam = zn.Camera(view=view, proj=proj)
        ref = oracle.scene_update(arrays, view, proj, threads=4)
        fused.SetCamera(cam)
        fused.Update()
        split.Cull(cam)
        assert same_values(split.World(), ref["world"])
        a, b = fused.Mvp(), split.Mvp()
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), eye          # bit for bit
        assert same_values(b, ref["mvp"])
        assert np.array_equal(split.Visible(), ref["visible"]) and np.array_equal(fused.Visible(), ref["visible"])
        assert split.VisibleCount() == ref["visible"].size
        assert np.array_equal(read_record(split), read_record(fused)), eye
        seen.add(ref["visible"].size)
    assert len(seen) >= 3                          # the cameras really see different sets
    fused.Close()
    split.Close()


def test_cull_into_caller_buffers_and_many_cameras_per_frame(oracle, gpu_ctx):
    """Three views of one frame, each into its own device buffers; the scene's own results stay put."""
    import torch
    import zenyth_b200 as zn
    arrays = ragged_scene(oracle)
    E = arrays["position"].shape[0]
    sc = zn.Scene(gpu_ctx, E, arrays["level_offsets"])
    sc.Load(arrays)
    main_view, main_proj = camera(oracle, eye=CAMERAS[1])
    sc.SetCamera(zn.Camera(view=main_view, proj=main_proj))
    sc.Update()
    main = oracle.scene_update(arrays, main_view, main_proj, threads=4)
    outs = []
    for eye in (CAMERAS[0], CAMERAS[2], CAMERAS[3]):
        mvp = torch.full((E, 16), float("nan"), dtype=torch.float32, device="cuda")
        vis = torch.full((E,), -1, dtype=torch.int32, device="cuda")
        cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
        outs.append((eye, mvp, vis, cnt))
    torch.cuda.synchronize()
    for eye, mvp, vis, cnt in outs:
        view, proj = camera(oracle, eye=eye)
        sc.Cull(zn.Camera(view=view, proj=proj), mvp_ptr=mvp.data_ptr(), visible_ptr=vis.data_ptr(), capacity=E, count_ptr=cnt.data_ptr())
    gpu_ctx.Sync()
    for eye, mvp, vis, cnt in outs:
        view, proj = camera(oracle, eye=eye)
        ref = oracle.scene_update(arrays, view, proj, threads=4)
        n = int(cnt.item())
        assert n == ref["visible"].size
        assert np.array_equal(vis[:n].cpu().numpy().astype(np.uint32), ref["visible"])
        assert same_values(mvp.cpu().numpy(), ref["mvp"])
    # the scene's own MVPs and list still belong to the main camera
    assert same_values(sc.Mvp(), main["mvp"]) and np.array_equal(sc.Visible(), main["visible"])
    # ... and a cull with index_base set renumbers like the fused path
    sc.SetIndexBase(5000)
    sc.Cull(zn.Camera(view=main_view, proj=main_proj))
    assert np.array_equal(sc.Visible(), main["visible"] + 5000)
    sc.Close()


def test_cull_after_fused_update_with_another_camera(oracle, gpu_ctx):
    """The camera moved between OnUpdate and OnRender: re-cull without recomposing the hierarchy."""
    import zenyth_b200 as zn
    arrays = oracle.synth_scene(BASE_SEED + 3, h8_geo_level_sizes(7, 5), fanout=8, center_range=0.25)
    E = arrays["position"].shape[0]
    sc = zn.Scene(gpu_ctx, E, arrays["level_offsets"])
    sc.Load(arrays)
    v0, p0 = camera(oracle, eye=CAMERAS[1])
    sc.SetCamera(zn.Camera(view=v0, proj=p0))
    for frame in range(5):
        sc.Update()
        v1, p1 = camera(oracle, eye=(10.0 * frame, 0.0, 1500.0 - 200.0 * frame))
        sc.Cull(zn.Camera(view=v1, proj=p1))
        ref = oracle.scene_update(arrays, v1, p1, threads=4)
        assert np.array_equal(sc.Visible(), ref["visible"]) and same_values(sc.Mvp(), ref["mvp"])
    sc.Close()


def test_cull_profiling_times(oracle, gpu_ctx):
    import zenyth_b200 as zn
    arrays = ragged_scene(oracle)
    sc = zn.Scene(gpu_ctx, arrays["position"].shape[0], arrays["level_offsets"])
    sc.Load(arrays)
    sc.UpdateWorld()
    sc.SetProfiling(True)
    view, proj = camera(oracle)
    sc.Cull(zn.Camera(view=view, proj=proj))
    cull_ms, emit_ms = sc.CullTimesMs()
    assert cull_ms > 0 and emit_ms > 0
    sc.SetProfiling(False)
    sc.Close()


@pytest.mark.parametrize("shape", ["flat", "hierarchy"])
def test_back_to_back_frames_overlap_without_mixing_results(oracle, gpu_ctx, shape):
    """Frames enqueued back to back with NO host synchronisation in between: level 0 of frame k+1 is launched
    programmatically behind frame k's emit kernel and — the record and the chunk totals being double / triple
    buffered — runs beside it.  Every frame uses another camera and writes its list into its own buffer; all
    lists are checked at the end.  Fused updates, world-only updates and cull passes are interleaved."""
    import torch
    import zenyth_b200 as zn
    if shape == "flat":
        arrays = oracle.synth_scene(BASE_SEED + 2, [300_007], fanout=1, center_range=0.5)
        arrays["parent"] = None
        arrays["level_offsets"] = None
    else:
        arrays = oracle.synth_scene(BASE_SEED + 3, h8_geo_level_sizes(7, 6), fanout=8, center_range=0.25)
    E = arrays["position"].shape[0]
    sc = zn.Scene(gpu_ctx, E, arrays.get("level_offsets"))
    sc.Load(arrays)
    frames = 14
    eyes = [(37.0 * k - 200.0, 11.0 * k, 2600.0 - 190.0 * k) for k in range(frames)]
    lists = [torch.full((E,), -1, dtype=torch.int32, device="cuda") for _ in range(frames)]
    counts = [torch.zeros(1, dtype=torch.int32, device="cuda") for _ in range(frames)]
    torch.cuda.synchronize()
    sc.SetCamera(zn.Camera(zn.CameraDesc(eye=eyes[0], center=(eyes[0][0], eyes[0][1], eyes[0][2] - 1.0))))
    sc.Update()
    gpu_ctx.Sync()                                     # from here on nothing synchronises until every frame is enqueued
    for k in range(frames):
        view, proj = camera(oracle, eye=eyes[k])
        cam = zn.Camera(view=view, proj=proj)
        if k % 3 == 2:                                 # the two-call form
            sc.UpdateWorld()
            sc.Cull(cam, visible_ptr=lists[k].data_ptr(), capacity=E, count_ptr=counts[k].data_ptr())
        else:                                          # the fused form, into this frame's own list
            sc.SetCamera(cam)
            sc.BindVisibleOutput(lists[k].data_ptr(), E, counts[k].data_ptr())
            sc.Update()
            sc.BindVisibleOutput(None)
    gpu_ctx.Sync()
    seen = set()
    for k in range(frames):
        view, proj = camera(oracle, eye=eyes[k])
        ref = oracle.scene_update(arrays, view, proj, threads=4)
        n = int(counts[k].item())
        assert n == ref["visible"].size, k
        assert np.array_equal(lists[k][:n].cpu().numpy().astype(np.uint32), ref["visible"]), k
        seen.add(n)
    assert len(seen) >= 6                              # the cameras really see different sets
    view, proj = camera(oracle, eye=eyes[-1])
    ref = oracle.scene_update(arrays, view, proj, threads=4)
    assert same_values(sc.World(), ref["world"]) and same_values(sc.Mvp(), ref["mvp"])
    sc.Close()
