"""-m gpu: scene_top_kernel — the tiny leading levels of a hierarchy composed by ONE launch that keeps their world
matrices in shared memory (zn_scene_set_level_fusion, default on) — against the one-launch-per-level form and
the CPU oracle.  Bar: world, MVP, visible list and visibility record bit for bit, fused update and world-only
update, over frames that overlap (no host synchronisation in between)."""
import numpy as np
import pytest

from conftest import same_values
from oracle.binding import BASE_SEED, camera, h8_geo_level_sizes
from test_gpu_cull import read_record

pytestmark = pytest.mark.gpu

SHAPES = {
    # level sizes                          what the top kernel takes
    "bench-like": [7, 56, 448, 3584, 28672],        # tiles 1,1,2,14,...: levels 0..2 (4 tiles)
    "sandbox": [1, 8, 64, 512, 4096],               # tiles 1,1,1,2,16: levels 0..3 (5 tiles)
    "ragged": [7, 61, 500, 4001, 30011],            # tiles 1,1,2,16: levels 0..2, every tile partial
    "whole-scene": [2, 5, 9],                       # the whole scene is the top of the tree
    "two-levels": [300, 700],                       # 2 + 3 tiles: exactly kTopTiles
    "too-wide": [3, 1500],                          # 1 + 6 tiles: nothing qualifies, plain level kernels
    "first-level-wide": [1200, 9000],               # level 0 alone is 5 tiles: a single level is never fused
}


def scene_arrays(oracle, name):
    return oracle.synth_scene(BASE_SEED + 21, SHAPES[name], fanout=8, center_range=0.25)


def run(gpu_ctx, arrays, view, proj, fusion, world_only_first=False):
    import zenyth_b200 as zn
    E = arrays["position"].shape[0]
    sc = zn.Scene(gpu_ctx, E, arrays["level_offsets"])
    sc.SetLevelFusion(fusion)
    sc.Load(arrays)
    if world_only_first:
        sc.UpdateWorld()
        world_only = sc.World().copy()
    sc.SetCamera(zn.Camera(view=view, proj=proj))
    sc.Update()
    out = dict(world=sc.World().copy(), mvp=sc.Mvp().copy(), visible=sc.Visible().copy(), record=read_record(sc))
    if world_only_first:
        out["world_only"] = world_only
    sc.Close()
    return out


@pytest.mark.parametrize("name", sorted(SHAPES))
def test_fused_top_levels_equal_per_level_launches_and_oracle(oracle, gpu_ctx, name):
    arrays = scene_arrays(oracle, name)
    for eye in [(0.0, 0.0, 1500.0), (0.0, 0.0, 0.0)]:
        view, proj = camera(oracle, eye=eye)
        ref = oracle.scene_update(arrays, view, proj, threads=4)
        fused = run(gpu_ctx, arrays, view, proj, True, world_only_first=True)
        plain = run(gpu_ctx, arrays, view, proj, False, world_only_first=True)
        for key in ("world", "mvp", "world_only"):
            assert np.array_equal(fused[key].view(np.uint32), plain[key].view(np.uint32)), (name, key)
        assert np.array_equal(fused["visible"], plain["visible"]) and np.array_equal(fused["record"], plain["record"])
        assert same_values(fused["world"], ref["world"]) and same_values(fused["world_only"], ref["world"])
        assert same_values(fused["mvp"], ref["mvp"])
        assert np.array_equal(fused["visible"], ref["visible"])


def test_parents_in_any_earlier_level_and_roots_below_level_zero(oracle, gpu_ctx):
    """The upload contract only asks that a parent lives in an EARLIER level: entities of level 2 and 3 hang off
    level 0 / level 1 here, and some entities of deeper levels are roots."""
    arrays = scene_arrays(oracle, "sandbox")
    off = arrays["level_offsets"]
    parent = arrays["parent"].copy()
    # the first entities of a level have the smallest parents: re-hanging them further up keeps the order by parent
    parent[off[2]: off[2] + 5] = 0                    # level 2 -> the root
    parent[off[3]: off[3] + 40] = np.repeat(np.arange(off[1], off[1] + 8, dtype=np.uint32), 5)   # level 3 -> level 1
    parent[off[3] + 40: off[3] + 43] = np.uint32(0xFFFFFFFF)                                     # roots in level 3
    parent[off[4]: off[4] + 3] = np.uint32(0xFFFFFFFF)                                           # ... and below the fused run
    arrays["parent"] = parent
    view, proj = camera(oracle)
    ref = oracle.scene_update(arrays, view, proj, threads=4)
    fused = run(gpu_ctx, arrays, view, proj, True)
    plain = run(gpu_ctx, arrays, view, proj, False)
    for key in ("world", "mvp"):
        assert np.array_equal(fused[key].view(np.uint32), plain[key].view(np.uint32)), key
        assert same_values(fused[key], ref[key]), key
    assert np.array_equal(fused["visible"], ref["visible"]) and np.array_equal(plain["visible"], ref["visible"])


def test_overlapping_frames_with_moving_roots(oracle, gpu_ctx):
    """Frames back to back without a host synchronisation: the top kernel of frame k+1 runs beside frame k's emit.
    Roots move between frames (a device-side upload makes that frame's first launch a normal one)."""
    import torch
    import zenyth_b200 as zn
    arrays = oracle.synth_scene(BASE_SEED + 3, h8_geo_level_sizes(7, 6), fanout=8, center_range=0.25)
    E = arrays["position"].shape[0]
    sc = zn.Scene(gpu_ctx, E, arrays["level_offsets"])
    sc.Load(arrays)
    frames = 9
    lists = [torch.full((E,), -1, dtype=torch.int32, device="cuda") for _ in range(frames)]
    counts = [torch.zeros(1, dtype=torch.int32, device="cuda") for _ in range(frames)]
    eyes = [(25.0 * k, -10.0 * k, 2400.0 - 250.0 * k) for k in range(frames)]
    pos = arrays["position"].copy()
    moved = []
    torch.cuda.synchronize()
    view, proj = camera(oracle, eye=eyes[0])
    sc.SetCamera(zn.Camera(view=view, proj=proj))
    sc.Update()
    gpu_ctx.Sync()
    for k in range(frames):
        if k % 3 == 1:                                  # every third frame the roots move
            pos = pos.copy()
            pos[:7] += np.float32([1.5 * k, 0.0, -2.0])
            sc.SetTransforms(position=pos[:7], first=0, count=7)
        moved.append(pos)
        view, proj = camera(oracle, eye=eyes[k])
        sc.SetCamera(zn.Camera(view=view, proj=proj))
        sc.BindVisibleOutput(lists[k].data_ptr(), E, counts[k].data_ptr())
        sc.Update()
        sc.BindVisibleOutput(None)
    gpu_ctx.Sync()
    for k in range(frames):
        view, proj = camera(oracle, eye=eyes[k])
        ref = oracle.scene_update(dict(arrays, position=moved[k]), view, proj, threads=4)
        n = int(counts[k].item())
        assert n == ref["visible"].size, k
        assert np.array_equal(lists[k][:n].cpu().numpy().astype(np.uint32), ref["visible"]), k
    assert same_values(sc.World(), ref["world"]) and same_values(sc.Mvp(), ref["mvp"])
    sc.Close()
