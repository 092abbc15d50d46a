"""-m gpu: the CUDA scene path, through the C ABI, against the CPU oracle on the same inputs."""
import numpy as np
import pytest

from conftest import rel_err, same_values
from oracle.binding import BASE_SEED, camera, h8_geo_level_sizes

pytestmark = pytest.mark.gpu

TOL = 1e-5   # north star: world/MVP within 1e-5 relative; the design target is exact equality


def run_gpu(gpu_ctx, arrays, view, proj, index_base=0):
    import zenyth_b200 as zn
    sc = zn.Scene(gpu_ctx, arrays["position"].shape[0], arrays.get("level_offsets"))
    sc.Load(arrays)
    sc.SetCamera(zn.Camera(view=view, proj=proj))
    if index_base:
        sc.SetIndexBase(index_base)
    sc.Update()
    out = {"world": sc.World(), "mvp": sc.Mvp(), "visible": sc.Visible(), "count": sc.VisibleCount()}
    sc.Close()
    return out


def check_against_oracle(oracle, gpu_ctx, arrays, view, proj, exact=True):
    ref = oracle.scene_update(arrays, view, proj, threads=4)
    got = run_gpu(gpu_ctx, arrays, view, proj)
    assert rel_err(got["world"], ref["world"]) <= TOL
    assert rel_err(got["mvp"], ref["mvp"]) <= TOL
    if exact:
        # same operation order on both sides: equal to the last bit (modulo the sign of zero)
        assert same_values(got["world"], ref["world"])
        assert same_values(got["mvp"], ref["mvp"])
    assert got["count"] == ref["visible"].size
    assert np.array_equal(got["visible"], ref["visible"])   # bit-exact, ascending
    return ref, got


@pytest.mark.parametrize("n", [1, 31, 32, 255, 256, 257, 1000, 4097])
def test_flat_ragged_sizes(oracle, gpu_ctx, n):
    arrays = oracle.synth_scene(BASE_SEED + 2, [n], fanout=1, center_range=0.5)
    arrays["parent"] = None
    arrays["level_offsets"] = None
    view, proj = camera(oracle)
    ref, _ = check_against_oracle(oracle, gpu_ctx, arrays, view, proj)


def test_flat_64k_three_cameras(oracle, gpu_ctx):
    arrays = oracle.synth_scene(BASE_SEED + 2, [65536], fanout=1, center_range=0.0)
    fractions = []
    for eye in [(0, 0, 0), (0, 0, 1500), (0, 0, 3000)]:
        view, proj = camera(oracle, eye=eye)
        ref, _ = check_against_oracle(oracle, gpu_ctx, arrays, view, proj)
        fractions.append(ref["visible"].size / 65536)
    assert fractions[0] < fractions[1] < fractions[2]


def test_hierarchy_h8_geo_5_levels(oracle, gpu_ctx):
    arrays = oracle.synth_scene(BASE_SEED + 3, h8_geo_level_sizes(7, 5), fanout=8, center_range=0.25)
    view, proj = camera(oracle)
    check_against_oracle(oracle, gpu_ctx, arrays, view, proj)


def test_hierarchy_chain_fanout1(oracle, gpu_ctx):
    arrays = oracle.synth_scene(BASE_SEED + 3, [3000] * 6, fanout=1, center_range=0.25)
    view, proj = camera(oracle)
    check_against_oracle(oracle, gpu_ctx, arrays, view, proj)


def test_hierarchy_scattered_parents_take_the_gather_path(oracle, gpu_ctx):
    # children are NOT sorted by parent: parent ranges per tile exceed the staging window
    rng = np.random.default_rng(7)
    sizes = [2000, 5000, 5000]
    arrays = oracle.synth_scene(BASE_SEED + 4, sizes, fanout=2, center_range=0.25)
    off = arrays["level_offsets"]
    for l in range(1, 3):
        arrays["parent"][off[l]:off[l + 1]] = rng.integers(off[l - 1], off[l], sizes[l], dtype=np.uint32)
    # a few roots inside deeper levels
    arrays["parent"][off[1] + 5] = 0xFFFFFFFF
    arrays["parent"][off[2] + 77] = 0xFFFFFFFF
    view, proj = camera(oracle)
    check_against_oracle(oracle, gpu_ctx, arrays, view, proj)


def test_matches_float64_shadow_outside_boundary_set(oracle, gpu_ctx):
    arrays = oracle.synth_scene(BASE_SEED + 3, h8_geo_level_sizes(7, 5), fanout=8, center_range=0.25)
    view, proj = camera(oracle)
    got = run_gpu(gpu_ctx, arrays, view, proj)
    d = oracle.scene_update64(arrays, view, proj, eps=1e-5)
    assert rel_err(got["world"], d["world"]) <= TOL
    assert rel_err(got["mvp"], d["mvp"]) <= TOL
    flags = np.zeros(arrays["position"].shape[0], np.uint8)
    flags[got["visible"]] = 1
    mismatch = flags != d["visible"]
    assert not np.any(mismatch & (d["boundary"] == 0))


def test_index_base_and_rerun_idempotent(oracle, gpu_ctx):
    import zenyth_b200 as zn
    arrays = oracle.synth_scene(BASE_SEED + 5, [300, 2400], fanout=8)
    view, proj = camera(oracle)
    ref = oracle.scene_update(arrays, view, proj)
    sc = zn.Scene(gpu_ctx, 2700, arrays["level_offsets"])
    sc.Load(arrays)
    sc.SetCamera(zn.Camera(view=view, proj=proj))
    sc.SetIndexBase(1000)
    sc.Update()
    a = sc.Visible()
    sc.Update()
    b = sc.Visible()
    assert np.array_equal(a, b)
    assert np.array_equal(a, ref["visible"] + 1000)
    sc.Close()


def test_vec3_stride16_upload(oracle, gpu_ctx):
    import zenyth_b200 as zn
    arrays = oracle.synth_scene(BASE_SEED + 6, [777], fanout=1)
    view, proj = camera(oracle)
    ref = oracle.scene_update(arrays, view, proj)
    pad = lambda a: np.concatenate([a, np.full((a.shape[0], 1), 123.0, np.float32)], axis=1)
    sc = zn.Scene(gpu_ctx, 777)
    sc.SetTransforms(pad(arrays["position"]), pad(arrays["rotation"]), pad(arrays["scale"]))
    sc.SetBounds(pad(arrays["aabb_center"]), pad(arrays["aabb_half"]))
    sc.SetCamera(zn.Camera(view=view, proj=proj))
    sc.Update()
    assert same_values(sc.World(), ref["world"])
    assert np.array_equal(sc.Visible(), ref["visible"])
    back = sc.Inputs()
    assert np.array_equal(back["position"], arrays["position"])
    assert np.array_equal(back["aabb_half"], arrays["aabb_half"])
    sc.Close()


def test_device_synth_matches_host_generator(oracle, gpu_ctx):
    import zenyth_b200 as zn
    sizes = h8_geo_level_sizes(7, 4)
    arrays = oracle.synth_scene(BASE_SEED + 3, sizes, fanout=8, center_range=0.25)
    sc = zn.Scene(gpu_ctx, sum(sizes), arrays["level_offsets"])
    sc.FillSynthetic(BASE_SEED + 3, 8, 0.25)
    back = sc.Inputs()
    for k in ("position", "rotation", "scale", "aabb_center", "aabb_half"):
        assert np.array_equal(back[k].view(np.uint32), arrays[k].view(np.uint32)), k
    assert np.array_equal(back["parent"], arrays["parent"])
    sc.Close()


def test_errors_are_loud(gpu_ctx):
    import zenyth_b200 as zn
    sc = zn.Scene(gpu_ctx, 16)
    with pytest.raises(zn.ZenythError):
        sc.Cull(zn.Camera())   # no world matrices yet
    with pytest.raises(zn.ZenythError):
        sc.SetParents(np.array([3], np.uint32), first=0)   # parent not in an earlier level
    with pytest.raises(zn.ZenythError):
        sc.World(first=10, count=10)
    sc.Close()


class _DevArray:
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i4", "data": (int(ptr), False), "version": 3}


def test_visibility_record_expands_to_the_same_list(oracle, gpu_ctx):
    """The 1-bit-per-entity record that multi-GPU runs exchange reproduces the compacted list."""
    import torch
    import zenyth_b200 as zn
    sizes = h8_geo_level_sizes(7, 5)
    E = sum(sizes)
    arrays = oracle.synth_scene(BASE_SEED + 3, sizes, fanout=8, center_range=0.25)
    view, proj = camera(oracle)
    sc = zn.Scene(gpu_ctx, E, arrays["level_offsets"])
    sc.Load(arrays)
    sc.SetCamera(zn.Camera(view=view, proj=proj))
    # the record may live in caller memory (a slot of an all-gather input)
    words = sc.DeviceBuffers().visibility_record_words
    rec = torch.zeros(3 * words, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    sc.BindVisibilityRecord(rec.data_ptr() + 4 * words)          # middle slot
    sc.Update()
    vis = sc.Visible()
    assert np.array_equal(vis, oracle.scene_update(arrays, view, proj)["visible"])
    rec[:words] = rec[words:2 * words]                           # three "ranks" with the same record
    rec[2 * words:] = rec[words:2 * words]
    lists = torch.full((3 * E,), -1, dtype=torch.int32, device="cuda")
    counts = torch.zeros(3, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    sc.ExpandVisible(rec.data_ptr(), 3, words, lists.data_ptr(), E, counts.data_ptr(), index_bases=[0, 1000, 5], skip_index=2)
    gpu_ctx.Sync()
    assert counts.tolist() == [vis.size, vis.size, 0]
    got = lists.view(3, E).cpu().numpy()
    assert np.array_equal(got[0, :vis.size], vis.astype(np.int32))
    assert np.array_equal(got[1, :vis.size], vis.astype(np.int32) + 1000)
    assert np.all(got[2] == -1)                                  # skipped record untouched
    sc.BindVisibilityRecord(None)
    sc.Update()
    assert np.array_equal(sc.Visible(), vis)
    sc.Close()


def test_pipelined_host_frames_match_the_oracle(oracle, gpu_ctx):
    """zn_scene_frame_host_submit/_wait: two frames in flight, each with its own inputs."""
    import zenyth_b200 as zn
    sizes = h8_geo_level_sizes(7, 4)
    E = sum(sizes)
    arrays = oracle.synth_scene(BASE_SEED + 3, sizes, fanout=8, center_range=0.25)
    view, proj = camera(oracle)
    sc = zn.Scene(gpu_ctx, E, arrays["level_offsets"])
    sc.Load(arrays)
    sc.SetCamera(zn.Camera(view=view, proj=proj))
    frames = []
    for f in range(5):   # every frame moves and spins the roots differently
        a = {k: v.copy() for k, v in arrays.items()}
        a["rotation"][:7, 1] += np.float32(0.3 * f)
        a["position"][:7, 0] += np.float32(100.0 * f)
        frames.append(a)
    expect = [oracle.scene_update(a, view, proj)["visible"] for a in frames]
    pinned = [[gpu_ctx.HostAlloc((E, 3), np.float32) for _ in range(3)] for _ in range(2)]
    out = [gpu_ctx.HostAlloc((E,), np.uint32) for _ in range(2)]
    got = []

    def submit(f):
        p, r, s = pinned[f % 2]
        p[:] = frames[f]["position"]; r[:] = frames[f]["rotation"]; s[:] = frames[f]["scale"]
        sc.SubmitFrameHost(p, r, s)

    submit(0)
    submit(1)
    with pytest.raises(zn.ZenythError):
        submit(2)                                   # a third frame in flight is refused
    for f in range(5):
        n = sc.WaitFrameHost(out[f % 2])
        got.append(out[f % 2][:n].copy())
        if f + 2 < 5:
            submit(f + 2)
    with pytest.raises(zn.ZenythError):
        sc.WaitFrameHost(out[0])                    # nothing left in flight
    for f in range(5):
        assert np.array_equal(got[f], expect[f]), f
    assert same_values(sc.World(), oracle.scene_update(frames[4], view, proj)["world"])
    sc.Close()


def test_build_draw_packs_visible_mvps_and_indirect_args(oracle, gpu_ctx):
    """N3: visible list -> per-instance MVP buffer + D3D12_DRAW_INDEXED_ARGUMENTS, on the device."""
    import torch
    import zenyth_b200 as zn
    sizes = h8_geo_level_sizes(7, 5)
    E = sum(sizes)
    arrays = oracle.synth_scene(BASE_SEED + 3, sizes, fanout=8, center_range=0.25)
    for eye, base in (((0, 0, 1500), 0), ((0, 0, 0), 4096)):
        view, proj = camera(oracle, eye=eye)
        ref = oracle.scene_update(arrays, view, proj)
        sc = zn.Scene(gpu_ctx, E, arrays["level_offsets"])
        sc.Load(arrays)
        sc.SetCamera(zn.Camera(view=view, proj=proj))
        sc.SetIndexBase(base)
        sc.Update()
        inst = torch.full((E, 16), float("nan"), dtype=torch.float32, device="cuda")
        args = torch.zeros(5, dtype=torch.int32, device="cuda")
        torch.cuda.synchronize()
        sc.BuildDraw(36, inst.data_ptr(), E, args.data_ptr())
        gpu_ctx.Sync()
        n = ref["visible"].size
        assert args.tolist() == [36, n, 0, 0, 0]
        got = inst.cpu().numpy()
        assert same_values(got[:n], ref["mvp"][ref["visible"]])          # packed in list order
        assert np.all(np.isnan(got[n:]))                                 # nothing written past the count
        sc.Close()


def test_frame_host_with_pageable_buffers_and_matrix_download(oracle, gpu_ctx):
    """zn_scene_frame_host: plain numpy (pageable) arrays in, world + MVP + visible list out."""
    import zenyth_b200 as zn
    sizes = h8_geo_level_sizes(7, 4)
    E = sum(sizes)
    arrays = oracle.synth_scene(BASE_SEED + 3, sizes, fanout=8, center_range=0.25)
    view, proj = camera(oracle)
    sc = zn.Scene(gpu_ctx, E, arrays["level_offsets"])
    sc.SetBounds(arrays["aabb_center"], arrays["aabb_half"])
    sc.SetParents(arrays["parent"])
    sc.SetCamera(zn.Camera(view=view, proj=proj))
    world = np.empty((E, 16), np.float32)
    mvp = np.empty((E, 16), np.float32)
    vis = np.empty(E, np.uint32)
    n = sc.FrameHost(arrays["position"], arrays["rotation"], arrays["scale"], world=world, mvp=mvp, visible=vis)
    ref = oracle.scene_update(arrays, view, proj)
    assert n == ref["visible"].size and np.array_equal(vis[:n], ref["visible"])
    assert same_values(world, ref["world"]) and same_values(mvp, ref["mvp"])
    # inputs omitted: the resident ones are reused; a single-entity edit goes through SetTransform
    t = zn.Transform(position=(5.0, 6.0, 7.0), rotation=(0.1, 0.2, 0.3), scale=(1.0, 2.0, 3.0))
    sc.SetTransform(9, t)
    arrays["position"][9], arrays["rotation"][9], arrays["scale"][9] = t.position, t.rotation, t.scale
    n = sc.FrameHost(visible=vis)
    ref = oracle.scene_update(arrays, view, proj)
    assert n == ref["visible"].size and np.array_equal(vis[:n], ref["visible"]) and same_values(sc.World(), ref["world"])
    sc.Close()


def test_scene_file_round_trip(oracle, gpu_ctx, tmp_path):
    """N4: the ZNSCENE1 dump — a file written with numpy loads into a scene that matches the oracle,
    and saving that scene reproduces the file byte for byte."""
    import zenyth_b200 as zn
    sizes = [5, 300, 1000]
    arrays = oracle.synth_scene(BASE_SEED + 14, sizes, fanout=4, center_range=0.25)
    path = tmp_path / "scene.znscene"
    with open(path, "wb") as f:
        f.write(b"ZNSCENE1")
        f.write(np.asarray([sum(sizes), len(sizes)], np.uint32).tobytes())
        f.write(arrays["level_offsets"].astype(np.uint32).tobytes())
        for k in ("position", "rotation", "scale", "aabb_center", "aabb_half"):
            f.write(arrays[k].astype(np.float32).tobytes())
        f.write(arrays["parent"].astype(np.uint32).tobytes())
    sc = zn.Scene.LoadFile(gpu_ctx, path)
    assert sc.entity_count == sum(sizes)
    view, proj = camera(oracle)
    sc.SetCamera(zn.Camera(view=view, proj=proj))
    sc.Update()
    ref = oracle.scene_update(arrays, view, proj)
    assert same_values(sc.World(), ref["world"]) and np.array_equal(sc.Visible(), ref["visible"])
    out = tmp_path / "again.znscene"
    sc.SaveFile(out)
    assert out.read_bytes() == path.read_bytes()
    sc.Close()
    bad = tmp_path / "bad.znscene"
    bad.write_bytes(b"NOTASCENE" + bytes(64))
    with pytest.raises(zn.ZenythError):
        zn.Scene.LoadFile(gpu_ctx, bad)
    with pytest.raises(zn.ZenythError):
        zn.Scene.LoadFile(gpu_ctx, tmp_path / "missing.znscene")


@pytest.mark.parametrize("world", [2, 5])
def test_one_scene_sharded_by_root_subtree(oracle, gpu_ctx, world):
    """SURVEY §8e, 'within one scene': hierarchy-closed shards (zenyth_b200.distributed.shard_scene_by_roots),
    each run as a scene of its own, reproduce the whole scene entity for entity and, merged, its visible list."""
    from zenyth_b200.distributed import merge_visible, shard_scene_by_roots
    sizes = h8_geo_level_sizes(7, 5)
    arrays = oracle.synth_scene(BASE_SEED + 3, sizes, fanout=8, center_range=0.25)
    arrays["parent"][sizes[0] + 5] = 0xFFFFFFFF             # a root below level 0, with its own subtree
    view, proj = camera(oracle)
    ref = oracle.scene_update(arrays, view, proj, threads=4)
    lists = []
    for shard in shard_scene_by_roots(arrays, world):
        got = run_gpu(gpu_ctx, shard.arrays, view, proj)
        assert same_values(got["world"], ref["world"][shard.global_index])
        assert same_values(got["mvp"], ref["mvp"][shard.global_index])
        lists.append(shard.to_global(got["visible"]))
    assert np.array_equal(merge_visible(lists), ref["visible"])


def test_scene_given_in_creation_order(oracle, gpu_ctx):
    """An engine's entities in arbitrary order: zn_hierarchy_sort lays them out, the GPU scene then
    matches the oracle entity for entity (mapped back through `order`)."""
    import zenyth_b200 as zn
    sizes = h8_geo_level_sizes(7, 4)
    arrays = oracle.synth_scene(BASE_SEED + 17, sizes, fanout=8, center_range=0.25)
    view, proj = camera(oracle)
    ref = oracle.scene_update(arrays, view, proj, threads=4)
    rng = np.random.default_rng(5)
    n = sum(sizes)
    new_of_old = rng.permutation(n).astype(np.uint32)            # shuffle into "creation order"
    old_of_new = np.argsort(new_of_old).astype(np.uint32)
    p = arrays["parent"][old_of_new]
    created_parent = np.where(p == 0xFFFFFFFF, 0xFFFFFFFF, new_of_old[np.minimum(p, n - 1)]).astype(np.uint32)
    order, new_parent, offsets = zn.SortHierarchy(created_parent)
    assert np.array_equal(np.diff(offsets), sizes)
    pick = old_of_new[order]                                      # index in `arrays` of every sorted entity
    laid_out = {k: np.ascontiguousarray(arrays[k][pick]) for k in ("position", "rotation", "scale", "aabb_center", "aabb_half")}
    laid_out.update(parent=new_parent, level_offsets=offsets)
    got = run_gpu(gpu_ctx, laid_out, view, proj)
    assert same_values(got["world"], ref["world"][pick]) and same_values(got["mvp"], ref["mvp"][pick])
    assert np.array_equal(np.sort(pick[got["visible"]]), ref["visible"])


def test_host_frames_with_ranges_and_absent_arrays(oracle, gpu_ctx):
    """zn_frame_host_io takes NULL arrays (keep what is resident) and a (first, count) range: a frame that only
    spins some entities uploads 12 bytes for each of them and nothing else.  Pipelined and one-at-a-time forms,
    ranges that straddle hierarchy levels, results against the oracle on the edited inputs."""
    import zenyth_b200 as zn
    sizes = h8_geo_level_sizes(7, 5)
    E = sum(sizes)
    arrays = oracle.synth_scene(BASE_SEED + 3, sizes, fanout=8, center_range=0.25)
    view, proj = camera(oracle)
    sc = zn.Scene(gpu_ctx, E, arrays["level_offsets"])
    sc.Load(arrays)
    sc.SetCamera(zn.Camera(view=view, proj=proj))
    cur = {k: v.copy() for k, v in arrays.items()}
    vis = gpu_ctx.HostAlloc((E,), np.uint32)
    rng = np.random.default_rng(5)
    # (1) pipelined: rotation only, ranges that start inside one level and end inside another
    edits = [(0, 7), (3, 500), (60, 4000), (4000, E - 4000), (0, E)]
    bufs = [gpu_ctx.HostAlloc((E, 3), np.float32) for _ in range(2)]
    expect = []
    for f, (first, count) in enumerate(edits):
        cur["rotation"][first:first + count] = rng.uniform(-3.1, 3.1, (count, 3)).astype(np.float32)
        expect.append(oracle.scene_update(cur, view, proj)["visible"].copy())
        if f >= 2:                                              # frame f-2 used the buffer frame f is about to fill
            n = sc.WaitFrameHost(vis)
            assert np.array_equal(vis[:n], expect[f - 2]), f - 2
        b = bufs[f % 2]
        b[:count] = cur["rotation"][first:first + count]
        sc.SubmitFrameHost(rotation=b[:count], first=first, count=count)
    for f in (len(edits) - 2, len(edits) - 1):
        n = sc.WaitFrameHost(vis)
        assert np.array_equal(vis[:n], expect[f]), f
    assert same_values(sc.World(), oracle.scene_update(cur, view, proj)["world"])
    # (2) one frame at a time: position only for a range, then nothing at all (a pure re-run)
    first, count = 50, 3000
    cur["position"][first:first + count] += np.float32([5.0, -3.0, 1.0])
    ref = oracle.scene_update(cur, view, proj)
    n = sc.FrameHost(position=np.ascontiguousarray(cur["position"][first:first + count]), visible=vis, first=first, count=count)
    assert np.array_equal(vis[:n], ref["visible"]) and same_values(sc.World(), ref["world"])
    sc.SubmitFrameHost()                                        # no array at all: just the frame
    n = sc.WaitFrameHost(vis)
    assert np.array_equal(vis[:n], ref["visible"])
    back = sc.Inputs()
    for k in ("position", "rotation", "scale"):
        assert np.array_equal(back[k], cur[k]), k
    with pytest.raises(zn.ZenythError):
        sc.SubmitFrameHost(rotation=bufs[0][:10], first=E - 5, count=10)   # range past the end
    sc.Close()
