"""Host-side logic of the multi-GPU path on CPU: world_size-2 gloo processes exchange variable-
length visible lists through zenyth_b200.distributed.VisibleGather (the same code bench.py runs
over NCCL), and the scene-sharding helpers are checked exhaustively.  No GPU, no CUDA library."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT


def test_shard_scenes_partitions_exactly():
    from zenyth_b200.distributed import shard_scenes
    for scenes in (1, 3, 8, 13):
        for world in (1, 2, 4, 8):
            owned = [shard_scenes(scenes, world, r) for r in range(world)]
            flat = [s for o in owned for s in o]
            assert flat == list(range(scenes))                       # every scene once, in order
            sizes = [len(o) for o in owned]
            assert max(sizes) - min(sizes) <= 1
    assert [len(shard_scenes(8, w, 0)) for w in (1, 2, 4, 8)] == [8, 4, 2, 1]   # BASELINE configs[4]
    with pytest.raises(ValueError):
        shard_scenes(8, 2, 2)


def test_scene_index_bases():
    from zenyth_b200.distributed import scene_index_bases
    assert scene_index_bases([5, 7, 1]) == [0, 5, 12]
    assert scene_index_bases([33554430] * 8)[-1] == 7 * 33554430
    with pytest.raises(ValueError):
        scene_index_bases([2 ** 31, 2 ** 31, 2 ** 31])


def _rank_list(rank: int, capacity: int):
    """Deterministic ascending 'visible list' of a rank; rank 1 of frame 2 is empty."""
    rng = np.random.default_rng(100 + rank)
    flags = rng.random(capacity) < (0.2 + 0.5 * rank)
    return np.flatnonzero(flags).astype(np.int32) + rank * capacity


def _worker(rank: int, world: int, port: int, capacity: int, out_dir: str):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    from zenyth_b200.distributed import VisibleGather
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = VisibleGather(capacity, world, "cpu", torch.int32)
    ok = True
    for frame in range(3):
        mine = _rank_list(rank, capacity)
        if frame == 2 and rank == 1:
            mine = mine[:0]                                   # an empty list must gather cleanly
        buf = torch.zeros(capacity, dtype=torch.int32)
        buf[: mine.size] = torch.from_numpy(mine)
        counts, lists = g.gather(buf, torch.tensor([mine.size], dtype=torch.int32))
        expect = [_rank_list(r, capacity) for r in range(world)]
        if frame == 2:
            expect[1] = expect[1][:0]
        ok &= counts == [e.size for e in expect]
        for r in range(world):
            ok &= bool(np.array_equal(lists[r, : counts[r]].numpy(), expect[r]))
        merged = VisibleGather.concatenate(counts, lists).numpy()
        ok &= bool(np.array_equal(merged, np.concatenate(expect)))
        ok &= bool(np.all(np.diff(merged) > 0))              # global list ascending
    # the pipelined exchange: two slots, sizes predicted from the last collected frame; frame 3
    # grows past the prediction on rank 0 and must be re-gathered exactly
    from zenyth_b200.distributed import VisibleExchange
    ex = VisibleExchange(capacity, world, "cpu", torch.int32, slots=2, headroom=0.0)
    results = {}

    def frame_list(frame, r):
        base = _rank_list(r, capacity)
        if frame == 3 and r == 0:
            return np.arange(capacity, dtype=np.int32)          # suddenly everything is visible
        if frame == 4 and r == 1:
            return base[:0]
        return base[: max(1, base.size - 10 * frame)]

    for frame in range(6):
        slot = frame % 2
        got = ex.acquire(slot)
        if got is not None:
            results[frame - 2] = [got[1][r, : got[0][r]].numpy().copy() for r in range(world)]
        mine = frame_list(frame, rank)
        ex.local[slot][: mine.size] = torch.from_numpy(mine)
        ex.local_count[slot][0] = mine.size
        ex.launch(slot)
    for frame, got in zip((4, 5), ex.drain()):       # slot 0 holds frame 4, slot 1 frame 5
        results[frame] = [got[1][r, : got[0][r]].numpy().copy() for r in range(world)]
    for frame in range(6):
        for r in range(world):
            ok &= bool(np.array_equal(results[frame][r], frame_list(frame, r)))
    ok &= ex.regathers >= 1
    with open(os.path.join(out_dir, f"rank{rank}.ok"), "w") as f:
        f.write("1" if ok else "0")
    dist.destroy_process_group()


def test_visible_gather_world_size_2_gloo(tmp_path):
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, 50_000, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        assert (tmp_path / f"rank{r}.ok").read_text() == "1"


# ---- PeerRecordExchange: the host protocol of the collective-free exchange, on fakes ----------------
class _FakeCtx:
    """Stands in for GpuContext: 'device memory' is a Python dict keyed by fake addresses shared
    through nothing at all — only the address arithmetic and the call protocol are under test."""

    def __init__(self, rank, fail_export=False, fail_open=False):
        self.rank, self.fail_export, self.fail_open = rank, fail_export, fail_open
        self.opened, self.closed, self.freed = [], [], []

    def DeviceAlloc(self, nbytes):
        self.nbytes = nbytes
        return 0x1000_0000 * (self.rank + 1)

    def DeviceFree(self, ptr):
        self.freed.append(ptr)

    def IpcExport(self, ptr):
        if self.fail_export:
            raise RuntimeError("export refused")
        return ptr.to_bytes(8, "little") + bytes(56)

    def IpcOpen(self, handle):
        if self.fail_open:
            raise RuntimeError("open refused")
        ptr = int.from_bytes(handle[:8], "little") + 0x8_0000_0000      # a peer maps at another address
        self.opened.append(ptr)
        return ptr

    def IpcClose(self, ptr):
        self.closed.append(ptr)


class _FakeScene:
    def __init__(self):
        self.calls = []

    def __getattr__(self, name):
        def record(*a, **k):
            self.calls.append((name, a, k))
        return record


def _peer_worker(rank: int, world: int, port: int, out_dir: str):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    from zenyth_b200.peer_exchange import FLAG_STRIDE_WORDS, PeerRecordExchange
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ok = True
    words, cap = 1000, 5000
    # (1) a rank that cannot export / map: EVERY rank raises, nobody hangs, nothing leaks
    for kw in ({"fail_export": rank == 1}, {"fail_open": rank == 0}):
        ctx = _FakeCtx(rank, **kw)
        try:
            PeerRecordExchange(ctx, words, cap, world, rank, "cpu", torch.int32)
            ok = False
        except RuntimeError as e:
            ok &= "peer memory unavailable" in str(e)
        ok &= sorted(ctx.closed) == sorted(ctx.opened) and len(ctx.freed) <= 1
    # (2) the protocol
    ctx, sc = _FakeCtx(rank), _FakeScene()
    ex = PeerRecordExchange(ctx, words, cap, world, rank, "cpu", torch.int32)
    own, peer = 0x1000_0000 * (rank + 1), 0x1000_0000 * (2 - rank) + 0x8_0000_0000
    slot_bytes = world * words * 4
    flags_off = (4 * slot_bytes + 127) // 128 * 128
    ok &= ctx.nbytes == flags_off + world * FLAG_STRIDE_WORDS * 4
    ok &= ex.flag_ptrs() == [(own if r == rank else peer) + flags_off + rank * 128 for r in range(world)]
    for k in range(6):
        ex.bind(sc, k)
        sc.Update()
        ex.launch(k)
        ok &= ex.expand(sc, k - 1) == (k >= 1)
    ok &= not ex.expand(sc, 4)                       # already expanded
    ok &= ex.expand(sc, 5) and not ex.expand(sc, 6)   # never issued
    names = [c[0] for c in sc.calls]
    ok &= names[:6] == ["BindPeerSignal", "BindVisibilityRecord", "BindPeerRecords", "BindVisibleOutput", "SetSignalValue", "Update"]
    ok &= names.count("BindPeerSignal") == 1 and names.count("ExpandVisibleSignalled") == 7   # 6 frames + the signal-only launch after frame 0
    binds = [c for c in sc.calls if c[0] == "BindVisibilityRecord"]
    peers = [c for c in sc.calls if c[0] == "BindPeerRecords"]
    for k in range(6):
        off = (k % 4) * slot_bytes + rank * words * 4
        ok &= binds[k][1] == (own + off,) and peers[k][1] == ([peer + off],)
    expands = [c for c in sc.calls if c[0] == "ExpandVisibleSignalled"]
    first = expands.pop(0)                         # after frame 0: every record skipped, nothing expected — it only raises the flags
    ok &= first[2]["skip_index"] == 0 and first[2]["skip_count"] == world and first[1][8] == 0
    for k, c in enumerate(expands):                 # frame k: slot k%4, list half k%2, waits for sequence k+1, skips itself
        a, kw = c[1], c[2]
        ok &= a[0] == own + (k % 4) * slot_bytes and a[1:3] == (world, words)
        ok &= a[3] == ex.lists.data_ptr() + (k % 2) * world * cap * 4 and a[6] == own + flags_off and a[8] == k + 1
        ok &= kw["skip_index"] == rank
    signal_values = [c[1][0] for c in sc.calls if c[0] == "SetSignalValue"]
    ok &= signal_values == [1, 2, 3, 4, 5, 6]
    ex.close(sc)
    ok &= ctx.closed == [peer] and ctx.freed == [own]
    # (3) several scenes per rank: rank r's scene j is record r*S + j; one flag per rank covers all of them
    ctx, a, b = _FakeCtx(rank), _FakeScene(), _FakeScene()
    ex = PeerRecordExchange(ctx, words, cap, world, rank, "cpu", torch.int32, records_per_rank=2)
    slot_bytes = world * 2 * words * 4
    for k in range(3):
        ex.bind([a, b], k)
        a.Update(), b.Update()
        ex.expand(a, k - 1)
    for j, sc in enumerate((a, b)):
        cells = [c[1][0] for c in sc.calls if c[0] == "BindVisibilityRecord"]
        ok &= cells == [own + (k % 4) * slot_bytes + (rank * 2 + j) * words * 4 for k in range(3)]
        outs = [c[1] for c in sc.calls if c[0] == "BindVisibleOutput"]
        ok &= [o[0] for o in outs] == [ex.lists.data_ptr() + ((k % 2) * world * 2 + rank * 2 + j) * cap * 4 for k in range(3)]
    ok &= [c[0] for c in b.calls].count("BindPeerSignal") == 0 and [c[0] for c in b.calls].count("SetSignalValue") == 0
    e = [c for c in a.calls if c[0] == "ExpandVisibleSignalled"][1:]     # [0] is the signal-only launch after frame 0
    ok &= len(e) == 2 and all(c[1][1] == world * 2 for c in e)
    ok &= all(c[2]["skip_index"] == rank * 2 and c[2]["skip_count"] == 2 and c[2]["records_per_flag"] == 2 for c in e)
    ex.close([a, b])
    with open(os.path.join(out_dir, f"peer{rank}.ok"), "w") as f:
        f.write("1" if ok else "0")
    dist.destroy_process_group()


def test_peer_record_exchange_protocol_world_size_2_gloo(tmp_path, monkeypatch):
    import torch
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_peer_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        assert (tmp_path / f"peer{r}.ok").read_text() == "1"


# ---- one scene split over ranks by root subtree (entity shards) ------------------------------------
def _forest(rng, level_sizes, orphan_rate=0.0):
    """A random level-ordered forest: children sorted by parent, optional extra roots below level 0."""
    offs = np.concatenate([[0], np.cumsum(level_sizes)]).astype(np.uint32)
    E = int(offs[-1])
    parent = np.full(E, 0xFFFFFFFF, np.uint32)
    for l in range(1, len(level_sizes)):
        lo, hi = int(offs[l]), int(offs[l + 1])
        p = np.sort(rng.integers(int(offs[l - 1]), lo, hi - lo)).astype(np.uint32)
        p[rng.random(hi - lo) < orphan_rate] = 0xFFFFFFFF
        parent[lo:hi] = p
    a = {"position": rng.uniform(-10, 10, (E, 3)), "rotation": rng.uniform(-3, 3, (E, 3)), "scale": rng.uniform(0.9, 1.1, (E, 3)),
         "aabb_center": rng.uniform(-0.2, 0.2, (E, 3)), "aabb_half": rng.uniform(0.1, 1.0, (E, 3))}
    a = {k: v.astype(np.float32) for k, v in a.items()}
    a["position"][: level_sizes[0]] *= 100.0
    a["parent"], a["level_offsets"] = parent, offs
    return a


@pytest.mark.parametrize("world", [1, 2, 3, 8])
@pytest.mark.parametrize("orphans", [0.0, 0.05])
def test_scene_sharded_by_root_subtree_equals_the_whole_scene(oracle, world, orphans):
    from oracle.binding import camera
    from zenyth_b200.distributed import merge_visible, root_of_entities, shard_scene_by_roots
    rng = np.random.default_rng(7 + world)
    scene = _forest(rng, [11, 90, 700, 3000, 1500], orphan_rate=orphans)
    view, proj = camera(oracle)
    whole = oracle.scene_update(scene, view, proj)
    shards = shard_scene_by_roots(scene, world)
    assert len(shards) == world and all(s.entity_count > 0 for s in shards)
    every = np.sort(np.concatenate([s.global_index for s in shards]))
    assert np.array_equal(every, np.arange(scene["parent"].size))            # a partition
    root = root_of_entities(scene["level_offsets"], scene["parent"])
    lists = []
    for s in shards:
        assert np.all(np.diff(s.global_index.astype(np.int64)) > 0)
        # hierarchy-closed: every parent is in the same shard, in an earlier level of it
        p = s.arrays["parent"]
        has = p != 0xFFFFFFFF
        assert np.array_equal(s.global_index[p[has]], scene["parent"][s.global_index[has]])
        offs = s.arrays["level_offsets"].astype(np.int64)
        level_of = np.searchsorted(offs, np.arange(s.entity_count), side="right") - 1
        assert np.all(level_of[p[has]] < level_of[has]) and np.all(np.diff(offs) > 0)
        for l in range(offs.size - 1):                                       # children stay sorted by parent
            seg = p[offs[l]:offs[l + 1]]
            seg = seg[seg != 0xFFFFFFFF].astype(np.int64)
            assert np.all(np.diff(seg) >= 0)
        assert np.unique(root[s.global_index]).size == np.count_nonzero(~has)
        got = oracle.scene_update(s.arrays, view, proj)
        assert np.array_equal(got["world"], whole["world"][s.global_index])  # bit-identical, not merely close
        assert np.array_equal(got["mvp"], whole["mvp"][s.global_index])
        lists.append(s.to_global(got["visible"]))
    assert np.array_equal(merge_visible(lists), whole["visible"])
    sizes = [s.entity_count for s in shards]
    assert max(sizes) <= 2.5 * (sum(sizes) / world) + 1


def test_scene_sharding_rejects_what_it_cannot_split():
    from zenyth_b200.distributed import root_of_entities, shard_scene_by_roots
    a = _forest(np.random.default_rng(1), [2, 10, 30])
    with pytest.raises(ValueError):
        shard_scene_by_roots(a, 3)                    # two root subtrees, three ranks
    with pytest.raises(ValueError):
        shard_scene_by_roots(a, 0)
    bad = dict(a, parent=a["parent"].copy())
    bad["parent"][5] = 20                             # a parent in a LATER level
    with pytest.raises(ValueError):
        root_of_entities(bad["level_offsets"], bad["parent"])
    flat = {k: v for k, v in _forest(np.random.default_rng(2), [100]).items() if k != "parent"}
    shards = shard_scene_by_roots(flat, 4)            # no parent array: every entity is a root
    assert [s.entity_count for s in shards] == [25, 25, 25, 25]
