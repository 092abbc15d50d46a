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
