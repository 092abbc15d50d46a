"""The collective-free exchange: records stored into extra ("peer") buffers by the kernels, the
completion flag raised by the emit kernel, the flag-gated expansion.  On one GPU the "peers" are
further local buffers, which exercises exactly the same kernels and entry points; the two-process
test over real CUDA IPC / NVLink runs where two GPUs are visible (bench.py --exchange peer checks
the same thing at N = 2..8 at full size)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle.binding import BASE_SEED, camera, h8_geo_level_sizes

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _scene(oracle, gpu_ctx, levels=5):
    import zenyth_b200 as zn
    sizes = h8_geo_level_sizes(7, levels)
    arrays = oracle.synth_scene(BASE_SEED + 3, sizes, fanout=8, center_range=0.25)
    view, proj = camera(oracle)
    sc = zn.Scene(gpu_ctx, sum(sizes), arrays["level_offsets"])
    sc.Load(arrays)
    sc.SetCamera(zn.Camera(view=view, proj=proj))
    return sc, sum(sizes), oracle.scene_update(arrays, view, proj)["visible"]


def _big_scene(gpu_ctx):
    """7 levels, the last one 1.8M entities: above the size from which a level kernel pushes its own masks."""
    import zenyth_b200 as zn
    sizes = h8_geo_level_sizes(7, 7)
    offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint32)
    sc = zn.Scene(gpu_ctx, sum(sizes), offs)
    sc.FillSynthetic(BASE_SEED + 3, 8, 0.25)
    sc.SetCamera(zn.Camera(zn.CameraDesc(eye=(0.0, 0.0, 1500.0), center=(0.0, 0.0, 1499.0))))
    sc.Update()
    return sc, sum(sizes), sc.Visible()      # the plain path is checked against the oracle elsewhere


@pytest.mark.parametrize("two_calls", [False, True], ids=["fused-update", "update-world+cull"])
@pytest.mark.parametrize("worklist", [False, True], ids=["expand-grid", "expand-worklist"])
@pytest.mark.parametrize("big", [False, True], ids=["small-levels-only", "with-a-big-level"])
def test_kernels_store_the_record_into_every_target_and_signal(oracle, gpu_ctx, big, worklist, two_calls):
    import torch
    import zenyth_b200 as zn
    sc, E, expect = _big_scene(gpu_ctx) if big else _scene(oracle, gpu_ctx)
    sc.SetExpandMode(worklist)
    cam = zn.Camera(zn.CameraDesc(eye=(0.0, 0.0, 1500.0), center=(0.0, 0.0, 1499.0)))
    assert 0 < expect.size < E
    words = sc.DeviceBuffers().visibility_record_words
    recs = torch.zeros(4 * words, dtype=torch.int32, device="cuda")          # own + 3 "peers"
    flags = torch.zeros(4 * 32, dtype=torch.int32, device="cuda")
    status = torch.zeros(1, dtype=torch.int32, device="cuda")
    lists = torch.full((4 * E,), -1, dtype=torch.int32, device="cuda")
    counts = torch.zeros(4, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    base, fbase = recs.data_ptr(), flags.data_ptr()
    sc.BindVisibilityRecord(base + 4 * words)                                 # own copy = record 1
    sc.BindPeerRecords([base, base + 8 * words, base + 12 * words])
    sc.BindPeerSignal([fbase + 128 * r for r in range(4)])
    for frame in range(3):                                                    # the done-counter resets every frame
        sc.SetSignalValue(frame + 1)
        if two_calls:                                                         # the cull pass stores the record into the targets itself
            sc.UpdateWorld()
            sc.Cull(cam)
        else:
            sc.Update()
        sc.ExpandVisibleSignalled(base, 4, words, lists.data_ptr(), E, counts.data_ptr(), fbase, 32, frame + 1, status.data_ptr(),
                                  index_bases=[0, 0, 7, 0], timeout_ms=5000)
        gpu_ctx.Sync()
        assert status.item() == 0
        assert flags.view(4, 32)[:, 0].tolist() == [frame + 1] * 4
    got_recs = recs.view(4, words).cpu().numpy()
    for r in (0, 2, 3):
        assert np.array_equal(got_recs[r], got_recs[1]), f"target {r} differs from the own record"
    assert np.array_equal(sc.Visible(), expect)
    assert counts.tolist() == [expect.size] * 4
    got = lists.view(4, E).cpu().numpy()
    for r in range(4):
        assert np.array_equal(got[r, :expect.size], expect.astype(np.int32) + (7 if r == 2 else 0))
    sc.BindPeerRecords([])
    sc.BindPeerSignal([])
    sc.BindVisibilityRecord(None)
    sc.Update()
    assert np.array_equal(sc.Visible(), expect)
    sc.Close()


@pytest.mark.parametrize("worklist", [False, True], ids=["expand-grid", "expand-worklist"])
def test_a_signal_that_never_comes_times_out_instead_of_hanging(oracle, gpu_ctx, worklist):
    import torch
    sc, E, expect = _scene(oracle, gpu_ctx, levels=3)
    sc.SetExpandMode(worklist)
    words = sc.DeviceBuffers().visibility_record_words
    recs = torch.zeros(2 * words, dtype=torch.int32, device="cuda")
    flags = torch.zeros(2 * 32, dtype=torch.int32, device="cuda")
    status = torch.zeros(1, dtype=torch.int32, device="cuda")
    lists = torch.zeros(2 * E, dtype=torch.int32, device="cuda")
    counts = torch.zeros(2, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    sc.BindVisibilityRecord(recs.data_ptr())
    sc.BindPeerSignal([flags.data_ptr()])                                     # only "rank 0" ever signals
    sc.SetSignalValue(1)
    sc.Update()
    sc.ExpandVisibleSignalled(recs.data_ptr(), 2, words, lists.data_ptr(), E, counts.data_ptr(), flags.data_ptr(), 32, 1,
                              status.data_ptr(), timeout_ms=20)
    gpu_ctx.Sync()
    assert status.item() == 2                                                 # 1 + the silent rank
    assert counts[0].item() == expect.size
    # the record whose owner never signalled is NOT expanded (it may be half written): count 0, list untouched
    assert counts[1].item() == 0
    assert int(lists[E:].abs().sum().item()) == 0
    assert np.array_equal(lists[:expect.size].cpu().numpy().astype(np.uint32), expect)
    # a flag already AHEAD of the expected sequence number passes at once
    flags[32] = 5
    status.zero_()
    torch.cuda.synchronize()
    sc.ExpandVisibleSignalled(recs.data_ptr(), 2, words, lists.data_ptr(), E, counts.data_ptr(), flags.data_ptr(), 32, 1,
                              status.data_ptr(), skip_index=0, timeout_ms=20)
    gpu_ctx.Sync()
    assert status.item() == 0
    sc.BindPeerSignal([])
    sc.BindVisibilityRecord(None)
    sc.Close()


_TWO_RANKS = r"""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["ZN_ROOT"])
import zenyth_b200 as zn
from zenyth_b200.peer_exchange import PeerExchange, PeerRecordExchange
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
# one GPU: both ranks share cuda:0 (CUDA IPC between processes on one device is legal) and gloo carries
# the handles — NCCL cannot place two ranks on one device; two or more GPUs: one rank each, NCCL.
# Ranks that share a GPU run in LOCKSTEP (device sync + host barrier after every frame): kernels of
# different processes on one device are not guaranteed to run at the same time, so no kernel may have
# to spin on a flag the other process has yet to raise (a spinning kernel can starve the one it waits
# for: Xid 109 on this driver).  IPC mapping, peer stores, flag stores / acquire loads and the slot
# rotation are exercised all the same; with a GPU per rank the device does the waiting, as in production.
shared = torch.cuda.device_count() < world
device_index = 0 if shared else rank
torch.cuda.set_device(device_index)
dev = torch.device("cuda", device_index)
if shared:
    dist.init_process_group("gloo")
else:
    dist.init_process_group("nccl", device_id=dev)
stream = torch.cuda.Stream(dev); torch.cuda.set_stream(stream)
ctx = zn.GpuContext(device_index, stream=stream.cuda_stream)
TIMEOUT_MS = 5000 if shared else 2000
sizes = [7 * 8 ** l for l in range(7)]
E = sum(sizes)
offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint32)
for signal, S in (("flags", 1), ("nccl", 1), ("flags", 2), ("library", 1), ("library", 2), ("library-overlap", 2)):
    if signal == "nccl" and shared:
        continue
    scenes, mine = [], []
    for j in range(S):
        sc = zn.Scene(ctx, E, offs)
        sc.FillSynthetic(0x5A454E59 + 11 + rank * S + j, 8, 0.25)
        sc.SetCamera(zn.Camera(zn.CameraDesc(eye=(0, 0, 1500.0), center=(0, 0, 1499.0))))
        sc.SetIndexBase((rank * S + j) * E)
        sc.Update()
        mine.append(sc.Visible().astype(np.int64))
        scenes.append(sc)
    if signal.startswith("library"):     # the C++ host layer (zn_exchange_*): torch only carries the 64-byte handles at set-up
        ex = PeerExchange(ctx, scenes[0].DeviceBuffers().visibility_record_words, E, world, rank, records_per_rank=S, timeout_ms=TIMEOUT_MS,
                          overlap=signal == "library-overlap")
    else:
        ex = PeerRecordExchange(ctx, scenes[0].DeviceBuffers().visibility_record_words, E, world, rank, dev, torch.int32, signal=signal,
                                records_per_rank=S, timeout_ms=TIMEOUT_MS)
    bases = np.arange(world * S, dtype=np.uint32) * E
    for k in range(9):
        ex.bind(scenes, k)
        for sc in scenes:
            sc.Update()
        if signal.startswith("library"):
            if shared:                  # raise the flags (no waiting), make sure every rank has, only then launch the kernel that waits
                ex.signal(scenes[0])
                dist.barrier()
            ex.expand(scenes[0], k if signal == "library-overlap" else k - 1, bases)
            if shared and ex._lib is not None and (signal == "library-overlap" or k >= 1):
                ex.result()
                dist.barrier()
        else:
            ex.launch(k)
            ex.expand(scenes[0], k - 1, bases)
            if shared:
                torch.cuda.synchronize(dev)
                dist.barrier()
    ex.expand(scenes[0], 8, bases)
    counts, lists = ex.result()
    every = [None] * world
    dist.all_gather_object(every, mine)
    every = [v for per_rank in every for v in per_rank]
    for r in range(world * S):
        assert counts[r] == every[r].size, (signal, S, r, counts, every[r].size)
        assert np.array_equal(lists[r, :counts[r]].cpu().numpy().astype(np.int64) & 0xFFFFFFFF, every[r]), (signal, S, r)
    ex.close(scenes)
    for sc in scenes:
        sc.Close()
ctx.Close()
dist.destroy_process_group()
print("peer exchange ok", rank)
"""


def test_two_processes_exchange_over_peer_memory(tmp_path):
    """Two processes over real CUDA IPC: on two GPUs one rank each (NVLink), on ONE GPU both on cuda:0."""
    script = tmp_path / "two_ranks.py"
    script.write_text(_TWO_RANKS)
    env = dict(os.environ, ZN_ROOT=ROOT)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", "29731", str(script)], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert out.stdout.count("peer exchange ok") == 2


@pytest.mark.parametrize("world,scenes_per_rank,overlap,worklist", [(2, 1, 0, 0), (2, 2, 0, 0), (3, 1, 0, 0), (2, 2, 1, 0), (2, 2, 0, 1), (2, 1, 1, 1)])
def test_cpp_ranks_exchange_without_torch_or_nccl(world, scenes_per_rank, overlap, worklist):
    """tests/cpp/exchange_ranks.cpp: fork()ed C++ processes, a socket hub as the host all-gather, Zenyth::PeerExchange,
    ShardScenes and SceneIndexBases of include/zenyth_b200.hpp — every gathered list of every frame is checked
    against the single-GPU result inside the program.  Ranks land on GPU rank % device_count."""
    from zenyth_b200 import build as zb
    exe = zb.build_cpp_tests()
    out = subprocess.run([exe, str(world), str(scenes_per_rank), "9", "7", str(overlap), str(worklist)], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert out.stdout.count("exchange ok rank") == world
