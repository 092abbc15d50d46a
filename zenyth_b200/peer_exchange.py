"""The exchange fused into the compute kernels over NVLink peer memory — no collective on the data path.

Two classes.  `PeerExchange` is the product path: a thin binding of the C++ host layer
(zn_exchange_* in include/zenyth_b200.h, zenyth_b200/csrc/zn_exchange.cpp) — buffers, CUDA IPC
rendezvous, per-frame binding arithmetic and the flag-gated expansion all live in the library; the
only thing Python supplies is the host all-gather callback for the 64-byte handles at set-up
(torch.distributed here; a C++ engine passes MPI_Allgather or a socket).  `PeerRecordExchange` is the
earlier torch-hosted statement of the same protocol, kept for the comparison variants bench.py can
still run (`--exchange peer-nccl`, `peer-waitkernel`) and for the gloo protocol test on fakes.
The protocol, as both implement it:

Every rank owns ONE plain device allocation

    records [slots][world * S][record_words]   gathered visibility records, slot = frame % slots;
                                               S = scenes per rank, rank r's scene j is record r*S + j
    flags   [world] (128-byte stride)          flags[r] = sequence number of the last frame rank r completed

exports its CUDA IPC handle and maps every peer's allocation (zn_ipc_export / zn_ipc_open).  A scene
bound to it stores each tile's visibility mask — and, from the emit kernel, the record header —
straight into slot [rank] of every peer's `records` while its kernels run: plain `st.global` on
mapped peer pointers, tile by tile (zn_scene_bind_visibility_record + zn_scene_bind_peer_records).

Completion travels the same way (signal="flags", the default): the first thread of the kernel that
FOLLOWS the update on the stream — the next zn_scene_expand_visible_signalled (or zn_scene_wait_flags)
launch; zn_scene_update itself raises nothing — does one system-scope fence and raises flags[rank] in
every rank's allocation to the frame's sequence number (zn_scene_bind_peer_signal), and the expansion
kernel of a frame waits on the owners' flags itself (bounded wait; a record whose owner never
signals is not expanded).  No NCCL call, no host synchronisation and no extra launch per frame.  The
LAST frame of a run must therefore be followed by one more expand() call (bench.py's finish()), or its
peers run into the timeout.  signal="nccl" keeps the earlier form — a 4-byte all-reduce per frame as
the cross-rank barrier — for comparison.

A rank's OWN lists are not expanded a second time: its scenes' emit kernels write them directly
into lists[frame % 2][rank*S + j] (zn_scene_bind_visible_output) and the expansion skips those records.

Pipeline per frame k:  bind(k); every owned scene's Update(); expand(k-1).
Slot reuse.  Update(k+1) on rank r stores into slot (k+1) % slots of every peer p.  It is
stream-ordered after r's expand(k-1), which waited for p's frame k-1 signal; p raised it at the end
of its update(k-1), i.e. after its expand(k-3).  So the youngest frame a peer is guaranteed to have
finished reading is k-3 and a slot may be reused after FOUR frames: slots = 4.
Lists are double-buffered: lists[k % 2] is complete (all ranks) once expand(k) has run and is
overwritten by update(k+2), so its consumer belongs on the stream between the two.
"""
from __future__ import annotations

from typing import List, Optional

FLAG_STRIDE_WORDS = 32          # one 128-byte line per writer


class PeerExchange:
    """zn_exchange_* behind the engine-idiom names (mirrors Zenyth::PeerExchange in zenyth_b200.hpp)."""

    def __init__(self, ctx, record_words: int, capacity: int, world_size: int, rank: int, records_per_rank: int = 1,
                 slots: int = 4, timeout_ms: int = 2000, allgather=None, group=None, overlap: bool = False):
        import ctypes as C

        from . import _capi
        self._lib, self._C, self._capi = _capi.load(), C, _capi
        self.ctx, self.world_size, self.rank = ctx, int(world_size), int(rank)
        self.records_per_rank, self.capacity = int(records_per_rank), int(capacity)
        self.total_records = self.world_size * self.records_per_rank
        self.timeout_ms = int(timeout_ms)
        if allgather is None and world_size > 1:
            allgather = self._torch_allgather(group)
        self._failed = None

        def trampoline(_user, send, recv, nbytes):
            try:
                mine = C.string_at(send, nbytes)
                parts = allgather(mine)
                assert len(parts) == self.world_size and all(len(p) == nbytes for p in parts)
                C.memmove(recv, b"".join(parts), nbytes * self.world_size)
                return 0
            except Exception as e:      # noqa: BLE001 - reported through the C layer's status
                self._failed = e
                return 1

        self._cb = _capi.ALLGATHER_FN(trampoline if world_size > 1 else (lambda *a: 0))   # kept alive with the object
        desc = _capi.ExchangeDesc(self.world_size, self.rank, self.records_per_rank, int(record_words), self.capacity, int(slots),
                                  self.timeout_ms, self._cb, None, 1 if overlap else 0)
        h = C.c_void_p()
        rc = self._lib.zn_exchange_create(ctx._h, C.byref(desc), C.byref(h))
        if rc != 0:
            raise RuntimeError(self._lib.zn_last_error().decode("utf-8", "replace"))
        self._h = h

    @staticmethod
    def _torch_allgather(group):
        import torch.distributed as dist

        def gather(mine: bytes):
            parts = [None] * dist.get_world_size(group)
            dist.all_gather_object(parts, mine, group=group)
            return parts
        return gather

    def _scene_array(self, scenes):
        scenes = list(scenes) if isinstance(scenes, (list, tuple)) else [scenes]
        return (self._C.c_void_p * max(len(scenes), 1))(*[s._h for s in scenes]), len(scenes)

    def bind(self, scenes, frame: int) -> None:
        arr, n = self._scene_array(scenes)
        self._capi.check(self._lib.zn_exchange_bind(self._h, arr, n, frame))

    def launch(self, frame: int) -> None:      # interface parity with the torch-hosted variants: nothing to launch
        return None

    def signal(self, scene) -> None:
        """Raise this rank's flags for the last bound frame without waiting for anyone; returns when they are up.
        Only ranks that SHARE one GPU need it (signal, host barrier, then expand)."""
        self._capi.check(self._lib.zn_exchange_signal(self._h, scene._h))

    def join(self) -> None:
        """Make the context stream wait for the last expansion (overlap=True: it runs on the side stream; else a no-op)."""
        self._capi.check(self._lib.zn_exchange_join(self._h))

    def expand(self, scene, frame: int, index_bases=None) -> None:
        """Asynchronous: raise this rank's flags for the last bound frame, wait on the device for the owners' flags of
        `frame`, derive their lists.  Default pattern: expand(k - 1) after Update(k); with overlap=True (side stream):
        expand(k) right after Update(k)."""
        import numpy as np
        bases = None if index_bases is None else np.ascontiguousarray(index_bases, np.uint32)
        self._keep = bases
        self._capi.check(self._lib.zn_exchange_expand(self._h, scene._h, frame, None if bases is None else bases.ctypes.data_as(self._C.c_void_p)))

    def result_raw(self):
        """(counts, device pointer of lists[total_records][capacity], stride) of the last expanded frame (synchronises)."""
        import numpy as np
        C = self._C
        counts = np.zeros(self.total_records, np.uint32)
        ptr, stride = C.c_void_p(), C.c_uint32()
        rc = self._lib.zn_exchange_result(self._h, counts.ctypes.data_as(C.c_void_p), C.byref(ptr), C.byref(stride))
        if rc != 0:
            raise RuntimeError(self._lib.zn_last_error().decode("utf-8", "replace"))
        return [int(c) for c in counts], int(ptr.value), int(stride.value)

    def result(self):
        """(counts, lists) with lists a zero-copy torch int32 view [total_records, capacity] of the library's buffer."""
        import torch
        counts, ptr, stride = self.result_raw()

        class _View:
            __cuda_array_interface__ = {"shape": (self.total_records * stride,), "typestr": "<i4", "data": (ptr, False), "version": 3}
        return counts, torch.as_tensor(_View(), device=f"cuda:{self.ctx.device}").view(self.total_records, stride)

    def peer_bytes_per_frame(self) -> int:
        n = self._C.c_uint64()
        self._capi.check(self._lib.zn_exchange_info(self._h, None, None, self._C.byref(n)))
        return int(n.value)

    def close(self, scenes=None) -> None:
        if getattr(self, "_h", None):
            arr, n = self._scene_array(scenes) if scenes is not None else (None, 0)
            self._lib.zn_exchange_destroy(self._h, arr, n)
            self._h = None


class PeerRecordExchange:
    def __init__(self, ctx, record_words: int, capacity: int, world_size: int, rank: int, device, dtype, slots: int = 4,
                 signal: str = "flags", timeout_ms: int = 2000, group=None, fused_wait: bool = True, records_per_rank: int = 1):
        import torch
        import torch.distributed as dist
        if signal not in ("flags", "nccl"):
            raise ValueError(f"signal must be 'flags' or 'nccl', not {signal!r}")
        if slots < 4:
            raise ValueError("slots must be >= 4 (see the module docstring)")
        self.ctx, self.world_size, self.rank, self.slots, self.group = ctx, world_size, rank, slots, group
        self.signal, self.timeout_ms, self.fused_wait = signal, int(timeout_ms), bool(fused_wait)
        self.device = torch.device(device)
        self.record_words, self.capacity = int(record_words), int(capacity)
        self.records_per_rank = int(records_per_rank)      # a rank that owns several scenes ships them together
        self.total_records = world_size * self.records_per_rank
        if self.total_records > 64:
            raise ValueError("at most 64 records per exchange (zn_scene_expand_visible)")
        self.slot_bytes = self.total_records * self.record_words * 4
        self.flags_offset = (slots * self.slot_bytes + 127) // 128 * 128
        # Set-up is collective; a rank that cannot allocate, export or map still takes part in every
        # rendezvous below, so all ranks learn of the failure together (and raise together) instead of
        # one of them hanging the others in a barrier.
        self.base, self.peer_base, problem = 0, [], ""
        try:
            self.base = ctx.DeviceAlloc(self.flags_offset + world_size * FLAG_STRIDE_WORDS * 4)
            handle = ctx.IpcExport(self.base)
        except Exception as e:      # noqa: BLE001 - reported to every rank below
            handle, problem = None, f"rank {rank}: {e}"
        handles: List[Optional[bytes]] = [None] * world_size
        dist.all_gather_object(handles, handle, group=group)
        if not problem and all(h is not None for h in handles):
            try:
                for r in range(world_size):
                    self.peer_base.append(self.base if r == rank else ctx.IpcOpen(handles[r]))
            except Exception as e:  # noqa: BLE001
                problem = f"rank {rank}: {e}"
        problems: List[Optional[str]] = [None] * world_size
        dist.all_gather_object(problems, problem, group=group)
        if any(problems) or any(h is None for h in handles):
            self._release()
            raise RuntimeError("peer memory unavailable: " + "; ".join(p for p in problems if p))
        self.lists = torch.empty(2 * self.total_records * self.capacity, dtype=dtype, device=device)
        self.counts = torch.zeros(2 * self.total_records, dtype=dtype, device=device)
        self.status = torch.zeros(1, dtype=torch.int32, device=device)
        self.token = torch.zeros(1, dtype=torch.int32, device=device)
        self.work = [None] * slots
        self.issued = -1            # last frame bound
        self.expanded = -1          # last frame expanded
        dist.barrier(group=group)   # every rank has mapped every buffer before anyone writes

    # ---- addresses -------------------------------------------------------------------------------
    def _record_offset(self, slot: int, j: int = 0) -> int:
        return slot * self.slot_bytes + (self.rank * self.records_per_rank + j) * self.record_words * 4

    def own_record_ptr(self, frame: int, j: int = 0) -> int:
        """Where the record of this rank's j-th scene for `frame` lives in its OWN gathered buffer."""
        return self.base + self._record_offset(frame % self.slots, j)

    def peer_record_ptrs(self, frame: int, j: int = 0) -> List[int]:
        """The same cell inside every PEER's gathered buffer (mapped pointers)."""
        off = self._record_offset(frame % self.slots, j)
        return [self.peer_base[r] + off for r in range(self.world_size) if r != self.rank]

    def flag_ptrs(self) -> List[int]:
        """This rank's flag cell in EVERY rank's allocation (its own included)."""
        off = self.flags_offset + self.rank * FLAG_STRIDE_WORDS * 4
        return [self.peer_base[r] + off for r in range(self.world_size)]

    def _list_ptr(self, frame: int, record: int = 0) -> int:
        return self.lists.data_ptr() + ((frame % 2) * self.total_records + record) * self.capacity * 4

    def _count_ptr(self, frame: int, record: int = 0) -> int:
        return self.counts.data_ptr() + ((frame % 2) * self.total_records + record) * 4

    # ---- per frame -------------------------------------------------------------------------------
    def bind(self, scenes, frame: int) -> None:
        """Point the next Update of this rank's scene(s) at frame `frame`'s cells (own + peers) and
        list slots.  `scenes`: one scene, or the list of the records_per_rank scenes this rank owns;
        the FIRST one is the scene expand() must be called with (it carries the signal)."""
        scenes = list(scenes) if isinstance(scenes, (list, tuple)) else [scenes]
        assert len(scenes) == self.records_per_rank, "one scene per record of this rank"
        assert frame == self.issued + 1, "frames are bound in order"
        if frame == 0 and self.signal == "flags":
            scenes[0].BindPeerSignal(self.flag_ptrs())
        for j, scene in enumerate(scenes):
            record = self.rank * self.records_per_rank + j
            scene.BindVisibilityRecord(self.own_record_ptr(frame, j))
            scene.BindPeerRecords(self.peer_record_ptrs(frame, j))
            scene.BindVisibleOutput(self._list_ptr(frame, record), self.capacity, self._count_ptr(frame, record))
        if self.signal == "flags":
            scenes[0].SetSignalValue(frame + 1)
        self.issued = frame

    def launch(self, frame: int) -> None:
        """signal="nccl" only: the cross-rank barrier ordered after this rank's update of `frame`."""
        if self.signal != "nccl":
            return
        import torch.distributed as dist
        slot = frame % self.slots
        assert self.work[slot] is None
        self.work[slot] = dist.all_reduce(self.token, group=self.group, async_op=True)

    def expand(self, scene, frame: int, index_bases=None) -> bool:
        """Derive every PEER's list of `frame` into lists[frame % 2] (asynchronous); False if the frame
        was never issued or is already expanded."""
        if frame < 0 and self.issued == 0 and self.expanded < 0 and self.signal == "flags":
            # frame 0 is updated, nothing to expand yet: a signal-only launch (every record skipped) raises its flags now
            scene.ExpandVisibleSignalled(self.base, self.total_records, self.record_words, self._list_ptr(0), self.capacity,
                                         self._count_ptr(0), self.base + self.flags_offset, FLAG_STRIDE_WORDS, 0,
                                         self.status.data_ptr(), skip_index=0, skip_count=self.total_records,
                                         records_per_flag=self.records_per_rank, timeout_ms=self.timeout_ms)
            return False
        if frame < 0 or frame > self.issued or frame <= self.expanded:
            return False
        assert frame == self.expanded + 1, "frames are expanded in order"
        records = self.base + (frame % self.slots) * self.slot_bytes
        S, first_own = self.records_per_rank, self.rank * self.records_per_rank
        plain_skip = first_own if S == 1 else -1        # the plain entry point can leave out one record only
        if self.signal == "flags" and not self.fused_wait:
            scene.WaitFlags(self.base + self.flags_offset, FLAG_STRIDE_WORDS, self.world_size, frame + 1, self.status.data_ptr(),
                            timeout_ms=self.timeout_ms)
            scene.ExpandVisible(records, self.total_records, self.record_words, self._list_ptr(frame), self.capacity,
                                self._count_ptr(frame), index_bases=index_bases, skip_index=plain_skip)
        elif self.signal == "flags":
            scene.ExpandVisibleSignalled(records, self.total_records, self.record_words, self._list_ptr(frame), self.capacity,
                                         self._count_ptr(frame), self.base + self.flags_offset, FLAG_STRIDE_WORDS, frame + 1,
                                         self.status.data_ptr(), index_bases=index_bases, skip_index=first_own, skip_count=S,
                                         records_per_flag=S, timeout_ms=self.timeout_ms)
        else:
            slot = frame % self.slots
            self.work[slot].wait()      # the current stream waits for the barrier
            self.work[slot] = None
            scene.ExpandVisible(records, self.total_records, self.record_words, self._list_ptr(frame), self.capacity,
                                self._count_ptr(frame), index_bases=index_bases, skip_index=plain_skip)
        self.expanded = frame
        return True

    def result(self):
        """(counts, lists[world * records_per_rank, capacity]) of the last expanded frame (synchronises)."""
        status = int(self.status.item())
        if status:
            raise RuntimeError(f"rank {self.rank}: rank {(status - 1) // self.records_per_rank} (record {status - 1}) never signalled "
                               f"a frame within {self.timeout_ms} ms")
        f = self.expanded % 2
        counts = [int(c) for c in self.counts[f * self.total_records: (f + 1) * self.total_records].tolist()]
        return counts, self.lists.view(2, self.total_records, self.capacity)[f]

    def close(self, scene=None) -> None:
        import torch
        import torch.distributed as dist
        if scene is not None:
            for sc in (list(scene) if isinstance(scene, (list, tuple)) else [scene]):
                sc.BindPeerRecords([])
                sc.BindPeerSignal([])
                sc.BindVisibilityRecord(None)
                sc.BindVisibleOutput(None)
        if self.device.type == "cuda":
            torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)          # nobody still writes into a buffer that is about to go away
        self._release()
        dist.barrier(group=self.group)

    def _release(self) -> None:
        for r, p in enumerate(self.peer_base):
            if r != self.rank:
                self.ctx.IpcClose(p)
        self.peer_base = []
        if self.base:
            self.ctx.DeviceFree(self.base)
            self.base = 0
