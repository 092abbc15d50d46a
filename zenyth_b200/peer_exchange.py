"""The exchange fused into the compute kernels over NVLink peer memory — no collective on the data path.

Every rank owns ONE plain device allocation

    records [slots][world * S][record_words]   gathered visibility records, slot = frame % slots;
                                               S = scenes per rank, rank r's scene j is record r*S + j
    flags   [world] (128-byte stride)          flags[r] = sequence number of the last frame rank r completed

exports its CUDA IPC handle and maps every peer's allocation (zn_ipc_export / zn_ipc_open).  A scene
bound to it stores each tile's visibility mask — and, from the emit kernel, the record header —
straight into slot [rank] of every peer's `records` while its kernels run: plain `st.global` on
mapped peer pointers, tile by tile (zn_scene_bind_visibility_record + zn_scene_bind_peer_records).

Completion travels the same way (signal="flags", the default): the last CTA of the emit kernel,
after a system-scope fence, raises flags[rank] in every rank's allocation to the frame's sequence
number (zn_scene_bind_peer_signal), and the expansion kernel of a frame waits on the owners' flags
itself (zn_scene_expand_visible_signalled, bounded wait).  No NCCL call, no host synchronisation and
no extra launch per frame.  signal="nccl" keeps the earlier form — a 4-byte all-reduce per frame as
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
            raise RuntimeError(f"rank {self.rank}: rank {status - 1} never signalled a frame within {self.timeout_ms} ms")
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
