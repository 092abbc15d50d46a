"""The exchange fused into the compute kernels over NVLink peer memory.

Every rank owns a gathered-records buffer [slots][world][record_words] in plain device memory,
exports its CUDA IPC handle and maps every peer's buffer (zn_ipc_export / zn_ipc_open).  A scene
bound to it (zn_scene_bind_visibility_record + zn_scene_bind_peer_records) stores each tile's
visibility mask — and, from the emit kernel, the record header — straight into slot [rank] of every
peer's buffer while the level kernels run: plain `st.global` on mapped peer pointers, tile by tile.
No collective moves the records; a 4-byte all-reduce is the cross-rank barrier that tells a rank
all peers' frame-k kernels (and therefore their stores into its memory) have completed.  Expansion
into the ordered index lists is the same `zn_scene_expand_visible` as in RecordExchange.

Pipelined like RecordExchange: per frame k a rank runs  update(k) -> barrier(k) [asynchronous] ->
wait barrier(k-1), expand(k-1).  Update(k+1) on rank r stores into slot (k+1) % slots of every
peer p.  It is ordered after barrier(k-1) (r waited for it in step k), and barrier(k-1) completes
only after p finished update(k-1), hence after p's expand(k-3).  So the youngest frame a peer is
guaranteed to be done expanding is k-3, and a slot may be reused after FOUR frames: slots = 4.
"""
from __future__ import annotations

from typing import List, Optional


class PeerRecordExchange:
    def __init__(self, ctx, record_words: int, capacity: int, world_size: int, rank: int, device, dtype, slots: int = 4, group=None):
        import torch
        import torch.distributed as dist
        self.ctx, self.world_size, self.rank, self.slots, self.group = ctx, world_size, rank, slots, group
        self.record_words, self.capacity = int(record_words), int(capacity)
        self.slot_bytes = world_size * self.record_words * 4
        self.base = ctx.DeviceAlloc(slots * self.slot_bytes)
        handles: List[Optional[bytes]] = [None] * world_size
        dist.all_gather_object(handles, ctx.IpcExport(self.base), group=group)
        self.peer_base = [self.base if r == rank else ctx.IpcOpen(handles[r]) for r in range(world_size)]
        self.lists = torch.empty(world_size * self.capacity, dtype=dtype, device=device)
        self.counts = torch.zeros(world_size, dtype=dtype, device=device)
        self.token = torch.zeros(1, dtype=torch.int32, device=device)
        self.work = [None] * slots
        dist.barrier(group=group)   # every rank has mapped every buffer before anyone writes

    def _offset(self, slot: int) -> int:
        return slot * self.slot_bytes + self.rank * self.record_words * 4

    def own_ptr(self, slot: int) -> int:
        """Where this rank's record of `slot` lives in its OWN gathered buffer."""
        return self.base + self._offset(slot)

    def peer_ptrs(self, slot: int) -> List[int]:
        """The same slot [rank] inside every PEER's gathered buffer (mapped pointers)."""
        return [self.peer_base[r] + self._offset(slot) for r in range(self.world_size) if r != self.rank]

    def bind(self, scene, slot: int) -> None:
        scene.BindVisibilityRecord(self.own_ptr(slot))
        scene.BindPeerRecords(self.peer_ptrs(slot))

    def launch(self, slot: int) -> None:
        """Cross-rank barrier ordered after this rank's update of `slot` (asynchronous)."""
        import torch.distributed as dist
        assert self.work[slot] is None
        self.work[slot] = dist.all_reduce(self.token, group=self.group, async_op=True)

    def ready(self, slot: int) -> bool:
        if self.work[slot] is None:
            return False
        self.work[slot].wait()      # the current stream waits for the barrier
        self.work[slot] = None
        return True

    def expand(self, scene, slot: int, index_bases=None) -> None:
        scene.ExpandVisible(self.base + slot * self.slot_bytes, self.world_size, self.record_words, self.lists.data_ptr(), self.capacity,
                            self.counts.data_ptr(), index_bases=index_bases)

    def result(self):
        counts = [int(c) for c in self.counts.tolist()]
        return counts, self.lists.view(self.world_size, self.capacity)

    def close(self, scene=None) -> None:
        import torch
        import torch.distributed as dist
        if scene is not None:
            scene.BindPeerRecords([])
            scene.BindVisibilityRecord(None)
        torch.cuda.synchronize()
        dist.barrier(group=self.group)          # nobody still writes into a buffer that is about to go away
        for r, p in enumerate(self.peer_base):
            if r != self.rank:
                self.ctx.IpcClose(p)
        dist.barrier(group=self.group)
        self.ctx.DeviceFree(self.base)
