"""Multi-GPU plumbing for the one exchange the path has: after every rank has
updated the scenes it owns, all-gather the compacted visible COUNTS, then the index LISTS.

One process per GPU; `torch.distributed` provides the collective (NCCL over NVLink/NVSwitch on the
GPU box, gloo in the CPU tests).  Scenes are independent, so nothing else ever crosses GPUs:
matrices stay where they were computed.  Everything here is host logic over torch tensors; it
does no arithmetic of the path.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple


def shard_scenes(num_scenes: int, world_size: int, rank: int) -> List[int]:
    """Scene ids owned by `rank`: contiguous blocks, sizes differing by at most one
    (8 scenes over 1/2/4/8 GPUs -> 8/4/2/1 per GPU; BASELINE.json configs[4])."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of {world_size}")
    base, extra = divmod(num_scenes, world_size)
    start = rank * base + min(rank, extra)
    return list(range(start, start + base + (1 if rank < extra else 0)))


def scene_index_bases(entity_counts: Sequence[int]) -> List[int]:
    """Global index base of every scene (exclusive prefix sum) so shard-local visible indices
    become global by a single add on the device (zn_scene_set_index_base)."""
    bases, acc = [], 0
    for n in entity_counts:
        bases.append(acc)
        acc += int(n)
    if acc > 2 ** 32:
        raise ValueError("global entity index does not fit in 32 bits")
    return bases


class VisibleGather:
    """All-gather of variable-length visible lists.

    gather(visible, count) -> (counts, lists): `visible` is this rank's device-resident index buffer
    (capacity entries, first `count` valid), `count` a 1-element tensor on the same device.  Step 1
    all-gathers the counts (4 bytes per rank); step 2 all-gathers the lists padded to the largest
    count, so a rank never ships more than max_count indices.  `lists[r, :counts[r]]` is rank r's
    ascending list.
    """

    def __init__(self, capacity: int, world_size: int, device, dtype, group=None):
        import torch
        self.world_size = world_size
        self.capacity = int(capacity)
        self.group = group
        self.counts = torch.zeros(world_size, dtype=dtype, device=device)
        self.buffer = torch.empty(world_size * self.capacity, dtype=dtype, device=device)

    def gather(self, visible, count) -> Tuple[List[int], "object"]:
        import torch.distributed as dist
        dist.all_gather_into_tensor(self.counts, count, group=self.group)
        counts = [int(c) for c in self.counts.tolist()]          # one small D2H: sizes the second step
        m = max(max(counts), 1)
        if m > self.capacity:
            raise RuntimeError(f"visible count {m} exceeds capacity {self.capacity}")
        out = self.buffer[: self.world_size * m]
        dist.all_gather_into_tensor(out, visible[:m], group=self.group)
        return counts, out.view(self.world_size, m)

    @staticmethod
    def concatenate(counts: Sequence[int], lists) -> "object":
        """Global visible list: rank lists trimmed to their counts, in rank order (ascending when
        every rank's index base is the sum of the entity counts before it)."""
        import torch
        return torch.cat([lists[r, : counts[r]] for r in range(len(counts))])
