"""Multi-GPU plumbing for the one exchange the path has: after every rank has
updated the scenes it owns, all-gather the compacted visible COUNTS, then the index LISTS.

STREAM CONTRACT of the NCCL-based classes here (VisibleGather / VisibleExchange / RecordExchange):
torch orders a collective against torch's CURRENT stream only.  A GpuContext runs Scene.Update on
its own non-blocking stream unless it was given torch's (`GpuContext(device, stream=torch.cuda.
current_stream().cuda_stream)`, as bench.py and the tests do) — otherwise the all-gather can read a
list or record the update has not written yet, silently.  Pass `ctx=` to the exchange constructors and
every launch() checks that the two streams are the same.  (The product path for N > 1,
zenyth_b200.peer_exchange.PeerExchange over zn_exchange_*, involves no torch stream at all.)

One process per GPU; `torch.distributed` provides the collective (NCCL over NVLink/NVSwitch on the
GPU box, gloo in the CPU tests).  Scenes are independent, so nothing else ever crosses GPUs:
matrices stay where they were computed.  Everything here is host logic over torch tensors; it
does no arithmetic of the path.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple


def _lib():
    from . import _capi
    return _capi.load()


def _check_value(rc: int) -> None:
    """Host-side argument errors of the C layer surface as ValueError here."""
    if rc != 0:
        raise ValueError(_lib().zn_last_error().decode("utf-8", "replace"))


def check_same_stream(ctx) -> None:
    """Raise if `ctx` (a GpuContext) enqueues its kernels on a stream other than torch's current one."""
    if ctx is None:
        return
    import torch
    mine, torchs = ctx.GetStream(), torch.cuda.current_stream().cuda_stream
    if mine != torchs:
        raise RuntimeError(f"the scene's context runs on stream {mine:#x} but torch's current stream is {torchs:#x}: a collective "
                           "enqueued now would not be ordered after Scene.Update (create the context with stream=torch's stream)")


def shard_scenes(num_scenes: int, world_size: int, rank: int) -> List[int]:
    """Scene ids owned by `rank`: contiguous blocks, sizes differing by at most one
    (8 scenes over 1/2/4/8 GPUs -> 8/4/2/1 per GPU; BASELINE.json configs[4]).  zn_shard_scenes."""
    import ctypes as C
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of {world_size}")
    first, count = C.c_uint32(), C.c_uint32()
    _check_value(_lib().zn_shard_scenes(num_scenes, world_size, rank, C.byref(first), C.byref(count)))
    return list(range(first.value, first.value + count.value))


def scene_index_bases(entity_counts: Sequence[int]) -> List[int]:
    """Global index base of every scene (exclusive prefix sum) so shard-local visible indices
    become global by a single add on the device (zn_scene_set_index_base).  zn_scene_index_bases."""
    import ctypes as C

    import numpy as np
    counts = np.ascontiguousarray(entity_counts, np.uint32)
    bases = np.zeros(counts.size, np.uint32)
    _check_value(_lib().zn_scene_index_bases(counts.ctypes.data_as(C.c_void_p), counts.size, bases.ctypes.data_as(C.c_void_p)))
    return [int(b) for b in bases]


NO_PARENT = 0xFFFFFFFF


class SceneShard:
    """One rank's part of a scene split by root subtree (shard_scene_by_roots).

    arrays         the shard as a scene of its own: position/rotation/scale/aabb_center/aabb_half
                   [n,3], parent [n] (shard-local indices, NO_PARENT for roots), level_offsets
                   [levels+1] — what Scene.Load / Scene(ctx, n, level_offsets) take.  A shard that owns
                   no entity of the deepest levels simply has fewer levels.
    global_index   [n] ascending: the scene-wide index of every shard entity.  A shard-local visible
                   list becomes global by global_index[visible]; because the map is monotonic the
                   result is still ascending, so the scene-wide list is the k-way merge (or, with
                   shard-major numbering, the concatenation) of the shards' lists.
    """

    def __init__(self, arrays: dict, global_index):
        self.arrays, self.global_index = arrays, global_index

    @property
    def entity_count(self) -> int:
        return int(self.global_index.size)

    def to_global(self, visible_local):
        return self.global_index[visible_local]


def root_of_entities(level_offsets, parent):
    """[E] scene-wide index of the root every entity hangs from (level-ordered scene: a parent
    always lives in an earlier level, so one pass per level resolves it)."""
    import numpy as np
    level_offsets = np.asarray(level_offsets, np.int64)
    parent = np.asarray(parent, np.uint32)
    root = np.arange(parent.size, dtype=np.int64)
    for l in range(level_offsets.size - 1):
        lo, hi = int(level_offsets[l]), int(level_offsets[l + 1])
        p = parent[lo:hi]
        has = p != NO_PARENT
        if has.any():
            if int(p[has].max()) >= lo:
                raise ValueError(f"level {l}: a parent index is not in an earlier level")
            seg = root[lo:hi]
            seg[has] = root[p[has].astype(np.int64)]
    return root


def shard_scene_by_roots(arrays: dict, world_size: int) -> List["SceneShard"]:
    """Split ONE scene into `world_size` hierarchy-closed shards ("within one scene,
    partition by root subtree, so no cross-GPU parent reads").

    Roots (entities without a parent, in any level) are dealt out in index order as contiguous runs
    whose subtree sizes balance the entity counts; every entity follows its root.  Within a shard the
    scene-wide order is kept, so levels stay level-ordered and children stay sorted by parent — the
    layout contract of zn_scene_set_parents holds for every shard without re-sorting.  The partition
    itself is computed by the C layer (zn_shard_scene_by_roots, the call a C++ engine makes); this
    wrapper only slices the numpy arrays accordingly.
    """
    import ctypes as C

    import numpy as np
    if world_size < 1:
        raise ValueError("world_size must be >= 1")
    level_offsets = np.ascontiguousarray(arrays["level_offsets"], np.uint32)
    E = int(level_offsets[-1])
    parent = arrays.get("parent")
    parent = np.full(E, NO_PARENT, np.uint32) if parent is None else np.ascontiguousarray(parent, np.uint32)
    owner = np.empty(E, np.uint32)
    local = np.empty(E, np.uint32)
    counts = np.empty(world_size, np.uint32)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)   # noqa: E731
    _check_value(_lib().zn_shard_scene_by_roots(vp(level_offsets), level_offsets.size - 1, vp(parent), world_size, vp(owner), vp(local), vp(counts)))
    level_of = np.searchsorted(level_offsets.astype(np.int64), np.arange(E), side="right") - 1
    shards = []
    for r in range(world_size):
        gidx = np.flatnonzero(owner == r)
        assert gidx.size == int(counts[r])
        p = parent[gidx]
        has = p != NO_PARENT
        p_local = np.full(gidx.size, NO_PARENT, np.uint32)
        p_local[has] = local[p[has].astype(np.int64)]
        per_level = np.bincount(level_of[gidx], minlength=level_offsets.size - 1)
        # a shard keeps the scene's level order; levels it owns nothing of are dropped (zn_scene_create
        # rejects empty levels) — parents still live in earlier levels, which is all the contract asks
        offs = np.concatenate([[0], np.cumsum(per_level[per_level > 0])]).astype(np.uint32)
        a = {k: np.ascontiguousarray(np.asarray(arrays[k])[gidx]) for k in ("position", "rotation", "scale", "aabb_center", "aabb_half")}
        a["parent"] = p_local
        a["level_offsets"] = offs
        shards.append(SceneShard(a, gidx.astype(np.uint32)))
    return shards


def merge_visible(lists) -> "object":
    """Scene-wide ascending visible list from the shards' (already global, ascending) lists."""
    import numpy as np
    lists = [np.asarray(v, np.uint32) for v in lists]
    return np.sort(np.concatenate(lists), kind="stable") if lists else np.empty(0, np.uint32)


class VisibleGather:
    """Blocking all-gather of variable-length visible lists.

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


class VisibleExchange:
    """Pipelined all-gather of visible lists: the exchange of frame k runs on the collective's own
    stream while frame k+1 is computed ("overlapped with the gather on a second
    stream").

    Per slot (two by default) it owns this rank's list/count buffers — the scene writes into them
    directly (zn_scene_bind_visible_output) — and the gathered buffers.  Protocol per frame k,
    slot = k % slots:

        ex.acquire(slot)      # frame k-slots' exchange is complete; its result may be collected
        scene.BindVisibleOutput(ex.local_ptr(slot), capacity, ex.count_ptr(slot)); scene.Update()
        ex.launch(slot)       # asynchronous: counts, then lists padded to a size all ranks agree on

    The list all-gather needs a size known on the host.  Reading this frame's counts would stall the
    GPU, so the size is the largest count of the most recently COLLECTED frame plus headroom (identical on all
    ranks, because every rank sees the same gathered counts).  collect() checks the real counts; if
    a list outgrew the guess the frame is re-gathered synchronously, so results are always exact.
    """

    def __init__(self, capacity: int, world_size: int, device, dtype, slots: int = 2, headroom: float = 0.02, group=None, ctx=None):
        import torch
        self.ctx = ctx
        self.world_size, self.capacity, self.slots, self.headroom, self.group = world_size, int(capacity), slots, headroom, group
        self.device = torch.device(device)
        self.cuda = self.device.type == "cuda"
        self.local = [torch.zeros(self.capacity, dtype=dtype, device=device) for _ in range(slots)]
        self.local_count = [torch.zeros(1, dtype=dtype, device=device) for _ in range(slots)]
        self.counts = [torch.zeros(world_size, dtype=dtype, device=device) for _ in range(slots)]
        self.gathered = [torch.empty(world_size * self.capacity, dtype=dtype, device=device) for _ in range(slots)]
        self.pending: List[Optional[tuple]] = [None] * slots
        self.guess: Optional[int] = None
        self.regathers = 0
        if self.cuda:
            # counts travel to the host on a side stream, so reading them never waits on compute
            self.side = torch.cuda.Stream(self.device)
            self.done = [torch.cuda.Event() for _ in range(slots)]
            self.host_counts = [torch.zeros(world_size, dtype=dtype).pin_memory() for _ in range(slots)]

    def local_ptr(self, slot: int) -> int:
        return self.local[slot].data_ptr()

    def count_ptr(self, slot: int) -> int:
        return self.local_count[slot].data_ptr()

    def _padded(self, n: int) -> int:
        n = int(n * (1.0 + self.headroom)) + 1024
        return max(1, min(self.capacity, (n + 1023) // 1024 * 1024))

    def launch(self, slot: int) -> None:
        import torch
        import torch.distributed as dist
        assert self.pending[slot] is None, "slot still holds an uncollected frame"
        if self.cuda:
            check_same_stream(self.ctx)
        w1 = dist.all_gather_into_tensor(self.counts[slot], self.local_count[slot], group=self.group, async_op=True)
        if self.guess is None:
            w1.wait()
            self.guess = self._padded(max(int(c) for c in self.counts[slot].tolist()))   # first frame only: a host sync
        m = self.guess
        out = self.gathered[slot][: self.world_size * m]
        w2 = dist.all_gather_into_tensor(out, self.local[slot][:m], group=self.group, async_op=True)
        if self.cuda:
            with torch.cuda.stream(self.side):
                w1.wait()
                w2.wait()
                self.host_counts[slot].copy_(self.counts[slot], non_blocking=True)
                self.done[slot].record(self.side)
        self.pending[slot] = (w1, w2, m)

    def collect(self, slot: int):
        """Result of the frame last launched on `slot`: (counts, lists[world, m]) or None."""
        import torch
        import torch.distributed as dist
        if self.pending[slot] is None:
            return None
        w1, w2, m = self.pending[slot]
        if self.cuda:
            self.done[slot].synchronize()                              # host: that exchange has landed
            torch.cuda.current_stream(self.device).wait_event(self.done[slot])   # compute stream may reuse the slot
            counts = [int(c) for c in self.host_counts[slot].tolist()]
        else:
            w1.wait()
            w2.wait()
            counts = [int(c) for c in self.counts[slot].tolist()]
        need = max(max(counts), 1)
        if need > m:   # a list outgrew the guess: redo this frame at the right size (rare, exact)
            self.regathers += 1
            m = self._padded(need)
            out = self.gathered[slot][: self.world_size * m]
            dist.all_gather_into_tensor(out, self.local[slot][:m], group=self.group)
        self.guess = self._padded(need)
        self.pending[slot] = None
        return counts, self.gathered[slot][: self.world_size * m].view(self.world_size, m)

    def acquire(self, slot: int):
        """Make `slot` reusable; returns the collected result of its previous frame (or None)."""
        return self.collect(slot)

    def drain(self):
        return [self.collect(s) for s in range(self.slots)]


class RecordExchange:
    """The exchange at 1 bit per entity (the default for N > 1).

    Instead of the index lists (4 bytes per VISIBLE entity, variable length) the ranks all-gather
    the frame's fixed-size visibility record (per-chunk counts + one mask bit per entity: 2.1 MB for
    a 16.7M-entity scene against 57 MB of indices at f = 0.86) and every rank expands all records
    into the ordered lists with the compaction kernel's own arithmetic
    (zn_scene_expand_visible).  Every rank ends up with exactly the lists and counts an
    all-gather of the lists would have delivered; NVLink carries 27x fewer bytes and, the size being
    fixed, no count has to reach the host before the collective can be enqueued.

    Double-buffered like VisibleExchange: the scene writes frame k's record into slot k%2
    (zn_scene_bind_visibility_record), the all-gather of frame k runs on the collective's stream
    while frame k+1 is computed, and frame k is expanded right after frame k+1's update.
    """

    def __init__(self, record_words: int, capacity: int, world_size: int, device, dtype, slots: int = 2, group=None,
                 records_per_rank: int = 1, ctx=None):
        import torch
        self.ctx = ctx
        self.world_size, self.record_words, self.capacity, self.slots, self.group = world_size, int(record_words), int(capacity), slots, group
        self.records_per_rank = int(records_per_rank)          # a rank that owns several scenes ships them together
        self.total_records = world_size * self.records_per_rank
        rank_words = self.record_words * self.records_per_rank
        self.local = [torch.zeros(rank_words, dtype=dtype, device=device) for _ in range(slots)]
        self.gathered = [torch.empty(world_size * rank_words, dtype=dtype, device=device) for _ in range(slots)]
        self.lists = torch.empty(self.total_records * self.capacity, dtype=dtype, device=device)
        self.counts = torch.zeros(self.total_records, dtype=dtype, device=device)
        self.work = [None] * slots

    def record_ptr(self, slot: int, j: int = 0) -> int:
        """Device address where this rank's j-th scene writes its record for `slot`."""
        return self.local[slot].data_ptr() + 4 * self.record_words * j

    def launch(self, slot: int) -> None:
        import torch.distributed as dist
        assert self.work[slot] is None, "slot still holds an unexpanded frame"
        if self.gathered[slot].is_cuda:
            check_same_stream(self.ctx)
        self.work[slot] = dist.all_gather_into_tensor(self.gathered[slot], self.local[slot], group=self.group, async_op=True)

    def ready(self, slot: int) -> bool:
        """Make the current stream wait for the slot's all-gather; False if nothing is pending."""
        if self.work[slot] is None:
            return False
        self.work[slot].wait()
        self.work[slot] = None
        return True

    def expand(self, scene, slot: int, index_bases=None) -> None:
        """Expand the gathered records of `slot` into self.lists / self.counts on the scene's stream."""
        scene.ExpandVisible(self.gathered[slot].data_ptr(), self.total_records, self.record_words, self.lists.data_ptr(), self.capacity,
                            self.counts.data_ptr(), index_bases=index_bases)

    def result(self):
        """(counts, lists[records, capacity]) of the most recently expanded frame (synchronises)."""
        counts = [int(c) for c in self.counts.tolist()]
        return counts, self.lists.view(self.total_records, self.capacity)
