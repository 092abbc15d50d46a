"""Property-based checks (hypothesis) of the host-side layout helpers: whatever forest an engine
throws at them, zn_hierarchy_sort and shard_scene_by_roots return layouts that honour the upload
contract of the scene path.  No GPU."""
import numpy as np
from hypothesis import given, settings
from hypothesis import strategies as st

from test_hierarchy_sort import NP, check_layout


@st.composite
def forests(draw, max_entities=200):
    """A random forest in arbitrary ('creation') order: every entity picks NO parent or any entity
    created before it, then the whole thing is renumbered by a random permutation."""
    n = draw(st.integers(1, max_entities))
    parent = [NP]
    for e in range(1, n):
        parent.append(NP if draw(st.integers(0, 9)) == 0 else draw(st.integers(0, e - 1)))
    perm = draw(st.permutations(range(n)))
    new_of_old = np.asarray(perm, np.uint32)
    old_of_new = np.argsort(new_of_old)
    p = np.asarray(parent, np.uint32)[old_of_new]
    return np.where(p == NP, NP, new_of_old[np.minimum(p, n - 1)]).astype(np.uint32)


@settings(max_examples=150, deadline=None)
@given(forests())
def test_hierarchy_sort_always_yields_a_valid_layout(parent):
    import zenyth_b200 as zn
    order, new_parent, offsets = zn.SortHierarchy(parent)
    check_layout(parent, order, new_parent, offsets)
    # idempotent: a sorted layout sorts to itself
    order2, parent2, offsets2 = zn.SortHierarchy(new_parent)
    assert np.array_equal(order2, np.arange(parent.size)) and np.array_equal(parent2, new_parent) and np.array_equal(offsets2, offsets)


@settings(max_examples=100, deadline=None)
@given(forests(), st.integers(1, 6))
def test_sharding_a_sorted_forest_partitions_it_into_closed_valid_shards(parent, world):
    import pytest
    import zenyth_b200 as zn
    from zenyth_b200.distributed import shard_scene_by_roots
    order, new_parent, offsets = zn.SortHierarchy(parent)
    n = parent.size
    fields = {k: np.arange(3 * n, dtype=np.float32).reshape(n, 3) + i for i, k in
              enumerate(("position", "rotation", "scale", "aabb_center", "aabb_half"))}
    scene = dict(fields, parent=new_parent, level_offsets=offsets)
    roots = int(np.count_nonzero(new_parent == NP))
    if roots < world:
        with pytest.raises(ValueError):
            shard_scene_by_roots(scene, world)
        return
    shards = shard_scene_by_roots(scene, world)
    assert len(shards) == world and all(s.entity_count > 0 for s in shards)
    assert np.array_equal(np.sort(np.concatenate([s.global_index for s in shards])), np.arange(n))
    for s in shards:
        g = s.global_index.astype(np.int64)
        assert np.all(np.diff(g) > 0)
        p = s.arrays["parent"]
        has = p != NP
        assert np.array_equal(g[p[has]], new_parent[g[has]])                       # closed under "parent of"
        assert np.array_equal(s.arrays["position"], fields["position"][g])          # data follows the entities
        # a shard is itself a sorted layout: sorting it again changes nothing
        o, p2, offs = zn.SortHierarchy(p)
        assert np.array_equal(o, np.arange(p.size)) and np.array_equal(offs, s.arrays["level_offsets"])
