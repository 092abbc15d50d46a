"""zn_hierarchy_sort: entities in creation order -> the level-ordered, sorted-by-parent layout the
scene path requires.  Host code of the product library; runs without a GPU."""
import ctypes as C

import numpy as np
import pytest

NP = 0xFFFFFFFF


def random_forest(rng, level_sizes, orphan_rate=0.0):
    """A valid layout to start from: level-ordered, children sorted by parent."""
    offs = np.concatenate([[0], np.cumsum(level_sizes)]).astype(np.uint32)
    parent = np.full(int(offs[-1]), NP, np.uint32)
    for l in range(1, len(level_sizes)):
        lo, hi = int(offs[l]), int(offs[l + 1])
        p = np.sort(rng.integers(int(offs[l - 1]), lo, hi - lo)).astype(np.uint32)
        p[rng.random(hi - lo) < orphan_rate] = NP
        parent[lo:hi] = p
    return parent, offs


def shuffled(rng, parent):
    """The same forest with its entities in random ('creation') order."""
    n = parent.size
    new_of_old = rng.permutation(n).astype(np.uint32)
    old_of_new = np.argsort(new_of_old).astype(np.uint32)
    p = parent[old_of_new]
    out = np.where(p == NP, NP, new_of_old[np.minimum(p, n - 1)]).astype(np.uint32)
    return out, old_of_new


def check_layout(parent, order, new_parent, offsets):
    n = parent.size
    assert np.array_equal(np.sort(order), np.arange(n))                       # a permutation
    new_of_old = np.empty(n, np.int64)
    new_of_old[order] = np.arange(n)
    p_old = parent[order]
    expect = np.where(p_old == NP, NP, new_of_old[np.minimum(p_old, n - 1)]).astype(np.uint32)
    assert np.array_equal(new_parent, expect)                                  # same forest, renumbered
    assert offsets[0] == 0 and offsets[-1] == n and np.all(np.diff(offsets.astype(np.int64)) > 0)
    level_of = np.searchsorted(offsets, np.arange(n), side="right") - 1
    has = new_parent != NP
    assert np.all(level_of[new_parent[has]] == level_of[has] - 1)              # a child sits exactly one level below its parent
    assert np.all(level_of[~has] == 0)                                          # every root in level 0
    for l in range(offsets.size - 1):
        lo, hi = int(offsets[l]), int(offsets[l + 1])
        seg = new_parent[lo:hi].astype(np.int64)
        assert np.all(np.diff(seg) >= 0)                                        # sorted by parent (roots, all NP, trivially)
        same = np.diff(seg) == 0
        assert np.all(np.diff(order[lo:hi].astype(np.int64))[same] > 0)         # siblings keep their original order


def test_small_known_case():
    import zenyth_b200 as zn
    #            0   1  2  3  4   5  6  7
    parent = [NP, 0, 0, 1, NP, 4, 3, 3]
    order, new_parent, offsets = zn.SortHierarchy(parent)
    assert order.tolist() == [0, 4, 1, 2, 5, 3, 6, 7]
    assert new_parent.tolist() == [NP, NP, 0, 0, 1, 2, 5, 5]
    assert offsets.tolist() == [0, 2, 5, 6, 8]


@pytest.mark.parametrize("seed,sizes,orphans", [(1, [1], 0.0), (2, [5, 40, 300, 2000, 900], 0.0), (3, [3, 3, 3, 3, 3, 3, 3], 0.3),
                                                (4, [1000, 10], 0.0), (5, [1] * 150, 0.0)])
def test_random_forests_in_creation_order(oracle, seed, sizes, orphans):
    import zenyth_b200 as zn
    from oracle.binding import camera
    rng = np.random.default_rng(seed)
    parent_sorted, offs = random_forest(rng, sizes, orphans)
    parent, old_of_new = shuffled(rng, parent_sorted)
    order, new_parent, offsets = zn.SortHierarchy(parent)
    check_layout(parent, order, new_parent, offsets)
    if not orphans:
        assert np.array_equal(np.diff(offsets), sizes)
    # the path computes the same thing for every entity, whatever order the engine created them in
    n = parent.size
    fields = {"position": rng.uniform(-50, 50, (n, 3)), "rotation": rng.uniform(-3, 3, (n, 3)), "scale": rng.uniform(0.8, 1.2, (n, 3)),
              "aabb_center": rng.uniform(-0.2, 0.2, (n, 3)), "aabb_half": rng.uniform(0.1, 1.0, (n, 3))}
    fields = {k: v.astype(np.float32) for k, v in fields.items()}           # indexed like parent_sorted
    view, proj = camera(oracle, eye=(0.0, 0.0, 150.0))
    if orphans:   # random_forest with orphans is not a valid start layout to compare against; the contract check above is the test
        return
    a = dict(fields, parent=parent_sorted, level_offsets=offs)
    ref = oracle.scene_update(a, view, proj)
    created = {k: v[old_of_new] for k, v in fields.items()}                 # the engine's creation order
    c = {k: v[order] for k, v in created.items()}
    c.update(parent=new_parent, level_offsets=offsets)
    got = oracle.scene_update(c, view, proj)
    original = old_of_new[order]                                            # index in the start layout of every sorted entity
    assert np.array_equal(got["world"], ref["world"][original])
    assert np.array_equal(got["mvp"], ref["mvp"][original])
    assert np.array_equal(np.sort(original[got["visible"]]), ref["visible"])


def test_sorting_a_sorted_layout_changes_nothing():
    import zenyth_b200 as zn
    parent, offs = random_forest(np.random.default_rng(9), [4, 30, 200, 50])
    order, new_parent, offsets = zn.SortHierarchy(parent)
    assert np.array_equal(order, np.arange(parent.size)) and np.array_equal(new_parent, parent) and np.array_equal(offsets, offs)


def test_bad_hierarchies_are_rejected():
    import zenyth_b200 as zn
    for bad in ([1, 0, NP], [NP, 1], [NP, 7], [1, 2, 0]):          # cycle beside a root, self-parent, out of range, nothing but a cycle
        with pytest.raises(zn.ZenythError):
            zn.SortHierarchy(bad)


def test_level_capacity_is_reported_not_overrun():
    from zenyth_b200 import _capi
    lib = _capi.load()
    parent = np.array([NP, 0, 1, 2, 3], np.uint32)                  # five levels need six offsets
    order = np.empty(5, np.uint32)
    offsets = np.full(8, 0xDEADBEEF, np.uint32)
    levels = C.c_uint32(0)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    assert lib.zn_hierarchy_sort(p(parent), 5, p(order), None, p(offsets), 3, C.byref(levels)) == 1
    assert levels.value == 5 and np.all(offsets[3:] == 0xDEADBEEF)
    assert lib.zn_hierarchy_sort(p(parent), 5, p(order), None, p(offsets), 6, C.byref(levels)) == 0
    assert offsets[:6].tolist() == [0, 1, 2, 3, 4, 5] and offsets[6] == 0xDEADBEEF
