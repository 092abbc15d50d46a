// zn_hierarchy.cpp — the layout step in front of the path: an engine keeps its entities in whatever
// order they were created, each naming its parent; zn_scene_set_parents wants them level by level
// (a parent always in an earlier level) with the children of a level sorted by parent, so that a
// tile's parents are one contiguous run of world matrices.  zn_hierarchy_sort derives that order.
//
// The reference has no Transform or Scene to take an order from; this is the
// counterpart of what a Scene::AddEntity / SetParent pair would maintain.  Host code only, O(E):
// children lists in CSR form, then a breadth-first walk from the roots — breadth-first order IS
// level order, and visiting the parents of a level in their new order lists their children sorted
// by (new) parent index.  Siblings keep their original relative order, so the result is unique.
#include "zn_internal.hpp"

#include <vector>

extern "C" {

int zn_hierarchy_sort(const uint32_t* parent, uint32_t count, uint32_t* out_order, uint32_t* out_parent,
                      uint32_t* out_level_offsets, uint32_t level_capacity, uint32_t* out_level_count) {
	ZN_REQUIRE(parent && out_order && out_level_count && count > 0, "zn_hierarchy_sort: NULL argument or count == 0");
	ZN_REQUIRE(out_level_offsets || level_capacity == 0, "zn_hierarchy_sort: out_level_offsets is NULL");
	// children of every entity, CSR: first_child[p] .. first_child[p+1], in original index order
	std::vector<uint32_t> first_child(size_t(count) + 1, 0u);
	uint32_t roots = 0;
	for (uint32_t e = 0; e < count; ++e) {
		const uint32_t p = parent[e];
		if (p == ZN_NO_PARENT) { ++roots; continue; }
		ZN_REQUIRE(p < count, "zn_hierarchy_sort: entity %u names parent %u, but there are only %u entities", e, p, count);
		ZN_REQUIRE(p != e, "zn_hierarchy_sort: entity %u is its own parent", e);
		++first_child[p + 1];
	}
	ZN_REQUIRE(roots > 0, "zn_hierarchy_sort: no entity without a parent (every entity sits on a cycle)");
	for (uint32_t p = 0; p < count; ++p) first_child[p + 1] += first_child[p];
	std::vector<uint32_t> children(count - roots);
	{
		std::vector<uint32_t> fill(first_child.begin(), first_child.end() - 1);
		for (uint32_t e = 0; e < count; ++e)
			if (parent[e] != ZN_NO_PARENT) children[fill[parent[e]]++] = e;
	}
	// breadth-first from the roots; out_order doubles as the queue (new index -> original index)
	std::vector<uint32_t> new_index(count, ZN_NO_PARENT);
	uint32_t tail = 0;
	for (uint32_t e = 0; e < count; ++e)
		if (parent[e] == ZN_NO_PARENT) { new_index[e] = tail; out_order[tail++] = e; }
	uint32_t levels = 0, level_begin = 0;
	while (level_begin < tail) {
		if (levels < level_capacity) out_level_offsets[levels] = level_begin;
		++levels;
		const uint32_t level_end = tail;
		for (uint32_t at = level_begin; at < level_end; ++at) {
			const uint32_t p = out_order[at];
			for (uint32_t k = first_child[p]; k < first_child[p + 1]; ++k) {
				const uint32_t c = children[k];
				new_index[c] = tail;
				out_order[tail++] = c;
			}
		}
		level_begin = level_end;
	}
	*out_level_count = levels;
	if (tail != count) {
		uint32_t stuck = 0;
		while (stuck < count && new_index[stuck] != ZN_NO_PARENT) ++stuck;
		return zn::fail(ZN_ERR_INVALID, "zn_hierarchy_sort: %u entities are not reachable from a root (parent cycle through entity %u)",
		                count - tail, stuck);
	}
	ZN_REQUIRE(levels + 1 <= level_capacity, "zn_hierarchy_sort: the hierarchy has %u levels; out_level_offsets holds %u entries, needs %u",
	           levels, level_capacity, levels + 1);
	out_level_offsets[levels] = count;
	if (out_parent)
		for (uint32_t n = 0; n < count; ++n) {
			const uint32_t p = parent[out_order[n]];
			out_parent[n] = p == ZN_NO_PARENT ? ZN_NO_PARENT : new_index[p];
		}
	return ZN_OK;
}

} // extern "C"
