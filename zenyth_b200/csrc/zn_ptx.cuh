// zn_ptx.cuh — thin inline-PTX wrappers for the sm_100a async-copy machinery used by the scene
// and mesh kernels: mbarrier, 1-D bulk copies (cp.async.bulk, SASS UBLKCP) and 2-D tensor-map
// stores (cp.async.bulk.tensor, SASS UTMASTG).
#pragma once
#include <cstdint>
#include <cuda.h>

namespace zn {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
	return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier ---------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t arrive_count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(arrive_count) : "memory");
}
// Makes a freshly initialised mbarrier visible to the async proxy (TMA unit).
__device__ __forceinline__ void fence_mbar_init() {
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
	asm volatile(
		"{\n\t"
		".reg .pred p;\n\t"
		"ZN_WAIT_%=:\n\t"
		"mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
		"@p bra ZN_DONE_%=;\n\t"
		"bra ZN_WAIT_%=;\n\t"
		"ZN_DONE_%=:\n\t"
		"}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// ---- proxy fence: generic-proxy smem writes -> visible to the async proxy ----------------------
__device__ __forceinline__ void fence_proxy_async_smem() {
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---- 1-D bulk copies (16-byte aligned addresses, size a multiple of 16) ------------------------
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
	             ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* gmem_dst, const void* smem_src, uint32_t bytes) {
	asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
	             ::"l"(gmem_dst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
}

// ---- 2-D tensor-map store: smem tile -> global, hardware un-swizzle and bounds clipping --------
__device__ __forceinline__ void tensor_store_2d(const CUtensorMap* map, const void* smem_src, int32_t c0, int32_t c1) {
	asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
	             ::"l"(map), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1) : "memory");
}
// 2-D tensor-map load: global -> smem tile in the map's swizzle pattern; rows past the tensor's
// extent arrive as zeros and still count towards the transaction bytes (always the full box).
__device__ __forceinline__ void tensor_load_2d(void* smem_dst, const CUtensorMap* map, int32_t c0, int32_t c1, uint64_t* bar) {
	asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
	             ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap* map) {
	asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// Wait until the sources (smem) of all committed bulk stores have been read.
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// ... of all but the latest N committed groups.
template <int N>
__device__ __forceinline__ void bulk_wait_read_but() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
// Wait until all committed bulk stores are complete (writes performed).
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ---- programmatic dependent launch (PDL) ---------------------------------------------------------
// wait: block until the preceding kernel in the stream has completed and its writes are visible.
// launch: allow the following kernel (launched with programmatic stream serialization) to start.
__device__ __forceinline__ void grid_dep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void grid_dep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---- per-thread asynchronous copies global -> shared (LDGSTS), 4 bytes each, completion by commit groups -----
__device__ __forceinline__ void cp_async_4(void* smem_dst, const void* gmem_src) {
	asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// A store that is PREDICATED, not branched around: nvcc turns `if (c) *p = v;` into a reconvergence
// region (BSSY / BRA / BSYNC) per store, which serialises a run of independent conditional stores.
__device__ __forceinline__ void st_global_if(uint32_t* p, uint32_t v, uint32_t cond) {
	asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q st.global.u32 [%0], %1;\n\t}" ::"l"(p), "r"(v), "r"(cond) : "memory");
}

// ---- cache-hinted global access ----------------------------------------------------------------
__device__ __forceinline__ float ld_stream(const float* p) {
	float v;
	asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
	return v;
}
__device__ __forceinline__ uint32_t ld_stream(const uint32_t* p) {
	uint32_t v;
	asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
	return v;
}

} // namespace zn
