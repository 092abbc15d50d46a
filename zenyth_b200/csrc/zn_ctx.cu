// zn_ctx.cu — context, error reporting, tensor-map encoding.
#include "zn_internal.hpp"

#include <cstring>

namespace zn {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(g_error, sizeof(g_error), fmt, ap);
	va_end(ap);
}

int fail(int code, const char* fmt, ...) {
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(g_error, sizeof(g_error), fmt, ap);
	va_end(ap);
	return code;
}

// cuTensorMapEncodeTiled is a driver-API entry point; fetch it through the runtime so the
// library does not link libcuda directly.
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
	static PFN_encodeTiled fn = nullptr;
	if (!fn) {
		void* p = nullptr;
		cudaDriverEntryPointQueryResult qres;
		if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
		    qres == cudaDriverEntryPointSuccess)
			fn = reinterpret_cast<PFN_encodeTiled>(p);
	}
	return fn;
}

int encode_matrix_tensor_map(CUtensorMap* out, void* base, uint64_t rows) {
	PFN_encodeTiled enc = get_encode_fn();
	if (!enc) return fail(ZN_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
	if (rows == 0) rows = 1;
	const cuuint64_t dims[2] = {16, rows};              // innermost first: 16 floats per matrix
	const cuuint64_t strides[1] = {64};                 // bytes between rows (dimension 1)
	const cuuint32_t box[2] = {16, static_cast<cuuint32_t>(kTile)};
	const cuuint32_t elem_strides[2] = {1, 1};
	const CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, dims, strides, box, elem_strides,
	                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
	                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	if (r != CUDA_SUCCESS) return fail(ZN_ERR_CUDA, "cuTensorMapEncodeTiled failed: CUresult %d", int(r));
	return ZN_OK;
}

} // namespace zn

extern "C" {

int zn_abi_version(void) { return ZN_ABI_VERSION; }

const char* zn_last_error(void) { return zn::g_error; }

int zn_device_count(int* out_count) {
	ZN_REQUIRE(out_count, "zn_device_count: out_count is NULL");
	int n = 0;
	const cudaError_t e = cudaGetDeviceCount(&n);
	if (e != cudaSuccess) {
		cudaGetLastError();
		*out_count = 0;
		return zn::fail(ZN_ERR_NO_DEVICE, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
	}
	*out_count = n;
	return ZN_OK;
}

int zn_ctx_create(int device_ordinal, zn_ctx** out_ctx) {
	ZN_REQUIRE(out_ctx, "zn_ctx_create: out_ctx is NULL");
	*out_ctx = nullptr;
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
		cudaGetLastError();
		return zn::fail(ZN_ERR_NO_DEVICE, "zn_ctx_create: no CUDA device is visible; this library has no CPU fallback");
	}
	ZN_REQUIRE(device_ordinal >= 0 && device_ordinal < n, "zn_ctx_create: device %d out of range [0,%d)", device_ordinal, n);
	cudaDeviceProp prop;
	ZN_CUDA(cudaGetDeviceProperties(&prop, device_ordinal));
	if (prop.major != 10)
		return zn::fail(ZN_ERR_NO_DEVICE, "zn_ctx_create: device %d is sm_%d%d; kernels are built for sm_100a only",
		                device_ordinal, prop.major, prop.minor);
	ZN_CUDA(cudaSetDevice(device_ordinal));
	zn_ctx* c = new zn_ctx();
	c->device = device_ordinal;
	c->sm_count = prop.multiProcessorCount;
	// The context's own stream gets the highest priority the device offers: side work that overlaps a frame
	// (the exchange's expansion kernels, zn_exchange_*) runs at the lowest, so that the block scheduler
	// hands freed SM slots to the frame's latency-bound small levels first.
	int least = 0, greatest = 0;
	cudaDeviceGetStreamPriorityRange(&least, &greatest);
	const cudaError_t e = cudaStreamCreateWithPriority(&c->own_stream, cudaStreamNonBlocking, greatest);
	if (e != cudaSuccess) {
		delete c;
		return zn::fail(ZN_ERR_CUDA, "zn_ctx_create: cudaStreamCreateWithPriority failed: %s", cudaGetErrorString(e));
	}
	c->stream = c->own_stream;
	*out_ctx = c;
	return ZN_OK;
}

int zn_ctx_destroy(zn_ctx* ctx) {
	if (!ctx) return ZN_OK;
	cudaSetDevice(ctx->device);
	cudaStreamSynchronize(ctx->stream);
	if (ctx->staging) cudaFree(ctx->staging);
	if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
	delete ctx;
	return ZN_OK;
}

int zn_ctx_set_stream(zn_ctx* ctx, void* cuda_stream) {
	ZN_REQUIRE(ctx, "zn_ctx_set_stream: ctx is NULL");
	ctx->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : ctx->own_stream;
	return ZN_OK;
}

int zn_ctx_get_stream(zn_ctx* ctx, void** out_cuda_stream) {
	ZN_REQUIRE(ctx && out_cuda_stream, "zn_ctx_get_stream: NULL argument");
	*out_cuda_stream = ctx->stream;
	return ZN_OK;
}

int zn_ctx_sync(zn_ctx* ctx) {
	ZN_REQUIRE(ctx, "zn_ctx_sync: ctx is NULL");
	ZN_CUDA(cudaSetDevice(ctx->device));
	ZN_CUDA(cudaStreamSynchronize(ctx->stream));
	return ZN_OK;
}

int zn_ctx_sm_count(zn_ctx* ctx, int* out_sm_count) {
	ZN_REQUIRE(ctx && out_sm_count, "zn_ctx_sm_count: NULL argument");
	*out_sm_count = ctx->sm_count;
	return ZN_OK;
}

int zn_host_alloc(zn_ctx* ctx, size_t bytes, void** out_ptr) {
	ZN_REQUIRE(ctx && out_ptr, "zn_host_alloc: NULL argument");
	ZN_CUDA(cudaSetDevice(ctx->device));
	ZN_CUDA(cudaHostAlloc(out_ptr, bytes ? bytes : 1, cudaHostAllocDefault));
	return ZN_OK;
}

int zn_host_free(zn_ctx* ctx, void* ptr) {
	ZN_REQUIRE(ctx, "zn_host_free: ctx is NULL");
	if (ptr) ZN_CUDA(cudaFreeHost(ptr));
	return ZN_OK;
}

// ---- peer memory: device allocations that other ranks (processes) of the node can map ----------
int zn_device_alloc(zn_ctx* ctx, size_t bytes, void** out_ptr) {
	ZN_REQUIRE(ctx && out_ptr, "zn_device_alloc: NULL argument");
	ZN_CUDA(cudaSetDevice(ctx->device));
	ZN_CUDA(cudaMalloc(out_ptr, bytes ? bytes : 1));
	ZN_CUDA(cudaMemsetAsync(*out_ptr, 0, bytes ? bytes : 1, ctx->stream));
	ZN_CUDA(cudaStreamSynchronize(ctx->stream));
	return ZN_OK;
}

int zn_device_free(zn_ctx* ctx, void* ptr) {
	ZN_REQUIRE(ctx, "zn_device_free: ctx is NULL");
	ZN_CUDA(cudaSetDevice(ctx->device));
	if (ptr) ZN_CUDA(cudaFree(ptr));
	return ZN_OK;
}

int zn_device_read(zn_ctx* ctx, const void* dev_ptr, void* host_dst, size_t bytes) {
	ZN_REQUIRE(ctx && (bytes == 0 || (dev_ptr && host_dst)), "zn_device_read: NULL argument");
	ZN_CUDA(cudaSetDevice(ctx->device));
	if (bytes) ZN_CUDA(cudaMemcpyAsync(host_dst, dev_ptr, bytes, cudaMemcpyDeviceToHost, ctx->stream));
	ZN_CUDA(cudaStreamSynchronize(ctx->stream));
	return ZN_OK;
}

int zn_device_write(zn_ctx* ctx, void* dev_ptr, const void* host_src, size_t bytes) {
	ZN_REQUIRE(ctx && (bytes == 0 || (dev_ptr && host_src)), "zn_device_write: NULL argument");
	ZN_CUDA(cudaSetDevice(ctx->device));
	if (bytes) ZN_CUDA(cudaMemcpyAsync(dev_ptr, host_src, bytes, cudaMemcpyHostToDevice, ctx->stream));
	ZN_CUDA(cudaStreamSynchronize(ctx->stream));
	return ZN_OK;
}

int zn_ipc_export(zn_ctx* ctx, void* dev_ptr, unsigned char out_handle[64]) {
	ZN_REQUIRE(ctx && dev_ptr && out_handle, "zn_ipc_export: NULL argument");
	static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
	ZN_CUDA(cudaSetDevice(ctx->device));
	cudaIpcMemHandle_t h;
	ZN_CUDA(cudaIpcGetMemHandle(&h, dev_ptr));
	memcpy(out_handle, &h, 64);
	return ZN_OK;
}

int zn_ipc_open(zn_ctx* ctx, const unsigned char handle[64], void** out_dev_ptr) {
	ZN_REQUIRE(ctx && handle && out_dev_ptr, "zn_ipc_open: NULL argument");
	ZN_CUDA(cudaSetDevice(ctx->device));
	cudaIpcMemHandle_t h;
	memcpy(&h, handle, 64);
	ZN_CUDA(cudaIpcOpenMemHandle(out_dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
	return ZN_OK;
}

int zn_ipc_close(zn_ctx* ctx, void* dev_ptr) {
	ZN_REQUIRE(ctx, "zn_ipc_close: ctx is NULL");
	ZN_CUDA(cudaSetDevice(ctx->device));
	if (dev_ptr) ZN_CUDA(cudaIpcCloseMemHandle(dev_ptr));
	return ZN_OK;
}

} // extern "C"
