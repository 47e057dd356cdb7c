// Host-side utilities: error reporting, device query, TMA descriptor encoding.
#include "common.cuh"

#include <cudaTypedefs.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

namespace nk {

static thread_local char g_last_error[512] = "";

void set_last_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
    va_end(ap);
}

const char* last_error() { return g_last_error; }

int check_cuda(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return NK_OK;
    set_last_error("CUDA error %d (%s) in %s", static_cast<int>(e), cudaGetErrorString(e), what);
    return NK_ERR_CUDA;
}

int device_sm_count() {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
    return n;
}

// ---- driver context of the calling thread -----------------------------------------------------------------------
// Entry points may be called from threads that never touched CUDA (PyTorch's autograd worker threads): the driver
// then has no current context and raw driver calls (TMA descriptor encoding) fail with CUDA_ERROR_INVALID_CONTEXT,
// while runtime calls would silently bind device 0.  Every entry point therefore adopts the context the caller's
// stream belongs to.
typedef CUresult (*PFN_ctxGetCurrent)(CUcontext*);
typedef CUresult (*PFN_ctxSetCurrent)(CUcontext);
typedef CUresult (*PFN_streamGetCtx)(CUstream, CUcontext*);

static void* driver_fn(const char* name) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
        return nullptr;
    return p;
}

cudaStream_t enter(void* stream) {
    static PFN_ctxGetCurrent get_cur = reinterpret_cast<PFN_ctxGetCurrent>(driver_fn("cuCtxGetCurrent"));
    static PFN_ctxSetCurrent set_cur = reinterpret_cast<PFN_ctxSetCurrent>(driver_fn("cuCtxSetCurrent"));
    static PFN_streamGetCtx stream_ctx = reinterpret_cast<PFN_streamGetCtx>(driver_fn("cuStreamGetCtx"));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    static CUcontext process_ctx = nullptr;  // one process per GPU: the first context seen is the process's context
    if (get_cur && set_cur) {
        CUcontext cur = nullptr;
        if (get_cur(&cur) == CUDA_SUCCESS) {
            if (cur != nullptr) {
                if (process_ctx == nullptr) process_ctx = cur;
            } else {
                CUcontext want = nullptr;
                if (st != nullptr && stream_ctx && stream_ctx(reinterpret_cast<CUstream>(st), &want) == CUDA_SUCCESS && want)
                    set_cur(want);
                else if (process_ctx != nullptr)
                    set_cur(process_ctx);  // legacy default stream: fall back to the context used so far
                else
                    cudaFree(nullptr);  // nothing known yet: bind the runtime's current device
            }
        }
    }
    return st;
}

static PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    if (fn) return fn;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !p) {
        set_last_error("cudaGetDriverEntryPoint(cuTensorMapEncodeTiled) failed: %d", static_cast<int>(e));
        return nullptr;
    }
    fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
    return fn;
}

int encode_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                     const uint64_t* strides_bytes, const uint32_t* box) {
    return encode_tmap(out, base, rank, dims, strides_bytes, box, 0, nullptr);
}

int encode_tmap(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                const uint32_t* box, int is_f32, const uint32_t* elem_strides) {
    PFN_cuTensorMapEncodeTiled_v12000 fn = get_encode_fn();
    if (!fn) return NK_ERR_CUDA;
    NK_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15u) == 0, NK_ERR_SHAPE,
               "tensor map: base %p not 16B aligned", base);
    for (int i = 0; i + 1 < rank; ++i)
        NK_REQUIRE((strides_bytes[i] & 15u) == 0 && strides_bytes[i] > 0, NK_ERR_SHAPE,
                   "tensor map: stride[%d]=%llu bytes not a positive multiple of 16", i,
                   static_cast<unsigned long long>(strides_bytes[i]));
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    cuuint64_t gdim[5];
    cuuint64_t gstr[4];
    cuuint32_t bdim[5];
    for (int i = 0; i < rank; ++i) {
        gdim[i] = dims[i];
        bdim[i] = box[i];
        if (elem_strides) estr[i] = elem_strides[i];
        NK_REQUIRE(dims[i] >= 1 && box[i] >= 1 && box[i] <= 256, NK_ERR_SHAPE,
                   "tensor map: dim[%d]=%llu box=%u out of range", i,
                   static_cast<unsigned long long>(dims[i]), box[i]);
    }
    for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
    CUresult r = fn(out, is_f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                    static_cast<cuuint32_t>(rank),
                    const_cast<void*>(base), gdim, gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_last_error(
            "cuTensorMapEncodeTiled failed (%d): dims [%llu,%llu,%llu,%llu] strides [%llu,%llu,%llu] "
            "box [%u,%u,%u,%u]",
            static_cast<int>(r), (unsigned long long)dims[0], (unsigned long long)dims[1],
            (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0),
            (unsigned long long)strides_bytes[0], (unsigned long long)(rank > 2 ? strides_bytes[1] : 0),
            (unsigned long long)(rank > 3 ? strides_bytes[2] : 0), box[0], box[1], rank > 2 ? box[2] : 0,
            rank > 3 ? box[3] : 0);
        return NK_ERR_CUDA;
    }
    return NK_OK;
}

}  // namespace nk
