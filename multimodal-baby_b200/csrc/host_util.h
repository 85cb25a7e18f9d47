// Host-side helpers shared by the C-ABI entry points: error reporting, TMA tensor-map
// construction (driver entry point fetched at run time, so the library links only cudart),
// launch helpers.
#pragma once
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <cuda.h>
#include <cuda_runtime.h>

namespace cvcl {

// thread-local message returned by cvcl_last_error()
inline char* last_error_buf() {
    static thread_local char buf[512] = {0};
    return buf;
}
inline int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(last_error_buf(), 512, fmt, ap);
    va_end(ap);
    return code;
}

// error codes: the macros of include/cvcl_b200.h (repeated here for stand-alone tools)
#ifndef CVCL_OK
#define CVCL_OK 0
#define CVCL_ERR_INVALID (-1)      // bad shape / null pointer / misalignment
#define CVCL_ERR_UNSUPPORTED (-2)  // valid request the kernels do not cover (never a fallback)
#define CVCL_ERR_CUDA (-3)         // CUDA runtime / driver error
#endif

#define CVCL_CHECK_CUDA(expr)                                                              \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess)                                                             \
            return ::cvcl::fail(CVCL_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,     \
                                cudaGetErrorString(_e), __FILE__, __LINE__);               \
    } while (0)

#define CVCL_REQUIRE(cond, ...)                                                            \
    do {                                                                                   \
        if (!(cond)) return ::cvcl::fail(CVCL_ERR_INVALID, __VA_ARGS__);           \
    } while (0)

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode_fn() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
                cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(p);
    }
    return fn;
}

// Row-major matrix [rows, cols] of bf16 (elem = 2) or fp32 (elem = 4) with leading dimension ld
// (elements); box = box_rows x box_cols, box_cols * elem == 128 bytes (one swizzle atom).
// Loads: out-of-bounds box elements read as zero; stores: they are clipped.
inline int make_tmap(CUtensorMap* map, const void* ptr, int elem, uint64_t rows, uint64_t cols,
                     uint64_t ld, uint32_t box_cols, uint32_t box_rows) {
    PFN_encodeTiled enc = get_encode_fn();
    if (!enc) return fail(CVCL_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    if ((reinterpret_cast<uintptr_t>(ptr) & 15) || ((ld * elem) & 15))
        return fail(CVCL_ERR_INVALID, "TMA tensor must be 16-byte aligned with a 16-byte row pitch "
                                      "(ptr=%p ld=%llu elem=%d)", ptr, (unsigned long long)ld, elem);
    if (rows == 0 || cols == 0) return fail(CVCL_ERR_INVALID, "empty TMA tensor");
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstr[1] = {ld * elem};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, elem == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                     const_cast<void*>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return fail(CVCL_ERR_CUDA, "cuTensorMapEncodeTiled failed (CUresult %d) rows=%llu cols=%llu "
                                   "ld=%llu box=%ux%u", (int)r, (unsigned long long)rows,
                    (unsigned long long)cols, (unsigned long long)ld, box_rows, box_cols);
    return CVCL_OK;
}

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// number of kernels this library has launched (all threads); read through cvcl_launch_count()
inline unsigned long long& launch_counter() {
    static unsigned long long n = 0;
    return n;
}
inline void count_launch() { __atomic_fetch_add(&launch_counter(), 1ull, __ATOMIC_RELAXED); }

}  // namespace cvcl
