// Host launcher for gemm_bf16_kernel: tensor maps, cluster attribute, dynamic smem opt-in.
#pragma once
#include "gemm_sm100.cuh"
#include "host_util.h"

namespace cvcl {

struct Mat {                    // a row-major matrix as it sits in memory
    const void* ptr; int rows, cols, ld;
};
inline Mat mat(const void* p, int rows, int cols, int ld) { Mat m; m.ptr = p; m.rows = rows; m.cols = cols; m.ld = ld; return m; }

struct GemmOperands {
    Mat A[2], B[2];             // per direction; K-major: [M or N, K]; MN-major: [K, M or N]
    Mat aux[2];                 // epilogue input tiles, bf16 [M, N]   (Epi::kNumAux of them)
    Mat out[2];                 // epilogue output per direction, [M, N] bf16 / fp32 (Epi::kOutElemBytes)
    int ndir;                   // 1 or 2
};

template <int BN, int STAGES, class Epi, bool A_MN, bool B_MN>
int launch_gemm(const GemmOperands& op, const GemmShape& gs, const typename Epi::Params& ep,
                int cluster_n, cudaStream_t stream) {
    using L = GemmSmem<BN, STAGES, Epi::kNumAux>;
    constexpr int smem_bytes = L::template total<Epi>();
    static_assert(smem_bytes <= 227 * 1024, "shared memory budget");
    auto kern = gemm_bf16_kernel<BN, STAGES, Epi, A_MN, B_MN>;
    static thread_local unsigned attr_mask = 0;     // per device: function attributes live in the device's context
    int attr_dev = 0;
    if (cudaGetDevice(&attr_dev) != cudaSuccess || attr_dev < 0 || attr_dev > 31) attr_dev = 0;
    if (!((attr_mask >> attr_dev) & 1u)) {
        CVCL_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
        attr_mask |= 1u << attr_dev;
    }
    GemmMaps maps;
    int rc;
    for (int z = 0; z < 2; ++z) {
        const int zz = z < op.ndir ? z : 0;
        const Mat& a = op.A[zz]; const Mat& b = op.B[zz];
        if (A_MN) rc = make_tmap(&maps.a[z], a.ptr, 2, a.rows, a.cols, a.ld, 64, kBK);
        else      rc = make_tmap(&maps.a[z], a.ptr, 2, a.rows, a.cols, a.ld, 64, kBM);
        if (rc) return rc;
        if (B_MN) rc = make_tmap(&maps.b[z], b.ptr, 2, b.rows, b.cols, b.ld, 64, kBK);
        else      rc = make_tmap(&maps.b[z], b.ptr, 2, b.rows, b.cols, b.ld, 64, BN);
        if (rc) return rc;
        if (Epi::kOutElemBytes) {
            const Mat& o = op.out[zz];
            rc = make_tmap(&maps.out[z], o.ptr, Epi::kOutElemBytes, o.rows, o.cols, o.ld,
                           128 / Epi::kOutElemBytes, kBM);
            if (rc) return rc;
        } else {
            maps.out[z] = maps.a[z];
        }
    }
    for (int i = 0; i < 2; ++i) {
        if (i < Epi::kNumAux) {
            const Mat& x = op.aux[i];
            rc = make_tmap(&maps.aux[i], x.ptr, 2, x.rows, x.cols, x.ld, 64, kBM);
            if (rc) return rc;
        } else {
            maps.aux[i] = maps.a[0];
        }
    }
    int max_m = gs.M[0], max_n = gs.N[0];
    if (op.ndir == 2) { max_m = max_m > gs.M[1] ? max_m : gs.M[1]; max_n = max_n > gs.N[1] ? max_n : gs.N[1]; }
    dim3 grid(ceil_div(max_m, gs.m_stride), ceil_div(max_n, gs.n_stride), gs.k_splits > 1 ? gs.k_splits : op.ndir);
    if (cluster_n > 1) grid.y = ceil_div((int)grid.y, cluster_n) * cluster_n;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(kGemmThreads);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = stream;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 1;
    at[0].val.clusterDim.y = cluster_n > 1 ? cluster_n : 1;
    at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // PDL: see ptx::pdl_wait()
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = 2;
    CVCL_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, maps, gs, ep));
    count_launch();
    return CVCL_OK;
}

inline int sm_count() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    }
    return n;
}

// Persistent launch (K-major operands, streaming epilogue): one CTA per SM over all tiles.
template <int BN, int STAGES, class Epi>
int launch_gemm_persistent(const GemmOperands& op, const GemmShape& gs, const typename Epi::Params& ep,
                           cudaStream_t stream) {
    using L = PersistSmem<BN, STAGES, Epi::kOutElemBytes, Epi::kScratchBytes>;
    constexpr int smem_bytes = L::total();
    static_assert(smem_bytes <= 227 * 1024, "shared memory budget");
    auto kern = gemm_bf16_persistent_kernel<BN, STAGES, Epi>;
    static thread_local unsigned attr_mask = 0;     // per device: function attributes live in the device's context
    int attr_dev = 0;
    if (cudaGetDevice(&attr_dev) != cudaSuccess || attr_dev < 0 || attr_dev > 31) attr_dev = 0;
    if (!((attr_mask >> attr_dev) & 1u)) {
        CVCL_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
        attr_mask |= 1u << attr_dev;
    }
    GemmMaps maps;
    TileGrid tg{};
    int rc;
    for (int z = 0; z < 2; ++z) {
        const int zz = z < op.ndir ? z : 0;
        if ((rc = make_tmap(&maps.a[z], op.A[zz].ptr, 2, op.A[zz].rows, op.A[zz].cols, op.A[zz].ld, 64, kBM))) return rc;
        if ((rc = make_tmap(&maps.b[z], op.B[zz].ptr, 2, op.B[zz].rows, op.B[zz].cols, op.B[zz].ld, 64, BN))) return rc;
        if (Epi::kOutElemBytes) {
            const Mat& o = op.out[zz];
            if ((rc = make_tmap(&maps.out[z], o.ptr, Epi::kOutElemBytes, o.rows, o.cols, o.ld, 128 / Epi::kOutElemBytes, kBM))) return rc;
        } else {
            maps.out[z] = maps.a[z];
        }
        tg.tiles_m[z] = z < op.ndir ? ceil_div(gs.M[z], gs.m_stride) : 0;
        tg.tiles_n[z] = z < op.ndir ? ceil_div(gs.N[z], gs.n_stride) : 0;
    }
    maps.aux[0] = maps.aux[1] = maps.a[0];
    tg.total = tg.tiles_m[0] * tg.tiles_n[0] + tg.tiles_m[1] * tg.tiles_n[1];
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(tg.total < sm_count() ? tg.total : sm_count());
    cfg.blockDim = dim3(64 + 128 * (BN / 128));
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    CVCL_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, maps, gs, tg, ep));
    count_launch();
    return CVCL_OK;
}

}  // namespace cvcl
