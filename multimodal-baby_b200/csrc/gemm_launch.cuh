// Host launcher for gemm_tn_bf16_kernel: tensor maps, cluster attribute, dynamic smem opt-in.
#pragma once
#include "gemm_sm100.cuh"
#include "host_util.h"

namespace cvcl {

struct GemmOperands {           // per direction: A [M,K] ld_a, B [N,K] ld_b (bf16, row-major)
    const void* A[2]; int ld_a[2];
    const void* B[2]; int ld_b[2];
    int ndir;                   // 1 or 2
};

template <int BN, int STAGES, class Epi>
int launch_gemm(const GemmOperands& op, const GemmShape& gs, const typename Epi::Params& ep,
                int cluster_n, cudaStream_t stream) {
    using L = GemmSmem<BN, STAGES>;
    constexpr int smem_bytes = L::template total<Epi>();
    static_assert(smem_bytes <= 227 * 1024, "shared memory budget");
    auto kern = gemm_tn_bf16_kernel<BN, STAGES, Epi>;
    static thread_local bool attr_done = false;
    if (!attr_done) {
        CVCL_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
        attr_done = true;
    }
    CUtensorMap maps[4];
    for (int z = 0; z < 2; ++z) {
        const int zz = z < op.ndir ? z : 0;
        int rc = make_tmap_bf16(&maps[2 * z], op.A[zz], gs.M[zz], gs.K, op.ld_a[zz], kBM);
        if (rc) return rc;
        rc = make_tmap_bf16(&maps[2 * z + 1], op.B[zz], gs.N[zz], gs.K, op.ld_b[zz], BN);
        if (rc) return rc;
    }
    int max_m = gs.M[0], max_n = gs.N[0];
    if (op.ndir == 2) { max_m = max_m > gs.M[1] ? max_m : gs.M[1]; max_n = max_n > gs.N[1] ? max_n : gs.N[1]; }
    dim3 grid(ceil_div(max_m, gs.m_stride), ceil_div(max_n, gs.n_stride), op.ndir);
    if (cluster_n > 1) grid.y = ceil_div((int)grid.y, cluster_n) * cluster_n;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(kGemmThreads);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 1;
    at[0].val.clusterDim.y = cluster_n > 1 ? cluster_n : 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    CVCL_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, maps[0], maps[1], maps[2], maps[3], gs, ep));
    count_launch();
    return CVCL_OK;
}

}  // namespace cvcl
