// Micro-benchmark of the tcgen05 engine (development tool): time vs K to separate the fixed
// per-kernel cost from the per-k-chunk cost.  usage: bench_gemm
#include "gemm_launch.cuh"
#include <vector>
using namespace cvcl;

template <int BN, int STAGES>
float run(int M, int N, int K, int reps, bool flush) {
    __nv_bfloat16 *dA, *dB; float* dC; char* fl;
    cudaMalloc(&dA, (size_t)M * K * 2); cudaMalloc(&dB, (size_t)N * K * 2); cudaMalloc(&dC, (size_t)M * N * 4);
    cudaMalloc(&fl, 256 << 20);
    cudaMemset(dA, 0, (size_t)M * K * 2); cudaMemset(dB, 0, (size_t)N * K * 2);
    GemmOperands op{}; op.ndir = 1;
    op.A[0] = mat(dA, M, K, K); op.B[0] = mat(dB, N, K, K); op.out[0] = mat(dC, M, N, N);
    GemmShape gs{}; gs.M[0] = gs.M[1] = M; gs.N[0] = gs.N[1] = N; gs.K = K; gs.m_stride = kBM; gs.n_stride = BN;
    EpiStoreF32::Params ep{}; ep.alpha = 1.f;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float tot = 0;
    for (int i = 0; i < reps + 3; ++i) {
        if (flush) cudaMemsetAsync(fl, i, 256 << 20, 0);
        cudaEventRecord(e0, 0);
        launch_gemm<BN, STAGES, EpiStoreF32, false, false>(op, gs, ep, 1, 0);
        cudaEventRecord(e1, 0);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (i >= 3) tot += ms;
    }
    cudaFree(dA); cudaFree(dB); cudaFree(dC); cudaFree(fl);
    return tot / reps * 1e3f;
}

int main() {
    printf("== stage / tile sweep (warm L2), time in us (TF/s)\n");
    const int shapes[3][3] = {{4096, 4096, 2048}, {32768, 512, 32768}, {16384, 16384, 512}};
    for (auto& sh : shapes) {
        const int M = sh[0], N = sh[1], K = sh[2];
        const double fl = 2.0 * M * N * K * 1e-6;
        float t;
        printf("M=%d N=%d K=%d:", M, N, K);
        t = run<128, 4>(M, N, K, 5, false); printf("  BN128/4st %8.1f (%4.0f)", t, fl / t);
        t = run<128, 3>(M, N, K, 5, false); printf("  BN128/3st %8.1f (%4.0f)", t, fl / t);
        t = run<128, 2>(M, N, K, 5, false); printf("  BN128/2st %8.1f (%4.0f)", t, fl / t);
        t = run<256, 3>(M, N, K, 5, false); printf("  BN256/3st %8.1f (%4.0f)", t, fl / t);
        printf("\n");
    }
    return 0;
}
