// Stand-alone self-test of the tcgen05 GEMM engine (development tool, built by `make selftest`).
// Prints max |err| of D = A * B^T against a host fp64 reference for several shapes.
#include "gemm_launch.cuh"
#include <vector>
#include <cmath>
#include <cstdlib>
using namespace cvcl;

static float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

template <int BN>
int run(int M, int N, int K, int ld_extra) {
    const int lda = K + ld_extra, ldb = K + ld_extra;
    std::vector<__nv_bfloat16> hA((size_t)M * lda), hB((size_t)N * ldb);
    std::vector<float> fA((size_t)M * K), fB((size_t)N * K);
    for (int i = 0; i < M; ++i) for (int k = 0; k < lda; ++k) {
        float v = bf16_round((rand() % 2001 - 1000) / 1000.f);
        hA[(size_t)i * lda + k] = __float2bfloat16_rn(k < K ? v : 77.f);
        if (k < K) fA[(size_t)i * K + k] = v;
    }
    for (int i = 0; i < N; ++i) for (int k = 0; k < ldb; ++k) {
        float v = bf16_round((rand() % 2001 - 1000) / 1000.f);
        hB[(size_t)i * ldb + k] = __float2bfloat16_rn(k < K ? v : 55.f);
        if (k < K) fB[(size_t)i * K + k] = v;
    }
    __nv_bfloat16 *dA, *dB; float* dC;
    const int ldc = (N + 3) / 4 * 4;
    cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dC, (size_t)M * ldc * 4);
    cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
    cudaMemset(dC, 0xff, (size_t)M * ldc * 4);
    GemmOperands op{}; op.A[0] = dA; op.ld_a[0] = lda; op.B[0] = dB; op.ld_b[0] = ldb; op.ndir = 1;
    GemmShape gs{}; gs.M[0] = M; gs.N[0] = N; gs.M[1] = M; gs.N[1] = N; gs.K = K; gs.m_stride = kBM; gs.n_stride = BN;
    EpiStoreF32::Params ep{}; ep.C[0] = dC; ep.ldc[0] = ldc; ep.C[1] = dC; ep.ldc[1] = ldc; ep.alpha = 1.f;
    int rc = launch_gemm<BN, 4, EpiStoreF32>(op, gs, ep, 1, 0);
    if (rc) { printf("launch rc=%d: %s\n", rc, last_error_buf()); return 1; }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("M=%d N=%d K=%d BN=%d: CUDA error %s\n", M, N, K, BN, cudaGetErrorString(e)); return 1; }
    std::vector<float> hC((size_t)M * ldc);
    cudaMemcpy(hC.data(), dC, hC.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0; int bad = 0;
    for (int i = 0; i < M; ++i) for (int j = 0; j < N; ++j) {
        double ref = 0; for (int k = 0; k < K; ++k) ref += (double)fA[(size_t)i * K + k] * fB[(size_t)j * K + k];
        double err = fabs(ref - hC[(size_t)i * ldc + j]);
        if (!(err <= 1e-3 * (1 + fabs(ref)))) { if (bad < 5) printf("  mismatch (%d,%d): got %f ref %f\n", i, j, hC[(size_t)i * ldc + j], ref); ++bad; }
        if (err > maxerr || err != err) maxerr = err;
    }
    printf("M=%d N=%d K=%d BN=%d ld+%d: max_err=%.3e bad=%d %s\n", M, N, K, BN, ld_extra, maxerr, bad, bad ? "FAIL" : "ok");
    cudaFree(dA); cudaFree(dB); cudaFree(dC);
    return bad != 0;
}

int main() {
    int fails = 0;
    fails += run<128>(128, 128, 64, 0);
    fails += run<128>(128, 128, 512, 0);
    fails += run<128>(512, 512, 2048, 0);
    fails += run<128>(200, 77, 520, 8);
    fails += run<128>(4, 1, 512, 0);
    fails += run<256>(256, 256, 512, 0);
    fails += run<256>(300, 500, 128, 0);
    fails += run<64>(130, 64, 192, 0);
    printf("selftest_gemm: %d failing shapes\n", fails);
    return fails != 0;
}
