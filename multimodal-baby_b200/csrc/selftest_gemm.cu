// Stand-alone self-test of the tcgen05 GEMM engine (development tool, `build.py::build_selftest`).
// D = A * B^T against a host fp64 reference for several shapes and all four operand-major
// combinations (K-major / MN-major A and B), output through the smem-staged TMA store.
#include "gemm_launch.cuh"
#include <vector>
#include <cmath>
#include <cstdlib>
using namespace cvcl;

static float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

// logical A [M,K], B [N,K]; stored K-major ([rows,K]) or MN-major ([K,rows])
template <int BN, bool A_MN, bool B_MN>
int run(int M, int N, int K) {
    auto pad8 = [](int x) { return (x + 7) / 8 * 8; };
    std::vector<float> fA((size_t)M * K), fB((size_t)N * K);
    for (auto& v : fA) v = bf16_round((rand() % 2001 - 1000) / 1000.f);
    for (auto& v : fB) v = bf16_round((rand() % 2001 - 1000) / 1000.f);
    const int lda = A_MN ? pad8(M) : pad8(K), ldb = B_MN ? pad8(N) : pad8(K);
    const int ra = A_MN ? K : M, rb = B_MN ? K : N;
    std::vector<__nv_bfloat16> hA((size_t)ra * lda, __float2bfloat16_rn(99.f)), hB((size_t)rb * ldb, __float2bfloat16_rn(77.f));
    for (int i = 0; i < M; ++i) for (int k = 0; k < K; ++k)
        hA[A_MN ? (size_t)k * lda + i : (size_t)i * lda + k] = __float2bfloat16_rn(fA[(size_t)i * K + k]);
    for (int i = 0; i < N; ++i) for (int k = 0; k < K; ++k)
        hB[B_MN ? (size_t)k * ldb + i : (size_t)i * ldb + k] = __float2bfloat16_rn(fB[(size_t)i * K + k]);
    __nv_bfloat16 *dA, *dB; float* dC;
    const int ldc = (N + 3) / 4 * 4;
    cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dC, (size_t)M * ldc * 4);
    cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
    cudaMemset(dC, 0xff, (size_t)M * ldc * 4);
    GemmOperands op{}; op.ndir = 1;
    op.A[0] = A_MN ? mat(dA, K, M, lda) : mat(dA, M, K, lda);
    op.B[0] = B_MN ? mat(dB, K, N, ldb) : mat(dB, N, K, ldb);
    op.out[0] = mat(dC, M, N, ldc);
    GemmShape gs{}; gs.M[0] = gs.M[1] = M; gs.N[0] = gs.N[1] = N; gs.K = K; gs.m_stride = kBM; gs.n_stride = BN;
    EpiStoreF32::Params ep{}; ep.alpha = 1.f;
    int rc = launch_gemm<BN, 4, EpiStoreF32, A_MN, B_MN>(op, gs, ep, 1, 0);
    if (rc) { printf("launch rc=%d: %s\n", rc, last_error_buf()); return 1; }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("M=%d N=%d K=%d BN=%d: CUDA error %s\n", M, N, K, BN, cudaGetErrorString(e)); return 1; }
    std::vector<float> hC((size_t)M * ldc);
    cudaMemcpy(hC.data(), dC, hC.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0; int bad = 0;
    for (int i = 0; i < M; ++i) for (int j = 0; j < N; ++j) {
        double ref = 0; for (int k = 0; k < K; ++k) ref += (double)fA[(size_t)i * K + k] * fB[(size_t)j * K + k];
        double err = fabs(ref - hC[(size_t)i * ldc + j]);
        if (!(err <= 1e-3 * (1 + fabs(ref)))) { if (bad < 3) printf("  mismatch (%d,%d): got %f ref %f\n", i, j, hC[(size_t)i * ldc + j], ref); ++bad; }
        if (err > maxerr || err != err) maxerr = err;
    }
    printf("A_%s B_%s M=%d N=%d K=%d BN=%d: max_err=%.3e bad=%d %s\n", A_MN ? "MN" : "K ", B_MN ? "MN" : "K ", M, N, K, BN,
           maxerr, bad, bad ? "FAIL" : "ok");
    cudaFree(dA); cudaFree(dB); cudaFree(dC);
    return bad != 0;
}

template <bool A_MN, bool B_MN>
int suite() {
    int f = 0;
    f += run<128, A_MN, B_MN>(128, 128, 64);
    f += run<128, A_MN, B_MN>(512, 512, 512);
    f += run<128, A_MN, B_MN>(200, 77, 520);
    f += run<128, A_MN, B_MN>(4, 1, 512);
    f += run<256, A_MN, B_MN>(300, 500, 136);
    return f;
}

int main() {
    int fails = 0;
    fails += suite<false, false>();
    fails += suite<false, true>();
    fails += suite<true, false>();
    fails += suite<true, true>();
    fails += run<128, true, true>(512, 2048, 512);     // dW shape: E x K over B
    printf("selftest_gemm: %d failing shapes\n", fails);
    return fails != 0;
}
