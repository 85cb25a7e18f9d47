// Collectives of the sharded contrastive step over NVLink PEER memory: the exchange
// steps (feature all-gather, LSE all-gather) and the gradient sum of the replicated head / embedding
// parameters run as single kernels that carry their own cross-rank barrier, instead of NCCL calls
// (measured at 2 GPUs: NCCL all-gather 17 us, all-reduce of the 9.6 MB gradient buffer 45 us).
//
// Every rank maps every other rank's block (torch symmetric memory, NVSwitch: full bandwidth to any
// peer).  Synchronisation is per block index: block b of rank r signals block b of every peer by
// storing the launch's epoch into the peer's flag word [phase][b][r] (st.release.sys over NVLink) and
// spins (ld.acquire.sys) on its own words [phase][b][peer].  Epochs only grow (one per launch, kept in
// a per-rank device array so CUDA-graph replays advance them), so flags are never reset.  All ranks
// must issue the same sequence of launches per channel (the grid may differ from launch to launch: flag
// words are indexed by a fixed per-phase stride and epochs are kept per block index); the grid is at
// most kPeerMaxBlocks CTAs of 512 threads, always co-resident on 148 SMs, so the spin cannot starve a
// block it waits for.
//
// Two families, same results (the PUSH kernels further down are the default of sharding.PeerExchange:
// all NVLink traffic is posted stores; measured 151.9 us vs 158.5 us (pull) vs 165.0 us (NCCL) per step at
// 2 GPUs and 193 us vs 256 us (NCCL) at 8 GPUs, profiles/r01_sharded_phases_*.txt):
//
//   all-gather : pull: start barrier (the peers' blocks are complete: a kernel only starts after all
//                earlier work of its stream) -> load every block with 16-byte loads, several in flight.
//                push: store the own block into every rank's gathered buffer -> barrier.
//   all-reduce : pull: start barrier -> rank r sums slice r of every rank's buffer in rank order (so the
//                result does not depend on who reduces it) and stores the sum into slice r of EVERY
//                rank's buffer -> end barrier.  push: scatter slice p into rank p's scratch -> barrier ->
//                local sum in rank order, store to every rank -> barrier.  In place, deterministic.
//
// A barrier that does not complete within timeout_ms (default 10 minutes, the order of NCCL's own
// watchdog: a rank may legitimately stall between steps -- data-loader start-up, a checkpoint) sets
// *status and traps (the step fails loudly instead of hanging the device); with CVCL_PEER_NO_TRAP or'ed into timeout_ms it only sets *status and
// goes on (the start-up self-test of sharding.PeerExchange uses this to decide for or against the path).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "ptx.cuh"

namespace cvcl {

constexpr int kPeerMaxWorld = 8;
constexpr int kPeerMaxBlocks = 64;
constexpr int kPeerThreads = 512;
// flag words one channel needs in every rank's flag area: [phase 2][block][source rank]
constexpr int kPeerFlagWords = 2 * kPeerMaxBlocks * kPeerMaxWorld;

struct PeerTable {
    void* data[kPeerMaxWorld];          // rank r's block (peer-mapped)
    void* aux[kPeerMaxWorld];           // rank r's scratch block (push all-reduce), peer-mapped
    uint32_t* flags[kPeerMaxWorld];     // rank r's flag area for this channel (peer-mapped)
    uint32_t* epoch;                    // local [kPeerMaxBlocks]
    int* status;                        // local, nullable
    unsigned int timeout_ms;
};

namespace peer {

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint4 ld_sys_v4(const uint4* p) {
    uint4 v;
    asm volatile("ld.global.relaxed.sys.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// All threads of the block call it.  Orders every earlier write of the block (to any rank) before
// the signal, and every later read after the peers' signals.
__device__ __forceinline__ void barrier(const PeerTable& t, int world, int rank, uint32_t e, int phase) {
    __threadfence_system();
    __syncthreads();
    const int peer = threadIdx.x;
    if (peer < world && peer != rank) {
        // fixed stride per phase (NOT gridDim.x): launches of one channel with different grids (a training
        // step, then a forward-only step) must never alias flag words that carry different epochs
        const int slot = (phase * kPeerMaxBlocks + blockIdx.x) * kPeerMaxWorld;
        st_release_sys(t.flags[peer] + slot + rank, e);
        const uint32_t* mine = t.flags[rank] + slot + peer;
        unsigned long long t0 = 0;
        unsigned int it = 0;
        while (static_cast<int32_t>(ld_acquire_sys(mine) - e) < 0) {
            if ((++it & 255u) == 0) {
                const unsigned long long now = globaltimer_ns();
                if (t0 == 0) t0 = now;
                else if (now - t0 > static_cast<unsigned long long>(t.timeout_ms & 0x7fffffffu) * 1000000ull) {
                    if (t.status) atomicExch(t.status, 1 + phase);
                    __threadfence_system();
                    if (t.timeout_ms & 0x80000000u) break;      // probe mode: report through *status, do not trap
                    __trap();
                }
            }
        }
    }
    __syncthreads();
}

__device__ __forceinline__ uint32_t epoch_begin(const PeerTable& t) { return t.epoch[blockIdx.x] + 1u; }
__device__ __forceinline__ void epoch_end(const PeerTable& t, uint32_t e) {
    if (threadIdx.x == 0) t.epoch[blockIdx.x] = e;
}

}  // namespace peer

// --------------------------------------------------------------------------------------
// all-gather: every rank's block holds `nseg` segments of seg16 16-byte words, segment s at word
// offset s * src_seg_stride16; dst gets segment s of rank r at s * dst_seg_stride16 + r * seg16
// (features: 1 segment; the two LSE vectors: 2 segments -> dst [2][world * b]).
// --------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kPeerThreads) peer_allgather_kernel(const PeerTable t, int world, int rank,
                                                                      long long seg16, int nseg,
                                                                      long long src_seg_stride16, uint4* dst,
                                                                      long long dst_seg_stride16) {
    const uint32_t e = peer::epoch_begin(t);
    peer::barrier(t, world, rank, e, 0);
    const long long per_rank = seg16 * nseg, total = per_rank * world;
    const long long stride = static_cast<long long>(gridDim.x) * kPeerThreads;
    constexpr int U = 4;
    for (long long base = static_cast<long long>(blockIdx.x) * kPeerThreads + threadIdx.x; base < total;
         base += U * stride) {
        uint4 v[U];
        long long out[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long idx = base + u * stride;
            out[u] = -1;
            if (idx < total) {
                const long long r = idx / per_rank, rem = idx - r * per_rank;
                const long long s = rem / seg16, i = rem - s * seg16;
                v[u] = peer::ld_sys_v4(static_cast<const uint4*>(t.data[r]) + s * src_seg_stride16 + i);
                out[u] = s * dst_seg_stride16 + r * seg16 + i;
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (out[u] >= 0) dst[out[u]] = v[u];
    }
    peer::epoch_end(t, e);
}

// --------------------------------------------------------------------------------------
// in-place all-reduce (sum) of n4 float4 words, two-shot over peer memory.
// --------------------------------------------------------------------------------------
template <int W>
__global__ void __launch_bounds__(kPeerThreads) peer_allreduce_f32_kernel(const PeerTable t, int rank, long long n4) {
    const uint32_t e = peer::epoch_begin(t);
    peer::barrier(t, W, rank, e, 0);
    const long long per = (n4 + W - 1) / W;
    const long long lo = rank * per;
    const long long hi = (lo + per < n4) ? lo + per : n4;
    const long long stride = static_cast<long long>(gridDim.x) * kPeerThreads;
    constexpr int U = (W >= 8) ? 1 : (W >= 4 ? 2 : 4);          // W * U loads in flight per thread
    for (long long base = lo + static_cast<long long>(blockIdx.x) * kPeerThreads + threadIdx.x; base < hi;
         base += U * stride) {
        uint4 v[U][W];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long i = base + u * stride;
            if (i < hi) {
#pragma unroll
                for (int p = 0; p < W; ++p) v[u][p] = peer::ld_sys_v4(static_cast<const uint4*>(t.data[p]) + i);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long i = base + u * stride;
            if (i < hi) {
                float4 acc = make_float4(__uint_as_float(v[u][0].x), __uint_as_float(v[u][0].y),
                                         __uint_as_float(v[u][0].z), __uint_as_float(v[u][0].w));
#pragma unroll
                for (int p = 1; p < W; ++p) {           // rank order: the sum does not depend on the reducer
                    acc.x += __uint_as_float(v[u][p].x); acc.y += __uint_as_float(v[u][p].y);
                    acc.z += __uint_as_float(v[u][p].z); acc.w += __uint_as_float(v[u][p].w);
                }
#pragma unroll
                for (int p = 0; p < W; ++p) static_cast<float4*>(t.data[p])[i] = acc;
            }
        }
    }
    peer::barrier(t, W, rank, e, 1);
    peer::epoch_end(t, e);
}

// barrier only (fences the reuse of the exchange blocks when no gradient all-reduce follows)
__global__ void __launch_bounds__(32) peer_barrier_kernel(const PeerTable t, int world, int rank) {
    const uint32_t e = peer::epoch_begin(t);
    peer::barrier(t, world, rank, e, 0);
    peer::epoch_end(t, e);
}

// --------------------------------------------------------------------------------------
// PUSH variants: all NVLink traffic is posted stores (no load round trips over the switch).
//
// all-gather: every rank stores its block into slot `rank` of EVERY rank's gathered buffer (t.data[p] =
// rank p's gathered buffer), then one barrier: when a peer's signal arrives its stores have landed.
// The gathered buffers may be overwritten because the previous step's all-reduce (or closing barrier)
// ended after every rank had finished reading them.
// --------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kPeerThreads) peer_allgather_push_kernel(const PeerTable t, int world, int rank,
                                                                           const uint4* src, long long seg16,
                                                                           int nseg, long long src_seg_stride16,
                                                                           long long dst_seg_stride16) {
    ptx::pdl_launch_dependents(); ptx::pdl_wait();
    const uint32_t e = peer::epoch_begin(t);
    const long long per_rank = seg16 * nseg;
    const long long stride = static_cast<long long>(gridDim.x) * kPeerThreads;
    constexpr int U = 4;
    for (long long base = static_cast<long long>(blockIdx.x) * kPeerThreads + threadIdx.x; base < per_rank;
         base += U * stride) {
        uint4 v[U];
        long long out[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long idx = base + u * stride;
            out[u] = -1;
            if (idx < per_rank) {
                const long long s = idx / seg16, i = idx - s * seg16;
                v[u] = src[s * src_seg_stride16 + i];
                out[u] = s * dst_seg_stride16 + rank * seg16 + i;
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (out[u] >= 0) {
                for (int p = 0; p < world; ++p) static_cast<uint4*>(t.data[p])[out[u]] = v[u];
            }
        }
    }
    peer::barrier(t, world, rank, e, 0);
    peer::epoch_end(t, e);
}

// --------------------------------------------------------------------------------------
// all-reduce, two-shot, push: (A) rank r stores slice p of its buffer into rank p's scratch block
// [src rank r][.] for every p != r; barrier; (B) rank r sums its slice over {own buffer, scratch[src]} in
// rank order and stores the sum into slice r of every rank's buffer; barrier.  The scratch block of a
// rank is rewritten by launch e+1 only after the end barrier of launch e, i.e. after it was consumed.
// --------------------------------------------------------------------------------------
template <int W>
__global__ void __launch_bounds__(kPeerThreads) peer_allreduce_push_f32_kernel(const PeerTable t, int rank, long long n4) {
    ptx::pdl_launch_dependents(); ptx::pdl_wait();
    const uint32_t e = peer::epoch_begin(t);
    const long long per = (n4 + W - 1) / W;
    const long long stride = static_cast<long long>(gridDim.x) * kPeerThreads;
    const long long tid = static_cast<long long>(blockIdx.x) * kPeerThreads + threadIdx.x;
    const uint4* mine = static_cast<const uint4*>(t.data[rank]);
    constexpr int U = (W >= 8) ? 1 : (W >= 4 ? 2 : 4);
    // (A) scatter my contributions
#pragma unroll U
    for (long long j = tid; j < per; j += stride) {
        uint4 v[W];
#pragma unroll
        for (int p = 0; p < W; ++p)
            if (p != rank && p * per + j < n4) v[p] = mine[p * per + j];
#pragma unroll
        for (int p = 0; p < W; ++p)
            if (p != rank && p * per + j < n4) static_cast<uint4*>(t.aux[p])[rank * per + j] = v[p];
    }
    peer::barrier(t, W, rank, e, 0);
    // (B) reduce my slice, broadcast
    const long long lo = rank * per;
    const long long cnt = (lo + per <= n4) ? per : (n4 > lo ? n4 - lo : 0);
    const uint4* scr = static_cast<const uint4*>(t.aux[rank]);
#pragma unroll U
    for (long long j = tid; j < cnt; j += stride) {
        uint4 v[W];
#pragma unroll
        for (int p = 0; p < W; ++p) v[p] = (p == rank) ? mine[lo + j] : peer::ld_sys_v4(scr + p * per + j);
        float4 acc = make_float4(__uint_as_float(v[0].x), __uint_as_float(v[0].y), __uint_as_float(v[0].z),
                                 __uint_as_float(v[0].w));
#pragma unroll
        for (int p = 1; p < W; ++p) {               // rank order: the sum does not depend on the reducer
            acc.x += __uint_as_float(v[p].x); acc.y += __uint_as_float(v[p].y);
            acc.z += __uint_as_float(v[p].z); acc.w += __uint_as_float(v[p].w);
        }
#pragma unroll
        for (int p = 0; p < W; ++p) static_cast<float4*>(t.data[p])[lo + j] = acc;
    }
    peer::barrier(t, W, rank, e, 1);
    peer::epoch_end(t, e);
}

}  // namespace cvcl
