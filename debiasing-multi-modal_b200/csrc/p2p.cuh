// One-shot all-reduce of the small per-step vectors over NVLink peer memory, fused into the kernels that produce and
// consume them (no collective launch, no kernel boundary):
//   producer kernel (k_reduce_stats: BatchNorm column sums; k_rows_train: dgamma / dbeta): the LAST CTA to finish (atomic
//     ticket) PUSHES the rank's fp64 partial vector into slot [channel][parity][rank] of EVERY rank's symmetric buffer with
//     plain peer stores, fences at system scope and raises flag [channel][rank] = instance + 1 on every rank;
//   consumer kernel (k_rows_train / k_wgrad_tc prologue): spins (bounded) on its LOCAL flags until all ranks have arrived,
//     then sums the `world` slots in rank order -- every rank adds the same numbers in the same order, so the replicas
//     stay bit-identical -- and CTA 0 writes the global vector back for the later kernels of the step.
// Buffers come from cudaMalloc + CUDA IPC (exchanged through the library's own NCCL communicator at dbmm_comm_init).
// Slots are double-buffered by instance parity: a rank can run at most one instance ahead of the slowest rank, because
// completing instance i needs every rank's flag i.  `base` (device counter) makes instance numbers unique across replays
// of the same epoch graph.
#pragma once
#include "common.cuh"

namespace dbmm {

constexpr int P2P_MAX_WORLD = 8, P2P_VEC = 512, P2P_CHANNELS = 2;
constexpr size_t P2P_SLOT_BYTES = sizeof(double) * P2P_CHANNELS * 2 * P2P_MAX_WORLD * P2P_VEC;       // 128 KB
constexpr size_t P2P_FLAG_STRIDE = 32;                                                                // uint32 per 128-byte line
constexpr size_t P2P_CTRL_BYTES = 4096;                                                               // flags | tickets | base
constexpr size_t P2P_BYTES = P2P_SLOT_BYTES + P2P_CTRL_BYTES;

struct P2pArgs {
    int world, rank;          // world == 0: disabled (single GPU, or NCCL all-reduce between the kernels)
    int step;                 // instance = *base + step
    char* peer[P2P_MAX_WORLD];
};

__device__ __forceinline__ double* p2p_slot(char* buf, int ch, int parity, int src) {
    return reinterpret_cast<double*>(buf) + (((size_t)ch * 2 + parity) * P2P_MAX_WORLD + src) * P2P_VEC;
}
__device__ __forceinline__ unsigned* p2p_flag(char* buf, int ch, int src) {
    return reinterpret_cast<unsigned*>(buf + P2P_SLOT_BYTES) + ((size_t)ch * P2P_MAX_WORLD + src) * P2P_FLAG_STRIDE;
}
__device__ __forceinline__ unsigned* p2p_ticket(char* buf, int ch) {
    return reinterpret_cast<unsigned*>(buf + P2P_SLOT_BYTES) + ((size_t)P2P_CHANNELS * P2P_MAX_WORLD + ch) * P2P_FLAG_STRIDE;
}
__device__ __forceinline__ unsigned* p2p_base(char* buf) {
    return reinterpret_cast<unsigned*>(buf + P2P_SLOT_BYTES) + ((size_t)P2P_CHANNELS * P2P_MAX_WORLD + P2P_CHANNELS) * P2P_FLAG_STRIDE;
}

// Called by EVERY thread of EVERY CTA at the end of the producer kernel (after its own accumulator atomics).
// `local` holds the rank's complete vector once all CTAs have passed; n <= P2P_VEC doubles.
__device__ __forceinline__ void p2p_push_when_last(const P2pArgs& p, int ch, const double* local, int n, unsigned total_ctas) {
    __shared__ unsigned s_last;
    __threadfence();                                   // this CTA's accumulator atomics are visible device-wide
    __syncthreads();
    char* me = p.peer[p.rank];
    if (threadIdx.x == 0) s_last = (atomicAdd(p2p_ticket(me, ch), 1u) == total_ctas - 1) ? 1u : 0u;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const unsigned inst = __ldcg(p2p_base(me)) + (unsigned)p.step;
    const int parity = inst & 1u;
    for (int e = threadIdx.x; e < n; e += blockDim.x) {
        const double v = __ldcg(local + e);
        for (int r = 0; r < p.world; ++r) p2p_slot(p.peer[r], ch, parity, p.rank)[e] = v;
    }
    __threadfence_system();
    __syncthreads();
    if ((int)threadIdx.x < p.world) {
        unsigned* f = p2p_flag(p.peer[threadIdx.x], ch, p.rank);
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f), "r"(inst + 1u) : "memory");
    }
    if (threadIdx.x == 0) *p2p_ticket(me, ch) = 0u;    // ready for the next instance (ordered by the kernel boundary)
}

// Called by every thread of a consumer CTA before it needs the global vector.  Returns the slot parity to read.
__device__ __forceinline__ int p2p_wait(const P2pArgs& p, int ch) {
    char* me = p.peer[p.rank];
    const unsigned inst = __ldcg(p2p_base(me)) + (unsigned)p.step;
    if ((int)threadIdx.x < p.world) {
        const unsigned* f = p2p_flag(me, ch, threadIdx.x);
        unsigned v = 0;
        for (unsigned spin = 0; spin < (1u << 28); ++spin) {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
            if ((int)(v - (inst + 1u)) >= 0) break;
        }
        if ((int)(v - (inst + 1u)) < 0) __trap();      // a rank never arrived: fail loudly instead of hanging the GPU
    }
    __syncthreads();
    return inst & 1u;
}
__device__ __forceinline__ double p2p_sum(const P2pArgs& p, int ch, int parity, int e) {
    char* me = p.peer[p.rank];
    double s = 0.0;
    for (int r = 0; r < p.world; ++r) s += __ldcg(p2p_slot(me, ch, parity, r) + e);
    return s;
}

__global__ void k_p2p_bump(char* buf, unsigned steps) { *p2p_base(buf) += steps; }

}  // namespace dbmm
