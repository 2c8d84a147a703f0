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

constexpr int P2P_MAX_WORLD = 8, P2P_VEC = 512, P2P_CHANNELS = 3;      // channel 2: flags / ticket of the fp32 S matrix (slots below)
constexpr size_t P2P_SLOT_BYTES = sizeof(double) * 2 * 2 * P2P_MAX_WORLD * P2P_VEC;                   // fp64 slots of channels 0, 1: 128 KB
constexpr size_t P2P_FLAG_STRIDE = 32;                                                                // uint32 per 128-byte line
constexpr size_t P2P_CTRL_BYTES = 4096;                                                               // flags | tickets | base
// Gradient exchange of the fused data-parallel step tail (no NCCL inside the step):
//   S slots  [parity][rank][P2P_S_FLOATS]   k_tail_w2 CTA c pushes slice c of the rank's S = [c*h | c | ds]^T [h | 1]
//   G slots  [parity][rank][P2P_G_FLOATS] x 2   k_tail_w1 pushes the rank's chunk-summed dW1 as LL words (32 data bits +
//                                               instance tag per 8-byte store): the consumer polls the data itself
constexpr size_t P2P_S_FLOATS = 20480, P2P_G_FLOATS = 128 * 2048;
constexpr int P2P_G_CTAS = 128;                                     // (size of the former flag region, kept as padding)
constexpr size_t P2P_S_OFF = P2P_SLOT_BYTES + P2P_CTRL_BYTES;
constexpr size_t P2P_G_OFF = P2P_S_OFF + sizeof(float) * 2 * P2P_MAX_WORLD * P2P_S_FLOATS;
constexpr size_t P2P_GF_OFF = P2P_G_OFF + 2 * sizeof(float) * 2 * P2P_MAX_WORLD * P2P_G_FLOATS;
// Small vectors (channels 0, 1) travel as "LL" words: every fp64 value is two 8-byte stores {32 data bits, instance + 1},
// each atomic over NVLink, so the consumer polls the data words themselves -- no system fence, no separate flag store,
// one NVLink write latency per exchange.      LL slots [channel][parity][rank][P2P_VEC][2] x 8 bytes
constexpr int P2P_S_CTAS = 64;                                     // S flags [P2P_S_CTAS][32 words]: k_tail_w2 CTA c pushes slice c of S
constexpr size_t P2P_SF_OFF = P2P_GF_OFF + (size_t)P2P_G_CTAS * 128;
constexpr size_t P2P_LL_OFF = P2P_SF_OFF + (size_t)P2P_S_CTAS * 128;
constexpr size_t P2P_BYTES = P2P_LL_OFF + (size_t)2 * 2 * P2P_MAX_WORLD * P2P_VEC * 16;

struct P2pArgs {
    int world, rank;          // world == 0: disabled (single GPU, or NCCL all-reduce between the kernels)
    int step;                 // instance = *base + step
    int skip;                 // timing experiments only (DBMM_DP_SKIP): bit 0/1 sum only the own LL slot of channel 0/1, bit 2 no
                              // wait for the dW1 slices, bit 3 no wait for S, bit 4 no system fence before the dW1 flags
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

__device__ __forceinline__ unsigned long long* p2p_ll_slot(char* buf, int ch, int parity, int src) {
    return reinterpret_cast<unsigned long long*>(buf + P2P_LL_OFF) + ((((size_t)ch * 2 + parity) * P2P_MAX_WORLD + src) * P2P_VEC) * 2;
}
__device__ __forceinline__ void p2p_ll_store(unsigned long long* slot, int e, double v, unsigned tag) {
    const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
    const unsigned long long w0 = (bits & 0xffffffffull) | ((unsigned long long)tag << 32);
    const unsigned long long w1 = (bits >> 32) | ((unsigned long long)tag << 32);
    asm volatile("st.relaxed.sys.global.v2.u64 [%0], {%1, %2};" ::"l"(slot + 2 * (size_t)e), "l"(w0), "l"(w1) : "memory");
}
__device__ __forceinline__ double p2p_ll_load(const unsigned long long* slot, int e, unsigned tag) {
    unsigned long long w0 = 0, w1 = 0;
    for (unsigned spin = 0; spin < (1u << 26); ++spin) {
        asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(slot + 2 * (size_t)e) : "memory");
        if ((unsigned)(w0 >> 32) == tag && (unsigned)(w1 >> 32) == tag) break;
    }
    if ((unsigned)(w0 >> 32) != tag || (unsigned)(w1 >> 32) != tag) __trap();      // a rank never arrived: fail loudly, do not hang
    return __longlong_as_double((long long)((w0 & 0xffffffffull) | (w1 << 32)));
}
__device__ __forceinline__ unsigned p2p_instance(const P2pArgs& p) { return __ldcg(p2p_base(p.peer[p.rank])) + (unsigned)p.step; }

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
    const unsigned inst = p2p_instance(p);
    const int parity = inst & 1u;
    for (int e = threadIdx.x; e < n; e += blockDim.x) {
        const double v = __ldcg(local + e);
        for (int r = 0; r < p.world; ++r) p2p_ll_store(p2p_ll_slot(p.peer[r], ch, parity, p.rank), e, v, inst + 1u);
    }
    if (threadIdx.x == 0) *p2p_ticket(me, ch) = 0u;    // ready for the next instance (ordered by the kernel boundary)
}

// Consumer side: the token (instance number) to pass to p2p_sum; the waiting happens per element there.
__device__ __forceinline__ int p2p_wait(const P2pArgs& p, int ch) {
    (void)ch;
    return (int)p2p_instance(p);
}
// Element e of the global vector: polls every rank's LL word pair until it carries this instance, sums in rank order
// (every rank adds the same numbers in the same order: the replicas stay bit-identical).
__device__ __forceinline__ double p2p_sum(const P2pArgs& p, int ch, int token, int e) {
    char* me = p.peer[p.rank];
    const unsigned inst = (unsigned)token;
    double s = 0.0;
    if (p.skip & (1 << ch)) return p2p_ll_load(p2p_ll_slot(me, ch, inst & 1u, p.rank), e, inst + 1u);
    for (int r = 0; r < p.world; ++r) s += p2p_ll_load(p2p_ll_slot(me, ch, inst & 1u, r), e, inst + 1u);
    return s;
}

__global__ void k_p2p_bump(char* buf, unsigned steps) { *p2p_base(buf) += steps; }

__device__ __forceinline__ float* p2p_s_slot(char* buf, int parity, int src) {
    return reinterpret_cast<float*>(buf + P2P_S_OFF) + ((size_t)parity * P2P_MAX_WORLD + src) * P2P_S_FLOATS;
}
__device__ __forceinline__ unsigned long long* p2p_g_ll(char* buf, int parity, int src) {        // 4 words per float4
    return reinterpret_cast<unsigned long long*>(buf + P2P_G_OFF) + ((size_t)parity * P2P_MAX_WORLD + src) * P2P_G_FLOATS;
}
__device__ __forceinline__ void p2p_g_store(unsigned long long* slot, int64_t i, float4 v, unsigned tag) {
    const unsigned long long t = (unsigned long long)tag << 32;
    const unsigned long long w0 = __float_as_uint(v.x) | t, w1 = __float_as_uint(v.y) | t, w2 = __float_as_uint(v.z) | t, w3 = __float_as_uint(v.w) | t;
    asm volatile("st.relaxed.sys.global.v2.u64 [%0], {%1, %2};" ::"l"(slot + 4 * i), "l"(w0), "l"(w1) : "memory");
    asm volatile("st.relaxed.sys.global.v2.u64 [%0], {%1, %2};" ::"l"(slot + 4 * i + 2), "l"(w2), "l"(w3) : "memory");
}
__device__ __forceinline__ float4 p2p_g_load(const unsigned long long* slot, int64_t i, unsigned tag, bool no_wait) {
    unsigned long long w0 = 0, w1 = 0, w2 = 0, w3 = 0;
    bool ok = false;
    for (unsigned spin = 0; spin < (no_wait ? 1u : (1u << 26)); ++spin) {
        asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(slot + 4 * i) : "memory");
        asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(w2), "=l"(w3) : "l"(slot + 4 * i + 2) : "memory");
        ok = (unsigned)(w0 >> 32) == tag && (unsigned)(w1 >> 32) == tag && (unsigned)(w2 >> 32) == tag && (unsigned)(w3 >> 32) == tag;
        if (ok) break;
    }
    if (!ok && !no_wait) __trap();                    // a rank never arrived: fail loudly, do not hang
    return make_float4(__uint_as_float((unsigned)w0), __uint_as_float((unsigned)w1), __uint_as_float((unsigned)w2), __uint_as_float((unsigned)w3));
}
__device__ __forceinline__ unsigned* p2p_s_flag(char* buf, int cta, int src) {
    return reinterpret_cast<unsigned*>(buf + P2P_SF_OFF) + (size_t)cta * 32 + src;
}
// Whole all-reduce of a small fp64 vector inside the producer kernel (k_reduce_stats: BatchNorm column sums): the LAST
// CTA to finish (atomic ticket) pushes the rank's vector to every OTHER rank as LL words, polls theirs, adds everything
// in rank order (own values from registers) and writes the global vector back over `local` -- the consumer kernels
// read plain memory and need no peer-memory code at all.
__device__ __forceinline__ void p2p_allreduce_when_last(const P2pArgs& p, int ch, double* local, int n, unsigned total_ctas) {
    __shared__ unsigned s_last3;
    __threadfence();
    __syncthreads();
    char* me = p.peer[p.rank];
    if (threadIdx.x == 0) s_last3 = (atomicAdd(p2p_ticket(me, ch), 1u) == total_ctas - 1) ? 1u : 0u;
    __syncthreads();
    if (!s_last3) return;
    __threadfence();
    const unsigned inst = p2p_instance(p);
    const int parity = inst & 1u;
    for (int e = threadIdx.x; e < n; e += blockDim.x) {
        const double v = __ldcg(local + e);
        for (int r = 0; r < p.world; ++r)
            if (r != p.rank) p2p_ll_store(p2p_ll_slot(p.peer[r], ch, parity, p.rank), e, v, inst + 1u);
        double s = 0.0;
        for (int r = 0; r < p.world; ++r)
            s += (r == p.rank || (p.skip & (1 << ch))) ? (r == p.rank ? v : 0.0) : p2p_ll_load(p2p_ll_slot(me, ch, parity, r), e, inst + 1u);
        local[e] = s;
    }
    if (threadIdx.x == 0) *p2p_ticket(me, ch) = 0u;
}

// k_wgrad_tc prologue: CTA (0, 0) pushes the rank's (dgamma, dbeta) to every rank INCLUDING itself; every CTA then sums
// all ranks' LL slots with p2p_sum (the row kernel stays free of peer-memory code).
__device__ __forceinline__ void p2p_push_now(const P2pArgs& p, int ch, const double* local, int n) {
    const unsigned inst = p2p_instance(p);
    for (int e = threadIdx.x; e < n; e += blockDim.x) {
        const double v = __ldcg(local + e);
        for (int r = 0; r < p.world; ++r) p2p_ll_store(p2p_ll_slot(p.peer[r], ch, inst & 1u, p.rank), e, v, inst + 1u);
    }
}

}  // namespace dbmm
