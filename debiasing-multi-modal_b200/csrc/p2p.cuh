// Data-parallel exchanges of the training step over NVLink peer memory, fused into the kernels that produce and consume
// the data: no collective launch, no kernel boundary, no NCCL inside the step.
//
//   channel 0  BatchNorm column sums [nad][2][H] fx64   k_reduce_stats: the LAST CTA to finish (atomic ticket) pushes the
//              rank's vector to the other ranks, polls theirs, adds in rank order, writes the global sums back in place
//   channel 1  (dgamma, dbeta) [2][H] fx64              k_wgrad_tc: CTA (0, 0) pushes in its prologue, every CTA polls
//   dW1        [H][D] fp32                              k_tail_w1: reduce-scatter + all-gather by OWNER: rank o owns a contiguous
//              1 / world of the quads; every thread pushes its chunk-summed quad to the owner only, the owner's thread sums the
//              ranks in rank order, does the SGD step and pushes the UPDATED weights to every rank (2 x (N-1)/N x |W1| per rank
//              instead of (N-1) x |W1|; momentum lives on the owner).  No cross-CTA synchronisation at all
//   S^T        [H+1][144] fp32                          k_sum_spart_g: every element travels as ONE LL word to every rank, summed in rank order
//   S          [H+1+C][H+1] fp32                        k_tail_w2 / k_p2p_sum_st (TN-GEMM path): CTA c stores slice c to every rank and raises flag
//              [c][rank] (release, system scope); the consumer waits for the slices it needs (off the critical path)
//
// Channels 0 / 1 and dW1 travel as "LL" words: every 32 data bits ride in an 8-byte store together with the instance tag;
// 8-byte stores are atomic over NVLink, so the consumer polls the data words themselves -- no system fence, no flag store,
// one NVLink write latency per exchange.  Everything is summed in rank order: every rank adds the same numbers in the same
// order, so the replicas stay bit-identical.  Slots are double-buffered by instance parity: a rank can run at most one
// instance ahead of the slowest rank, because completing instance i needs every rank's data of instance i.  `base`
// (device counter, bumped by k_p2p_bump after every epoch) makes instance numbers unique across replays of an epoch graph.
// Waits are bounded by WALL CLOCK (%globaltimer; P2pArgs::timeout_ns, default 300 s, DBMM_P2P_TIMEOUT_S overrides): ordinary
// rank skew (one rank writing a checkpoint, a first-time graph capture, a longer validation) just waits; a rank that never
// arrives makes the waiter raise the error word of its own buffer and carry on, and the host reports it from
// dbmm_comm_check() (the CUDA context stays usable).  Ranks must otherwise run in lock step: one epoch call per rank.
// Buffers come from cudaMalloc + CUDA IPC (handles exchanged through the library's own NCCL communicator, dbmm_comm_init).
#pragma once
#include "common.cuh"

namespace dbmm {

constexpr int P2P_MAX_WORLD = 8, P2P_VEC = 512, P2P_CHANNELS = 2;
constexpr size_t P2P_CTRL_BYTES = 4096;                    // tickets [channel] | base, one 128-byte line each
constexpr size_t P2P_S_FLOATS = 20480, P2P_G_FLOATS = 128 * 2048;      // capacity of one S slot / one dW1 slot (H * D <= 256 K)
constexpr int P2P_S_CTAS = 64;                             // S flags [P2P_S_CTAS][32 words]: one line per k_tail_w2 CTA, word r = rank r
// layout of the symmetric buffer
constexpr size_t P2P_S_OFF = P2P_CTRL_BYTES;                                                       // S slots [parity][rank][P2P_S_FLOATS] fp32
constexpr size_t P2P_G_OFF = P2P_S_OFF + sizeof(float) * 2 * P2P_MAX_WORLD * P2P_S_FLOATS;          // dW1 LL slots [parity][rank][P2P_G_FLOATS] x 8 B
constexpr size_t P2P_SF_OFF = P2P_G_OFF + 8 * (size_t)2 * P2P_MAX_WORLD * P2P_G_FLOATS;             // S flags
constexpr size_t P2P_LL_OFF = P2P_SF_OFF + (size_t)P2P_S_CTAS * 128;                               // fp64 LL slots [channel][parity][rank][P2P_VEC] x 16 B
constexpr size_t P2P_W_OFF = P2P_LL_OFF + (size_t)P2P_CHANNELS * 2 * P2P_MAX_WORLD * P2P_VEC * 16;   // updated-W1 LL slots [parity][P2P_G_FLOATS] x 8 B
constexpr size_t P2P_ST_FLOATS = 20480;                                                            // >= (H + 1) * 144 elements of S^T
constexpr size_t P2P_ST_OFF = P2P_W_OFF + 8 * (size_t)2 * P2P_G_FLOATS;                            // S^T LL slots [parity][rank][P2P_ST_FLOATS] x 8 B
constexpr size_t P2P_BYTES = P2P_ST_OFF + 8 * (size_t)2 * P2P_MAX_WORLD * P2P_ST_FLOATS;

struct P2pArgs {
    unsigned long long timeout_ns;   // wall-clock bound of every wait
    int world, rank;          // world == 0: disabled (single GPU, or NCCL all-reduces between the kernels)
    int step;                 // instance = *base + step
    int skip;                 // timing experiments only (DBMM_DP_SKIP): bit 0/1 sum only the own value of channel 0/1, bit 2 no
                              // wait for the dW1 quads, bit 3 no wait for S
    char* peer[P2P_MAX_WORLD];
};

__device__ __forceinline__ unsigned* p2p_ticket(char* buf, int ch) { return reinterpret_cast<unsigned*>(buf) + (size_t)ch * 32; }
__device__ __forceinline__ unsigned* p2p_base(char* buf) { return reinterpret_cast<unsigned*>(buf) + (size_t)P2P_CHANNELS * 32; }
__device__ __forceinline__ unsigned* p2p_error(char* buf) { return reinterpret_cast<unsigned*>(buf) + (size_t)(P2P_CHANNELS + 1) * 32; }
__device__ __forceinline__ unsigned long long p2p_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// true when the wait that started at t0 has run out of time (raises the error word; checked every 1024 polls)
__device__ __forceinline__ bool p2p_expired(const P2pArgs& p, unsigned long long& t0, unsigned spin) {
    if ((spin & 1023u) != 1023u) return false;
    const unsigned long long now = p2p_now();
    if (t0 == 0) { t0 = now; return false; }
    if (now - t0 < p.timeout_ns) return false;
    atomicExch(p2p_error(p.peer[p.rank]), 1u);
    return true;
}
__device__ __forceinline__ unsigned p2p_instance(const P2pArgs& p) { return __ldcg(p2p_base(p.peer[p.rank])) + (unsigned)p.step; }
__global__ void k_p2p_bump(char* buf, unsigned steps) { *p2p_base(buf) += steps; }

// ---- fp64 vectors as LL words
__device__ __forceinline__ unsigned long long* p2p_ll_slot(char* buf, int ch, int parity, int src) {
    return reinterpret_cast<unsigned long long*>(buf + P2P_LL_OFF) + ((((size_t)ch * 2 + parity) * P2P_MAX_WORLD + src) * P2P_VEC) * 2;
}
__device__ __forceinline__ void p2p_ll_store(unsigned long long* slot, int e, long long v, unsigned tag) {
    const unsigned long long bits = (unsigned long long)v;
    const unsigned long long w0 = (bits & 0xffffffffull) | ((unsigned long long)tag << 32);
    const unsigned long long w1 = (bits >> 32) | ((unsigned long long)tag << 32);
    asm volatile("st.relaxed.sys.global.v2.u64 [%0], {%1, %2};" ::"l"(slot + 2 * (size_t)e), "l"(w0), "l"(w1) : "memory");
}
__device__ __forceinline__ long long p2p_ll_load(const P2pArgs& p, const unsigned long long* slot, int e, unsigned tag) {
    unsigned long long w0 = 0, w1 = 0, t0 = 0;
    for (unsigned spin = 0;; ++spin) {
        asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(slot + 2 * (size_t)e) : "memory");
        if ((unsigned)(w0 >> 32) == tag && (unsigned)(w1 >> 32) == tag) break;
        if (p2p_expired(p, t0, spin)) break;           // a rank never arrived: error word raised, the host reports it
    }
    return (long long)((w0 & 0xffffffffull) | (w1 << 32));
}

// Consumer side: the token (instance number) to pass to p2p_sum; the waiting happens per element there.
__device__ __forceinline__ int p2p_wait(const P2pArgs& p, int ch) {
    (void)ch;
    return (int)p2p_instance(p);
}
// Element e of the global vector: polls every rank's LL word pair until it carries this instance and adds the fixed-point
// values (integer sums: every rank gets the same bits whatever the order; the replicas stay bit-identical).
__device__ __forceinline__ long long p2p_sum(const P2pArgs& p, int ch, int token, int e) {
    char* me = p.peer[p.rank];
    const unsigned inst = (unsigned)token;
    long long s = 0;
    if (p.skip & (1 << ch)) return p2p_ll_load(p, p2p_ll_slot(me, ch, inst & 1u, p.rank), e, inst + 1u);
    for (int r = 0; r < p.world; ++r) s += p2p_ll_load(p, p2p_ll_slot(me, ch, inst & 1u, r), e, inst + 1u);
    return s;
}

// Whole all-reduce of a small fp64 vector inside the producer kernel (k_reduce_stats: BatchNorm column sums): the LAST
// CTA to finish (atomic ticket) pushes the rank's vector to every OTHER rank as LL words, polls theirs, adds everything
// in rank order (own values from registers) and writes the global vector back over `local` -- the consumer kernels
// read plain memory and need no peer-memory code at all.
__device__ __forceinline__ void p2p_allreduce_when_last(const P2pArgs& p, int ch, fx64* local, int n, unsigned total_ctas) {
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
        const long long v = __ldcg(&local[e].v);
        for (int r = 0; r < p.world; ++r)
            if (r != p.rank) p2p_ll_store(p2p_ll_slot(p.peer[r], ch, parity, p.rank), e, v, inst + 1u);
        long long s = 0;
        for (int r = 0; r < p.world; ++r)
            s += (r == p.rank || (p.skip & (1 << ch))) ? (r == p.rank ? v : 0ll) : p2p_ll_load(p, p2p_ll_slot(me, ch, parity, r), e, inst + 1u);
        local[e].v = s;
    }
    if (threadIdx.x == 0) *p2p_ticket(me, ch) = 0u;
}

// k_wgrad_tc prologue: CTA (0, 0) pushes the rank's (dgamma, dbeta) to every rank INCLUDING itself; every CTA then sums
// all ranks' LL slots with p2p_sum (the row kernel stays free of peer-memory code).
__device__ __forceinline__ void p2p_push_now(const P2pArgs& p, int ch, const fx64* local, int n) {
    const unsigned inst = p2p_instance(p);
    for (int e = threadIdx.x; e < n; e += blockDim.x) {
        const long long v = __ldcg(&local[e].v);
        for (int r = 0; r < p.world; ++r) p2p_ll_store(p2p_ll_slot(p.peer[r], ch, inst & 1u, p.rank), e, v, inst + 1u);
    }
}

// ---- dW1 quads as LL words, S slices with flags
__device__ __forceinline__ unsigned long long* p2p_g_ll(char* buf, int parity, int src) {        // 4 words per float4
    return reinterpret_cast<unsigned long long*>(buf + P2P_G_OFF) + ((size_t)parity * P2P_MAX_WORLD + src) * P2P_G_FLOATS;
}
__device__ __forceinline__ void p2p_g_store(unsigned long long* slot, int64_t i, float4 v, unsigned tag) {
    const unsigned long long t = (unsigned long long)tag << 32;
    const unsigned long long w0 = __float_as_uint(v.x) | t, w1 = __float_as_uint(v.y) | t, w2 = __float_as_uint(v.z) | t, w3 = __float_as_uint(v.w) | t;
    asm volatile("st.relaxed.sys.global.v2.u64 [%0], {%1, %2};" ::"l"(slot + 4 * i), "l"(w0), "l"(w1) : "memory");
    asm volatile("st.relaxed.sys.global.v2.u64 [%0], {%1, %2};" ::"l"(slot + 4 * i + 2), "l"(w2), "l"(w3) : "memory");
}
__device__ __forceinline__ float4 p2p_g_load(const P2pArgs& p, const unsigned long long* slot, int64_t i, unsigned tag, bool no_wait) {
    unsigned long long w0 = 0, w1 = 0, w2 = 0, w3 = 0, t0 = 0;
    bool ok = false;
    for (unsigned spin = 0;; ++spin) {
        asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(slot + 4 * i) : "memory");
        asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(w2), "=l"(w3) : "l"(slot + 4 * i + 2) : "memory");
        ok = (unsigned)(w0 >> 32) == tag && (unsigned)(w1 >> 32) == tag && (unsigned)(w2 >> 32) == tag && (unsigned)(w3 >> 32) == tag;
        if (ok || no_wait || p2p_expired(p, t0, spin)) break;
    }
    return make_float4(__uint_as_float((unsigned)w0), __uint_as_float((unsigned)w1), __uint_as_float((unsigned)w2), __uint_as_float((unsigned)w3));
}
// the owner's updated W1 quads (all-gather half of the dW1 exchange): one slot per parity, written by the owner of each quad
__device__ __forceinline__ unsigned long long* p2p_w_ll(char* buf, int parity) {
    return reinterpret_cast<unsigned long long*>(buf + P2P_W_OFF) + (size_t)parity * P2P_G_FLOATS;
}
// ---- single floats as LL words (S^T elements, exchanged inside k_sum_spart_g: one NVLink write latency, no fence, no flag kernel)
__device__ __forceinline__ unsigned long long* p2p_st_ll(char* buf, int parity, int src) {
    return reinterpret_cast<unsigned long long*>(buf + P2P_ST_OFF) + ((size_t)parity * P2P_MAX_WORLD + src) * P2P_ST_FLOATS;
}
__device__ __forceinline__ void p2p_f_store(unsigned long long* slot, int e, float v, unsigned tag) {
    const unsigned long long w = (unsigned long long)__float_as_uint(v) | ((unsigned long long)tag << 32);
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(slot + e), "l"(w) : "memory");
}
__device__ __forceinline__ float p2p_f_load(const P2pArgs& p, const unsigned long long* slot, int e, unsigned tag) {
    unsigned long long w = 0, t0 = 0;
    for (unsigned spin = 0;; ++spin) {
        asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(w) : "l"(slot + e) : "memory");
        if ((unsigned)(w >> 32) == tag || (p.skip & 8) || p2p_expired(p, t0, spin)) break;
    }
    return __uint_as_float((unsigned)w);
}
__device__ __forceinline__ float* p2p_s_slot(char* buf, int parity, int src) {
    return reinterpret_cast<float*>(buf + P2P_S_OFF) + ((size_t)parity * P2P_MAX_WORLD + src) * P2P_S_FLOATS;
}
__device__ __forceinline__ unsigned* p2p_s_flag(char* buf, int cta, int src) {
    return reinterpret_cast<unsigned*>(buf + P2P_SF_OFF) + (size_t)cta * 32 + src;
}

}  // namespace dbmm
