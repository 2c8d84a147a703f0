// GEMM-1 on the 5th-generation tensor cores:   A[b][j] = sum_k X[row(b)][k] * W1[j][k] + b1[j]
//
//   * tcgen05.mma kind::tf32, M = 128 rows per CTA, N = BN hidden units, K = 8 per instruction; fp32 accumulator in TMEM.
//   * X is used as stored (fp32 containers; CLIP embeddings are fp16-valued, hence exact in tf32).  W1 is split
//     W1 = hi + lo with hi = W1 & 0xffffe000 (tf32-exact) and lo = W1 - hi, and two MMAs accumulate x*hi + x*lo,
//     which restores fp32-level accuracy (the tensor core truncates lo to its top 19 bits: ~2^-22 relative to W1).
//   * operands are staged in shared memory as SWIZZLE_128B K-major tiles (128-byte rows = 32 floats of K) by four
//     producer warps with 16-byte cp.async (rows are gathered through the batch's index list), multi-stage ring
//     with full/empty mbarriers; one thread of a fifth warp issues the MMAs; the four producer warps then drain
//     TMEM (tcgen05.ld 32x32b: one accumulator row per thread), add the bias, write A and reduce the BatchNorm
//     column sums (fp64 across CTAs).
#pragma once
#include "common.cuh"
#include "ptx_sm100.cuh"
#include "p2p.cuh"

namespace dbmm {

constexpr int G1_BM = 128, G1_BK = 32, G1_PRODUCERS = 128, G1_THREADS = 160, G1_LAG = 2;

struct Gemm1TcArgs {
    const float* X; int64_t ldx; const int32_t* idx; int64_t pos0;
    int B, D, H, nad;
    const float* Whi[2]; const float* Wlo[2]; const float* b1[2];   // per adapter, [H][D] / [H]
    float* A;          // [nad][B][H]
    fx64* colsum;      // [nad][2][H] or nullptr (fixed point, FX_COLSUM)
    // eval epilogue (hhi != nullptr): h = relu(BN_running(a)) split into tf32 hi + lo, [nad][B][H] each, instead of A
    float* hhi; float* hlo;
    const float* bn_mean[2]; const float* bn_var[2]; const float* bn_gamma[2]; const float* bn_beta[2];
    int ksplit;        // > 1: blockIdx.z owns a slice of D and stores its raw partial tile (no bias, no sums) to
    float* part;       //      part[kpart][nad][B][H]; k_reduce_stats finishes the job
    int pack;          // data parallel: request only the used part of the stage ring (two CTAs may share an SM)
    int stages;        // set by the launcher: min(ring depth of the configuration, k-blocks per CTA)
    fx64* zero_colsum; int zero_colsum_n;       // ksplit > 1 only: CTA 0 resets the column sums k_reduce_stats will accumulate
};

__global__ void __launch_bounds__(256) k_split_tf32(const float* __restrict__ w, float* __restrict__ hi,
                                                    float* __restrict__ lo, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float v = w[i];
        const float h = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
        hi[i] = h;
        lo[i] = v - h;
    }
}

template <int BN, int TERMS>
struct G1Cfg {
    static constexpr int A_BYTES = G1_BM * 128;
    static constexpr int B_BYTES = BN * 128;
    static constexpr int STAGE_BYTES = A_BYTES + TERMS * B_BYTES;
    static constexpr int STAGES = (STAGE_BYTES * 6 <= 196608) ? 6 : (196608 / STAGE_BYTES);
    static constexpr int TMEM_COLS = BN < 32 ? 32 : BN;
    static constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

// kz: D-slice of this CTA (blockIdx.z of the single-run launch; the batched launch keeps the member index there)
template <int BN, int TERMS>
__device__ __forceinline__ void gemm1_tc_body(const Gemm1TcArgs& a, const int kz) {
    using Cfg = G1Cfg<BN, TERMS>;
    const int S = a.stages;                             // stage ring depth: the launch sizes the shared memory for it
    extern __shared__ uint8_t g1_smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)g1_smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* full = (uint64_t*)(smem + (size_t)S * Cfg::STAGE_BYTES);
    uint64_t* empty = full + S;
    uint64_t* tmem_full = empty + S;
    uint32_t* tmem_ptr = (uint32_t*)(tmem_full + 1);
    __shared__ int64_t sRowOff[G1_BM];
    __shared__ unsigned long long sCol[2][BN];          // fixed-point column sums of this CTA (integer adds commute)

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = blockIdx.x * G1_BM;
    const int slices = a.H / BN;                        // hidden slices per adapter
    const int ad = blockIdx.y / slices;
    const int n0 = (blockIdx.y - ad * slices) * BN;     // first hidden unit of this CTA
    const int KB_all = a.D / G1_BK;
    const int kb_per = (KB_all + a.ksplit - 1) / a.ksplit;
    const int kb_lo = kz * kb_per;
    const int KB = max(0, min(KB_all, kb_lo + kb_per) - kb_lo);      // k-blocks of this CTA (host guarantees >= 1)

    if (tid < G1_BM) {
        int m = m0 + tid;
        if (m >= a.B) m = a.B - 1;                      // clamp: tail rows read a valid row, masked in the epilogue
        const int64_t r = a.idx ? (int64_t)a.idx[a.pos0 + m] : (a.pos0 + m);
        sRowOff[tid] = r * a.ldx;
        // pull this CTA's share of the row into L2 with ONE sequential request; the k-blocked 128-byte copies below then
        // hit L2 instead of opening a DRAM page per slice (measured: 1.4 -> see profiles/ TB/s on the eval stream)
        if (blockIdx.y == 0 && m0 + tid < a.B) ptx::prefetch_l2_bulk(a.X + r * a.ldx + (size_t)kb_lo * G1_BK, (uint32_t)KB * G1_BK * 4);
    }
    if (tid < BN) { sCol[0][tid] = 0ull; sCol[1][tid] = 0ull; }
    if (tid == 0) {
        for (int s = 0; s < S; ++s) { ptx::mbar_init(&full[s], G1_PRODUCERS); ptx::mbar_init(&empty[s], 1); }
        ptx::mbar_init(tmem_full, 1);
        ptx::fence_mbar_init();
    }
    if (warp == 4) ptx::tmem_alloc<Cfg::TMEM_COLS>(tmem_ptr);
    ptx::tc_fence_before_sync();
    __syncthreads();
    ptx::tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_ptr;
    // The embedding rows do not depend on the previous kernel: when the CTA's whole operand set fits the stage ring
    // (training: 2 k-blocks), the producers put every X tile in flight BEFORE the dependency wait, so the HBM fetch
    // overlaps the tail of the previous step (programmatic dependent launch); the weight slices follow after the wait.
    const bool x_early = KB <= S;
    if (warp < 4 && x_early) {
        for (int kb = 0; kb < KB; ++kb) {
            const uint32_t sA = ptx::smem_u32(smem + (size_t)kb * Cfg::STAGE_BYTES);
            const int k0 = (kb_lo + kb) * G1_BK;
#pragma unroll
            for (int id = tid; id < G1_BM * 8; id += G1_PRODUCERS) {
                const int row = id >> 3, c = id & 7;
                ptx::cp_async16(sA + ptx::sw128_offset(row, c), a.X + sRowOff[row] + k0 + c * 4);
            }
        }
    }
    ptx::pdl_wait();                // W1 hi / lo come from the previous step's update kernel
    if (a.ksplit > 1) DBMM_TL_WAIT(TL_GEMM1);
    if (a.zero_colsum && blockIdx.x == 0 && blockIdx.y == 0 && kz == 0)       // (its chores CTA read the column sums)
        for (int e = tid; e < a.zero_colsum_n; e += G1_THREADS) a.zero_colsum[e].v = 0;
    ptx::pdl_launch();

    if (warp < 4) {
        // ===================== producers: gather X rows + W1 hi/lo slices into swizzled K-major tiles =====================
        const float* whi = a.Whi[ad] + (size_t)n0 * a.D;
        const float* wlo = a.Wlo[ad] + (size_t)n0 * a.D;
        auto signal = [&](int kb) {
            ptx::fence_proxy_async_smem();
            ptx::mbar_arrive(&full[kb % S]);
        };
        for (int kb = 0; kb < KB; ++kb) {
            const int s = kb % S;
            ptx::mbar_wait(&empty[s], ((kb / S) & 1) ^ 1);
            uint8_t* stage = smem + (size_t)s * Cfg::STAGE_BYTES;
            const uint32_t sA = ptx::smem_u32(stage);
            const int k0 = (kb_lo + kb) * G1_BK;
            if (!x_early) {
#pragma unroll
                for (int id = tid; id < G1_BM * 8; id += G1_PRODUCERS) {
                    const int row = id >> 3, c = id & 7;
                    ptx::cp_async16(sA + ptx::sw128_offset(row, c), a.X + sRowOff[row] + k0 + c * 4);
                }
            }
#pragma unroll
            for (int t = 0; t < TERMS; ++t) {
                const uint32_t sB = sA + Cfg::A_BYTES + t * Cfg::B_BYTES;
                const float* w = t == 0 ? whi : wlo;
#pragma unroll
                for (int id = tid; id < BN * 8; id += G1_PRODUCERS) {
                    const int row = id >> 3, c = id & 7;
                    ptx::cp_async16(sB + ptx::sw128_offset(row, c), w + (size_t)row * a.D + k0 + c * 4);
                }
            }
            ptx::cp_async_commit();
            if (kb >= G1_LAG) { ptx::cp_async_wait<G1_LAG>(); signal(kb - G1_LAG); }
        }
        // drain the last LAG groups
        if (KB >= 2) { ptx::cp_async_wait<1>(); signal(KB - 2); }
        ptx::cp_async_wait<0>();
        signal(KB - 1);

        // ===================== epilogue: TMEM -> registers -> A (+ bias) and BatchNorm column sums =====================
        ptx::mbar_wait(tmem_full, 0);
        ptx::tc_fence_after_sync();
        float* scr = (float*)smem + warp * (32 * 33);            // stage memory is free once every MMA has completed
        const int m = m0 + warp * 32 + lane;
        const bool row_ok = m < a.B;
        const int rows_here = min(32, max(0, a.B - (m0 + warp * 32)));
        const float* bias = a.b1[ad] + n0;
        float* arow = a.A + ((size_t)ad * a.B + (row_ok ? m : 0)) * a.H + n0;
#pragma unroll 1
        for (int ch = 0; ch < BN / 32; ++ch) {
            uint32_t r[32];
            ptx::tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(warp * 32) << 16) + ch * 32, r);
            ptx::tmem_ld_wait();
            float v[32];
            if (a.ksplit > 1) {
                if (row_ok) {
                    float* prow = a.part + (((size_t)kz * a.nad + ad) * a.B + m) * a.H + n0 + ch * 32;
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        *reinterpret_cast<float4*>(prow + j) = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                                                          __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
                }
                continue;
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) + __ldg(bias + ch * 32 + j);
            if (a.hhi) {
                if (row_ok) {
                    const size_t off = ((size_t)ad * a.B + m) * a.H + n0 + ch * 32;
                    float hi[32], lo[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int c = n0 + ch * 32 + j;
                        const float ah = (v[j] - __ldg(a.bn_mean[ad] + c)) * (1.0f / sqrtf(__ldg(a.bn_var[ad] + c) + DBMM_BN_EPS));
                        const float h = fmaxf(fmaf(ah, __ldg(a.bn_gamma[ad] + c), __ldg(a.bn_beta[ad] + c)), 0.f);
                        hi[j] = __uint_as_float(__float_as_uint(h) & 0xffffe000u);
                        lo[j] = h - hi[j];
                    }
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        *reinterpret_cast<float4*>(a.hhi + off + j) = make_float4(hi[j], hi[j + 1], hi[j + 2], hi[j + 3]);
                        *reinterpret_cast<float4*>(a.hlo + off + j) = make_float4(lo[j], lo[j + 1], lo[j + 2], lo[j + 3]);
                    }
                }
                continue;
            }
            if (row_ok) {
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    *reinterpret_cast<float4*>(arow + ch * 32 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            }
            if (a.colsum) {
#pragma unroll
                for (int j = 0; j < 32; ++j) scr[lane * 33 + j] = v[j];
                __syncwarp();
                float s1 = 0.f, s2 = 0.f;
                for (int rr = 0; rr < rows_here; ++rr) { const float x = scr[rr * 33 + lane]; s1 += x; s2 = fmaf(x, x, s2); }
                __syncwarp();
                atomicAdd(&sCol[0][ch * 32 + lane], fx_bits<FX_COLSUM>((double)s1));
                atomicAdd(&sCol[1][ch * 32 + lane], fx_bits<FX_COLSUM>((double)s2));
            }
        }
        ptx::tc_fence_before_sync();
        asm volatile("bar.sync 1, 128;" ::: "memory");          // the four epilogue warps only
        if (a.ksplit == 1 && a.colsum && tid < BN) {
            atomicAdd(reinterpret_cast<unsigned long long*>(&a.colsum[((size_t)ad * 2 + 0) * a.H + n0 + tid].v), sCol[0][tid]);
            atomicAdd(reinterpret_cast<unsigned long long*>(&a.colsum[((size_t)ad * 2 + 1) * a.H + n0 + tid].v), sCol[1][tid]);
        }
    } else {
        // ===================== MMA issuer: one thread =====================
        if (lane == 0) {
            constexpr uint32_t idesc = ptx::umma_idesc(/*tf32*/ 2, G1_BM, BN, 0, 0);
            for (int kb = 0; kb < KB; ++kb) {
                const int s = kb % S;
                ptx::mbar_wait(&full[s], (kb / S) & 1);
                ptx::tc_fence_after_sync();
                const uint32_t sA = ptx::smem_u32(smem + (size_t)s * Cfg::STAGE_BYTES);
#pragma unroll
                for (int kk = 0; kk < G1_BK / 8; ++kk) {
                    const uint64_t adesc = ptx::umma_smem_desc(sA + kk * 32, 0, 1024);
#pragma unroll
                    for (int t = 0; t < TERMS; ++t) {
                        const uint64_t bdesc = ptx::umma_smem_desc(sA + Cfg::A_BYTES + t * Cfg::B_BYTES + kk * 32, 0, 1024);
                        ptx::mma_tf32_ss(tmem_base, adesc, bdesc, idesc, (kb | kk | t) != 0 ? 1u : 0u);
                    }
                }
                ptx::mma_commit(&empty[s]);          // smem slot reusable once these MMAs have read it
            }
            ptx::mma_commit(tmem_full);              // accumulator complete
        }
        __syncwarp();
    }
    __syncthreads();
    if (warp == 4) ptx::tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
}

template <int BN, int TERMS>
__global__ void __launch_bounds__(G1_THREADS, 1) k_gemm1_tc(Gemm1TcArgs a) { DBMM_TL_SCOPE(TL_GEMM1); gemm1_tc_body<BN, TERMS>(a, (int)blockIdx.z); }

template <int BN, int TERMS>
static int launch_gemm1_tc_impl(const Gemm1TcArgs& a, cudaStream_t st) {
    using Cfg = G1Cfg<BN, TERMS>;
    auto kern = k_gemm1_tc<BN, TERMS>;
    DBMM_CUDA(set_smem(kern, Cfg::SMEM));
    dim3 grid(ceil_div(a.B, G1_BM), a.nad * (a.H / BN), a.ksplit);
    // a D-sliced training CTA owns 2 k-blocks of the 4-deep ring; see train_smem_bytes for what is requested
    Gemm1TcArgs b = a;
    const int kb_per = (a.D / G1_BK + a.ksplit - 1) / a.ksplit;
    b.stages = kb_per < Cfg::STAGES ? kb_per : Cfg::STAGES;
    const size_t smem = train_smem_bytes((size_t)b.stages * Cfg::STAGE_BYTES + 1024 + 256, Cfg::SMEM, a.pack);
    g_plain_next_launch = pdl_off_for("gemm1");
    DBMM_CUDA(launch_pdl(kern, grid, dim3(G1_THREADS), smem, st, b));
    return DBMM_OK;
}

// BN: hidden units per CTA.  Small batches use narrow slices so that more SMs pull operands concurrently.
static int launch_gemm1_tc(const Gemm1TcArgs& a, int bn, cudaStream_t st) {
    DBMM_CHECK_SHAPE(a.D % G1_BK == 0, "tensor-core GEMM-1 needs D %% 32 == 0 (D=%d)", a.D);
    DBMM_CHECK_ARG(a.ksplit >= 1 && (a.ksplit - 1) * ((a.D / G1_BK + a.ksplit - 1) / a.ksplit) < a.D / G1_BK,
                   "GEMM-1 split of D into %d parts leaves an empty part", a.ksplit);
    DBMM_CHECK_SHAPE(a.H % bn == 0, "H=%d not divisible by the hidden slice %d", a.H, bn);
    switch (bn) {
        case 128: return launch_gemm1_tc_impl<128, 2>(a, st);
        case 64: return launch_gemm1_tc_impl<64, 2>(a, st);
        case 32: return launch_gemm1_tc_impl<32, 2>(a, st);
        default: set_error("unsupported hidden slice %d", bn); return DBMM_ERR_UNSUPPORTED_SHAPE;
    }
}

}  // namespace dbmm

namespace dbmm {

// A[ad][b][:] = b1 + sum over the D-slices' partial tiles, plus the fp64 BatchNorm column sums (train only).
struct ReduceStatsArgs {
    const float* part; int ksplit, nad, B, H;
    const float* b1[2];
    float* A;          // [nad][B][H]
    fx64* colsum;      // [nad][2][H] or nullptr (fixed point, FX_COLSUM)
    P2pArgs p2p;       // data parallel over peer memory: the last CTA all-reduces the column sums in place (channel 0)
    fx64* zero_dgb; int zero_dgb_n;             // fused step tail: CTA 0 resets (dgamma, dbeta), which the row kernel
    float* zero_S; int zero_S_n;                //                  S -- the row kernel accumulates both next
};
constexpr int RS_ROWS = 16, RS_MAXK = 16, RS_THREADS = 256;

// 256 threads = 8 row lanes x 32 column quads (H <= 128), two rows per thread; the D-slice partials of one element are
// fetched together (RS_MAXK independent 16-byte loads in flight) before they are summed.  Few, fat CTAs on purpose:
// the column-sum atomics (64-bit fixed point, see fx64) of all CTAs hit the same 2H addresses and the L2 serialises them
// (~40 cycles each).
__global__ void __launch_bounds__(RS_THREADS) k_reduce_stats(ReduceStatsArgs a) {
    __shared__ float sS[2][8][DBMM_MAX_H];
    const int H = a.H, H4 = H >> 2;
    const int ad = blockIdx.y;
    const int tid = threadIdx.x, q = tid >> 5, c = tid & 31;
    const size_t plane = (size_t)a.nad * a.B * H;
    float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
    DBMM_TL_SCOPE(TL_REDUCE);
    ptx::pdl_wait();
    DBMM_TL_WAIT(TL_REDUCE);
    ptx::pdl_launch();
    if (a.zero_dgb && blockIdx.x == 0 && blockIdx.y == 0)
        for (int e = tid; e < a.zero_dgb_n; e += RS_THREADS) a.zero_dgb[e].v = 0;
    if (a.zero_S) {
        const int cta = blockIdx.y * gridDim.x + blockIdx.x, ncta = gridDim.x * gridDim.y;
        for (int e = cta * RS_THREADS + tid; e < a.zero_S_n; e += ncta * RS_THREADS) a.zero_S[e] = 0.f;
    }
    if (c < H4) {
        const float4 bias = __ldg(reinterpret_cast<const float4*>(a.b1[ad]) + c);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int r = blockIdx.x * RS_ROWS + q + 8 * half;
            if (r >= a.B) continue;
            const size_t off = ((size_t)ad * a.B + r) * H + c * 4;
            float4 v[RS_MAXK];
#pragma unroll
            for (int kp = 0; kp < RS_MAXK; ++kp)
                v[kp] = kp < a.ksplit ? __ldcg(reinterpret_cast<const float4*>(a.part + kp * plane + off)) : make_float4(0.f, 0.f, 0.f, 0.f);
            float4 acc = bias;
#pragma unroll
            for (int kp = 0; kp < RS_MAXK; ++kp) { acc.x += v[kp].x; acc.y += v[kp].y; acc.z += v[kp].z; acc.w += v[kp].w; }
            *reinterpret_cast<float4*>(a.A + off) = acc;
            s1[0] += acc.x; s1[1] += acc.y; s1[2] += acc.z; s1[3] += acc.w;
            s2[0] = fmaf(acc.x, acc.x, s2[0]); s2[1] = fmaf(acc.y, acc.y, s2[1]);
            s2[2] = fmaf(acc.z, acc.z, s2[2]); s2[3] = fmaf(acc.w, acc.w, s2[3]);
        }
    }
    if (!a.colsum) return;
    if (c < H4) {
#pragma unroll
        for (int e = 0; e < 4; ++e) { sS[0][q][c * 4 + e] = s1[e]; sS[1][q][c * 4 + e] = s2[e]; }
    }
    __syncthreads();
    for (int e = tid; e < 2 * H; e += RS_THREADS) {
        const int which = e / H, j = e - which * H;
        double v = 0.0;
#pragma unroll
        for (int rr = 0; rr < 8; ++rr) v += (double)sS[which][rr][j];
        fx_add<FX_COLSUM>(&a.colsum[((size_t)ad * 2 + which) * H + j], v);
    }
    if (a.p2p.world) p2p_allreduce_when_last(a.p2p, 0, a.colsum, a.nad * 2 * H, gridDim.x * gridDim.y);
}

}  // namespace dbmm
