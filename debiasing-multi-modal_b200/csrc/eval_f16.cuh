// Eval forward over an fp16-resident embedding matrix (validate / validate_zs, final_main.py:655-803): CLIP emits fp16, the
// pack stores it losslessly (pack.py), and this path reads those 2 bytes per element straight into kind::f16 tensor-core MMAs
// (twice the tf32 rate, half the HBM and L2 bytes of the fp32 path).
//
//   k_f16_gemm<G1>   a = x W1^T, W1 = Wh + 2^-11 Wl as two fp16 operands (Wl pre-scaled by 2^11 so that it stays out of the
//                    fp16 subnormals: 22 significant bits, the same as tf32 hi + lo) into TWO TMEM accumulators;
//                    epilogue: h = relu(BatchNorm_running(a + b1)), split the same way, stored as fp16 pairs (512 B per row).
//   k_f16_gemm<HS>   t = h Q (Q = W2^T W2, the symmetric block of the Gram matrix; three scaled terms, two accumulators);
//                    epilogue: n^2 = h.(t + q) + h.q + b2.b2, prompt scores s_c = h.M_c + m0_c on the CUDA cores, then the
//                    whole finish -- logits, CE, argmax, per-group counters (update_dict, final_main.py:383-391) -- in place.
// Both are persistent (one CTA per SM walks the 128-row tiles), TMA-fed (SWIZZLE_128B boxes of 64 halfs x 128 rows) with the
// accumulator pair double-buffered in TMEM (2 x 2 x 128 columns) so the epilogue of tile i overlaps the MMAs of tile i + 1.
#pragma once
#include <cuda_fp16.h>
#include "tc_gemm.cuh"

namespace dbmm {

constexpr int EF_THREADS = 192, EF_THREADS_HS = 320, EF_BM = 128, EF_BN = 128, EF_BK = 64;          // BK in halfs: 128-byte rows
constexpr int EF_TILE_BYTES = 128 * 128;
constexpr float EF_LO_SCALE = 2048.f, EF_LO_INV = 1.0f / 2048.f;
enum { EF_G1 = 0, EF_HS = 1 };

struct EvalF16Args {
    int64_t M; int K;                         // rows, contraction length (D for G1, H = 128 for HS)
    // G1 epilogue
    const float2* bn_affine;                  // [128]: h = relu(acc * x + y)   (bias, running statistics, gamma, beta folded)
    __half* h_hi; __half* h_lo;               // [M][128] each
    // HS epilogue
    const float* gb;                          // [128]: Gram row H (q), added to t
    const float* gt;                          // [128][8]: per hidden unit j: {q_j, M_j0, M_j1, M_j2, M_j3, 0, 0, 0}
    const float* scal;                        // [8]: G[H][H] = b2.b2, then G[H][H+1+c] = b2.That_c
    int C, G, nad_pass;                       // nad_pass: 0 = only / trainable adapter (finish), 1 = frozen adapter (park n^2, s_c)
    float* park; const float* parked;         // [M][8]: n^2, s_0..s_3 of the frozen adapter (written in pass 1, read by the finish)
    float w_old, inv_tau;
    const int32_t* y; const int32_t* grp; int64_t pos0, batch_size;
    double* loss_sum; int64_t* counts; float* logits_out; int32_t* pred_out;
};

static int make_tmap_2d_f16(CUtensorMap* map, const __half* base, int64_t rows, int64_t cols, int64_t ld) {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        DBMM_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
        DBMM_CHECK_ARG(p != nullptr && q == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled is not available in this driver");
        fn = (PFN_encodeTiled)p;
    }
    DBMM_CHECK_ARG(((uintptr_t)base & 15) == 0 && ld % 8 == 0, "TMA operands need 16-byte aligned rows (ld=%lld halfs)", (long long)ld);
    const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t gstr[1] = {(cuuint64_t)ld * sizeof(__half)};
    const cuuint32_t box[2] = {EF_BK, 128};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void*)base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    DBMM_CHECK_ARG(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (fp16) failed with code %d", (int)r);
    return DBMM_OK;
}

template <int MODE>
struct EfCfg {
    static constexpr int TILES = MODE == EF_G1 ? 3 : 2;                 // G1: x, Wh, Wl;  HS: h_hi, h_lo (Q stays resident)
    static constexpr int STAGE_BYTES = TILES * EF_TILE_BYTES;
    static constexpr int STAGES = 4;
    static constexpr int RESIDENT_BYTES = MODE == EF_G1 ? 0 : 4 * EF_TILE_BYTES;      // HS: Q_hi kb0, kb1, Q_lo kb0, kb1
    static constexpr int CONST_BYTES = 12288;
    static constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + RESIDENT_BYTES + CONST_BYTES + 1024 + 256;
};

template <int MODE>
__global__ void __launch_bounds__(MODE == EF_HS ? EF_THREADS_HS : EF_THREADS, 1)
k_f16_gemm(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapAlo,
           const __grid_constant__ CUtensorMap mapBhi, const __grid_constant__ CUtensorMap mapBlo, EvalF16Args a) {
    using Cfg = EfCfg<MODE>;
    constexpr int S = Cfg::STAGES;
    extern __shared__ uint8_t ef_smem_raw[];
    uint8_t* smem = ef_smem_raw + ((1024u - (ptx::smem_u32(ef_smem_raw) & 1023u)) & 1023u);
    uint8_t* sQ = smem + (size_t)S * Cfg::STAGE_BYTES;                  // HS: resident Q operand
    float* sConst = (float*)(sQ + Cfg::RESIDENT_BYTES);                 // G1: [128] float2 affine;  HS: [128][8] gt, [128] gb, [128][8] partials
    uint64_t* full = (uint64_t*)((uint8_t*)sConst + Cfg::CONST_BYTES);
    uint64_t* empty = full + S;
    uint64_t* tmem_full = empty + S;
    uint64_t* tmem_empty = tmem_full + 2;
    uint64_t* q_full = tmem_empty + 2;
    uint32_t* tmem_ptr = (uint32_t*)(q_full + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int total_tiles = (int)((a.M + EF_BM - 1) / EF_BM);
    const int KB = (a.K + EF_BK - 1) / EF_BK;
    constexpr int NT = MODE == EF_HS ? EF_THREADS_HS : EF_THREADS;
    constexpr int EPI_THREADS = NT - 64;

    if (tid == 0) {
        // HS: a stage is also held by the epilogue threads, which read h back from it (count: the MMA commit + every epilogue thread)
        for (int s = 0; s < S; ++s) { ptx::mbar_init(&full[s], 1); ptx::mbar_init(&empty[s], MODE == EF_HS ? 1 + EPI_THREADS : 1); }
        ptx::mbar_init(q_full, 1);
        for (int s = 0; s < 2; ++s) { ptx::mbar_init(&tmem_full[s], 1); ptx::mbar_init(&tmem_empty[s], EPI_THREADS); }
        ptx::fence_mbar_init();
        ptx::tma_prefetch_desc(&mapA); ptx::tma_prefetch_desc(&mapBhi); ptx::tma_prefetch_desc(&mapBlo);
        if (MODE == EF_HS) ptx::tma_prefetch_desc(&mapAlo);
    }
    if (MODE == EF_G1) {
        for (int e = tid; e < 128; e += NT) reinterpret_cast<float2*>(sConst)[e] = a.bn_affine[e];
    } else {
        for (int e = tid; e < 128 * 8; e += NT) sConst[e] = a.gt[e];
        for (int e = tid; e < 128; e += NT) sConst[128 * 8 + e] = a.gb[e];
    }
    if (warp == 0) ptx::tmem_alloc<512>(tmem_ptr);
    ptx::tc_fence_before_sync();
    __syncthreads();
    ptx::tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        // ===================== TMA producer: one thread =====================
        if (lane == 0) {
            if (MODE == EF_HS) {                                     // Q (K = 128: two k-blocks, hi and lo) once per CTA
                ptx::mbar_arrive_expect_tx(q_full, Cfg::RESIDENT_BYTES);
                ptx::tma_load_2d(&mapBhi, q_full, sQ, 0, 0);
                ptx::tma_load_2d(&mapBhi, q_full, sQ + EF_TILE_BYTES, EF_BK, 0);
                ptx::tma_load_2d(&mapBlo, q_full, sQ + 2 * EF_TILE_BYTES, 0, 0);
                ptx::tma_load_2d(&mapBlo, q_full, sQ + 3 * EF_TILE_BYTES, EF_BK, 0);
            }
            uint32_t g = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int m0 = tile * EF_BM;
                for (int kb = 0; kb < KB; ++kb, ++g) {
                    const int s = g % S;
                    ptx::mbar_wait(&empty[s], ((g / S) & 1) ^ 1);
                    uint8_t* st = smem + (size_t)s * Cfg::STAGE_BYTES;
                    ptx::mbar_arrive_expect_tx(&full[s], Cfg::STAGE_BYTES);
                    const int k0 = kb * EF_BK;
                    int t = 0;
                    ptx::tma_load_2d(&mapA, &full[s], st + (t++) * EF_TILE_BYTES, k0, m0);
                    if (MODE == EF_HS) ptx::tma_load_2d(&mapAlo, &full[s], st + (t++) * EF_TILE_BYTES, k0, m0);
                    else {
                        ptx::tma_load_2d(&mapBhi, &full[s], st + (t++) * EF_TILE_BYTES, k0, 0);
                        ptx::tma_load_2d(&mapBlo, &full[s], st + (t++) * EF_TILE_BYTES, k0, 0);
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer: one thread =====================
        if (lane == 0) {
            constexpr uint32_t idesc = ptx::umma_idesc(/*f16*/ 0, EF_BM, EF_BN, 0, 0);
            uint32_t g = 0, it = 0;
            if (MODE == EF_HS) { ptx::mbar_wait(q_full, 0); ptx::tc_fence_after_sync(); }
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
                const uint32_t as = it & 1u;
                const uint32_t acc0 = tmem_base + as * 256, acc1 = acc0 + 128;
                ptx::mbar_wait(&tmem_empty[as], ((it >> 1) & 1) ^ 1);
                ptx::tc_fence_after_sync();
                for (int kb = 0; kb < KB; ++kb, ++g) {
                    const int s = g % S;
                    ptx::mbar_wait(&full[s], (g / S) & 1);
                    ptx::tc_fence_after_sync();
                    const uint32_t base = ptx::smem_u32(smem + (size_t)s * Cfg::STAGE_BYTES);
                    const uint32_t sAhi = base, sAlo = base + EF_TILE_BYTES;
                    const uint32_t sBhi = MODE == EF_HS ? ptx::smem_u32(sQ) + (uint32_t)kb * EF_TILE_BYTES : base + EF_TILE_BYTES;
                    const uint32_t sBlo = MODE == EF_HS ? sBhi + 2 * EF_TILE_BYTES : sBhi + EF_TILE_BYTES;
#pragma unroll
                    for (int kk = 0; kk < EF_BK / 16; ++kk) {
                        const uint64_t ahi = ptx::umma_smem_desc(sAhi + kk * 32, 0, 1024);
                        const uint64_t bhi = ptx::umma_smem_desc(sBhi + kk * 32, 0, 1024);
                        const uint64_t blo = ptx::umma_smem_desc(sBlo + kk * 32, 0, 1024);
                        const uint32_t acc = (kb | kk) != 0 ? 1u : 0u;
                        ptx::mma_f16_ss(acc0, ahi, bhi, idesc, acc);
                        ptx::mma_f16_ss(acc1, ahi, blo, idesc, acc);
                        if (MODE == EF_HS) {
                            const uint64_t alo = ptx::umma_smem_desc(sAlo + kk * 32, 0, 1024);
                            ptx::mma_f16_ss(acc1, alo, bhi, idesc, 1u);
                        }
                    }
                    ptx::mma_commit(&empty[s]);
                }
                ptx::mma_commit(&tmem_full[as]);
            }
        }
        __syncwarp();
    } else {
        // ===================== epilogue: warps 2..5, TMEM lane quarter = warp % 4 =====================
        const int q = warp & 3;
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
            const int64_t m = (int64_t)tile * EF_BM + q * 32 + lane;
            const bool row_ok = m < a.M;
            const uint32_t as = it & 1u;
            const uint32_t acc0 = tmem_base + as * 256 + ((uint32_t)(q * 32) << 16), acc1 = acc0 + 128;
            ptx::mbar_wait(&tmem_full[as], (it >> 1) & 1);
            ptx::tc_fence_after_sync();
            if (MODE == EF_G1) {
                const float2* aff = reinterpret_cast<const float2*>(sConst);
#pragma unroll 1
                for (int ch = 0; ch < 4; ++ch) {
                    uint32_t r0[32], r1[32];
                    ptx::tmem_ld_32x32b_x32(acc0 + ch * 32, r0);
                    ptx::tmem_ld_32x32b_x32(acc1 + ch * 32, r1);
                    ptx::tmem_ld_wait();
                    if (!row_ok) continue;
                    uint32_t hp[16], lp[16];
#pragma unroll
                    for (int j = 0; j < 32; j += 2) {
                        float hv[2];
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const float2 c2 = aff[ch * 32 + j + e];
                            const float acc = fmaf(__uint_as_float(r1[j + e]), EF_LO_INV, __uint_as_float(r0[j + e]));
                            hv[e] = fmaxf(fmaf(acc, c2.x, c2.y), 0.f);
                        }
                        const __half2 hh = __floats2half2_rn(hv[0], hv[1]);
                        const float2 back = __half22float2(hh);
                        const __half2 ll = __floats2half2_rn((hv[0] - back.x) * EF_LO_SCALE, (hv[1] - back.y) * EF_LO_SCALE);
                        hp[j >> 1] = *reinterpret_cast<const uint32_t*>(&hh);
                        lp[j >> 1] = *reinterpret_cast<const uint32_t*>(&ll);
                    }
                    uint4* dh = reinterpret_cast<uint4*>(a.h_hi + (size_t)m * 128 + ch * 32);
                    uint4* dl = reinterpret_cast<uint4*>(a.h_lo + (size_t)m * 128 + ch * 32);
#pragma unroll
                    for (int v = 0; v < 4; ++v) {
                        dh[v] = make_uint4(hp[4 * v], hp[4 * v + 1], hp[4 * v + 2], hp[4 * v + 3]);
                        dl[v] = make_uint4(lp[4 * v], lp[4 * v + 1], lp[4 * v + 2], lp[4 * v + 3]);
                    }
                }
            } else {
                const float* gt = sConst;
                const float* gb = sConst + 128 * 8;
                float* sPart = sConst + 128 * 8 + 128;             // [128 rows][8]: partial sums of the upper column half
                const int ch0 = warp >= 6 ? 2 : 0;                 // warps 2..5: hidden units 0..63, warps 6..9: 64..127
                float dot = 0.f, tq = 0.f, sc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int cc = 0; cc < 2; ++cc) {
                    const int ch = ch0 + cc;
                    uint32_t r0[32], r1[32];
                    ptx::tmem_ld_32x32b_x32(acc0 + ch * 32, r0);
                    ptx::tmem_ld_32x32b_x32(acc1 + ch * 32, r1);
                    // this row's h = hi + 2^-11 lo, straight from the operand tiles TMA put in shared memory (k-block ch / 2;
                    // SWIZZLE_128B: 16-byte chunk c of row r sits at r * 128 + ((c ^ (r & 7)) << 4))
                    uint4 hq4[4], lq4[4];
                    {
                        const uint32_t gk = 2 * it + (uint32_t)(ch >> 1);
                        ptx::mbar_wait(&full[gk % S], (gk / S) & 1);               // (already complete: acquires the TMA writes)
                        const uint8_t* stg = smem + (size_t)(gk % S) * Cfg::STAGE_BYTES;
                        const uint32_t r = (uint32_t)(q * 32 + lane);
#pragma unroll
                        for (int v = 0; v < 4; ++v) {
                            const uint32_t c16 = (uint32_t)((ch & 1) * 4 + v);
                            const uint32_t off = r * 128u + ((c16 ^ (r & 7u)) << 4);
                            hq4[v] = *reinterpret_cast<const uint4*>(stg + off);
                            lq4[v] = *reinterpret_cast<const uint4*>(stg + EF_TILE_BYTES + off);
                        }
                    }
                    ptx::tmem_ld_wait();
                    const uint32_t* hw = reinterpret_cast<const uint32_t*>(hq4);
                    const uint32_t* lw = reinterpret_cast<const uint32_t*>(lq4);
#pragma unroll
                    for (int j = 0; j < 32; j += 2) {
                        const float2 hh = __half22float2(*reinterpret_cast<const __half2*>(&hw[j >> 1]));
                        const float2 ll = __half22float2(*reinterpret_cast<const __half2*>(&lw[j >> 1]));
                        const float hx[2] = {fmaf(ll.x, EF_LO_INV, hh.x), fmaf(ll.y, EF_LO_INV, hh.y)};
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int jj = ch * 32 + j + e;
                            const float t = fmaf(__uint_as_float(r1[j + e]), EF_LO_INV, __uint_as_float(r0[j + e])) + gb[jj];
                            const float4 g0 = *reinterpret_cast<const float4*>(gt + jj * 8);
                            const float g4 = gt[jj * 8 + 4];
                            dot = fmaf(t, hx[e], dot);
                            tq = fmaf(hx[e], g0.x, tq);
                            sc[0] = fmaf(hx[e], g0.y, sc[0]); sc[1] = fmaf(hx[e], g0.z, sc[1]); sc[2] = fmaf(hx[e], g0.w, sc[2]);
                            sc[3] = fmaf(hx[e], g4, sc[3]);
                        }
                    }
                }
                // the accumulator and the h tiles have been read: hand the TMEM stage and both shared-memory stages back
                ptx::tc_fence_before_sync();
                ptx::mbar_arrive(&tmem_empty[as]);
                ptx::mbar_arrive(&empty[(2 * it) % S]);
                ptx::mbar_arrive(&empty[(2 * it + 1) % S]);
                const int prow = q * 32 + lane;
                if (warp >= 6) {
                    float4* pp = reinterpret_cast<float4*>(sPart + prow * 8);
                    pp[0] = make_float4(dot, tq, sc[0], sc[1]);
                    pp[1] = make_float4(sc[2], sc[3], 0.f, 0.f);
                }
                asm volatile("bar.sync 3, 256;" ::: "memory");
                if (warp >= 6) { asm volatile("bar.sync 4, 256;" ::: "memory"); continue; }
                {
                    const float4 p0 = *reinterpret_cast<const float4*>(sPart + prow * 8), p1 = *reinterpret_cast<const float4*>(sPart + prow * 8 + 4);
                    dot += p0.x; tq += p0.y; sc[0] += p0.z; sc[1] += p0.w; sc[2] += p1.x; sc[3] += p1.y;
                }
                asm volatile("bar.sync 4, 256;" ::: "memory");       // sPart may be rewritten for the next tile
                const float n2 = dot + tq + __ldg(a.scal);
#pragma unroll
                for (int c = 0; c < 4; ++c) sc[c] += __ldg(a.scal + 1 + c);
                if (a.nad_pass == 1) {
                    if (row_ok) {
                        float4* p = reinterpret_cast<float4*>(a.park + (size_t)m * 8);
                        p[0] = make_float4(n2, sc[0], sc[1], sc[2]);
                        p[1] = make_float4(sc[3], 0.f, 0.f, 0.f);
                    }
                } else {
                    int gv = -1, corr = 0; float nll = 0.f;
                    const int64_t pos = a.pos0 + m;
                    if (row_ok) {
                        const int yv = a.y ? a.y[pos] : -1;
                        gv = a.grp ? a.grp[pos] : 0;
                        const float inv_n1 = 1.0f / sqrtf(n2);
                        float so[4] = {0.f, 0.f, 0.f, 0.f}, inv_n0 = 0.f;
                        const bool two = a.parked != nullptr;
                        if (two) {
                            const float4 p0 = __ldcg(reinterpret_cast<const float4*>(a.parked + (size_t)m * 8));
                            const float4 p1 = __ldcg(reinterpret_cast<const float4*>(a.parked + (size_t)m * 8) + 1);
                            inv_n0 = 1.0f / sqrtf(p0.x); so[0] = p0.y; so[1] = p0.z; so[2] = p0.w; so[3] = p1.x;
                        }
                        const float coef = two ? (1.0f - a.w_old) : 1.0f;
                        float l[4], mx = -INFINITY; int am = 0;
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            l[c] = -INFINITY;
                            if (c < a.C) {
                                const float lnew = a.inv_tau * sc[c] * inv_n1;
                                float v = lnew;
                                if (two) v = fmaf(coef, lnew, a.w_old * a.inv_tau * so[c] * inv_n0);
                                l[c] = v;
                                if (v > mx) { mx = v; am = c; }
                                if (a.logits_out) a.logits_out[(size_t)pos * a.C + c] = v;
                            }
                        }
                        float se = 0.f, ly = 0.f;
#pragma unroll
                        for (int c = 0; c < 4; ++c) if (c < a.C) { se += expf(l[c] - mx); if (c == yv) ly = l[c]; }
                        nll = a.y ? (logf(se) + mx - ly) : 0.f;
                        corr = am == yv;
                        if (a.pred_out) a.pred_out[pos] = am;
                    }
                    const int64_t slot = row_ok ? pos / a.batch_size : -1;
                    const int64_t slot0 = __shfl_sync(0xffffffffu, slot, 0);
                    const bool uniform = __all_sync(0xffffffffu, !row_ok || slot == slot0);
                    if (uniform) {
                        const double tot = warp_sum((double)nll);
                        if (lane == 0 && a.loss_sum && slot0 >= 0) atomicAdd(&a.loss_sum[slot0], tot);
                        const unsigned cmask = __ballot_sync(0xffffffffu, corr != 0);
                        for (int g = 0; g < a.G; ++g) {
                            const unsigned gm = __ballot_sync(0xffffffffu, gv == g);
                            if (lane == 0 && gm && a.counts && slot0 >= 0) {
                                int64_t* cnt = a.counts + (size_t)slot0 * 2 * a.G;
                                const int nc = __popc(gm & cmask);
                                if (nc) atomicAdd((unsigned long long*)&cnt[g], (unsigned long long)nc);
                                atomicAdd((unsigned long long*)&cnt[a.G + g], (unsigned long long)__popc(gm));
                            }
                        }
                    } else if (row_ok) {
                        if (a.loss_sum) atomicAdd(&a.loss_sum[slot], (double)nll);
                        if (a.counts && gv >= 0 && gv < a.G) {
                            int64_t* cnt = a.counts + (size_t)slot * 2 * a.G;
                            if (corr) atomicAdd((unsigned long long*)&cnt[gv], 1ull);
                            atomicAdd((unsigned long long*)&cnt[a.G + gv], 1ull);
                        }
                    }
                }
            }
            if (MODE == EF_G1) {
                ptx::tc_fence_before_sync();
                ptx::mbar_arrive(&tmem_empty[as]);
            }
        }
    }
    __syncthreads();
    if (warp == 0) ptx::tmem_dealloc<512>(tmem_base);
}

// W1 -> fp16 pair (hi, (w - hi) * 2^11);  [H][D] each
__global__ void __launch_bounds__(256) k_split_f16(const float* __restrict__ w, __half* __restrict__ hi, __half* __restrict__ lo, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float v = w[i];
        const __half h = __float2half_rn(v);
        hi[i] = h;
        lo[i] = __float2half_rn((v - __half2float(h)) * EF_LO_SCALE);
    }
}

// Everything the two kernels need from one adapter besides W1: the folded BatchNorm affine, the Q block of the Gram matrix as
// an fp16 pair (K-major: Q is symmetric), and the fp32 side columns.
struct EvalF16Prep {
    const float* gram; int H, C;              // [H+1][H+1+C]
    const float* b1; const float* mean; const float* var; const float* gamma; const float* beta;
    float2* affine; __half* q_hi; __half* q_lo; float* gt; float* gb; float* scal;       // scal[0] = G[H][H], scal[1 + c] = G[H][H+1+c]
};
__global__ void __launch_bounds__(256) k_eval_f16_prep(EvalF16Prep p) {
    const int H = p.H, ldg = H + 1 + p.C;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < H * H; e += gridDim.x * blockDim.x) {
        const float v = p.gram[(size_t)(e / H) * ldg + (e % H)];
        const __half h = __float2half_rn(v);
        p.q_hi[e] = h;
        p.q_lo[e] = __float2half_rn((v - __half2float(h)) * EF_LO_SCALE);
    }
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < H; j += gridDim.x * blockDim.x) {
        const float rstd = 1.0f / sqrtf(p.var[j] + DBMM_BN_EPS);
        const float x = rstd * p.gamma[j];
        p.affine[j] = make_float2(x, (p.b1[j] - p.mean[j]) * x + p.beta[j]);
        p.gb[j] = p.gram[(size_t)H * ldg + j];
        float* g = p.gt + (size_t)j * 8;
        g[0] = p.gram[(size_t)j * ldg + H];
        for (int c = 0; c < 4; ++c) g[1 + c] = c < p.C ? p.gram[(size_t)j * ldg + H + 1 + c] : 0.f;
        g[5] = g[6] = g[7] = 0.f;
    }
    if (blockIdx.x == 0 && threadIdx.x < 8) {
        const int t = threadIdx.x;
        p.scal[t] = t <= p.C ? p.gram[(size_t)H * ldg + H + t] : 0.f;
    }
}

template <int MODE>
static int launch_f16_gemm(const __half* A, const __half* Alo, int64_t lda, const __half* Bhi, const __half* Blo, int64_t ldb,
                           const EvalF16Args& a, cudaStream_t st) {
    using Cfg = EfCfg<MODE>;
    CUtensorMap mA, mAlo, mBhi, mBlo;
    if (int rc = make_tmap_2d_f16(&mA, A, a.M, a.K, lda)) return rc;
    if (int rc = make_tmap_2d_f16(&mAlo, MODE == EF_HS ? Alo : A, a.M, a.K, lda)) return rc;
    if (int rc = make_tmap_2d_f16(&mBhi, Bhi, 128, a.K, ldb)) return rc;
    if (int rc = make_tmap_2d_f16(&mBlo, Blo, 128, a.K, ldb)) return rc;
    auto kern = k_f16_gemm<MODE>;
    DBMM_CUDA(set_smem(kern, Cfg::SMEM));
    int grid = (int)((a.M + EF_BM - 1) / EF_BM);
    if (grid > 148) grid = 148;
    kern<<<grid, MODE == EF_HS ? EF_THREADS_HS : EF_THREADS, Cfg::SMEM, st>>>(mA, mAlo, mBhi, mBlo, a);
    DBMM_LAUNCH_CHECK();
    return DBMM_OK;
}

}  // namespace dbmm
