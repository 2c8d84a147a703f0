// Training row phase on the 5th-generation tensor cores (H-space formulation; reference math:
// final_main.py:66-80 / 121-140 forward, autograd backward of final_main.py:465/542/622).
//
// One CTA owns HR_ROWS = 64 batch rows.  Per adapter in the forward (frozen + trainable in stage 2):
//   P0   h = relu(BatchNorm_batch(a)) is formed from the pre-activations and split hi + lo into a K-major SWIZZLE_128B
//        operand tile (K = hidden unit); the Gram matrix G = [W2 | b2]^T [W2 | b2 | That] becomes the B operand
//        (N = Gram column, 144 with padding; K = Gram row < H; row H, the bias row of [h, 1], is added in the epilogue).
//   MMA1 T = h G  as 3xTF32 (hi*hi + lo*hi + hi*lo, fp32 accumulate in TMEM): 48 tcgen05.mma of M128 N144 K8 -- the upper 64
//        accumulator lanes belong to operand rows nobody wrote and are never read.
//   E1   one thread per batch row reads its TMEM lane twice: pass 1 -> n^2 = h.t + t_H, prompt scores, logits, CE, argmax,
//        counters, the row's gradient coefficients (c, ds); pass 2 -> dh = c t + ds M^T, ReLU / BatchNorm-backward inputs
//        (dahat out; dgamma, dbeta as fixed-point sums), and the operands of the second contraction.
//   MMA2 S^T_part = h^T L  with L = [c*h | c | ds] (K = the CTA's 64 batch rows; both operands written TRANSPOSED by the row
//        threads so that they are plain K-major tiles): 24 MMAs; the tile goes to Spart[tile] with plain stores.
// No atomics on the large reductions, no CUDA-core contraction: round 1's k_rows_train spent 11 us per 1024-row step on the
// 129 x 131 contraction (8 rows per CTA, 128 CTAs) and 2.2 M fp32 atomics; here a step is 16 CTAs, and 64 sweep members
// fill the machine (batched.cuh).  k_sum_spart adds the tiles in tile order (deterministic) into S = L^T [h | 1].
#pragma once
#include "p2p.cuh"
#include "kernels_simt.cuh"
#include "ptx_sm100.cuh"

namespace dbmm {

constexpr int HR_ROWS = 64, HR_THREADS = 512, HR_N = 144, HR_H = 128;
constexpr int HR_B_KT_BYTES = HR_N * 128;                    // one k-tile (32 floats of K) of a 144-row operand
constexpr int HR_A_KT_BYTES = HR_ROWS * 128;                 // one k-tile of the 64-row h operand (MMA1)
constexpr int HR_A2_KT_BYTES = 128 * 128;                    // one k-tile of the 128-row h^T operand (MMA2)
constexpr int HR_SB_BYTES = 2 * 4 * HR_B_KT_BYTES;           // hi | lo, 4 k-tiles                     = 147,456
constexpr int HR_SA_BYTES = 2 * 4 * HR_A_KT_BYTES;           // hi | lo, 4 k-tiles of 64 rows          =  65,536
constexpr int HR_CONST_BYTES = 11264;                        // constants; also absorbs the M = 128 over-read of the last A k-tile
constexpr size_t HR_SMEM = (size_t)HR_SB_BYTES + HR_SA_BYTES + HR_CONST_BYTES + 1024 + 256;
constexpr int HR_SP_LD = HR_N;                               // row stride of an S^T partial tile [H+1][144]
constexpr int HR_CT = 4;                                     // prompt columns kept in shared memory / registers (C <= 4 on this path)

struct HsRowsArgs {
    int B; int64_t Bg;
    const int32_t* idx; const int32_t* y; const int32_t* grp;
    int H, C, G, nad;
    const float* A; int64_t strideA;      // [nad][B][H]
    const float* gram;                    // [nad][H+1][H+1+C]
    const fx64* colsum;                   // [nad][2][H]
    AdapterView ad[2];
    float w_old, inv_tau, inv_B;
    double* loss_sum; int64_t* counts; int64_t slot;
    float* dahat; fx64* dgb;              // [B][H]; [2][H] (+=, FX_DGB)
    float* logits_out; const float* dlogits_in;
    float* Spart;                         // [tiles][H+1][HR_SP_LD]: rows j < H: sum_rows h_j L; row H: sum_rows L
};

static inline bool hs_rows_supported(int H, int C) { return H == HR_H && C >= 1 && C <= HR_CT; }
static inline size_t hs_spart_floats(int B) { return (size_t)((B + HR_ROWS - 1) / HR_ROWS) * (HR_H + 1) * HR_SP_LD; }

#ifdef DBMM_PHASE_TIMERS
__device__ long long g_phase_clk[2][16];        // [kernel][phase]: clock64 of CTA 0 (development builds only)
#define HR_TICK(K, I) do { if (blockIdx.x == 0 && blockIdx.y == 0 && (threadIdx.x & 31) == 0 && (threadIdx.x >> 5) == ((I) < 0 ? 2 : 0)) g_phase_clk[K][(I) < 0 ? -(I) : (I)] = clock64(); } while (0)
#else
#define HR_TICK(K, I) do { } while (0)
#endif

__device__ __forceinline__ void hr_split(float x, float& hi, float& lo) {
    hi = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
    lo = x - hi;
}

// Sum over the 32 lanes of v[c] for every column c < 32; lane l returns the total of column l.  Butterfly: 31 shuffles + adds
// (a shared-memory transpose costs 64 accesses and a buffer per warp).
__device__ __forceinline__ float warp_colsum32(float (&v)[32], int lane) {
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        const bool up = (lane & s) != 0;
#pragma unroll
        for (int i = 0; i < s; ++i) {
            const float keep = up ? v[i + s] : v[i];
            const float send = up ? v[i] : v[i + s];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
        }
    }
    return v[0];
}

// A warp holds a 32 x 32 block row-per-thread (lane = row, v[q] = column q).  Written out through a 32 x 33 shared-memory
// transpose so that every store instruction covers 128 contiguous bytes of ONE row (row-per-thread 16-byte stores touch 32
// sectors per instruction).  rows: valid rows of the block.
__device__ __forceinline__ void warp_store_block32(const float (&v)[32], float* scr, float* gbase, size_t ld, int rows, int lane) {
    // scr: 32 rows x 36 floats (16-byte aligned rows; a quarter warp's 16-byte stores hit 32 distinct banks)
#pragma unroll
    for (int q = 0; q < 32; q += 4) *reinterpret_cast<float4*>(scr + lane * 36 + q) = make_float4(v[q], v[q + 1], v[q + 2], v[q + 3]);
    __syncwarp();
    const int rsub = lane >> 3, c4 = (lane & 7) * 4;                     // 8 lanes cover the 128 bytes of one row, 4 rows per instruction
#pragma unroll
    for (int r0 = 0; r0 < 32; r0 += 4) {
        const int rr = r0 + rsub;
        if (rr < rows) *reinterpret_cast<float4*>(gbase + (size_t)rr * ld + c4) = *reinterpret_cast<const float4*>(scr + rr * 36 + c4);
    }
    __syncwarp();
}

__device__ __forceinline__ void hs_rows_body(const HsRowsArgs& a) {
    extern __shared__ uint8_t hr_smem_raw[];
    uint8_t* smem = hr_smem_raw + ((1024u - (ptx::smem_u32(hr_smem_raw) & 1023u)) & 1023u);     // (offset, not a cast: keeps the
    uint8_t* sB = smem;                                      // shared address space)  MMA1 B operand; MMA2: B' (first 73,728 B)
    uint8_t* sA = smem + HR_SB_BYTES;                        // MMA1 A operand; MMA2: A'
    float4* sBN = (float4*)(smem + HR_SB_BYTES + HR_SA_BYTES);       // [nad][H]: {rstd, -mean * rstd, gamma, beta}: ahat = a * x + y
    float* sGb = (float*)(sBN + 2 * HR_H);                   // [nad][HR_N]: Gram row H (the bias row of [h, 1])
    float4* sM = (float4*)(sGb + 2 * HR_N);                  // [H]: Gram columns H+1+c of the trainable adapter, c < 4
    float* sRow = (float*)(sM + HR_H);                       // [2][HR_N] per row warp: column sums of L (row H of the S^T tile)
    float* sX = sRow + 2 * HR_N;                             // [nad][3][64] partial h.t of the column quarters 1..3
    float* sXc = sX + 2 * 3 * 64;                            // [64][5] c, ds of the row for the other column quarters
    uint64_t* bars = (uint64_t*)(sXc + 64 * 5);              // [0..1]: MMA1 done (per adapter); [2]: MMA2 operands ready; [3]: MMA2 done
    uint32_t* tmem_ptr = (uint32_t*)(bars + 4);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int H = HR_H, C = a.C, ldg = H + 1 + C, NAD = a.nad;
    const int m0 = blockIdx.x * HR_ROWS;
    const int rows_here = min(HR_ROWS, a.B - m0);

    if (tid == 0) {
        ptx::mbar_init(&bars[0], 1); ptx::mbar_init(&bars[1], 1); ptx::mbar_init(&bars[2], 256); ptx::mbar_init(&bars[3], 1);
        ptx::fence_mbar_init();
    }
    if (warp == 2) ptx::tmem_alloc<512>(tmem_ptr);
    // ---- row threads: warps with (warp & 2) == 0, i.e. 0, 1, 4, 5, 8, 9, 12, 13: batch rows rw * 32 + lane, hidden units of the
    // column quarter qt = warp / 4.  A warp may only touch the TMEM lanes 32 * (warp % 4) .. +31, hence this assignment.
    const bool ep = (warp & 2) == 0;
    const int qt = warp >> 2, rw = warp & 1, j0 = qt * 32;
    const int my_r = m0 + rw * 32 + lane;
    const bool valid = ep && my_r < a.B;
    const uint32_t tlane = ((uint32_t)(rw * 32)) << 16;
    const int xrow = rw * 32 + lane;
    // labels of this thread's row (constants): requested before the dependency wait
    int yv = -1, gval = -1;
    if (valid && qt == 0) {
        const int64_t dsrow = a.idx ? (int64_t)a.idx[my_r] : (int64_t)my_r;
        yv = a.y ? a.y[dsrow] : -1;
        gval = a.grp ? a.grp[dsrow] : 0;
    }
    HR_TICK(0, 0);
    ptx::pdl_wait();                // A / column sums: k_reduce_stats (or the GEMM-1 epilogue); the Gram matrix: the W2 branch
    ptx::pdl_launch();
    HR_TICK(0, 1);

    // B operand of one adapter: B[n][k]: n < H+1: G[n][k] (the Gram block is symmetric: Q = W2^T W2, row H = column H = W2^T b2);
    // n = H+1+c: column H+1+c of G (W2^T That_c);  n >= ldg: zeros.   144 x 32 tasks of 4 floats, 9 per thread, loads first.
    auto build_b = [&](int ad) {
        const float* G = a.gram + (size_t)ad * (H + 1) * ldg;
        float x[9][4];
#pragma unroll
        for (int u = 0; u < 9; ++u) {
            const int task = tid + u * HR_THREADS, n = task >> 5, k = (task & 31) * 4;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float v = 0.f;
                if (n <= H) v = __ldcg(G + (size_t)n * ldg + k + q);
                else if (n < ldg) v = __ldcg(G + (size_t)(k + q) * ldg + n);
                x[u][q] = v;
            }
        }
#pragma unroll
        for (int u = 0; u < 9; ++u) {
            const int task = tid + u * HR_THREADS, n = task >> 5, c16 = task & 31;
            float hi[4], lo[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) hr_split(x[u][q], hi[q], lo[q]);
            const uint32_t off = (uint32_t)(c16 >> 3) * HR_B_KT_BYTES + ptx::sw128_offset(n, c16 & 7);
            *reinterpret_cast<float4*>(sB + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<float4*>(sB + 4 * HR_B_KT_BYTES + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
        }
    };
    // A operand: A[row][k] = h = relu(gamma * ahat + beta), rows past the batch end: zeros.   64 x 32 tasks, 4 per thread.
    auto build_a = [&](int ad) {
        const float4* bn = sBN + (size_t)ad * H;
        float4 av[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int task = tid + u * HR_THREADS, row = task >> 5, k = (task & 31) * 4;
            av[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row < rows_here) av[u] = __ldcg(reinterpret_cast<const float4*>(a.A + (size_t)ad * a.strideA + (size_t)(m0 + row) * H + k));
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int task = tid + u * HR_THREADS, row = task >> 5, c16 = task & 31, k = c16 * 4;
            float hi[4] = {0.f, 0.f, 0.f, 0.f}, lo[4] = {0.f, 0.f, 0.f, 0.f};
            if (row < rows_here) {
                const float ax[4] = {av[u].x, av[u].y, av[u].z, av[u].w};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 c4 = bn[k + q];
                    const float h = fmaxf(fmaf(fmaf(ax[q], c4.x, c4.y), c4.z, c4.w), 0.f);
                    hr_split(h, hi[q], lo[q]);
                }
            }
            const uint32_t off = (uint32_t)(c16 >> 3) * HR_A_KT_BYTES + ptx::sw128_offset(row, c16 & 7);
            *reinterpret_cast<float4*>(sA + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<float4*>(sA + 4 * HR_A_KT_BYTES + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
        }
    };

    build_b(0);                                      // the Gram loads are in flight while the constants are formed
    for (int e = tid; e < NAD * H; e += HR_THREADS) {
        const int ad = e / H, j = e - ad * H;
        const double s1 = fx_get<FX_COLSUM>(&a.colsum[((size_t)ad * 2 + 0) * H + j]), s2 = fx_get<FX_COLSUM>(&a.colsum[((size_t)ad * 2 + 1) * H + j]);
        const double m = s1 / (double)a.Bg;
        double v = s2 / (double)a.Bg - m * m;
        if (v < 0.0) v = 0.0;
        const AdapterView& av = a.ad[NAD == 2 ? ad : 1];
        const float mean = (float)m, rstd = 1.0f / sqrtf((float)v + DBMM_BN_EPS);
        sBN[e] = make_float4(rstd, -mean * rstd, av.gamma[j], av.beta[j]);
    }
    for (int e = tid; e < NAD * HR_N; e += HR_THREADS) {
        const int ad = e / HR_N, n = e - ad * HR_N;
        sGb[e] = n < ldg ? __ldcg(a.gram + ((size_t)ad * (H + 1) + H) * ldg + n) : 0.f;
    }
    if (tid < H) {
        const float* Gt = a.gram + (size_t)(NAD - 1) * (H + 1) * ldg + (size_t)tid * ldg + H + 1;
        sM[tid] = make_float4(C > 0 ? __ldcg(Gt) : 0.f, C > 1 ? __ldcg(Gt + 1) : 0.f, C > 2 ? __ldcg(Gt + 2) : 0.f, C > 3 ? __ldcg(Gt + 3) : 0.f);
    }
    ptx::tc_fence_before_sync();
    __syncthreads();
    ptx::tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_ptr;
    constexpr uint32_t TM_D1 = 0, TM_D1B = 160, TM_D2 = 320;    // accumulator column offsets: adapter 0, adapter 1, S^T tile
    constexpr uint32_t idesc = ptx::umma_idesc(/*tf32*/ 2, 128, HR_N, 0, 0);
    HR_TICK(0, 2);

    auto issue_mma1 = [&](int ad) {            // one thread
        const uint32_t d = tmem_base + (ad == 0 ? TM_D1 : TM_D1B);
        const uint32_t bA = ptx::smem_u32(sA), bB = ptx::smem_u32(sB);
#pragma unroll
        for (int kt = 0; kt < 4; ++kt)
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const uint64_t ahi = ptx::umma_smem_desc(bA + kt * HR_A_KT_BYTES + kk * 32, 0, 1024);
                const uint64_t alo = ptx::umma_smem_desc(bA + 4 * HR_A_KT_BYTES + kt * HR_A_KT_BYTES + kk * 32, 0, 1024);
                const uint64_t bhi = ptx::umma_smem_desc(bB + kt * HR_B_KT_BYTES + kk * 32, 0, 1024);
                const uint64_t blo = ptx::umma_smem_desc(bB + 4 * HR_B_KT_BYTES + kt * HR_B_KT_BYTES + kk * 32, 0, 1024);
                ptx::mma_tf32_ss(d, alo, bhi, idesc, (kt | kk) != 0 ? 1u : 0u);
                ptx::mma_tf32_ss(d, ahi, blo, idesc, 1u);
                ptx::mma_tf32_ss(d, ahi, bhi, idesc, 1u);
            }
        ptx::mma_commit(&bars[ad]);
    };
    auto ep_sync = [&]() { asm volatile("bar.sync 2, 256;" ::: "memory"); };
    // the row's 32 pre-activations of this column quarter, one burst of 16-byte loads
    auto load_a_quarter = [&](int ad, float (&av)[32]) {
        const float* src = a.A + (size_t)ad * a.strideA + (size_t)(valid ? my_r : 0) * H + j0;
#pragma unroll
        for (int q = 0; q < 32; q += 4) {
            const float4 v = __ldcg(reinterpret_cast<const float4*>(src + q));
            av[q] = v.x; av[q + 1] = v.y; av[q + 2] = v.z; av[q + 3] = v.w;
        }
    };
    // pass 1 of one adapter: n^2 and the prompt scores of the row (complete on the qt == 0 threads)
    auto row_pass1 = [&](int ad, const float (&av)[32], float& n2, float (&sc)[HR_CT]) {
        const uint32_t d = tmem_base + (ad == 0 ? TM_D1 : TM_D1B) + tlane;
        const float4* bn = sBN + (size_t)ad * H;
        const float* gb = sGb + (size_t)ad * HR_N;
        float dot = 0.f;
        {
            uint32_t r[32];
            ptx::tmem_ld_32x32b_x32(d + j0, r);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int q = 0; q < 32; ++q) {
                const float4 c4 = bn[j0 + q];
                const float h = fmaxf(fmaf(fmaf(av[q], c4.x, c4.y), c4.z, c4.w), 0.f);
                dot = fmaf(h, __uint_as_float(r[q]) + gb[j0 + q], dot);
            }
        }
        if (qt != 0) sX[(ad * 3 + qt - 1) * 64 + xrow] = dot;
        ep_sync();
        n2 = 1.f;
#pragma unroll
        for (int c = 0; c < HR_CT; ++c) sc[c] = 0.f;
        if (qt == 0) {
            uint32_t r[32];
            ptx::tmem_ld_32x32b_x32(d + 128, r);
            ptx::tmem_ld_wait();
            n2 = ((dot + sX[(ad * 3 + 0) * 64 + xrow]) + (sX[(ad * 3 + 1) * 64 + xrow] + sX[(ad * 3 + 2) * 64 + xrow])) + (__uint_as_float(r[0]) + gb[H]);
#pragma unroll
            for (int c = 0; c < HR_CT; ++c) sc[c] = c < C ? __uint_as_float(r[1 + c]) + gb[H + 1 + c] : 0.f;
        }
    };

    float lo_logit[HR_CT] = {0.f, 0.f, 0.f, 0.f};
    build_a(0);
    HR_TICK(0, 3);
    ptx::fence_proxy_async_smem();
    ptx::tc_fence_before_sync();
    __syncthreads();
    ptx::tc_fence_after_sync();
    if (tid == 64) issue_mma1(0);
    if (NAD == 2) {
        ptx::mbar_wait(&bars[0], 0);               // operand tiles are free again
        ptx::tc_fence_after_sync();
        build_b(1);
        build_a(1);
        ptx::fence_proxy_async_smem();
        ptx::tc_fence_before_sync();
        __syncthreads();
        ptx::tc_fence_after_sync();
        if (tid == 64) issue_mma1(1);
        if (ep) {                                   // frozen adapter's share of the logits, while MMA1 of the trainable one runs
            float av0[32], n2, sc[HR_CT];
            load_a_quarter(0, av0);
            row_pass1(0, av0, n2, sc);
            const float inv_n = 1.0f / sqrtf(n2);
#pragma unroll
            for (int c = 0; c < HR_CT; ++c) lo_logit[c] = a.w_old * a.inv_tau * sc[c] * inv_n;
        }
    }
    const int ADT = NAD - 1;
    float av[32];
    if (ep) load_a_quarter(ADT, av);                // in flight while MMA1 finishes
    ptx::mbar_wait(&bars[ADT], 0);
    ptx::tc_fence_after_sync();
    HR_TICK(0, 4);

    // MMA2 operand tiles (the MMA1 tiles are dead now): A'[j][k] = h, B'[n][k] = L, K = batch row k of this CTA
    uint8_t* sA2 = sA;                               // [hi | lo][2 k-tiles][128 rows x 128 B]
    uint8_t* sB2 = sB;                               // [hi | lo][2 k-tiles][144 rows x 128 B]

    if (ep) {
        const float coef = NAD == 2 ? (1.0f - a.w_old) : 1.0f;
        float n2, sc[HR_CT];
        row_pass1(ADT, av, n2, sc);
        float cc = 0.f, dsv[HR_CT] = {0.f, 0.f, 0.f, 0.f};
        if (qt == 0) {
            float nll = 0.f; int corr = 0;
            if (valid) {
                const float inv_n = 1.0f / sqrtf(n2);
                float lnew[HR_CT], l[HR_CT];
                float mx = -INFINITY; int am = 0;
#pragma unroll
                for (int c = 0; c < HR_CT; ++c) {
                    lnew[c] = 0.f; l[c] = -INFINITY;
                    if (c < C) {
                        lnew[c] = a.inv_tau * sc[c] * inv_n;
                        l[c] = NAD == 2 ? fmaf(coef, lnew[c], lo_logit[c]) : lnew[c];
                        if (l[c] > mx) { mx = l[c]; am = c; }
                    }
                }
                float se = 0.f, ly = 0.f, p[HR_CT];
#pragma unroll
                for (int c = 0; c < HR_CT; ++c) {
                    p[c] = 0.f;
                    if (c < C) { p[c] = expf(l[c] - mx); se += p[c]; if (c == yv) ly = l[c]; }
                }
                nll = logf(se) + mx - ly;
                corr = am == yv ? 1 : 0;
                const float inv_se = 1.0f / se;
                float dot = 0.f;
#pragma unroll
                for (int c = 0; c < HR_CT; ++c) {
                    if (c < C) {
                        if (a.logits_out) a.logits_out[(size_t)my_r * C + c] = l[c];
                        const float dl = a.dlogits_in ? __ldg(a.dlogits_in + (size_t)my_r * C + c)
                                                      : (p[c] * inv_se - (c == yv ? 1.f : 0.f)) * a.inv_B;
                        dot = fmaf(dl, lnew[c], dot);
                        dsv[c] = coef * dl * a.inv_tau * inv_n;
                    }
                }
                cc = -coef * dot / n2;
            }
            sXc[xrow * 5 + 0] = cc;
#pragma unroll
            for (int c = 0; c < HR_CT; ++c) sXc[xrow * 5 + 1 + c] = dsv[c];
            // loss / group counters of this warp's 32 rows (update_dict, final_main.py:383-391)
            const double tot = warp_sum((double)nll);
            if (lane == 0 && a.loss_sum) atomicAdd(&a.loss_sum[a.slot], tot);
            const unsigned cmask = __ballot_sync(0xffffffffu, corr != 0);
            for (int g = 0; g < a.G; ++g) {
                const unsigned gm = __ballot_sync(0xffffffffu, valid && gval == g);
                if (lane == 0 && gm && a.counts) {
                    int64_t* cnt = a.counts + (size_t)a.slot * 2 * a.G;
                    const int nc = __popc(gm & cmask);
                    if (nc) atomicAdd((unsigned long long*)&cnt[g], (unsigned long long)nc);
                    atomicAdd((unsigned long long*)&cnt[a.G + g], (unsigned long long)__popc(gm));
                }
            }
        }
        HR_TICK(0, 5);
        ep_sync();
        if (qt != 0) {
            cc = sXc[xrow * 5 + 0];
#pragma unroll
            for (int c = 0; c < HR_CT; ++c) dsv[c] = sXc[xrow * 5 + 1 + c];
        }
        // pass 2: dh, BatchNorm-backward inputs, and the transposed operands of S^T_part = h^T L (this thread: 32 hidden units)
        const uint32_t d = tmem_base + (ADT == 0 ? TM_D1 : TM_D1B) + tlane;
        const float4* bn = sBN + (size_t)ADT * H;
        const float* gb = sGb + (size_t)ADT * HR_N;
        const int k = xrow;                                                // K index of this row in the MMA2 tiles
        const uint32_t kofsA = (uint32_t)(k >> 5) * HR_A2_KT_BYTES, kofsB = (uint32_t)(k >> 5) * HR_B_KT_BYTES;
        const uint32_t kc = (uint32_t)(k & 31) >> 2, kw = ((uint32_t)k & 3u) * 4u;
        float dgv[32], dbv[32], clv[32];
        float* scr = (float*)(sB + 4 * HR_B_KT_BYTES) + (qt * 2 + rw) * (32 * 36);          // per-warp transpose scratch (above B')
        {
            uint32_t r[32];
            ptx::tmem_ld_32x32b_x32(d + j0, r);
            ptx::tmem_ld_wait();
            float dav[32];
#pragma unroll
            for (int q4 = 0; q4 < 32; q4 += 4) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int q = q4 + u, j = j0 + q;
                    const float4 c4 = bn[j];
                    const float4 m4 = sM[j];
                    const float ah = fmaf(av[q], c4.x, c4.y);
                    const float pre = fmaf(ah, c4.z, c4.w);
                    const float h = valid ? fmaxf(pre, 0.f) : 0.f;
                    float dh = cc * (__uint_as_float(r[q]) + gb[j]);
                    dh = fmaf(dsv[0], m4.x, dh); dh = fmaf(dsv[1], m4.y, dh); dh = fmaf(dsv[2], m4.z, dh); dh = fmaf(dsv[3], m4.w, dh);
                    const float dpre = (valid && pre > 0.f) ? dh : 0.f;
                    dav[q] = dpre * c4.z;
                    const float lj = cc * h;
                    dgv[q] = dpre * ah; dbv[q] = dpre; clv[q] = lj;
                    float hh, hl, lh, ll;
                    hr_split(h, hh, hl);
                    hr_split(lj, lh, ll);
                    const uint32_t off = (uint32_t)j * 128u + ((kc ^ ((uint32_t)j & 7u)) << 4) + kw;
                    *reinterpret_cast<float*>(sA2 + kofsA + off) = hh;
                    *reinterpret_cast<float*>(sA2 + 2 * HR_A2_KT_BYTES + kofsA + off) = hl;
                    *reinterpret_cast<float*>(sB2 + kofsB + off) = lh;
                    *reinterpret_cast<float*>(sB2 + 2 * HR_B_KT_BYTES + kofsB + off) = ll;
                }
            }
            warp_store_block32(dav, scr, a.dahat + (size_t)(m0 + rw * 32) * H + j0, H, min(32, max(0, a.B - (m0 + rw * 32))), lane);
        }
        // L columns H (c), H+1+c (ds), zero padding up to HR_N, for this row k
        if (qt == 0) {
            float hh, hl;
#pragma unroll 1
            for (int n = H; n < HR_N; ++n) {
                float v = 0.f;
                if (n == H) v = cc;
                else if (n < ldg) {
#pragma unroll
                    for (int c = 0; c < HR_CT; ++c) if (n - H - 1 == c) v = dsv[c];
                }
                hr_split(v, hh, hl);
                const uint32_t offB = kofsB + (uint32_t)n * 128u + ((kc ^ ((uint32_t)n & 7u)) << 4) + kw;
                *reinterpret_cast<float*>(sB2 + offB) = hh;
                *reinterpret_cast<float*>(sB2 + 2 * HR_B_KT_BYTES + offB) = hl;
            }
        }
        HR_TICK(0, 6);
        ptx::fence_proxy_async_smem();
        ptx::mbar_arrive(&bars[2]);
        // column sums over this warp's 32 rows (lane = column afterwards): dgamma / dbeta as one fixed-point atomic per element
        // and warp; the column sums of L are row H of the S^T tile
        {
            const float sdg = warp_colsum32(dgv, lane), sdb = warp_colsum32(dbv, lane), scl = warp_colsum32(clv, lane);
            fx_add<FX_DGB>(&a.dgb[j0 + lane], (double)sdg);
            fx_add<FX_DGB>(&a.dgb[H + j0 + lane], (double)sdb);
            sRow[rw * HR_N + j0 + lane] = scl;
        }
        if (qt == 0) {
            const float s_c = warp_sum(cc);
            float s_ds[HR_CT];
#pragma unroll
            for (int c = 0; c < HR_CT; ++c) s_ds[c] = warp_sum(dsv[c]);
            if (lane < HR_N - H) {
                float v = 0.f;
                if (lane == 0) v = s_c;
#pragma unroll
                for (int c = 0; c < HR_CT; ++c) if (lane == 1 + c && c < C) v = s_ds[c];
                sRow[rw * HR_N + H + lane] = v;
            }
        }
    }
    // ---- MMA2: S^T_part[j][n] = sum_k h[k][j] L[k][n]
    if (tid == 64) {
        ptx::mbar_wait(&bars[2], 0);
        ptx::tc_fence_after_sync();
        const uint32_t d2 = tmem_base + TM_D2;
        const uint32_t bA = ptx::smem_u32(sA2), bB = ptx::smem_u32(sB2);
#pragma unroll
        for (int kt = 0; kt < 2; ++kt)
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const uint64_t ahi = ptx::umma_smem_desc(bA + kt * HR_A2_KT_BYTES + kk * 32, 0, 1024);
                const uint64_t alo = ptx::umma_smem_desc(bA + 2 * HR_A2_KT_BYTES + kt * HR_A2_KT_BYTES + kk * 32, 0, 1024);
                const uint64_t bhi = ptx::umma_smem_desc(bB + kt * HR_B_KT_BYTES + kk * 32, 0, 1024);
                const uint64_t blo = ptx::umma_smem_desc(bB + 2 * HR_B_KT_BYTES + kt * HR_B_KT_BYTES + kk * 32, 0, 1024);
                ptx::mma_tf32_ss(d2, alo, bhi, idesc, (kt | kk) != 0 ? 1u : 0u);
                ptx::mma_tf32_ss(d2, ahi, blo, idesc, 1u);
                ptx::mma_tf32_ss(d2, ahi, bhi, idesc, 1u);
            }
        ptx::mma_commit(&bars[3]);
    }
    HR_TICK(0, 7);
    __syncthreads();                                  // sRow complete; everybody past the MMA1 accumulators
    HR_TICK(0, 8);
    float* tile = a.Spart + (size_t)blockIdx.x * (H + 1) * HR_SP_LD;
    if (tid < HR_N) tile[(size_t)H * HR_SP_LD + tid] = sRow[tid] + sRow[HR_N + tid];      // row H: the two row warps, in warp order
    {   // rows j < H from TMEM: warp w owns lanes 32 * (w % 4) .. +31 and the column chunk w / 4 (chunk 4, 16 wide: warps 0..3 again)
        ptx::mbar_wait(&bars[3], 0);
        ptx::tc_fence_after_sync();
        const int j = (warp & 3) * 32 + lane;
        float* out = tile + (size_t)j * HR_SP_LD;
        const uint32_t d2 = tmem_base + TM_D2 + (((uint32_t)((warp & 3) * 32)) << 16);
        float* scr2 = (float*)(sB + 4 * HR_B_KT_BYTES) + warp * (32 * 36);                 // (MMA2 has completed: everything past B' is free)
#pragma unroll 1
        for (int ch = warp >> 2; ch < 5; ch += 4) {
            uint32_t r[32];
            ptx::tmem_ld_32x32b_x32(d2 + ch * 32, r);
            ptx::tmem_ld_wait();
            float v[32];
#pragma unroll
            for (int q = 0; q < 32; ++q) v[q] = __uint_as_float(r[q]);
            if (ch < 4) warp_store_block32(v, scr2, tile + (size_t)((warp & 3) * 32) * HR_SP_LD + ch * 32, HR_SP_LD, 32, lane);
            else {
#pragma unroll
                for (int q = 0; q < HR_N - 128; q += 4) *reinterpret_cast<float4*>(out + 128 + q) = make_float4(v[q], v[q + 1], v[q + 2], v[q + 3]);
            }
        }
        ptx::tc_fence_before_sync();
    }
    HR_TICK(0, 9);
    __syncthreads();
    HR_TICK(0, 10);
    if (warp == 2) ptx::tmem_dealloc<512>(tmem_base);
}

__global__ void __launch_bounds__(HR_THREADS, 1) k_hs_rows(HsRowsArgs a) { hs_rows_body(a); }

// S[l][j] = sum over the row tiles of S^T_part[tile][j][l]   (l < H+1+C rows of L, j < H+1 columns [h | 1]; row stride s_stride(H),
// padding columns zero): tiles added in tile order by ONE thread per element -- deterministic.
struct SumSpartArgs { const float* Spart; int tiles, H, C; float* S; float* ST; };   // ST (optional): the same sums as [j][l], row stride HR_SP_LD
__device__ __forceinline__ void sum_spart_body(const SumSpartArgs& a) {
    const int H = a.H, ldl = H + 1 + a.C, SP = s_stride(H);
    ptx::pdl_wait();
    ptx::pdl_launch();
    const int e = blockIdx.x * 256 + threadIdx.x;                // e = j * ldl + l: reads coalesced along l
    if (e >= (H + 1) * ldl) {
        const int p = e - (H + 1) * ldl;                         // the padding columns of S
        const int npad = SP - (H + 1);
        if (p < ldl * npad) a.S[(size_t)(p / npad) * SP + (H + 1) + p % npad] = 0.f;
        return;
    }
    const int j = e / ldl, l = e - j * ldl;
    float v = 0.f;
    const float* src = a.Spart + (size_t)j * HR_SP_LD + l;
    const size_t ts = (size_t)(H + 1) * HR_SP_LD;
    int t = 0;
    for (; t + 4 <= a.tiles; t += 4) {
        const float p0 = __ldcg(src + (size_t)t * ts), p1 = __ldcg(src + (size_t)(t + 1) * ts);
        const float p2 = __ldcg(src + (size_t)(t + 2) * ts), p3 = __ldcg(src + (size_t)(t + 3) * ts);
        v += p0; v += p1; v += p2; v += p3;
    }
    for (; t < a.tiles; ++t) v += __ldcg(src + (size_t)t * ts);
    a.S[(size_t)l * SP + j] = v;
    if (a.ST) a.ST[(size_t)j * HR_SP_LD + l] = v;
}
__global__ void __launch_bounds__(256) k_sum_spart(SumSpartArgs a) { sum_spart_body(a); }

// S^T [j][l] (row stride HR_SP_LD, padding columns zero) = sum over the CUDA-core row kernel's per-CTA shares (rows_train.cuh:
// up to 296 shares of 8 batch rows), added in CTA order: a thread owns one element and one of four contiguous GROUPS of shares
// (all loads of a pass in flight together, added in order), the four group sums are combined in group order -- a fixed order,
// so the result does not depend on timing.
constexpr int SG_ELEMS = 64, SG_GROUPS = 4, SG_MAXT = 296, SG_PER = (SG_MAXT + SG_GROUPS - 1) / SG_GROUPS;
// Data parallel (P2P): the element sum of this rank goes to every other rank as ONE LL word (value + instance tag in an 8-byte
// store, p2p.cuh) and the ranks' values are added in rank order, so every rank holds the same bits; this replaces the separate
// S^T exchange kernel (k_p2p_sum_st: slices + system fence + flags, 12 us of the 2-GPU step, profiles/r3_step_timeline.md).
// A thread pushes before it polls and a push depends on nothing, so a CTA only ever waits for peer CTAs that will be scheduled
// without its help: the 291 CTAs of 256 threads are far below the resident capacity of the GPU (148 SMs x 8), i.e. the pollers
// can never hold every slot of a rank; the wait is bounded by wall clock like every other peer-memory wait (p2p_expired).
struct SumSpartGArgs { const float* Spart; int tiles, H, C; float* ST; P2pArgs p2p; };
template <bool P2P>
__global__ void __launch_bounds__(SG_ELEMS * SG_GROUPS) k_sum_spart_g(SumSpartGArgs a) {
    __shared__ float sg[SG_GROUPS][SG_ELEMS];
    const int H = a.H, ldl = H + 1 + a.C;
    const int el = threadIdx.x & (SG_ELEMS - 1), grp = threadIdx.x / SG_ELEMS;
    const int e = blockIdx.x * SG_ELEMS + el;                    // e = j * HR_SP_LD + l over the PADDED row: padding written as zeros
    const int j = e / HR_SP_LD, l = e - j * HR_SP_LD;
    const bool live = j <= H && l < ldl;
    DBMM_TL_SCOPE(TL_TN);           // (timeline builds: the slot of the TN GEMM this kernel replaced)
    ptx::pdl_wait();
    DBMM_TL_WAIT(TL_TN);
    ptx::pdl_launch();
    const int per = (a.tiles + SG_GROUPS - 1) / SG_GROUPS, t0 = grp * per, t1 = min(a.tiles, t0 + per);
    const size_t ts = (size_t)(H + 1) * HR_SP_LD;
    const float* src = a.Spart + (size_t)j * HR_SP_LD + l;
    float v = 0.f;
    if (live) {
        for (int tb = t0; tb < t1; tb += 16) {
            float p[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) p[u] = tb + u < t1 ? __ldcg(src + (size_t)(tb + u) * ts) : 0.f;
#pragma unroll
            for (int u = 0; u < 16; ++u) v += p[u];
        }
    }
    sg[grp][el] = v;
    __syncthreads();
    if (grp == 0 && j <= H) {
        float tot = live ? ((sg[0][el] + sg[1][el]) + sg[2][el]) + sg[3][el] : 0.f;
        if constexpr (P2P) {
            if (live) {
                const unsigned inst = p2p_instance(a.p2p);
                const int parity = inst & 1u;
                char* me = a.p2p.peer[a.p2p.rank];
                for (int r = 0; r < a.p2p.world; ++r)
                    if (r != a.p2p.rank) p2p_f_store(p2p_st_ll(a.p2p.peer[r], parity, a.p2p.rank), e, tot, inst + 1u);
                float sum = 0.f;
                for (int r = 0; r < a.p2p.world; ++r)
                    sum += r == a.p2p.rank ? tot : p2p_f_load(a.p2p, p2p_st_ll(me, parity, r), e, inst + 1u);
                tot = sum;
            }
        }
        a.ST[(size_t)j * HR_SP_LD + l] = tot;
    }
}
static int launch_sum_spart_g(const float* Spart, int tiles, int H, int C, float* ST, cudaStream_t st, const P2pArgs* p2p = nullptr) {
    DBMM_CHECK_ARG(tiles >= 1 && tiles <= SG_MAXT, "k_sum_spart_g: %d shares", tiles);
    SumSpartGArgs s;
    memset(&s, 0, sizeof(s));
    s.Spart = Spart; s.tiles = tiles; s.H = H; s.C = C; s.ST = ST;
    const dim3 grid(ceil_div((H + 1) * HR_SP_LD, SG_ELEMS)), block(SG_ELEMS * SG_GROUPS);
    if (p2p && p2p->world > 1) {
        DBMM_CHECK_ARG((size_t)(H + 1) * HR_SP_LD <= P2P_ST_FLOATS, "S^T exceeds the peer-memory slot");
        s.p2p = *p2p;
        DBMM_CUDA(launch_pdl(k_sum_spart_g<true>, grid, block, 0, st, s));
    } else DBMM_CUDA(launch_pdl(k_sum_spart_g<false>, grid, block, 0, st, s));
    return DBMM_OK;
}

static int launch_hs_rows(const HsRowsArgs& a, cudaStream_t st) {
    DBMM_CHECK_SHAPE(hs_rows_supported(a.H, a.C), "tensor-core row kernel needs H == 128 and C <= 4 (H=%d C=%d)", a.H, a.C);
    DBMM_CUDA(set_smem(k_hs_rows, HR_SMEM));
    DBMM_CUDA(launch_pdl(k_hs_rows, dim3(ceil_div(a.B, HR_ROWS)), dim3(HR_THREADS), HR_SMEM, st, a));
    return DBMM_OK;
}
static int launch_sum_spart(const float* Spart, int B, int H, int C, float* S, float* ST, cudaStream_t st, bool pdl) {
    SumSpartArgs s; s.Spart = Spart; s.tiles = ceil_div(B, HR_ROWS); s.H = H; s.C = C; s.S = S; s.ST = ST;
    const int n = (H + 1) * (H + 1 + C) + (H + 1 + C) * (s_stride(H) - (H + 1));
    if (pdl) DBMM_CUDA(launch_pdl(k_sum_spart, dim3(ceil_div(n, 256)), dim3(256), 0, st, s));
    else { k_sum_spart<<<ceil_div(n, 256), 256, 0, st>>>(s); DBMM_LAUNCH_CHECK(); }
    return DBMM_OK;
}

}  // namespace dbmm
