// Training row kernel (H-space, see kernels_simt.cuh header):  BatchNorm -> ReLU -> t = [h,1] G -> logits -> CE ->
// row backward (dL/d ahat) -> per-batch reductions (dgamma, dbeta) and the operand rows of S = [c*h | c | ds]^T [h | 1] in ONE launch.
//
// A CTA owns RT_ROWS = 8 batch rows and 16 warps.  The (H+1)-long contraction t = [h,1] G is split over the WARPS
// (9 Gram rows each, all 8 batch rows at once, 40 accumulators per lane) and combined through shared memory, so the
// dependent-FMA chain per warp is 9 long instead of 129.  Afterwards warp w < 8 finishes row w (softmax-CE, argmax,
// group counters, dh, BatchNorm-backward inputs), and all warps store the CTA's 8 operand rows of S.
#pragma once
#include "kernels_simt.cuh"
#include "ptx_sm100.cuh"
#include "tn_gemm.cuh"

namespace dbmm {

constexpr int RT_ROWS = 8;      // batch rows per CTA; NW warps per CTA (16 for one adapter, 8 when two Gram matrices fill the smem)
constexpr int RT_SP_LD = 144;   // row stride of an S^T share (== HR_SP_LD of hs_rows.cuh: the operand layout of k_hs_w2)
constexpr int RT_MAX_FUSED_CTAS = 148 * 2;

struct RowsTrainArgs {
    int B; int64_t Bg;
    const int32_t* idx; const int32_t* y; const int32_t* grp;
    int H, C, G;
    const float* A; int64_t strideA;      // [nad][B][H]
    const float* gram;                    // [nad][H+1][H+1+C]
    const fx64* colsum;                   // [nad][2][H] (fixed point, FX_COLSUM)
    AdapterView ad[2];
    float w_old, inv_tau, inv_B;
    double* loss_sum; int64_t* counts; int64_t slot;
    float* dahat; fx64* dgb;              // outputs: [B][H], [2][H] (+=, FX_DGB)
    float* logits_out;                    // optional [B][C]: the batch's logits (train-mode nn.Module forward, modules.py)
    const float* dlogits_in;              // optional [B][C]: upstream dL/dlogits replaces the fused CE gradient (autograd backward)
    float* Lrows; float* Hrows;           // outputs: [B][l_stride] rows [c*h | c | ds] and [B][s_stride] rows [h | 1]: the operands
                                          // of S = L^T [h | 1] (k_tn_gemm on the second graph branch)
    float* Spart;                         // != nullptr (one pass over the batch: gridDim.x * RT_ROWS >= B): instead of those rows the CTA
                                          // writes its share S^T_part[cta][i][j] = sum over its 8 rows of [h | 1]_i [c*h | c | ds]_j
                                          // ([H+1][RT_SP_LD], one mma.sync k-step per 16 x 8 tile, 3xTF32); k_sum_spart_g adds the
                                          // CTAs' shares in CTA order.  The TN GEMM this replaces was the longest kernel of the W2
                                          // branch (14.5 us of a 43.7 us step, profiles/r3_step_timeline.md)
};

static inline size_t rows_train_smem_bytes(int H, int C, int nad, int CT, int RT_WARPS) {
    const size_t ldg = H + 1 + C;
    size_t fl = rows_gram_floats(H, C, nad) + (size_t)nad * 4 * H + (size_t)RT_ROWS * (H + 1)      // sG, sBN, sH
              + (size_t)RT_WARPS * RT_ROWS * RK_NSLOT * 32                                          // sT
              + (size_t)RT_ROWS * ldg + 3 * 32 + (size_t)RT_ROWS * CT + 2 * (size_t)H;              // sL, row stats, sLo, sDgb
    return fl * 4 + 16;
}

template <int NAD, int CT, int NW>
__device__ __forceinline__ void rows_train_body(const RowsTrainArgs& a) {
    constexpr int RT_WARPS = NW, RT_THREADS = NW * 32;
    extern __shared__ __align__(16) float dyn_smem[];
    const int H = a.H, C = a.C, ldg = H + 1 + C, HP = H + 1;
    const int HS = (H + 31) >> 5, NS = (ldg + 31) >> 5;
    float* sG = dyn_smem;
    float* sBN = sG + rows_gram_floats_dev(H, C, NAD);
    float* sH = sBN + (size_t)NAD * 4 * H;                               // [RT_ROWS][H+1]
    float* sT = sH + (size_t)RT_ROWS * HP;                               // [RT_WARPS][RT_ROWS][RK_NSLOT*32]
    float* sL = sT + (size_t)RT_WARPS * RT_ROWS * RK_NSLOT * 32;         // [RT_ROWS][ldg]
    float* sRowNll = sL + (size_t)RT_ROWS * ldg;
    int* sRowG = reinterpret_cast<int*>(sRowNll + 32);
    int* sRowCorr = sRowG + 32;
    float* sLo = reinterpret_cast<float*>(sRowCorr + 32);                // [RT_ROWS][CT]
    constexpr int TSTR = RK_NSLOT * 32;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int ADT = NAD - 1;

    {
        const int n = NAD * HP * ldg, n4 = n >> 2;
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(sG);
        for (int e = tid; e < n4; e += RT_THREADS)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + e * 16), "l"(a.gram + e * 4) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
        for (int e = (n4 << 2) + tid; e < n; e += RT_THREADS) sG[e] = a.gram[e];
    }
    // labels of the first row group (constants) are requested before the dependency wait as well
    int y_pre = -1, g_pre = -1;
    if (warp < RT_ROWS && blockIdx.x * RT_ROWS + warp < a.B) {
        const int r = blockIdx.x * RT_ROWS + warp;
        const int64_t dsrow = a.idx ? (int64_t)a.idx[r] : (int64_t)r;
        y_pre = a.y ? a.y[dsrow] : -1;
        g_pre = a.grp ? a.grp[dsrow] : 0;
    }
    ptx::pdl_wait();                // A / column sums come from k_reduce_stats; the Gram matrix is two kernels upstream
    DBMM_TL_WAIT(TL_ROWS);
    ptx::pdl_launch();
    // first row group's activations (adapter 0): in flight while the Gram matrix and the statistics arrive
    float av_pre[RK_HSLOT];
    {
        const int r = blockIdx.x * RT_ROWS + warp;
#pragma unroll
        for (int s = 0; s < RK_HSLOT; ++s) {
            const int j = lane + 32 * s;
            av_pre[s] = (warp < RT_ROWS && s < HS && j < H && r < a.B) ? __ldcg(a.A + (size_t)r * H + j) : 0.f;
        }
    }
    for (int e = tid; e < NAD * H; e += RT_THREADS) {
        const int ad = e / H, j = e - ad * H;
        const double s1 = fx_get<FX_COLSUM>(&a.colsum[((size_t)ad * 2 + 0) * H + j]), s2 = fx_get<FX_COLSUM>(&a.colsum[((size_t)ad * 2 + 1) * H + j]);
        const double m = s1 / (double)a.Bg;
        double v = s2 / (double)a.Bg - m * m;
        if (v < 0.0) v = 0.0;
        float* bn = sBN + (size_t)ad * 4 * H;
        bn[j] = (float)m; bn[H + j] = 1.0f / sqrtf((float)v + DBMM_BN_EPS);
        bn[2 * H + j] = a.ad[ad].gamma[j]; bn[3 * H + j] = a.ad[ad].beta[j];
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();

    float dg_acc[RK_HSLOT], db_acc[RK_HSLOT];
#pragma unroll
    for (int s = 0; s < RK_HSLOT; ++s) { dg_acc[s] = 0.f; db_acc[s] = 0.f; }
    const float coef = (NAD == 2) ? (1.0f - a.w_old) : 1.0f;
    const int IR = (HP + RT_WARPS - 1) / RT_WARPS;             // Gram rows per warp
    const int MR = (ldg + RT_WARPS - 1) / RT_WARPS;            // S rows per warp

    for (int base = blockIdx.x * RT_ROWS; base < a.B; base += gridDim.x * RT_ROWS) {
        const int r = base + warp;                             // this warp's batch row (warps >= RT_ROWS own none)
        const bool owner = warp < RT_ROWS;
        const bool valid = owner && r < a.B;
        int yv = -1, gval = -1;                                // labels of this row: requested before the long phases
        if (valid) {
            if (base == (int)blockIdx.x * RT_ROWS) { yv = y_pre; gval = g_pre; }
            else {
                const int64_t dsrow = a.idx ? (int64_t)a.idx[r] : (int64_t)r;
                yv = a.y ? a.y[dsrow] : -1;
                gval = a.grp ? a.grp[dsrow] : 0;
            }
        }
        float t[RK_NSLOT], hv[RK_HSLOT], ahat[RK_HSLOT];
        unsigned prepos = 0u;
        float n2 = 1.f, sc[CT];

#pragma unroll
        for (int ad = 0; ad < NAD; ++ad) {
            const float* bn = sBN + (size_t)ad * 4 * H;
            const float* G = sG + (size_t)ad * HP * ldg;
            // ---- phase 1: h = relu(BN(a)) for row `warp`
            prepos = 0u;
            float* hrow = sH + (size_t)warp * HP;
#pragma unroll
            for (int s = 0; s < RK_HSLOT; ++s) {
                const int j = lane + 32 * s;
                float h = 0.f, ah = 0.f;
                if (s < HS && j < H && valid) {
                    const float av = (ad == 0 && base == (int)blockIdx.x * RT_ROWS)
                                         ? av_pre[s] : __ldcg(a.A + (size_t)ad * a.strideA + (size_t)r * H + j);
                    ah = (av - bn[j]) * bn[H + j];
                    const float pre = fmaf(ah, bn[2 * H + j], bn[3 * H + j]);
                    if (pre > 0.f) { h = pre; prepos |= (1u << s); }
                }
                hv[s] = h; ahat[s] = ah;
                if (owner && s < HS && j < H) hrow[j] = h;
            }
            if (owner && lane == 0) hrow[H] = valid ? 1.0f : 0.f;
            __syncthreads();
            // ---- phase 2: partial t over this warp's Gram rows, all RT_ROWS batch rows
            {
                float acc[RT_ROWS][RK_NSLOT];
#pragma unroll
                for (int rr = 0; rr < RT_ROWS; ++rr)
#pragma unroll
                    for (int s = 0; s < RK_NSLOT; ++s) acc[rr][s] = 0.f;
                const int i0 = warp * IR, i1 = min(HP, i0 + IR);
#pragma unroll 2
                for (int i = i0; i < i1; ++i) {
                    float g[RK_NSLOT];
#pragma unroll
                    for (int s = 0; s < RK_NSLOT; ++s) {
                        const int j = lane + 32 * s;
                        g[s] = (s < NS && j < ldg) ? G[(size_t)i * ldg + j] : 0.f;
                    }
#pragma unroll
                    for (int rr = 0; rr < RT_ROWS; ++rr) {
                        const float hh = sH[(size_t)rr * HP + i];
#pragma unroll
                        for (int s = 0; s < RK_NSLOT; ++s) acc[rr][s] = fmaf(hh, g[s], acc[rr][s]);
                    }
                }
#pragma unroll
                for (int rr = 0; rr < RT_ROWS; ++rr)
#pragma unroll
                    for (int s = 0; s < RK_NSLOT; ++s) sT[((size_t)warp * RT_ROWS + rr) * TSTR + lane + 32 * s] = acc[rr][s];
            }
            __syncthreads();
#pragma unroll
            for (int s = 0; s < RK_NSLOT; ++s) {
                float v = 0.f;
                if (owner) {
#pragma unroll
                    for (int w2 = 0; w2 < RT_WARPS; ++w2) v += sT[((size_t)w2 * RT_ROWS + warp) * TSTR + lane + 32 * s];
                }
                t[s] = v;
            }
            // ---- n^2 and the C prompt scores of this adapter (row `warp`)
            {
                float part = 0.f;
#pragma unroll
                for (int s = 0; s < RK_HSLOT; ++s) part = fmaf(t[s], hv[s], part);
#pragma unroll
                for (int s = 0; s < RK_NSLOT; ++s)
                    if (lane + 32 * s == H) part += t[s];
                n2 = warp_sum(part);
#pragma unroll
                for (int c = 0; c < CT; ++c) {
                    float v = 0.f;
                    if (c < C) {
                        const int col = H + 1 + c;
#pragma unroll
                        for (int s = 0; s < RK_NSLOT; ++s) {
                            const float tmp = __shfl_sync(0xffffffffu, t[s], col & 31);
                            if (s == (col >> 5)) v = tmp;
                        }
                    }
                    sc[c] = v;
                }
                if (NAD == 2 && ad == 0) {
                    const float inv_n = 1.0f / sqrtf(n2);
                    if (owner && lane < CT) {
                        float v = 0.f;
#pragma unroll
                        for (int c = 0; c < CT; ++c) if (c == lane) v = sc[c];
                        sLo[(size_t)warp * CT + lane] = a.w_old * a.inv_tau * v * inv_n;
                    }
                    __syncthreads();          // sH / sT are rewritten by the next adapter
                }
            }
        }

        // ---- phase 3: row epilogue (trainable adapter)
        float nll = 0.f; int corr = 0;
        float cc = 0.f, dsv[CT];
#pragma unroll
        for (int c = 0; c < CT; ++c) dsv[c] = 0.f;
        if (valid) {
            const float inv_n = 1.0f / sqrtf(n2);
            float lnew[CT], l[CT];
            float mx = -INFINITY; int am = 0;
#pragma unroll
            for (int c = 0; c < CT; ++c) {
                lnew[c] = 0.f; l[c] = -INFINITY;
                if (c < C) {
                    lnew[c] = a.inv_tau * sc[c] * inv_n;
                    l[c] = (NAD == 2) ? fmaf(coef, lnew[c], sLo[(size_t)warp * CT + c]) : lnew[c];
                    if (l[c] > mx) { mx = l[c]; am = c; }
                }
            }
            float se = 0.f, ly = 0.f, p[CT];
#pragma unroll
            for (int c = 0; c < CT; ++c) {
                p[c] = 0.f;
                if (c < C) { p[c] = expf(l[c] - mx); se += p[c]; if (c == yv) ly = l[c]; }
            }
            nll = logf(se) + mx - ly;
            corr = (am == yv) ? 1 : 0;
            const float inv_se = 1.0f / se;
            float dot = 0.f;
            if (a.logits_out && lane < C) {
                float v = 0.f;
#pragma unroll
                for (int c = 0; c < CT; ++c) if (c == lane) v = l[c];
                a.logits_out[(size_t)r * C + lane] = v;
            }
#pragma unroll
            for (int c = 0; c < CT; ++c) {
                if (c < C) {
                    const float dl = a.dlogits_in ? __ldg(a.dlogits_in + (size_t)r * C + c)
                                                  : (p[c] * inv_se - (c == yv ? 1.f : 0.f)) * a.inv_B;
                    dot = fmaf(dl, lnew[c], dot);
                    dsv[c] = coef * dl * a.inv_tau * inv_n;
                }
            }
            cc = -coef * dot / n2;
            const float* G = sG + (size_t)ADT * HP * ldg;
            const float* bn = sBN + (size_t)ADT * 4 * H;
#pragma unroll
            for (int s = 0; s < RK_HSLOT; ++s) {
                const int j = lane + 32 * s;
                if (s < HS && j < H) {
                    float dh = cc * t[s];
#pragma unroll
                    for (int c = 0; c < CT; ++c)
                        if (c < C) dh = fmaf(dsv[c], G[(size_t)j * ldg + H + 1 + c], dh);
                    const float dpre = ((prepos >> s) & 1u) ? dh : 0.f;
                    dg_acc[s] = fmaf(dpre, ahat[s], dg_acc[s]);
                    db_acc[s] += dpre;
                    a.dahat[(size_t)r * H + j] = dpre * bn[2 * H + j];
                }
            }
        }
        // L row = [c*h | c | ds] for the S reduction (zeros for rows past the batch end)
        if (owner) {
            float* lrow = sL + (size_t)warp * ldg;
#pragma unroll
            for (int s = 0; s < RK_HSLOT; ++s) {
                const int j = lane + 32 * s;
                if (s < HS && j < H) lrow[j] = cc * hv[s];
            }
            if (lane == 0) lrow[H] = cc;
            if (lane < C) {
                float v = 0.f;
#pragma unroll
                for (int c = 0; c < CT; ++c) if (c == lane) v = dsv[c];
                lrow[H + 1 + lane] = v;
            }
        }
        if (owner && lane == 0) { sRowNll[warp] = nll; sRowG[warp] = gval; sRowCorr[warp] = corr; }
        __syncthreads();

        // ---- loss / group counters of these 8 rows: one atomic per group (update_dict, final_main.py:383-391)
        if (warp == 0) {
            const int gv = lane < RT_ROWS ? sRowG[lane] : -1;
            const int cr = lane < RT_ROWS ? sRowCorr[lane] : 0;
            const float nl = lane < RT_ROWS ? sRowNll[lane] : 0.f;
            const double tot = warp_sum((double)nl);
            if (lane == 0 && a.loss_sum) atomicAdd(&a.loss_sum[a.slot], tot);
            const unsigned cmask = __ballot_sync(0xffffffffu, cr != 0);
            for (int g = 0; g < a.G; ++g) {
                const unsigned gm = __ballot_sync(0xffffffffu, gv == g);
                if (lane == 0 && gm && a.counts) {
                    int64_t* cnt = a.counts + (size_t)a.slot * 2 * a.G;
                    const int nc = __popc(gm & cmask);
                    if (nc) atomicAdd((unsigned long long*)&cnt[g], (unsigned long long)nc);
                    atomicAdd((unsigned long long*)&cnt[a.G + g], (unsigned long long)__popc(gm));
                }
            }
        }
        // ---- phase 4: the rows' operands of S = L^T [h | 1] go to global memory, coalesced (the batch reduction itself is a
        // small tensor-core GEMM off the critical path; round 1 issued 17 K fp32 atomics per CTA here)
        if (a.Spart) {
            // S^T share of these 8 rows on the warp-level tensor cores: A[i][k] = [h | 1] of row k at hidden index i, B[k][j] = the
            // row's [c*h | c | ds]; K = 8 rows = ONE m16n8k8 step per output tile; tiles round-robin over the warps
            const int g = lane >> 2, t4 = lane & 3;
            const int MT = (HP + 15) >> 4, NT = (ldg + 7) >> 3;
            float* outp = a.Spart + (size_t)blockIdx.x * HP * RT_SP_LD;
            for (int tt = warp; tt < MT * NT; tt += RT_WARPS) {
                const int mt = tt / NT, nt = tt - mt * NT, i0 = mt * 16 + g, j0 = nt * 8 + g;
                uint32_t ah[4], al[4], bh[2], bl[2];
                tf32_split(i0 < HP ? sH[(size_t)t4 * HP + i0] : 0.f, ah[0], al[0]);
                tf32_split(i0 + 8 < HP ? sH[(size_t)t4 * HP + i0 + 8] : 0.f, ah[1], al[1]);
                tf32_split(i0 < HP ? sH[(size_t)(t4 + 4) * HP + i0] : 0.f, ah[2], al[2]);
                tf32_split(i0 + 8 < HP ? sH[(size_t)(t4 + 4) * HP + i0 + 8] : 0.f, ah[3], al[3]);
                tf32_split(j0 < ldg ? sL[(size_t)t4 * ldg + j0] : 0.f, bh[0], bl[0]);
                tf32_split(j0 < ldg ? sL[(size_t)(t4 + 4) * ldg + j0] : 0.f, bh[1], bl[1]);
                float c4[4] = {0.f, 0.f, 0.f, 0.f};
                mma_3xtf32(c4, ah, al, bh, bl);
                const int jc = nt * 8 + 2 * t4;
                if (i0 < HP) *reinterpret_cast<float2*>(outp + (size_t)i0 * RT_SP_LD + jc) = make_float2(c4[0], c4[1]);
                if (i0 + 8 < HP) *reinterpret_cast<float2*>(outp + (size_t)(i0 + 8) * RT_SP_LD + jc) = make_float2(c4[2], c4[3]);
            }
        } else {
            const int LDL = l_stride(H, C), LDH = s_stride(H);
            for (int e = tid; e < RT_ROWS * ldg; e += RT_THREADS) {
                const int rr = e / ldg, c = e - rr * ldg;
                if (base + rr < a.B) a.Lrows[(size_t)(base + rr) * LDL + c] = sL[(size_t)rr * ldg + c];
            }
            for (int e = tid; e < RT_ROWS * HP; e += RT_THREADS) {
                const int rr = e / HP, c = e - rr * HP;
                if (base + rr < a.B) a.Hrows[(size_t)(base + rr) * LDH + c] = sH[(size_t)rr * HP + c];
            }
        }
        __syncthreads();
    }

    // (dgamma, dbeta): one slot per warp (sT is free now), summed over the warps in a fixed order, then one fixed-point
    // atomic per element and CTA -- nothing here depends on the order in which warps or CTAs arrive
#pragma unroll
    for (int s = 0; s < RK_HSLOT; ++s) {
        const int j = lane + 32 * s;
        if (s < HS && j < H) { sT[(size_t)warp * 2 * H + j] = dg_acc[s]; sT[(size_t)warp * 2 * H + H + j] = db_acc[s]; }
    }
    __syncthreads();
    for (int e = tid; e < 2 * H; e += RT_THREADS) {
        float v = 0.f;
        for (int w2 = 0; w2 < RT_WARPS; ++w2) v += sT[(size_t)w2 * 2 * H + e];
        fx_add<FX_DGB>(&a.dgb[e], (double)v);
    }
}

template <int NAD, int CT, int NW>
__global__ void __launch_bounds__(NW * 32) k_rows_train(RowsTrainArgs a) { DBMM_TL_SCOPE(TL_ROWS); rows_train_body<NAD, CT, NW>(a); }

static int launch_rows_train(const RowsTrainArgs& ra, int nad, cudaStream_t st) {
    const int CT = ra.C <= 4 ? 4 : 16;
    // 8 warps: half the register file per CTA, so the W2 branch's light CTAs can share the SM (43.3 vs 44.0 us / step at batch
    // 1024, 65.6 vs 66.1 at 8 GPUs) although the kernel alone is a little slower than with 16; DBMM_ROWS_NW=16 restores that
    static const int nw_env = getenv("DBMM_ROWS_NW") ? atoi(getenv("DBMM_ROWS_NW")) : 0;          // tuning switch
    const int nw = (nad == 1 && nw_env == 16) ? 16 : 8;
    const size_t smem = rows_train_smem_bytes(ra.H, ra.C, nad, CT, nw);
    DBMM_CHECK_SHAPE(smem <= 227 * 1024, "train row kernel needs %zu bytes of shared memory", smem);
    int grid = ceil_div(ra.B, RT_ROWS);
    if (grid > 148 * 2) grid = 148 * 2;
#define DBMM_RT_CASE(NAD_, CT_, NW_)                                                                      \
    do {                                                                                                  \
        auto kern = k_rows_train<NAD_, CT_, NW_>;                                                         \
        DBMM_CUDA(set_smem(kern, smem));    \
        DBMM_CUDA(launch_pdl(kern, dim3(grid), dim3(NW_ * 32), smem, st, ra));                            \
    } while (0)
    if (nad == 1 && CT == 4 && nw == 8) DBMM_RT_CASE(1, 4, 8);
    else if (nad == 1 && CT == 4) DBMM_RT_CASE(1, 4, 16);
    else if (nad == 1) DBMM_RT_CASE(1, 16, 16);
    else if (CT == 4) DBMM_RT_CASE(2, 4, 8);
    else DBMM_RT_CASE(2, 16, 8);
#undef DBMM_RT_CASE
    DBMM_LAUNCH_CHECK();
    return DBMM_OK;
}

}  // namespace dbmm
