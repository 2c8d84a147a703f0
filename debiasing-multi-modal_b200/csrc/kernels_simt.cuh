// Round-1 kernels of the adapter hot path: fp32 SIMT, H-space formulation.
//
// After GEMM-1 (a = x W1^T + b1) nothing D-wide is materialised.  With W2a = [W2 | b2] (D x (H+1)),
// h_a = [h, 1] and the per-step Gram matrix  G = W2a^T [W2 | b2 | That]  ((H+1) x (H+1+C)):
//     t = h_a G            t[0:H] = hQ + q,  t[H] = h.q + b2.b2,  t[H+1+c] = s_c = z . That[:,c]
//     n^2 = t[0:H].h + t[H]                     (= ||W2 h + b2||^2)
//     logit_c = s_c / (n tau)
// and the backward needs only  dh = ds M^T + c t[0:H]  plus two small batch reductions
//     S = [c*h | c | ds]^T [h | 1]   ->   dW2a = [W2 | b2 | That] S          (see k_w2grad).
#pragma once
#include <cuda_fp16.h>
#include "gemm_simt.cuh"

namespace dbmm {

// ------------------------------------------------------------------------------------------------
// text prompts: column L2 normalisation (final_main.py:77)
// ------------------------------------------------------------------------------------------------
__global__ void k_normalize_text(const float* __restrict__ T, float* __restrict__ That, int D, int C) {
    const int c = blockIdx.x;
    __shared__ float red[32];
    float s = 0.f;
    for (int d = threadIdx.x; d < D; d += blockDim.x) { float v = T[(size_t)d * C + c]; s = fmaf(v, v, s); }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
        v = warp_sum(v);
        if (threadIdx.x == 0) red[0] = sqrtf(v);
    }
    __syncthreads();
    const float nrm = red[0];
    for (int d = threadIdx.x; d < D; d += blockDim.x) That[(size_t)d * C + c] = T[(size_t)d * C + c] / nrm;
}

// ------------------------------------------------------------------------------------------------
// GEMM-1:  A[ad][b][j] = sum_k X[row(b)][k] W1_ad[j][k] + b1_ad[j]   (+ fp64 column sums for BatchNorm)
// ------------------------------------------------------------------------------------------------
struct Gemm1Args {
    const float* X; int64_t ldx; const int32_t* idx; int64_t pos0;
    int B, D, H, nad;
    const float* W1[2]; const float* b1[2];
    float* A;          // [nad][B][H]
    fx64* colsum;      // [nad][2][H] or nullptr (fixed point, FX_COLSUM)
};

__global__ void __launch_bounds__(GT_THREADS) k_gemm1(Gemm1Args a) {
    __shared__ __align__(16) float smem[GT_SMEM_FLOATS];
    __shared__ int64_t sRow[GT_BM];
    __shared__ unsigned long long sCol[2][GT_BN];      // fixed-point column sums (integer adds commute: deterministic)
    const int m0 = blockIdx.x * GT_BM, n0 = blockIdx.y * GT_BN;
    const int NT = a.nad * a.H;
    for (int i = threadIdx.x; i < GT_BM; i += GT_THREADS) {
        int m = m0 + i;
        int64_t r = -1;
        if (m < a.B) r = a.idx ? (int64_t)a.idx[a.pos0 + m] : (a.pos0 + m);
        sRow[i] = r;
        if (i < GT_BN) { sCol[0][i] = 0ull; sCol[1][i] = 0ull; }
    }
    __syncthreads();
    auto fa = [&](int m, int k) -> float {
        int64_t r = sRow[m - m0];
        return r >= 0 ? __ldg(a.X + r * a.ldx + k) : 0.f;
    };
    auto fb = [&](int k, int n) -> float {
        if (n >= NT) return 0.f;
        int ad = n / a.H, j = n - ad * a.H;
        return __ldg(a.W1[ad] + (size_t)j * a.D + k);
    };
    float acc[GT_TM][GT_TN];
    simt_gemm_tile<true, false>(acc, m0, n0, 0, a.D, fa, fb, smem);

    float s1[GT_TN], s2[GT_TN];
#pragma unroll
    for (int j = 0; j < GT_TN; ++j) {
        s1[j] = 0.f; s2[j] = 0.f;
        const int n = gt_col(n0, j);
        if (n < NT) {
            const int ad = n / a.H, jj = n - ad * a.H;
            const float bias = a.b1[ad][jj];
#pragma unroll
            for (int i = 0; i < GT_TM; ++i) {
                const int m = gt_row(m0, i);
                if (m < a.B) {
                    const float v = acc[i][j] + bias;
                    a.A[((size_t)ad * a.B + m) * a.H + jj] = v;
                    s1[j] += v; s2[j] = fmaf(v, v, s2[j]);
                }
            }
        }
    }
    if (a.colsum) {
#pragma unroll
        for (int j = 0; j < GT_TN; ++j) {
            // lanes l and l^16 hold the same column (tx = tid % 16)
            float v1 = s1[j] + __shfl_xor_sync(0xffffffffu, s1[j], 16);
            float v2 = s2[j] + __shfl_xor_sync(0xffffffffu, s2[j], 16);
            if ((threadIdx.x & 16) == 0) {
                const int cl = (threadIdx.x % (GT_BN / GT_TN)) * GT_TN + j;
                atomicAdd(&sCol[0][cl], fx_bits<FX_COLSUM>((double)v1));
                atomicAdd(&sCol[1][cl], fx_bits<FX_COLSUM>((double)v2));
            }
        }
        __syncthreads();
        for (int i = threadIdx.x; i < GT_BN; i += GT_THREADS) {
            const int n = n0 + i;
            if (n < NT) {
                const int ad = n / a.H, jj = n - ad * a.H;
                atomicAdd(reinterpret_cast<unsigned long long*>(&a.colsum[((size_t)ad * 2 + 0) * a.H + jj].v), sCol[0][i]);
                atomicAdd(reinterpret_cast<unsigned long long*>(&a.colsum[((size_t)ad * 2 + 1) * a.H + jj].v), sCol[1][i]);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Gram matrices:  gram[ad] = [W2 | b2]^T [W2 | b2 | That]      ((H+1) x (H+1+C), K = D)
// split-K over gridDim.z / nad partial products, accumulated with fp32 atomics into a zeroed buffer.
// ------------------------------------------------------------------------------------------------
struct GramArgs {
    const float* W2[2]; const float* b2[2]; const float* That;
    float* gram; int D, H, C, nad, ksplit;
};

__global__ void __launch_bounds__(GT_THREADS) k_gram(GramArgs a) {
    __shared__ __align__(16) float smem[GT_SMEM_FLOATS];
    const int ad = blockIdx.z / a.ksplit, part = blockIdx.z - ad * a.ksplit;
    const int H = a.H, C = a.C, M = H + 1, N = H + 1 + C;
    const int m0 = blockIdx.x * GT_BM, n0 = blockIdx.y * GT_BN;
    const float* W2 = ad == 0 ? a.W2[0] : a.W2[1];
    const float* b2 = ad == 0 ? a.b2[0] : a.b2[1];
    const float* That = a.That;
    auto fa = [=](int i, int d) -> float {
        const float* p = (i < H) ? (W2 + (size_t)d * H + i) : (b2 + d);
        return (i <= H) ? __ldg(p) : 0.f;
    };
    auto fb = [=](int d, int j) -> float {
        const float* p = (j < H) ? (W2 + (size_t)d * H + j) : (j == H ? (b2 + d) : (That + (size_t)d * C + (j - H - 1)));
        return (j < N) ? __ldg(p) : 0.f;
    };
    int k0, k1;
    gt_split_k(a.D, a.ksplit, part, k0, k1);
    float acc[GT_TM][GT_TN];
    simt_gemm_tile<false, true>(acc, m0, n0, k0, k1, fa, fb, smem);
    float* out = a.gram + (size_t)ad * M * N;
#pragma unroll
    for (int i = 0; i < GT_TM; ++i) {
        const int m = gt_row(m0, i);
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < GT_TN; ++j) {
            const int n = gt_col(n0, j);
            if (n < N) {
                if (a.ksplit > 1) atomicAdd(&out[(size_t)m * N + n], acc[i][j]);
                else out[(size_t)m * N + n] = acc[i][j];
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Row kernel: BatchNorm -> ReLU -> H-space logits -> CE / argmax / group counters (-> row backward)
// ------------------------------------------------------------------------------------------------
constexpr int RK_NSLOT = 5, RK_HSLOT = 4;     // 32-wide column slots: Gram columns (<= 160) / hidden units (<= 128)

__host__ __device__ inline size_t rows_gram_floats_dev(int H, int C, int nad) {
    return ((size_t)nad * (H + 1) * (H + 1 + C) + 3) / 4 * 4;
}

struct RowsArgs {
    int64_t N;            // rows in this launch
    int64_t pos0;         // position of row 0 inside the caller's row list
    const int32_t* idx; const int32_t* y; const int32_t* grp;
    int H, C, G;
    const float* A; int64_t strideA;      // [nad][N][H]
    const float* gram;                    // [nad][H+1][H+1+C]
    const double* colsum; int64_t Bg;     // train: batch statistics from the fp64 column sums
    AdapterView ad[2];
    float w_old, inv_tau, inv_B;
    float* logits_out; int32_t* pred_out;
    double* loss_sum; int64_t* counts; int64_t batch_size; int64_t slot_fixed;
    float* hbuf; float* dahat; float* cvec; float* ds; double* dgb;
};

static inline size_t rows_gram_floats(int H, int C, int nad) {
    return ((size_t)nad * (H + 1) * (H + 1 + C) + 3) / 4 * 4;          // padded to 16 bytes
}
static inline size_t rows_smem_bytes(int H, int C, int nad, int CT, int rows_per_cta) {
    size_t fl = rows_gram_floats(H, C, nad) + (size_t)nad * 4 * H + (size_t)rows_per_cta * (H + 1)
              + 32 + 32 + 32 + (size_t)rows_per_cta * CT + 2 * (size_t)H;
    return fl * 4;
}

// RK_RB rows per warp (they share every Gram element read from shared memory), RK_WARPS warps per CTA.
template <bool TRAIN, int NAD, int CT, int RK_RB, int RK_WARPS>
__global__ void __launch_bounds__(RK_WARPS * 32) k_rows(RowsArgs a) {
    constexpr int RK_ROWS = RK_RB * RK_WARPS;
    static_assert(RK_ROWS <= 32, "the group reduction maps one CTA row to one lane of warp 0");
    extern __shared__ __align__(16) float dyn_smem[];
    const int H = a.H, C = a.C, ldg = H + 1 + C;
    const int HS = (H + 31) >> 5;             // slots holding hidden units
    const int NS = (ldg + 31) >> 5;           // slots holding Gram columns
    float* sG = dyn_smem;                                     // [NAD][H+1][ldg]
    float* sBN = sG + rows_gram_floats_dev(H, C, NAD);        // [NAD][4][H]: mu, rstd, gamma, beta
    float* sH = sBN + (size_t)NAD * 4 * H;                    // [RK_ROWS][H+1]
    float* sRowNll = sH + (size_t)RK_ROWS * (H + 1);          // [32]
    int* sRowG = reinterpret_cast<int*>(sRowNll + 32);        // [32]
    int* sRowCorr = sRowG + 32;                               // [32]
    float* sLo = reinterpret_cast<float*>(sRowCorr + 32);     // [RK_ROWS][CT] old adapter's share of the logits
    float* sDgb = sLo + (size_t)RK_ROWS * CT;                 // [2][H]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int ADT = NAD - 1;              // the trainable / only adapter

    {   // Gram matrices -> shared memory: 16-byte cp.async, everything in flight at once
        const int n = NAD * (H + 1) * ldg, n4 = n >> 2;
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(sG);
        for (int e = tid; e < n4; e += blockDim.x)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + e * 16), "l"(a.gram + e * 4) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
        for (int e = (n4 << 2) + tid; e < n; e += blockDim.x) sG[e] = a.gram[e];
    }
    for (int e = tid; e < NAD * H; e += blockDim.x) {
        const int ad = e / H, j = e - ad * H;
        float mu, var;
        if (TRAIN) {
            const double s1 = a.colsum[((size_t)ad * 2 + 0) * H + j], s2 = a.colsum[((size_t)ad * 2 + 1) * H + j];
            const double m = s1 / (double)a.Bg;
            double v = s2 / (double)a.Bg - m * m;
            if (v < 0.0) v = 0.0;
            mu = (float)m; var = (float)v;
        } else {
            mu = a.ad[ad].running_mean[j]; var = a.ad[ad].running_var[j];
        }
        float* bn = sBN + (size_t)ad * 4 * H;
        bn[j] = mu; bn[H + j] = 1.0f / sqrtf(var + DBMM_BN_EPS);
        bn[2 * H + j] = a.ad[ad].gamma[j]; bn[3 * H + j] = a.ad[ad].beta[j];
    }
    if (TRAIN) for (int e = tid; e < 2 * H; e += blockDim.x) sDgb[e] = 0.f;
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();

    float dg_acc[RK_HSLOT], db_acc[RK_HSLOT];
#pragma unroll
    for (int s = 0; s < RK_HSLOT; ++s) { dg_acc[s] = 0.f; db_acc[s] = 0.f; }

    const float coef = (NAD == 2) ? (1.0f - a.w_old) : 1.0f;

    for (int64_t base = (int64_t)blockIdx.x * RK_ROWS; base < a.N; base += (int64_t)gridDim.x * RK_ROWS) {
        float t[RK_RB][RK_NSLOT];
        float hv[RK_RB][RK_HSLOT];
        float ahat[RK_RB][RK_HSLOT];
        unsigned prepos[RK_RB];
        float n2[RK_RB];
        float sc[RK_RB][CT];

#pragma unroll
        for (int ad = 0; ad < NAD; ++ad) {
            const float* bn = sBN + (size_t)ad * 4 * H;
            const float* G = sG + (size_t)ad * (H + 1) * ldg;
#pragma unroll
            for (int rb = 0; rb < RK_RB; ++rb) {
                const int64_t r = base + warp * RK_RB + rb;
                float* hrow = sH + (size_t)(warp * RK_RB + rb) * (H + 1);
                prepos[rb] = 0u;
#pragma unroll
                for (int s = 0; s < RK_HSLOT; ++s) {
                    const int j = lane + 32 * s;
                    float h = 0.f, ah = 0.f;
                    if (s < HS && j < H && r < a.N) {
                        const float av = a.A[(size_t)ad * a.strideA + (size_t)r * H + j];
                        ah = (av - bn[j]) * bn[H + j];
                        const float pre = fmaf(ah, bn[2 * H + j], bn[3 * H + j]);
                        if (pre > 0.f) { h = pre; prepos[rb] |= (1u << s); }
                    }
                    hv[rb][s] = h; ahat[rb][s] = ah;
                    if (s < HS && j < H) hrow[j] = h;
                }
                if (lane == 0) hrow[H] = 1.0f;
            }
            __syncwarp();
#pragma unroll
            for (int rb = 0; rb < RK_RB; ++rb)
#pragma unroll
                for (int s = 0; s < RK_NSLOT; ++s) t[rb][s] = 0.f;
            const float* hbase = sH + (size_t)(warp * RK_RB) * (H + 1);
#pragma unroll 4
            for (int i = 0; i <= H; ++i) {
                float g[RK_NSLOT];
#pragma unroll
                for (int s = 0; s < RK_NSLOT; ++s) {
                    const int j = lane + 32 * s;
                    g[s] = (s < NS && j < ldg) ? G[(size_t)i * ldg + j] : 0.f;
                }
#pragma unroll
                for (int rb = 0; rb < RK_RB; ++rb) {
                    const float hh = hbase[(size_t)rb * (H + 1) + i];
#pragma unroll
                    for (int s = 0; s < RK_NSLOT; ++s) t[rb][s] = fmaf(hh, g[s], t[rb][s]);
                }
            }
            __syncwarp();
            // n^2 and the C prompt scores of this adapter
#pragma unroll
            for (int rb = 0; rb < RK_RB; ++rb) {
                float part = 0.f;
#pragma unroll
                for (int s = 0; s < RK_HSLOT; ++s) part = fmaf(t[rb][s], hv[rb][s], part);   // hv = 0 outside j < H
#pragma unroll
                for (int s = 0; s < RK_NSLOT; ++s)
                    if (lane + 32 * s == H) part += t[rb][s];
                n2[rb] = warp_sum(part);
#pragma unroll
                for (int c = 0; c < CT; ++c) {
                    float v = 0.f;
                    if (c < C) {
                        const int col = H + 1 + c;
#pragma unroll
                        for (int s = 0; s < RK_NSLOT; ++s) {
                            const float tmp = __shfl_sync(0xffffffffu, t[rb][s], col & 31);
                            if (s == (col >> 5)) v = tmp;
                        }
                    }
                    sc[rb][c] = v;
                }
                if (NAD == 2 && ad == 0) {
                    // old adapter contributes  w * s_old / (n_old tau)  to the logits, as a constant
                    const float inv_n = 1.0f / sqrtf(n2[rb]);
                    if (lane < CT)  {
                        float v = 0.f;
#pragma unroll
                        for (int c = 0; c < CT; ++c) if (c == lane) v = sc[rb][c];
                        sLo[(size_t)(warp * RK_RB + rb) * CT + lane] = a.w_old * a.inv_tau * v * inv_n;
                    }
                }
            }
            __syncwarp();
        }

        // ---- per-row epilogue on the trainable adapter's t / n2 / sc
#pragma unroll
        for (int rb = 0; rb < RK_RB; ++rb) {
            const int64_t r = base + warp * RK_RB + rb;
            const int lr = warp * RK_RB + rb;
            const bool valid = r < a.N;
            float nll = 0.f; int gval = -1, corr = 0;
            if (valid) {
                const int64_t pos = a.pos0 + r;
                const int64_t dsrow = a.idx ? (int64_t)a.idx[pos] : pos;
                const int yv = a.y ? a.y[dsrow] : -1;     // y == NULL: logits / argmax only
                gval = a.grp ? a.grp[dsrow] : 0;
                const float nn = sqrtf(n2[rb]);
                const float inv_n = 1.0f / nn;
                float lnew[CT], l[CT];
                float mx = -INFINITY; int am = 0;
#pragma unroll
                for (int c = 0; c < CT; ++c) {
                    lnew[c] = 0.f; l[c] = -INFINITY;
                    if (c < C) {
                        lnew[c] = a.inv_tau * sc[rb][c] * inv_n;
                        l[c] = (NAD == 2) ? fmaf(coef, lnew[c], sLo[(size_t)lr * CT + c]) : lnew[c];
                        if (l[c] > mx) { mx = l[c]; am = c; }
                    }
                }
                float se = 0.f, ly = 0.f;
                float p[CT];
#pragma unroll
                for (int c = 0; c < CT; ++c) {
                    p[c] = 0.f;
                    if (c < C) { p[c] = expf(l[c] - mx); se += p[c]; if (c == yv) ly = l[c]; }
                }
                nll = logf(se) + mx - ly;
                corr = (am == yv) ? 1 : 0;
                if (a.logits_out && lane < C) {
                    float v = 0.f;
#pragma unroll
                    for (int c = 0; c < CT; ++c) if (c == lane) v = l[c];
                    a.logits_out[(size_t)pos * C + lane] = v;
                }
                if (a.pred_out && lane == 0) a.pred_out[pos] = am;

                if (TRAIN) {
                    const float inv_se = 1.0f / se;
                    float dsv[CT];
                    float dot = 0.f;
#pragma unroll
                    for (int c = 0; c < CT; ++c) {
                        dsv[c] = 0.f;
                        if (c < C) {
                            const float dl = (p[c] * inv_se - (c == yv ? 1.f : 0.f)) * a.inv_B;
                            dot = fmaf(dl, lnew[c], dot);
                            dsv[c] = coef * dl * a.inv_tau * inv_n;
                        }
                    }
                    const float cc = -coef * dot / n2[rb];
                    const float* G = sG + (size_t)ADT * (H + 1) * ldg;
                    const float* bn = sBN + (size_t)ADT * 4 * H;
#pragma unroll
                    for (int s = 0; s < RK_HSLOT; ++s) {
                        const int j = lane + 32 * s;
                        if (s < HS && j < H) {
                            float dh = cc * t[rb][s];
#pragma unroll
                            for (int c = 0; c < CT; ++c)
                                if (c < C) dh = fmaf(dsv[c], G[(size_t)j * ldg + H + 1 + c], dh);
                            const float dpre = ((prepos[rb] >> s) & 1u) ? dh : 0.f;
                            dg_acc[s] = fmaf(dpre, ahat[rb][s], dg_acc[s]);
                            db_acc[s] += dpre;
                            a.dahat[(size_t)r * H + j] = dpre * bn[2 * H + j];
                            a.hbuf[(size_t)r * H + j] = hv[rb][s];
                        }
                    }
                    if (lane == 0) a.cvec[r] = cc;
                    if (lane < C) {
                        float v = 0.f;
#pragma unroll
                        for (int c = 0; c < CT; ++c) if (c == lane) v = dsv[c];
                        a.ds[(size_t)r * C + lane] = v;
                    }
                }
            }
            if (lane == 0) { sRowNll[lr] = nll; sRowG[lr] = gval; sRowCorr[lr] = corr; }
        }
        __syncthreads();
        // ---- warp-shuffle group reduction over this CTA's 32 rows: one atomic per group
        if (warp == 0) {
            const int64_t r = base + lane;
            const bool valid = lane < RK_ROWS && r < a.N;
            const int64_t slot = a.slot_fixed >= 0 ? a.slot_fixed : (valid ? (a.pos0 + r) / a.batch_size : -1);
            const int64_t slot0 = __shfl_sync(0xffffffffu, slot, 0);
            const bool uniform = __all_sync(0xffffffffu, !valid || slot == slot0);
            const int gv = lane < RK_ROWS ? sRowG[lane] : -1;
            const int cr = lane < RK_ROWS ? sRowCorr[lane] : 0;
            const float nl = lane < RK_ROWS ? sRowNll[lane] : 0.f;
            if (uniform) {
                const double tot = warp_sum((double)nl);
                if (lane == 0 && a.loss_sum) atomicAdd(&a.loss_sum[slot0], tot);
                const unsigned cmask = __ballot_sync(0xffffffffu, cr != 0);
                for (int g = 0; g < a.G; ++g) {
                    const unsigned gm = __ballot_sync(0xffffffffu, gv == g);
                    if (lane == 0 && gm && a.counts) {
                        int64_t* cnt = a.counts + (size_t)slot0 * 2 * a.G;
                        const int nc = __popc(gm & cmask);
                        if (nc) atomicAdd((unsigned long long*)&cnt[g], (unsigned long long)nc);
                        atomicAdd((unsigned long long*)&cnt[a.G + g], (unsigned long long)__popc(gm));
                    }
                }
            } else if (valid) {
                if (a.loss_sum) atomicAdd(&a.loss_sum[slot], (double)nl);
                if (a.counts && gv >= 0 && gv < a.G) {
                    int64_t* cnt = a.counts + (size_t)slot * 2 * a.G;
                    if (cr) atomicAdd((unsigned long long*)&cnt[gv], 1ull);
                    atomicAdd((unsigned long long*)&cnt[a.G + gv], 1ull);
                }
            }
        }
        __syncthreads();
    }

    if (TRAIN) {
#pragma unroll
        for (int s = 0; s < RK_HSLOT; ++s) {
            const int j = lane + 32 * s;
            if (s < HS && j < H) { atomicAdd(&sDgb[j], dg_acc[s]); atomicAdd(&sDgb[H + j], db_acc[s]); }
        }
        __syncthreads();
        for (int e = tid; e < 2 * H; e += blockDim.x) atomicAdd(&a.dgb[e], (double)sDgb[e]);
    }
}

// ------------------------------------------------------------------------------------------------
// dW1 = da^T X, fp32 SIMT tiles (generic shapes; the tensor-core kernel in wgrad_tc.cuh covers D % 128 == 0)
// ------------------------------------------------------------------------------------------------
struct WgradArgs {
    const float* X; int64_t ldx; const int32_t* idx;
    int B; int64_t Bg; int D, H;
    const float* A; const float* dahat;
    const fx64* colsum; const fx64* dgb; const float* gamma;      // fixed point (FX_COLSUM / FX_DGB)
    float* gW1;
    int tiles_m, tiles_n, ksplit;
};

__global__ void __launch_bounds__(GT_THREADS) k_wgrad(WgradArgs a) {
    __shared__ __align__(16) float smem[GT_SMEM_FLOATS];
    __shared__ float sCst[4][GT_BM];      // mu, rstd, m1, m2 for this tile's hidden units
    const int H = a.H;
    const int tile = blockIdx.x;
    int k0, k1;
    gt_split_k(a.B, a.ksplit, blockIdx.y, k0, k1);
    float acc[GT_TM][GT_TN];
    const int m0 = (tile / a.tiles_n) * GT_BM, n0 = (tile % a.tiles_n) * GT_BN;
    for (int i = threadIdx.x; i < GT_BM; i += GT_THREADS) {
        const int j = m0 + i;
        float mu = 0.f, rstd = 0.f, m1 = 0.f, m2 = 0.f;
        if (j < H) {
            const double m = fx_get<FX_COLSUM>(&a.colsum[j]) / (double)a.Bg;
            double v = fx_get<FX_COLSUM>(&a.colsum[H + j]) / (double)a.Bg - m * m;
            if (v < 0.0) v = 0.0;
            mu = (float)m; rstd = 1.0f / sqrtf((float)v + DBMM_BN_EPS);
            const double gm = (double)a.gamma[j];
            m1 = (float)(gm * fx_get<FX_DGB>(&a.dgb[H + j]) / (double)a.Bg);     // mean_B(dahat)       = gamma * dbeta  / B
            m2 = (float)(gm * fx_get<FX_DGB>(&a.dgb[j]) / (double)a.Bg);         // mean_B(dahat*ahat)  = gamma * dgamma / B
        }
        sCst[0][i] = mu; sCst[1][i] = rstd; sCst[2][i] = m1; sCst[3][i] = m2;
    }
    __syncthreads();
    const float* Ap = a.A; const float* dap = a.dahat; const float* Xp = a.X; const int32_t* idxp = a.idx;
    const int64_t ldx = a.ldx; const int Dd = a.D;
    auto fa = [=](int j, int b) -> float {          // da[b][j]  (sCst is static shared: referenced directly)
        const int jj = j < H ? j : H - 1;
        const int i = jj - m0;
        const float ah = (__ldg(Ap + (size_t)b * H + jj) - sCst[0][i]) * sCst[1][i];
        const float v = (__ldg(dap + (size_t)b * H + jj) - sCst[2][i] - ah * sCst[3][i]) * sCst[1][i];
        return j < H ? v : 0.f;
    };
    auto fb = [=](int b, int k) -> float {
        const int64_t r = idxp ? (int64_t)__ldg(idxp + b) : (int64_t)b;
        const int kk = k < Dd ? k : Dd - 1;
        const float v = __ldg(Xp + r * ldx + kk);
        return k < Dd ? v : 0.f;
    };
    simt_gemm_tile<false, true>(acc, m0, n0, k0, k1, fa, fb, smem);
#pragma unroll
    for (int i = 0; i < GT_TM; ++i) {
        const int m = gt_row(m0, i);
        if (m >= H) continue;
#pragma unroll
        for (int j = 0; j < GT_TN; ++j) {
            const int n = gt_col(n0, j);
            if (n < a.D) {
                if (a.ksplit > 1) atomicAdd(&a.gW1[(size_t)m * a.D + n], acc[i][j]);
                else a.gW1[(size_t)m * a.D + n] = acc[i][j];
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// SGD (torch.optim.SGD semantics, demo/util.py:118-136) + BatchNorm running-stat update
// ------------------------------------------------------------------------------------------------
struct SgdArgs {
    float* p[6]; int64_t off[7];
    const float* g; float* v;
    float lr, momentum, wd; int first; int vec4;
    int nad, H; int64_t Bg; const double* colsum;
    float* rm[2]; float* rv[2]; long long* nbt[2];
};

__global__ void __launch_bounds__(256) k_sgd(SgdArgs a) {
    const int64_t n = a.off[6];
    if (a.vec4) {          // every tensor size (and so every flat offset) is a multiple of 4: 16-byte accesses
        for (int64_t i4 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i4 * 4 < n; i4 += (int64_t)gridDim.x * blockDim.x) {
            const int64_t i = i4 * 4;
            int seg = 0;
#pragma unroll
            for (int s = 1; s < 6; ++s) seg += (i >= a.off[s]) ? 1 : 0;
            float4* pp = reinterpret_cast<float4*>(a.p[seg] + (i - a.off[seg]));
            const float4 pv = *pp;
            const float4 gv = *reinterpret_cast<const float4*>(a.g + i);
            float4 vv = a.first ? make_float4(0.f, 0.f, 0.f, 0.f) : *reinterpret_cast<const float4*>(a.v + i);
            const float px[4] = {pv.x, pv.y, pv.z, pv.w}, gx[4] = {gv.x, gv.y, gv.z, gv.w}, vx[4] = {vv.x, vv.y, vv.z, vv.w};
            float po[4], vo[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float g = gx[q] + a.wd * px[q];
                vo[q] = a.first ? g : (a.momentum * vx[q] + g);
                po[q] = px[q] - a.lr * vo[q];
            }
            *reinterpret_cast<float4*>(a.v + i) = make_float4(vo[0], vo[1], vo[2], vo[3]);
            *pp = make_float4(po[0], po[1], po[2], po[3]);
        }
    } else
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int seg = 0;
#pragma unroll
        for (int s = 1; s < 6; ++s) seg += (i >= a.off[s]) ? 1 : 0;
        float* pp = a.p[seg] + (i - a.off[seg]);
        const float pv = *pp;
        const float g = a.g[i] + a.wd * pv;
        const float v = a.first ? g : (a.momentum * a.v[i] + g);
        a.v[i] = v;
        *pp = pv - a.lr * v;
    }
    if (blockIdx.x == gridDim.x - 1 && a.colsum) {
        for (int e = threadIdx.x; e < a.nad * a.H; e += blockDim.x) {
            const int ad = e / a.H, j = e - ad * a.H;
            const double m = a.colsum[((size_t)ad * 2 + 0) * a.H + j] / (double)a.Bg;
            double var = a.colsum[((size_t)ad * 2 + 1) * a.H + j] / (double)a.Bg - m * m;
            if (var < 0.0) var = 0.0;
            const float unbiased = (float)(var * (double)a.Bg / (double)(a.Bg - 1));
            a.rm[ad][j] = (1.0f - DBMM_BN_MOMENTUM) * a.rm[ad][j] + DBMM_BN_MOMENTUM * (float)m;
            a.rv[ad][j] = (1.0f - DBMM_BN_MOMENTUM) * a.rv[ad][j] + DBMM_BN_MOMENTUM * unbiased;
        }
        if (threadIdx.x < a.nad) *a.nbt[threadIdx.x] += 1;
    }
}

__global__ void __launch_bounds__(256) k_sgd_flat(float* __restrict__ p, const float* __restrict__ g,
                                                  float* __restrict__ v, int64_t n, float lr, float momentum,
                                                  float wd, int first) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float pv = p[i];
        const float gg = g[i] + wd * pv;
        const float vv = first ? gg : (momentum * v[i] + gg);
        v[i] = vv;
        p[i] = pv - lr * vv;
    }
}

// ------------------------------------------------------------------------------------------------
// Ingest: fp16 -> fp32 rows (exact).  Streaming: 16-byte loads of 8 halves, two 16-byte stores; D % 8 == 0 and 16-byte
// aligned rows take the vector path, anything else the scalar one.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_widen_f16(const __half* __restrict__ src, int64_t ld_src, float* __restrict__ dst,
                                                   int64_t ld_dst, int64_t n_rows, int D, int vec) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x, t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (vec) {
        const int d8 = D >> 3;
        const int64_t n8 = n_rows * d8;
        for (int64_t i = t0; i < n8; i += stride) {
            const int64_t r = i / d8; const int c = (int)(i - r * d8) << 3;
            const uint4 raw = __ldcs(reinterpret_cast<const uint4*>(src + r * ld_src + c));
            const __half2* h = reinterpret_cast<const __half2*>(&raw);
            const float2 a = __half22float2(h[0]), b = __half22float2(h[1]), c2 = __half22float2(h[2]), d = __half22float2(h[3]);
            float4* o = reinterpret_cast<float4*>(dst + r * ld_dst + c);
            o[0] = make_float4(a.x, a.y, b.x, b.y);
            o[1] = make_float4(c2.x, c2.y, d.x, d.y);
        }
    } else {
        const int64_t n = n_rows * D;
        for (int64_t i = t0; i < n; i += stride) {
            const int64_t r = i / D; const int c = (int)(i - r * D);
            dst[r * ld_dst + c] = __half2float(src[r * ld_src + c]);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// update_dict on given logits (final_main.py:383-391): thread per row, ballot per group
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_group_counts(const float* __restrict__ logits, const int32_t* __restrict__ y,
                                                      const int32_t* __restrict__ grp, int64_t N, int C, int G,
                                                      int64_t batch_size, double* loss_sum, int64_t* counts,
                                                      int32_t* pred_out) {
    const int lane = threadIdx.x & 31;
    const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t base = warp_global * 32; base < N; base += nwarps * 32) {
        const int64_t r = base + lane;
        const bool valid = r < N;
        int gv = -1, corr = 0; float nll = 0.f;
        if (valid) {
            const float* l = logits + (size_t)r * C;
            float mx = l[0]; int am = 0;
            for (int c = 1; c < C; ++c) if (l[c] > mx) { mx = l[c]; am = c; }
            float se = 0.f;
            for (int c = 0; c < C; ++c) se += expf(l[c] - mx);
            const int yv = y[r];
            nll = logf(se) + mx - l[yv];
            corr = am == yv;
            gv = grp ? grp[r] : 0;
            if (pred_out) pred_out[r] = am;
        }
        const int64_t slot = valid ? r / batch_size : -1;
        const int64_t slot0 = __shfl_sync(0xffffffffu, slot, 0);
        const bool uniform = __all_sync(0xffffffffu, !valid || slot == slot0);
        if (uniform) {
            const double tot = warp_sum((double)nll);
            if (lane == 0 && loss_sum) atomicAdd(&loss_sum[slot0], tot);
            const unsigned cmask = __ballot_sync(0xffffffffu, corr != 0);
            for (int g = 0; g < G; ++g) {
                const unsigned gm = __ballot_sync(0xffffffffu, gv == g);
                if (lane == 0 && gm && counts) {
                    int64_t* cnt = counts + (size_t)slot0 * 2 * G;
                    const int nc = __popc(gm & cmask);
                    if (nc) atomicAdd((unsigned long long*)&cnt[g], (unsigned long long)nc);
                    atomicAdd((unsigned long long*)&cnt[G + g], (unsigned long long)__popc(gm));
                }
            }
        } else if (valid) {
            if (loss_sum) atomicAdd(&loss_sum[slot], (double)nll);
            if (counts && gv >= 0 && gv < G) {
                int64_t* cnt = counts + (size_t)slot * 2 * G;
                if (corr) atomicAdd((unsigned long long*)&cnt[gv], 1ull);
                atomicAdd((unsigned long long*)&cnt[G + gv], 1ull);
            }
        }
    }
}

}  // namespace dbmm
