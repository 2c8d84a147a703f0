// Contrastive-adapter training step (`--tl_method contrastive_adapter`; the reference accepts the flag, final_main.py:230, but has
// no runnable branch for it -- the formula and the epoch come from demo/visualizer_supcon.py:412-508 / 1522-1587 and
// workspace/jinsu/SupCon.ipynb:109-113):
//
//     x' = x / |x|  (ca_pre_norm)   z = adapter(x')   u = z / |z|  (forward_ca, head = identity)
//     L  = w * SupCon_all_anchors(u, labels; tau_cl)          (the B x B tcgen05 GEMMs of head_supcon.cuh / dbmm_supcon_*)
//
// and the D-WIDE backward the H-space collapse of the CE path cannot provide (the loss is not a function of a few prompt
// scores): dz = (du - u (u.du)) / |z|, dW2 = dz^T h, db2 = sum dz, dh = dz W2, ReLU / BatchNorm backward, dW1 = da^T x'.
// The three dense contractions (z = h W2^T, dh = dz W2, dW2 = dz^T h) run on k_tc_gemm_nt (TMA + tcgen05, 3xTF32), GEMM-1 and
// dW1 on the kernels of the CE step (k_gemm1_tc, k_wgrad_tc with a per-row scale for the input normalisation); the kernels
// below are the row-wise glue.  Everything is deterministic (fixed-order column sums, no floating-point atomics).
#pragma once
#include "common.cuh"
#include "kernels_simt.cuh"

namespace dbmm {

// inv_xn[b] = 1 / |x_row(b)| (or 1), labels_b[b] = labels[row(b)]; one warp per batch row
__global__ void __launch_bounds__(256) k_ca_rows_in(const float* __restrict__ X, int64_t ldx, const int32_t* __restrict__ idx, int B, int D,
                                                    const int32_t* __restrict__ labels, int pre_norm, float* __restrict__ inv_xn,
                                                    int32_t* __restrict__ labels_b) {
    const int b = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (b >= B) return;
    const int64_t row = idx ? (int64_t)idx[b] : (int64_t)b;
    float s = 0.f;
    if (pre_norm) {
        const float* x = X + row * ldx;
        for (int k = lane; k < D; k += 32) { const float v = x[k]; s = fmaf(v, v, s); }
        s = warp_sum(s);
    }
    if (lane == 0) { inv_xn[b] = pre_norm ? 1.0f / sqrtf(s) : 1.0f; labels_b[b] = labels[row]; }
}

// a' = (A - b1) * inv_xn + b1 in place (A = x W1^T + b1 from GEMM-1) and the BatchNorm column sums of a' as fixed-point
// integers (the format k_wgrad_tc reads).  One CTA per 32 hidden units, 8 row lanes, fixed summation order.
__global__ void __launch_bounds__(256) k_ca_bn_stats(float* __restrict__ A, const float* __restrict__ b1, const float* __restrict__ inv_xn,
                                                     int B, int H, fx64* __restrict__ colsum) {
    __shared__ double s1[8][32], s2[8][32];
    const int j = blockIdx.x * 32 + (threadIdx.x & 31), rl = threadIdx.x >> 5;
    double a1 = 0.0, a2 = 0.0;
    if (j < H) {
        const float bj = b1[j];
        for (int b = rl; b < B; b += 8) {
            const float v = fmaf(A[(size_t)b * H + j] - bj, inv_xn[b], bj);
            A[(size_t)b * H + j] = v;
            a1 += (double)v; a2 += (double)v * (double)v;
        }
    }
    s1[rl][threadIdx.x & 31] = a1; s2[rl][threadIdx.x & 31] = a2;
    __syncthreads();
    if (rl == 0 && j < H) {
        double t1 = 0.0, t2 = 0.0;
        for (int r = 0; r < 8; ++r) { t1 += s1[r][threadIdx.x]; t2 += s2[r][threadIdx.x]; }
        colsum[j].v = (long long)fx_bits<FX_COLSUM>(t1);
        colsum[H + j].v = (long long)fx_bits<FX_COLSUM>(t2);
    }
}

// h = relu(gamma * ahat + beta), also split hi + lo for the tensor-core GEMM-2
__global__ void __launch_bounds__(256) k_ca_hidden(const float* __restrict__ A, const fx64* __restrict__ colsum, const float* __restrict__ gamma,
                                                   const float* __restrict__ beta, int B, int H, float* __restrict__ h,
                                                   float* __restrict__ h_hi, float* __restrict__ h_lo) {
    const int64_t n = (int64_t)B * H;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        const int j = (int)(e % H);
        const double m = fx_get<FX_COLSUM>(&colsum[j]) / (double)B;
        double v = fx_get<FX_COLSUM>(&colsum[H + j]) / (double)B - m * m;
        if (v < 0.0) v = 0.0;
        const float ah = (A[e] - (float)m) * (1.0f / sqrtf((float)v + DBMM_BN_EPS));
        const float hv = fmaxf(fmaf(ah, gamma[j], beta[j]), 0.f);
        const float hi = __uint_as_float(__float_as_uint(hv) & 0xffffe000u);
        h[e] = hv; h_hi[e] = hi; h_lo[e] = hv - hi;
    }
}

// u = (Z + b2) / |Z + b2|, inv_n = 1 / |Z + b2|; one warp per row (Z is overwritten by u)
__global__ void __launch_bounds__(256) k_ca_normalize(float* __restrict__ Z, const float* __restrict__ b2, int B, int D, float* __restrict__ inv_n) {
    const int b = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (b >= B) return;
    float* z = Z + (size_t)b * D;
    float s = 0.f;
    for (int k = lane; k < D; k += 32) { const float v = z[k] + b2[k]; z[k] = v; s = fmaf(v, v, s); }
    s = warp_sum(s);
    const float inv = 1.0f / sqrtf(s);
    for (int k = lane; k < D; k += 32) z[k] *= inv;
    if (lane == 0) inv_n[b] = inv;
}

// dz = w * (du - u (u.du)) * inv_n with du = dU_a + dU_b (anchor-role + contrast-role gradients of the supcon kernels);
// (dU_b may be NULL: dU_a is the whole gradient); written over dU_a, also split hi + lo
__global__ void __launch_bounds__(256) k_ca_dz(const float* __restrict__ U, float* __restrict__ dUa, const float* __restrict__ dUb,
                                               const float* __restrict__ inv_n, float weight, int B, int D, float* __restrict__ dz_hi,
                                               float* __restrict__ dz_lo) {
    const int b = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (b >= B) return;
    const float* u = U + (size_t)b * D;
    float* da = dUa + (size_t)b * D;
    const float* db = dUb + (size_t)b * D;
    float dot = 0.f;
    for (int k = lane; k < D; k += 32) dot = fmaf(u[k], da[k] + (dUb ? db[k] : 0.f), dot);
    dot = warp_sum(dot);
    const float sc = weight * inv_n[b];
    for (int k = lane; k < D; k += 32) {
        const float v = sc * ((da[k] + (dUb ? db[k] : 0.f)) - u[k] * dot);
        const float hi = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
        da[k] = v; dz_hi[(size_t)b * D + k] = hi; dz_lo[(size_t)b * D + k] = v - hi;
    }
}

// out[j] = sum_b M[b][j]  (fixed order): one CTA per 32 columns
__global__ void __launch_bounds__(256) k_colsum_rows(const float* __restrict__ M, int B, int ld, int ncols, float* __restrict__ out) {
    __shared__ double s[8][32];
    const int j = blockIdx.x * 32 + (threadIdx.x & 31), rl = threadIdx.x >> 5;
    double a = 0.0;
    if (j < ncols) for (int b = rl; b < B; b += 8) a += (double)M[(size_t)b * ld + j];
    s[rl][threadIdx.x & 31] = a;
    __syncthreads();
    if (rl == 0 && j < ncols) {
        double t = 0.0;
        for (int r = 0; r < 8; ++r) t += s[r][threadIdx.x];
        out[j] = (float)t;
    }
}

// ReLU / BatchNorm backward inputs from dh: dahat = dh [pre > 0] gamma (for k_wgrad_tc), dgamma = sum dpre ahat, dbeta = sum dpre
// (fixed-point sums for k_wgrad_tc, fp32 into the flat gradient), gb1 = 0 (vanishes identically under BatchNorm)
__global__ void __launch_bounds__(256) k_ca_bn_bwd(const float* __restrict__ dh, const float* __restrict__ A, const fx64* __restrict__ colsum,
                                                   const float* __restrict__ gamma, const float* __restrict__ beta, int B, int H,
                                                   float* __restrict__ dahat, fx64* __restrict__ dgb, float* __restrict__ g_b1,
                                                   float* __restrict__ g_gamma, float* __restrict__ g_beta) {
    __shared__ double s1[8][32], s2[8][32];
    const int j = blockIdx.x * 32 + (threadIdx.x & 31), rl = threadIdx.x >> 5;
    double dg = 0.0, db = 0.0;
    if (j < H) {
        const double m = fx_get<FX_COLSUM>(&colsum[j]) / (double)B;
        double v = fx_get<FX_COLSUM>(&colsum[H + j]) / (double)B - m * m;
        if (v < 0.0) v = 0.0;
        const float mu = (float)m, rstd = 1.0f / sqrtf((float)v + DBMM_BN_EPS), ga = gamma[j], be = beta[j];
        for (int b = rl; b < B; b += 8) {
            const float ah = (A[(size_t)b * H + j] - mu) * rstd;
            const float pre = fmaf(ah, ga, be);
            const float dpre = pre > 0.f ? dh[(size_t)b * H + j] : 0.f;
            dahat[(size_t)b * H + j] = dpre * ga;
            dg += (double)dpre * (double)ah; db += (double)dpre;
        }
    }
    s1[rl][threadIdx.x & 31] = dg; s2[rl][threadIdx.x & 31] = db;
    __syncthreads();
    if (rl == 0 && j < H) {
        double t1 = 0.0, t2 = 0.0;
        for (int r = 0; r < 8; ++r) { t1 += s1[r][threadIdx.x]; t2 += s2[r][threadIdx.x]; }
        dgb[j].v = (long long)fx_bits<FX_DGB>(t1);
        dgb[H + j].v = (long long)fx_bits<FX_DGB>(t2);
        g_gamma[j] = (float)t1; g_beta[j] = (float)t2; g_b1[j] = 0.f;
    }
}

// gW1 = sum over the batch chunks of k_wgrad_tc's partial tiles, in chunk order
__global__ void __launch_bounds__(256) k_sum_chunks(const float* __restrict__ part, int nchunk, int64_t n4, float* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 a = __ldcg(reinterpret_cast<const float4*>(part) + i);
        for (int c = 1; c < nchunk; ++c) {
            const float4 p = __ldcg(reinterpret_cast<const float4*>(part) + (size_t)c * n4 + i);
            a.x += p.x; a.y += p.y; a.z += p.z; a.w += p.w;
        }
        reinterpret_cast<float4*>(out)[i] = a;
    }
}

// loss_out += weight * loss_sum / max(n_valid, 1); n_out = n_valid   (device scalars)
__global__ void k_ca_loss_out(const double* loss_sum, const int* n_valid, float weight, double* loss_out, int* n_out) {
    if (loss_out) *loss_out += (double)weight * loss_sum[0] / (double)max(n_valid[0], 1);
    if (n_out) *n_out = n_valid[0];
}

}  // namespace dbmm
