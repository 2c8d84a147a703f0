// Adapted-embedding export (demo/demo_visualization.ipynb:1117-1215, validate_adapter_with_return):
// the D-wide features the H-space evaluation never materialises, written out for the visualisation notebooks together
// with the notebook's two logit sets (features @ That / tau for the class and the spurious prompts).
//   single adapter:    out = z = W2 relu(BN_running(a)) + b2            (UN-normalised, as the notebook does for "adapter")
//   MultipleAdapter:   out = w * z_old / |z_old| + (1 - w) * z_new / |z_new|
// Input: a = x W1^T + b1 from the GEMM-1 kernels ([nad][B][H]).  A CTA owns ER rows; thread t owns the output columns
// d = t, t + 256, ...; W2 rows come straight from L2 as 16-byte vectors.  Not a hot path: fp32 SIMT, no tensor cores.
#pragma once
#include "common.cuh"

namespace dbmm {

constexpr int EX_ROWS = 8, EX_THREADS = 256, EX_MAXDPT = 8;        // D <= EX_THREADS * EX_MAXDPT = 2048

struct ExportArgs {
    int B, D, H, nad;
    const float* A; int64_t strideA;          // [nad][B][H]
    AdapterView ad[2];                        // [0] = old (or the only) adapter, [1] = trainable adapter
    float w_old; int normalize_single;
    const float* That_a; int Ca; const float* That_b; int Cb; float inv_tau;      // prompt sets [D][C], column-normalised; may be NULL
    float* out; int64_t ld_out; int64_t pos0;                                      // rows pos0 .. pos0 + B of the outputs
    float* logits_a; float* logits_b;                                              // [N][Ca], [N][Cb] or NULL
};

__global__ void __launch_bounds__(EX_THREADS) k_export_rows(ExportArgs a) {
    extern __shared__ __align__(16) float ex_smem[];
    const int H = a.H, D = a.D, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float* sH = ex_smem;                                  // [EX_ROWS][H]
    float* sRed = sH + EX_ROWS * H;                       // [EX_ROWS][8 warps]
    float* sLog = sRed + EX_ROWS * 8;                     // [EX_ROWS][32]
    const int r0 = blockIdx.x * EX_ROWS;
    const int ndpt = (D + EX_THREADS - 1) / EX_THREADS;
    float outv[EX_MAXDPT][EX_ROWS];
#pragma unroll
    for (int j = 0; j < EX_MAXDPT; ++j)
#pragma unroll
        for (int r = 0; r < EX_ROWS; ++r) outv[j][r] = 0.f;

    for (int ad = 0; ad < a.nad; ++ad) {
        const AdapterView& v = a.ad[a.nad == 2 ? ad : 1];
        __syncthreads();
        for (int e = tid; e < EX_ROWS * H; e += EX_THREADS) {         // h = relu(gamma * (a - rm) / sqrt(rv + eps) + beta)
            const int r = e / H, k = e - r * H;
            float h = 0.f;
            if (r0 + r < a.B) {
                const float av = a.A[(size_t)ad * a.strideA + (size_t)(r0 + r) * H + k];
                const float pre = v.gamma[k] * ((av - v.running_mean[k]) * (1.0f / sqrtf(v.running_var[k] + DBMM_BN_EPS))) + v.beta[k];
                h = fmaxf(pre, 0.f);
            }
            sH[e] = h;
        }
        __syncthreads();
        float z[EX_MAXDPT][EX_ROWS];
        float n2[EX_ROWS];
#pragma unroll
        for (int r = 0; r < EX_ROWS; ++r) n2[r] = 0.f;
#pragma unroll
        for (int j = 0; j < EX_MAXDPT; ++j) {
            const int d = tid + j * EX_THREADS;
#pragma unroll
            for (int r = 0; r < EX_ROWS; ++r) z[j][r] = 0.f;
            if (j < ndpt && d < D) {
                const float bias = v.b2[d];
                const float4* wrow = reinterpret_cast<const float4*>(v.W2 + (size_t)d * H);       // H % 4 == 0
                for (int k4 = 0; k4 < (H >> 2); ++k4) {
                    const float4 w4 = __ldg(wrow + k4);
#pragma unroll
                    for (int r = 0; r < EX_ROWS; ++r) {
                        const float4 h4 = *reinterpret_cast<const float4*>(sH + r * H + k4 * 4);
                        z[j][r] = fmaf(w4.x, h4.x, fmaf(w4.y, h4.y, fmaf(w4.z, h4.z, fmaf(w4.w, h4.w, z[j][r]))));
                    }
                }
#pragma unroll
                for (int r = 0; r < EX_ROWS; ++r) { z[j][r] += bias; n2[r] = fmaf(z[j][r], z[j][r], n2[r]); }
            }
        }
        float scale[EX_ROWS];
        const bool need_norm = a.nad == 2 || a.normalize_single;
        if (need_norm) {                                             // row norms: warp shuffle, then across the 8 warps
#pragma unroll
            for (int r = 0; r < EX_ROWS; ++r) {
                const float s = warp_sum(n2[r]);
                if (lane == 0) sRed[r * 8 + warp] = s;
            }
            __syncthreads();
#pragma unroll
            for (int r = 0; r < EX_ROWS; ++r) {
                float s = 0.f;
#pragma unroll
                for (int w = 0; w < 8; ++w) s += sRed[r * 8 + w];
                scale[r] = 1.0f / sqrtf(s);                           // no epsilon in the reference (final_main.py:68)
            }
        }
        const float mixw = a.nad == 2 ? (ad == 0 ? a.w_old : 1.0f - a.w_old) : 1.0f;
#pragma unroll
        for (int j = 0; j < EX_MAXDPT; ++j)
#pragma unroll
            for (int r = 0; r < EX_ROWS; ++r) outv[j][r] = fmaf(mixw * (need_norm ? scale[r] : 1.0f), z[j][r], outv[j][r]);
    }
    // ---- features out (coalesced: consecutive threads own consecutive columns)
#pragma unroll
    for (int j = 0; j < EX_MAXDPT; ++j) {
        const int d = tid + j * EX_THREADS;
        if (j < ndpt && d < D)
#pragma unroll
            for (int r = 0; r < EX_ROWS; ++r)
                if (r0 + r < a.B) a.out[(size_t)(a.pos0 + r0 + r) * a.ld_out + d] = outv[j][r];
    }
    // ---- logits = features @ That / tau for up to two prompt sets (Ca + Cb <= 32)
    const int Ct = (a.logits_a ? a.Ca : 0) + (a.logits_b ? a.Cb : 0);
    if (Ct == 0) return;
    __syncthreads();
    for (int e = tid; e < EX_ROWS * 32; e += EX_THREADS) sLog[e] = 0.f;
    __syncthreads();
    for (int c = 0; c < Ct; ++c) {
        const bool first = a.logits_a && c < a.Ca;
        const float* T = first ? a.That_a : a.That_b;
        const int Cn = first ? a.Ca : a.Cb, cc = first ? c : c - (a.logits_a ? a.Ca : 0);
        float part[EX_ROWS];
#pragma unroll
        for (int r = 0; r < EX_ROWS; ++r) part[r] = 0.f;
#pragma unroll
        for (int j = 0; j < EX_MAXDPT; ++j) {
            const int d = tid + j * EX_THREADS;
            if (j < ndpt && d < D) {
                const float tv = __ldg(T + (size_t)d * Cn + cc);
#pragma unroll
                for (int r = 0; r < EX_ROWS; ++r) part[r] = fmaf(outv[j][r], tv, part[r]);
            }
        }
#pragma unroll
        for (int r = 0; r < EX_ROWS; ++r) {
            const float s = warp_sum(part[r]);
            if (lane == 0) atomicAdd(&sLog[r * 32 + c], s);
        }
    }
    __syncthreads();
    for (int e = tid; e < EX_ROWS * Ct; e += EX_THREADS) {
        const int r = e / Ct, c = e - r * Ct;
        if (r0 + r >= a.B) continue;
        const bool first = a.logits_a && c < a.Ca;
        const float val = sLog[r * 32 + c] * a.inv_tau;
        if (first) a.logits_a[(size_t)(a.pos0 + r0 + r) * a.Ca + c] = val;
        else a.logits_b[(size_t)(a.pos0 + r0 + r) * a.Cb + (c - (a.logits_a ? a.Ca : 0))] = val;
    }
}

static int launch_export_rows(const ExportArgs& a, cudaStream_t st) {
    DBMM_CHECK_SHAPE(a.H % 4 == 0 && a.D <= EX_THREADS * EX_MAXDPT, "export kernel: unsupported D=%d H=%d", a.D, a.H);
    DBMM_CHECK_SHAPE((a.logits_a ? a.Ca : 0) + (a.logits_b ? a.Cb : 0) <= 32, "export kernel: more than 32 prompt columns");
    const size_t smem = sizeof(float) * ((size_t)EX_ROWS * a.H + EX_ROWS * 8 + EX_ROWS * 32);
    k_export_rows<<<ceil_div(a.B, EX_ROWS), EX_THREADS, smem, st>>>(a);
    DBMM_LAUNCH_CHECK();
    return DBMM_OK;
}

}  // namespace dbmm
