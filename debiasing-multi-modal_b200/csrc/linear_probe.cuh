// Linear probing (LinearClassifier, final_main.py:43-49 with --tl_method linear_probing): logits = x W^T + b on the raw
// embeddings, softmax-CE, gradients dW = dl^T X / db = sum dl, per-group counters -- one kernel per step; the SGD
// update reuses k_sgd_flat.  C <= 16 classes; not a tensor-core shape (C = 2 in the reference).
#pragma once
#include "common.cuh"

namespace dbmm {

constexpr int LP_ROWS = 8, LP_THREADS = 256, LP_MAXC = 16;

struct LinearStepArgs {
    const float* X; int64_t ldx; const int32_t* idx; const int32_t* y; const int32_t* grp;
    int B, D, C, G;
    const float* W; const float* b;        // [C, D], [C]
    float* gW; float* gb;                  // zeroed by the caller; += dl^T X, += sum dl
    double* loss_sum; int64_t* counts; int64_t slot;
};

__global__ void __launch_bounds__(LP_THREADS) k_linear_train(LinearStepArgs a) {
    __shared__ float sDl[LP_ROWS][LP_MAXC];
    __shared__ int64_t sRow[LP_ROWS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int D = a.D, C = a.C;
    const int r = blockIdx.x * LP_ROWS + warp;
    const bool valid = r < a.B;
    int gv = -1, corr = 0; float nll = 0.f;
    int64_t row = 0;
    if (valid) {
        row = a.idx ? (int64_t)a.idx[r] : (int64_t)r;
        const float* x = a.X + row * a.ldx;
        float l[LP_MAXC];
#pragma unroll
        for (int c = 0; c < LP_MAXC; ++c) l[c] = 0.f;
        for (int d = lane; d < D; d += 32) {
            const float xv = __ldg(x + d);
#pragma unroll
            for (int c = 0; c < LP_MAXC; ++c)
                if (c < C) l[c] = fmaf(xv, __ldg(a.W + (size_t)c * D + d), l[c]);
        }
        float mx = -INFINITY; int am = 0;
#pragma unroll
        for (int c = 0; c < LP_MAXC; ++c)
            if (c < C) { l[c] = warp_sum(l[c]) + a.b[c]; if (l[c] > mx) { mx = l[c]; am = c; } }
        const int yv = a.y[row];
        gv = a.grp ? a.grp[row] : 0;
        float se = 0.f, ly = 0.f;
#pragma unroll
        for (int c = 0; c < LP_MAXC; ++c)
            if (c < C) { se += expf(l[c] - mx); if (c == yv) ly = l[c]; }
        nll = logf(se) + mx - ly;
        corr = am == yv;
        if (lane < C) {
            float lv = 0.f;
#pragma unroll
            for (int c = 0; c < LP_MAXC; ++c) if (c == lane) lv = l[c];
            sDl[warp][lane] = (expf(lv - mx) / se - (lane == yv ? 1.f : 0.f)) / (float)a.B;
        }
    } else if (lane < C) sDl[warp][lane] = 0.f;
    if (lane == 0) sRow[warp] = valid ? row : -1;
    __syncthreads();
    // dW[c][d] += sum over this CTA's rows of dl[r][c] x[r][d]   (threads own columns d, one red per (c, d) per CTA)
    for (int d = tid; d < D; d += LP_THREADS) {
        float acc[LP_MAXC];
#pragma unroll
        for (int c = 0; c < LP_MAXC; ++c) acc[c] = 0.f;
        for (int rr = 0; rr < LP_ROWS; ++rr) {
            if (sRow[rr] < 0) continue;
            const float xv = __ldg(a.X + sRow[rr] * a.ldx + d);
#pragma unroll
            for (int c = 0; c < LP_MAXC; ++c)
                if (c < C) acc[c] = fmaf(sDl[rr][c], xv, acc[c]);
        }
#pragma unroll
        for (int c = 0; c < LP_MAXC; ++c)
            if (c < C) atomicAdd(a.gW + (size_t)c * D + d, acc[c]);
    }
    if (tid < C) {
        float s = 0.f;
        for (int rr = 0; rr < LP_ROWS; ++rr) s += sDl[rr][tid];
        atomicAdd(a.gb + tid, s);
    }
    // loss / group counters of the CTA's rows (update_dict, final_main.py:383-391)
    __shared__ float sN[LP_ROWS]; __shared__ int sG[LP_ROWS], sC[LP_ROWS];
    if (lane == 0) { sN[warp] = nll; sG[warp] = gv; sC[warp] = corr; }
    __syncthreads();
    if (warp == 0) {
        const int g2 = lane < LP_ROWS ? sG[lane] : -1, c2 = lane < LP_ROWS ? sC[lane] : 0;
        const double tot = warp_sum((double)(lane < LP_ROWS ? sN[lane] : 0.f));
        if (lane == 0 && a.loss_sum) atomicAdd(&a.loss_sum[a.slot], tot);
        const unsigned cmask = __ballot_sync(0xffffffffu, c2 != 0);
        for (int g = 0; g < a.G; ++g) {
            const unsigned gm = __ballot_sync(0xffffffffu, g2 == g);
            if (lane == 0 && gm && a.counts) {
                int64_t* cnt = a.counts + (size_t)a.slot * 2 * a.G;
                const int nc = __popc(gm & cmask);
                if (nc) atomicAdd((unsigned long long*)&cnt[g], (unsigned long long)nc);
                atomicAdd((unsigned long long*)&cnt[a.G + g], (unsigned long long)__popc(gm));
            }
        }
    }
}

}  // namespace dbmm
