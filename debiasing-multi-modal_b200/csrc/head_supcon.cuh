// Row kernels around the tensor-core NT GEMM (tc_gemm.cuh):
//   * zero-shot head (config 4; validate_zs on raw embeddings, final_main.py:757-768): row L2 norms, and the combine of
//     the per-column-tile online-softmax partials into loss / argmax / per-group counters (update_dict, 383-391);
//   * contrastive regulariser (config 3; formula of demo/visualizer_supcon.py:1532-1571 applied to every anchor of the
//     batch at once): per-anchor masked log-sum-exp over the similarity row, loss, and the similarity gradient G.
#pragma once
#include "tc_gemm.cuh"

namespace dbmm {

// inv_norm[r] = 1 / ||X[row(r)]||_2     (one warp per row, 16-byte loads)
__global__ void __launch_bounds__(256) k_row_inv_norm(const float* __restrict__ X, int64_t ldx, const int32_t* __restrict__ idx,
                                                      int64_t pos0, int64_t n, int D, float* __restrict__ inv_norm) {
    const int lane = threadIdx.x & 31;
    const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp_global; r < n; r += nwarps) {
        const int64_t row = idx ? (int64_t)idx[pos0 + r] : (pos0 + r);
        const float4* p = reinterpret_cast<const float4*>(X + row * ldx);
        float s = 0.f;
        for (int c = lane; c < D / 4; c += 32) {
            const float4 v = __ldg(p + c);
            s = fmaf(v.x, v.x, s); s = fmaf(v.y, v.y, s); s = fmaf(v.z, v.z, s); s = fmaf(v.w, v.w, s);
        }
        s = warp_sum(s);
        if (lane == 0) inv_norm[r] = 1.0f / sqrtf(s);
    }
}

// Gather rows through an index list into a dense [n, D] matrix (TMA needs a regular tensor).
__global__ void __launch_bounds__(256) k_gather_rows(const float* __restrict__ X, int64_t ldx, const int32_t* __restrict__ idx,
                                                     int64_t pos0, int64_t n, int D, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp_global; r < n; r += nwarps) {
        const float4* p = reinterpret_cast<const float4*>(X + (int64_t)idx[pos0 + r] * ldx);
        float4* o = reinterpret_cast<float4*>(out + r * D);
        for (int c = lane; c < D / 4; c += 32) o[c] = __ldg(p + c);
    }
}

// Combine the column-tile partials of every row: nll = log sum exp - target logit, argmax (first maximum), counters.
__global__ void __launch_bounds__(256) k_head_finish(const SoftmaxPart* __restrict__ part, int ntile, int64_t n, int64_t pos0,
                                                     const int32_t* __restrict__ idx, const int32_t* __restrict__ grp, int G,
                                                     int64_t batch_size, double* loss_sum, int64_t* counts, int32_t* pred_out,
                                                     const int32_t* __restrict__ y) {
    const int lane = threadIdx.x & 31;
    const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t base = warp_global * 32; base < n; base += nwarps * 32) {
        const int64_t r = base + lane;
        const bool valid = r < n;
        int gv = -1, corr = 0; float nll = 0.f;
        if (valid) {
            float mx = -INFINITY, se = 0.f, ly = -INFINITY; int am = 0;
            for (int t = 0; t < ntile; ++t) {
                const SoftmaxPart p = part[(size_t)t * n + r];                 // [column tile][row of this launch]
                if (p.mx > mx) { se = se * expf(mx - p.mx) + p.se; mx = p.mx; am = p.am; }
                else se += p.se * expf(p.mx - mx);
                ly = fmaxf(ly, p.ly);
            }
            const int64_t pos = pos0 + r;
            const int64_t dsrow = idx ? (int64_t)idx[pos] : pos;
            const int yv = y ? y[dsrow] : -1;
            nll = y ? (logf(se) + mx - ly) : 0.f;
            corr = am == yv;
            gv = grp ? grp[dsrow] : 0;
            if (pred_out) pred_out[pos] = am;
        }
        const int64_t slot = valid ? (pos0 + r) / batch_size : -1;
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

// ------------------------------------------------------------------------------------------------
// Contrastive rows.  S[i][j] = z_i . z_j / tau_cl for local anchor i (global index row0 + i) against all Bg rows.
//   P_i = {j != self : label_j == label_i},  N_i = {j : label_j != label_i};  anchor valid iff both non-empty
//   loss_i = log sum_{j != self} exp(S_ij)  -  mean_{j in P_i} S_ij
//   G_ij   = softmax_{j != self}(S_i)_j - [j in P_i] / |P_i|          (0 for j = self and for invalid anchors)
// G overwrites S; the 1 / n_valid factor (a global count) is applied by the gradient GEMMs through scale_dev.
// One CTA of 256 threads per anchor row.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_supcon_rows(float* __restrict__ S, int64_t lds, int Bl, int Bg, int64_t row0,
                                                     const int32_t* __restrict__ labels, double* loss_sum, int* n_valid,
                                                     float* __restrict__ row_loss, unsigned* gmax_bits = nullptr) {
    __shared__ float red[8];
    __shared__ int redi[8];
    const int i = blockIdx.x;
    if (i >= Bl) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t self = row0 + i;
    const int li = labels[self];
    float* row = S + (size_t)i * lds;
    auto block_max = [&](float v) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
        if (lane == 0) red[warp] = v;
        __syncthreads();
        float r = red[0];
#pragma unroll
        for (int w = 1; w < 8; ++w) r = fmaxf(r, red[w]);
        __syncthreads();
        return r;
    };
    auto block_sum = [&](float v) {
        v = warp_sum(v);
        if (lane == 0) red[warp] = v;
        __syncthreads();
        float r = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) r += red[w];
        __syncthreads();
        return r;
    };
    float mx = -INFINITY, psum = 0.f; int npos = 0;
    for (int j = tid; j < Bg; j += 256) {
        if (j == self) continue;
        const float s = row[j];
        mx = fmaxf(mx, s);
        if (labels[j] == li) { psum += s; ++npos; }
    }
    mx = block_max(mx);
    psum = block_sum(psum);
    {
        int c = npos;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
        if (lane == 0) redi[warp] = c;
        __syncthreads();
        npos = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) npos += redi[w];
        __syncthreads();
    }
    const int nneg = Bg - 1 - npos;
    const bool valid = npos > 0 && nneg > 0;
    float se = 0.f;
    for (int j = tid; j < Bg; j += 256)
        if (j != self) se += expf(row[j] - mx);
    se = block_sum(se);
    const float inv_se = valid ? 1.0f / se : 0.f, inv_np = valid ? 1.0f / (float)npos : 0.f;
    float gm = 0.f;
    for (int j = tid; j < Bg; j += 256) {
        float g = 0.f;
        if (valid && j != self) g = expf(row[j] - mx) * inv_se - (labels[j] == li ? inv_np : 0.f);
        row[j] = g;
        gm = fmaxf(gm, fabsf(g));
    }
    if (gmax_bits) {                                   // largest |G| (bit pattern of a non-negative float): scale of the fp16 pair split
        gm = block_max(gm);
        if (tid == 0 && gm > 0.f) atomicMax(gmax_bits, __float_as_uint(gm));
    }
    if (tid == 0) {
        const float l = valid ? (logf(se) + mx - psum * inv_np) : 0.f;
        if (row_loss) row_loss[i] = l;
        if (valid) { atomicAdd(loss_sum, (double)l); atomicAdd(n_valid, 1); }
    }
}

// scale[0] = factor / max(n_valid, 1)
__global__ void k_supcon_scale(const int* n_valid, float factor, float* scale) { scale[0] = factor / (float)max(n_valid[0], 1); }

}  // namespace dbmm

namespace dbmm {

// ------------------------------------------------------------------------------------------------
// Eval forward, tensor-core path (validate / validate_zs through the adapter): finishing kernel.
//   n^2 = rowdot + tail[0] (= t[H]);  s_c = tail[1 + c];  logit_c = s_c / (n tau)  (two adapters: 0.5 / 0.5 mix of the two
//   normalised outputs, final_main.py:121-140), then CE / argmax / per-group counters as update_dict (final_main.py:383-391).
// ------------------------------------------------------------------------------------------------
struct EvalFinishArgs {
    int64_t n, pos0; int nad, C, G; int tail_ld;
    const float* rowdot[2]; const float* tail[2];
    float w_old, inv_tau;
    const int32_t* idx; const int32_t* y; const int32_t* grp;
    int64_t batch_size; double* loss_sum; int64_t* counts; float* logits_out; int32_t* pred_out;
};

__global__ void __launch_bounds__(256) k_eval_finish(EvalFinishArgs a) {
    const int lane = threadIdx.x & 31;
    const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t base = warp_global * 32; base < a.n; base += nwarps * 32) {
        const int64_t r = base + lane;
        const bool valid = r < a.n;
        int gv = -1, corr = 0; float nll = 0.f;
        if (valid) {
            const int64_t pos = a.pos0 + r;
            const int64_t dsrow = a.idx ? (int64_t)a.idx[pos] : pos;
            const int yv = a.y ? a.y[dsrow] : -1;
            gv = a.grp ? a.grp[dsrow] : 0;
            float l[DBMM_MAX_C];
            float mx = -INFINITY; int am = 0;
            const float inv_n1 = 1.0f / sqrtf(a.rowdot[a.nad - 1][r] + a.tail[a.nad - 1][(size_t)r * a.tail_ld]);
            const float inv_n0 = a.nad == 2 ? 1.0f / sqrtf(a.rowdot[0][r] + a.tail[0][(size_t)r * a.tail_ld]) : 0.f;
            const float coef = a.nad == 2 ? (1.0f - a.w_old) : 1.0f;
            for (int c = 0; c < a.C; ++c) {
                const float lnew = a.inv_tau * a.tail[a.nad - 1][(size_t)r * a.tail_ld + 1 + c] * inv_n1;
                float v = lnew;
                if (a.nad == 2) v = fmaf(coef, lnew, a.w_old * a.inv_tau * a.tail[0][(size_t)r * a.tail_ld + 1 + c] * inv_n0);
                l[c] = v;
                if (v > mx) { mx = v; am = c; }
                if (a.logits_out) a.logits_out[(size_t)pos * a.C + c] = v;
            }
            float se = 0.f, ly = 0.f;
            for (int c = 0; c < a.C; ++c) { se += expf(l[c] - mx); if (c == yv) ly = l[c]; }
            nll = a.y ? (logf(se) + mx - ly) : 0.f;
            corr = am == yv;
            if (a.pred_out) a.pred_out[pos] = am;
        }
        const int64_t slot = valid ? (a.pos0 + r) / a.batch_size : -1;
        const int64_t slot0 = __shfl_sync(0xffffffffu, slot, 0);
        const bool uniform = __all_sync(0xffffffffu, !valid || slot == slot0);
        if (uniform) {
            const double tot = warp_sum((double)nll);
            if (lane == 0 && a.loss_sum) atomicAdd(&a.loss_sum[slot0], tot);
            const unsigned cmask = __ballot_sync(0xffffffffu, corr != 0);
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
            if (a.loss_sum) atomicAdd(&a.loss_sum[slot], (double)nll);
            if (a.counts && gv >= 0 && gv < a.G) {
                int64_t* cnt = a.counts + (size_t)slot * 2 * a.G;
                if (corr) atomicAdd((unsigned long long*)&cnt[gv], 1ull);
                atomicAdd((unsigned long long*)&cnt[a.G + gv], 1ull);
            }
        }
    }
}

// B operand of the H-space GEMM from the Gram matrix G [(H+1)][ldg]:  Bt[n][k] = G[k][n] (k < H), split hi + lo;
// bias[n] = G[H][n] (the row that multiplies the constant 1 of [h, 1]).
__global__ void __launch_bounds__(256) k_gram_operand(const float* __restrict__ G, int H, int ldg, float* __restrict__ bt_hi,
                                                      float* __restrict__ bt_lo, float* __restrict__ bias) {
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < ldg * H; e += gridDim.x * blockDim.x) {
        const int n = e / H, k = e - n * H;
        const float v = G[(size_t)k * ldg + n];
        const float h = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
        bt_hi[e] = h; bt_lo[e] = v - h;
    }
    for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < ldg; n += gridDim.x * blockDim.x) bias[n] = G[(size_t)H * ldg + n];
}

}  // namespace dbmm
