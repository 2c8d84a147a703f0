// Fused tail of a single-GPU training step: gradient finalisation AND optimizer step in one kernel, split into two
// roles that the epoch graph launches on two streams so that the W2 half overlaps the tensor-core dW1 GEMM.
//
//   role bit 0 (after k_wgrad_tc):  dW1 = sum of the batch-chunk partial tiles -> SGD on W1 -> tf32 split of the new W1
//                                   (everything elementwise: one pass, 16 partial loads + p + v in flight per thread)
//   role bit 1 (after k_rows_train, concurrent with k_wgrad_tc):
//       W2 CTAs (64 embedding rows each, 512 threads: D / 64 = 16 CTAs fit the SMs the 128-CTA tensor-core kernels of
//       the main branch leave free -- those kernels fill a whole SM's shared memory, so a co-scheduled W2 CTA would
//       push them into a second wave): dW2a = [W2 | b2 | That] S -> SGD on W2 / b2 -> the slice's share of NEXT step's
//       Gram matrix G = [W2 | b2]^T [W2 | b2 | That] (coalesced fp32 reds into the other half of a double buffer; this
//       step's half, consumed by k_rows_train, is re-zeroed here).
//   One more CTA of role bit 0 (k_wgrad_tc reads gamma): dgamma / dbeta / db1 -> SGD on b1 / gamma / beta, BatchNorm
//       running statistics (final_main.py:122,574: the frozen adapter drifts too).
// The per-step accumulators are re-zeroed by the NEXT step's head kernels (column sums: k_gemm1_tc; dgamma / dbeta and S:
// k_reduce_stats), never here: k_wgrad_tc reads them concurrently.  SGD as torch.optim.SGD (demo/util.py:118-136):
// g += wd*p; v = momentum*v + g; p -= lr*v (the caller zeroes v before the optimizer's first step, which makes v = g).
// Data-parallel epochs keep k_finalize_grads + k_update (the gradient all-reduce sits between them).
#pragma once
#include "common.cuh"
#include "ptx_sm100.cuh"

namespace dbmm {

constexpr int ST_THREADS = 256, ST_MAXCHUNK = 16;                 // W1 role
constexpr int ST2_THREADS = 512, ST2_ROWS = 64, ST_NSLOT = 5;     // W2 role
constexpr int ST2_LP = 32 * ST_NSLOT;                             // row stride of the staged [W2 | b2 | That] rows (zero padded)

struct StepTailArgs {
    int roles;
    float* W1; float* b1; float* gamma; float* beta; float* W2; float* b2;    // trainable adapter (updated in place)
    float* g; float* v;                       // flat gradient (written: callers may read it back) / momentum
    const float* lr_dev; float lr; float momentum, wd;
    const float* part; int nchunk;            // [nchunk][H][D] batch-chunk partials of dW1
    float* whi; float* wlo;                   // tf32 split of the new W1
    const float* That; const float* S;        // [D][C]; [H+1+C][s_stride(H)]
    float* gram_next; float* gram_zero;       // trainable adapter's Gram matrix: next step's (+=) and this step's (reset)
    const double* dgb; const double* colsum;  // [2][H]; [nad][2][H]
    int D, H, C, nad; int64_t Bg;
    float* rm[2]; float* rv[2]; long long* nbt[2];
    int n_w1_ctas, n_w2_ctas;
};

static inline size_t step_tail_smem_bytes(int H, int C) {
    const size_t KP = (H + 1 + C + 3) & ~3, NP = (H + 1 + 3) & ~3;
    return sizeof(float) * (KP * NP + (size_t)ST2_ROWS * ST2_LP) + 16;
}

// ---- role bit 0: W1 CTAs + one chores CTA
__global__ void __launch_bounds__(ST_THREADS) k_tail_w1(StepTailArgs a) {
    const int H = a.H, D = a.D;
    const int tid = threadIdx.x;
    const float lr = a.lr_dev ? __ldg(a.lr_dev) : a.lr;
    const size_t oW1 = 0, ob1 = (size_t)H * D;
    const int bid = blockIdx.x;
    if (bid < a.n_w1_ctas) {
        // chunk sum + SGD + tf32 split, 16-byte accesses (H * D is a multiple of 4)
        const int64_t n4 = (int64_t)H * D / 4;
        const size_t plane4 = (size_t)H * D / 4;
        for (int64_t i = (int64_t)bid * ST_THREADS + tid; i < n4; i += (int64_t)a.n_w1_ctas * ST_THREADS) {
            float4 pt[ST_MAXCHUNK];
#pragma unroll
            for (int c = 0; c < ST_MAXCHUNK; ++c)
                pt[c] = c < a.nchunk ? __ldcg(reinterpret_cast<const float4*>(a.part) + c * plane4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 pv = reinterpret_cast<const float4*>(a.W1)[i];
            const float4 vv = reinterpret_cast<const float4*>(a.v + oW1)[i];
            float4 gs = pt[0];
#pragma unroll
            for (int c = 1; c < ST_MAXCHUNK; ++c) { gs.x += pt[c].x; gs.y += pt[c].y; gs.z += pt[c].z; gs.w += pt[c].w; }
            const float px[4] = {pv.x, pv.y, pv.z, pv.w}, gx[4] = {gs.x, gs.y, gs.z, gs.w}, vx[4] = {vv.x, vv.y, vv.z, vv.w};
            float po[4], vo[4], hi[4], lo[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float g = gx[q] + a.wd * px[q];
                vo[q] = a.momentum * vx[q] + g;
                po[q] = px[q] - lr * vo[q];
                hi[q] = __uint_as_float(__float_as_uint(po[q]) & 0xffffe000u);
                lo[q] = po[q] - hi[q];
            }
            reinterpret_cast<float4*>(a.v + oW1)[i] = make_float4(vo[0], vo[1], vo[2], vo[3]);
            reinterpret_cast<float4*>(a.W1)[i] = make_float4(po[0], po[1], po[2], po[3]);
            reinterpret_cast<float4*>(a.whi)[i] = make_float4(hi[0], hi[1], hi[2], hi[3]);
            reinterpret_cast<float4*>(a.wlo)[i] = make_float4(lo[0], lo[1], lo[2], lo[3]);
            reinterpret_cast<float4*>(a.g + oW1)[i] = gs;
        }
        return;
    }
    // ---- chores CTA (this role because k_wgrad_tc reads gamma and must have finished): b1 / gamma / beta (db1 = sum_B da
    // vanishes identically under BatchNorm: b1 moves by weight decay only, see k_finalize_grads), BatchNorm running
    // statistics of every adapter in the forward
    for (int e = tid; e < 3 * H; e += ST_THREADS) {
        const int seg = e / H, j = e - seg * H;
        float* pp = (seg == 0 ? a.b1 : (seg == 1 ? a.gamma : a.beta)) + j;
        const size_t fo = ob1 + e;
        const float graw = seg == 0 ? 0.f : (float)a.dgb[(size_t)(seg - 1) * H + j];
        const float pv = *pp;
        const float g = graw + a.wd * pv;
        const float vn = a.momentum * a.v[fo] + g;
        a.g[fo] = graw;
        a.v[fo] = vn;
        *pp = pv - lr * vn;
    }
    for (int e = tid; e < a.nad * H; e += ST_THREADS) {
        const int ad = e / H, j = e - ad * H;
        const double m = a.colsum[((size_t)ad * 2 + 0) * H + j] / (double)a.Bg;
        double var = a.colsum[((size_t)ad * 2 + 1) * H + j] / (double)a.Bg - m * m;
        if (var < 0.0) var = 0.0;
        const float unbiased = (float)(var * (double)a.Bg / (double)(a.Bg - 1));
        a.rm[ad][j] = (1.0f - DBMM_BN_MOMENTUM) * a.rm[ad][j] + DBMM_BN_MOMENTUM * (float)m;
        a.rv[ad][j] = (1.0f - DBMM_BN_MOMENTUM) * a.rv[ad][j] + DBMM_BN_MOMENTUM * unbiased;
    }
    if (tid < a.nad) *a.nbt[tid] += 1;
}

// ---- role bit 1: W2 / b2 rows [d0, d0 + ST2_ROWS):  dW2a[d][n] = sum_k L[d][k] S[k][n],  L = [W2 | b2 | That]
__global__ void __launch_bounds__(ST2_THREADS) k_tail_w2(StepTailArgs a) {
    extern __shared__ __align__(16) float st_smem[];
    const int H = a.H, D = a.D, C = a.C;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float lr = a.lr_dev ? __ldg(a.lr_dev) : a.lr;
    const size_t oW2 = (size_t)H * D + 3 * (size_t)H, ob2 = oW2 + (size_t)D * H;
    const int K = H + 1 + C, N = H + 1, KP = (K + 3) & ~3, NP = (N + 3) & ~3;
    const int w2 = blockIdx.x;
    {   // this step's Gram matrix has been consumed by k_rows_train: reset it for the step after next
        const int nz = N * K;
        for (int e = w2 * ST2_THREADS + tid; e < nz; e += (int)gridDim.x * ST2_THREADS) a.gram_zero[e] = 0.f;
    }
    float* sS = st_smem;                        // [KP][NP]
    float* sL = sS + (size_t)KP * NP;           // [ST2_ROWS][LP]; after the update: the NEW [W2 | b2 | That] rows
    constexpr int LP = ST2_LP;
    const int d0 = w2 * ST2_ROWS;
    {   // S is stored with row stride NP: whole 16-byte chunks, everything in flight at once
        const int n4 = K * NP / 4;
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(sS);
        for (int e = tid; e < n4; e += ST2_THREADS) ptx::cp_async16(dst + e * 16, a.S + e * 4);
        for (int e = K * NP + tid; e < KP * NP; e += ST2_THREADS) sS[e] = 0.f;
    }
    {   // W2 rows: 16-byte copies (H % 4 == 0, KP % 4 == 0); b2 / That / padding: scalars
        const int h4 = H >> 2;
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(sL);
        for (int e = tid; e < ST2_ROWS * h4; e += ST2_THREADS) {
            const int r = e / h4, q = e - r * h4, d = d0 + r;
            if (d < D) ptx::cp_async16(dst + ((size_t)r * LP + q * 4) * 4, a.W2 + (size_t)d * H + q * 4);
            else *reinterpret_cast<float4*>(sL + (size_t)r * LP + q * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        ptx::cp_async_commit();
        const int tail = LP - H;
        for (int e = tid; e < ST2_ROWS * tail; e += ST2_THREADS) {
            const int r = e / tail, k = H + (e - r * tail), d = d0 + r;
            float v = 0.f;
            if (d < D && k < K) v = k == H ? a.b2[d] : __ldg(a.That + (size_t)d * C + (k - H - 1));
            sL[(size_t)r * LP + k] = v;
        }
    }
    // 4 x 4 register tiles: warp rq owns rows d0 + 4 rq .. + 3, lane cq columns 4 cq .. + 3 (< H); column H (b2) afterwards
    const int rq = warp, cq = lane;
    const bool worker = cq * 4 < H;
    float4 vv[4];                                // momentum of the owned elements: requested before the contraction
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int d = d0 + rq * 4 + i;
        vv[i] = (worker && d < D) ? *reinterpret_cast<const float4*>(a.v + oW2 + (size_t)d * H + cq * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float vb2 = 0.f;
    if (lane < 4 && d0 + rq * 4 + lane < D) vb2 = a.v[ob2 + d0 + rq * 4 + lane];
    ptx::cp_async_wait<0>();
    __syncthreads();
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    if (worker) {
        for (int k = 0; k < KP; k += 4) {
            float4 sv[4], lv[4];
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) sv[kk] = *reinterpret_cast<const float4*>(sS + (size_t)(k + kk) * NP + cq * 4);
#pragma unroll
            for (int i = 0; i < 4; ++i) lv[i] = *reinterpret_cast<const float4*>(sL + (size_t)(rq * 4 + i) * LP + k);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float l[4] = {lv[i].x, lv[i].y, lv[i].z, lv[i].w};
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    acc[i][0] = fmaf(l[kk], sv[kk].x, acc[i][0]); acc[i][1] = fmaf(l[kk], sv[kk].y, acc[i][1]);
                    acc[i][2] = fmaf(l[kk], sv[kk].z, acc[i][2]); acc[i][3] = fmaf(l[kk], sv[kk].w, acc[i][3]);
                }
            }
        }
    }
    // db2 of the warp's four rows: lane-strided dot products with column H of S
    float gb2[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float p = 0.f;
        for (int k = lane; k < K; k += 32) p = fmaf(sL[(size_t)(rq * 4 + i) * LP + k], sS[(size_t)k * NP + H], p);
        gb2[i] = warp_sum(p);
    }
    __syncthreads();                             // every thread is done reading the OLD rows in sL
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int d = d0 + rq * 4 + i;
        if (d >= D) continue;
        float* lrow = sL + (size_t)(rq * 4 + i) * LP;
        if (worker) {
            const float4 p4 = *reinterpret_cast<const float4*>(lrow + cq * 4);
            const float px[4] = {p4.x, p4.y, p4.z, p4.w}, vx[4] = {vv[i].x, vv[i].y, vv[i].z, vv[i].w};
            float po[4], vo[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float g = acc[i][j] + a.wd * px[j];
                vo[j] = a.momentum * vx[j] + g;
                po[j] = px[j] - lr * vo[j];
            }
            const size_t fo = (size_t)d * H + cq * 4;
            *reinterpret_cast<float4*>(a.W2 + fo) = make_float4(po[0], po[1], po[2], po[3]);
            *reinterpret_cast<float4*>(a.v + oW2 + fo) = make_float4(vo[0], vo[1], vo[2], vo[3]);
            *reinterpret_cast<float4*>(a.g + oW2 + fo) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
            *reinterpret_cast<float4*>(lrow + cq * 4) = make_float4(po[0], po[1], po[2], po[3]);
        }
        const float vb = __shfl_sync(0xffffffffu, vb2, i);
        if (lane == 0) {
            const float pv = lrow[H];
            const float g = gb2[i] + a.wd * pv;
            const float vn = a.momentum * vb + g;
            const float pn = pv - lr * vn;
            a.b2[d] = pn; a.v[ob2 + d] = vn; a.g[ob2 + d] = gb2[i];
            lrow[H] = pn;
        }
    }
    __syncthreads();
    // ---- Gram share of the new rows: warp w owns m in [8w, 8w + 8) and the last warp also the row m = H
    const int ldg = K;
    constexpr int MW = 8;
    float gacc[MW][ST_NSLOT], gx[ST_NSLOT];
#pragma unroll
    for (int s = 0; s < ST_NSLOT; ++s) gx[s] = 0.f;
#pragma unroll
    for (int m = 0; m < MW; ++m)
#pragma unroll
        for (int s = 0; s < ST_NSLOT; ++s) gacc[m][s] = 0.f;
    const int m0 = warp * MW;                    // the rows are zero padded to LP columns: no guards in the loop (columns
    const float* rowp = sL + lane;               // m >= H of a short hidden layer produce products that are never stored)
    const float* rowm = sL + m0;
#pragma unroll 2
    for (int r = 0; r < ST2_ROWS; ++r) {
        float bv[ST_NSLOT];
#pragma unroll
        for (int s = 0; s < ST_NSLOT; ++s) bv[s] = rowp[r * LP + 32 * s];
        const float ax = sL[r * LP + H];
        const float4 a0 = *reinterpret_cast<const float4*>(rowm + r * LP), a1 = *reinterpret_cast<const float4*>(rowm + r * LP + 4);
        const float amx[MW] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
        for (int s = 0; s < ST_NSLOT; ++s) gx[s] = fmaf(ax, bv[s], gx[s]);
#pragma unroll
        for (int q = 0; q < MW; ++q)
#pragma unroll
            for (int s = 0; s < ST_NSLOT; ++s) gacc[q][s] = fmaf(amx[q], bv[s], gacc[q][s]);
    }
#pragma unroll
    for (int m = 0; m < MW; ++m) {
        if (m0 + m >= H) break;
#pragma unroll
        for (int s = 0; s < ST_NSLOT; ++s) {
            const int n = lane + 32 * s;
            if (n < ldg) atomicAdd(&a.gram_next[(size_t)(m0 + m) * ldg + n], gacc[m][s]);
        }
    }
    if (warp == ST2_THREADS / 32 - 1) {
#pragma unroll
        for (int s = 0; s < ST_NSLOT; ++s) {
            const int n = lane + 32 * s;
            if (n < ldg) atomicAdd(&a.gram_next[(size_t)H * ldg + n], gx[s]);
        }
    }
}

static inline bool step_tail_supported(int D, int H, int C) {
    return H % 4 == 0 && D % 4 == 0 && H <= 128 && (H + 1 + C) <= 32 * ST_NSLOT && step_tail_smem_bytes(H, C) <= 227 * 1024;
}

static int launch_step_tail(StepTailArgs a, cudaStream_t st) {
    DBMM_CHECK_SHAPE(step_tail_supported(a.D, a.H, a.C), "step tail kernels: unsupported D=%d H=%d C=%d", a.D, a.H, a.C);
    DBMM_CHECK_ARG(a.nchunk <= ST_MAXCHUNK, "at most %d batch chunks (got %d)", ST_MAXCHUNK, a.nchunk);
    static const int skip_roles = getenv("DBMM_TAIL_SKIP") ? atoi(getenv("DBMM_TAIL_SKIP")) : 0;      // timing experiments only
    a.roles &= ~skip_roles;
    a.n_w1_ctas = ceil_div((int64_t)a.H * a.D / 4, ST_THREADS);
    if (a.n_w1_ctas > 128) a.n_w1_ctas = 128;
    a.n_w2_ctas = ceil_div(a.D, ST2_ROWS);
    if (a.roles & 1) {
        k_tail_w1<<<a.n_w1_ctas + 1, ST_THREADS, 0, st>>>(a);
        DBMM_LAUNCH_CHECK();
    }
    if (a.roles & 2) {
        const size_t smem = step_tail_smem_bytes(a.H, a.C);
        DBMM_CUDA(set_smem(k_tail_w2, smem));
        k_tail_w2<<<a.n_w2_ctas, ST2_THREADS, smem, st>>>(a);
        DBMM_LAUNCH_CHECK();
    }
    return DBMM_OK;
}

}  // namespace dbmm
