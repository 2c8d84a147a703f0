// Optimizer step of the trainable adapter (torch.optim.SGD semantics, demo/util.py:118-136: g += wd*p;
// v = momentum*v + g; p -= lr*v -- the caller zeroes v before the optimizer's first step, which makes v = g) fused
// with everything the NEXT training step needs from the new weights:
//   * W1 CTAs also emit the tf32 split W1 = hi + lo consumed by the tensor-core GEMM-1,
//   * (the Gram matrix G = [W2 | b2]^T [W2 | b2 | That] of the new weights follows as a k_tn_gemm launch),
//   * the last CTA updates b1 / gamma / beta, the BatchNorm running statistics of every adapter in the forward
//     (the frozen one drifts too, final_main.py:122,574) and re-zeroes the per-step accumulators.
#pragma once
#include "common.cuh"
#include "ptx_sm100.cuh"

namespace dbmm {

constexpr int UP_THREADS = 256, UP_ROWS = 16, UP_NSLOT = 5;

struct UpdateArgs {
    float* W1; float* b1; float* gamma; float* beta; float* W2; float* b2;    // trainable adapter
    const float* g; float* v;                 // flat gradient / momentum: W1 | b1 | gamma | beta | W2 | b2
    const float* lr_dev; float lr;            // lr_dev != nullptr: learning rate read from device memory
    float momentum, wd;
    float* whi; float* wlo;                   // [H][D] tf32 split of the new W1, or nullptr
    const float* That;                        // [D][C]
    int D, H, C, nad; int64_t Bg;
    fx64* colsum; fx64* dgb;                  // per-step fixed-point accumulators (read for the running stats, then zeroed)
    int zero_accum;
    float* rm[2]; float* rv[2]; long long* nbt[2];
    int n_w1_ctas, n_w2_ctas;
};

static inline size_t update_smem_bytes(int H, int C) { return sizeof(float) * (size_t)UP_ROWS * ((H + 1 + C + 3) & ~3) + 16; }

__global__ void __launch_bounds__(UP_THREADS) k_update(UpdateArgs a) {
    extern __shared__ __align__(16) float up_smem[];
    const int H = a.H, D = a.D, C = a.C;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float lr = a.lr_dev ? __ldg(a.lr_dev) : a.lr;
    const size_t oW1 = 0, ob1 = (size_t)H * D, og = ob1 + H, obeta = og + H, oW2 = obeta + H, ob2 = oW2 + (size_t)D * H;
    const int bid = blockIdx.x;
    ptx::pdl_wait();                // the flat gradient comes from k_finalize_grads
    ptx::pdl_launch();

    if (bid < a.n_w1_ctas) {
        // ---- W1: 16-byte SGD + tf32 split
        const int64_t n4 = (int64_t)H * D / 4;
        for (int64_t i = (int64_t)bid * UP_THREADS + tid; i < n4; i += (int64_t)a.n_w1_ctas * UP_THREADS) {
            const float4 pv = reinterpret_cast<const float4*>(a.W1)[i];
            const float4 gv = __ldcg(reinterpret_cast<const float4*>(a.g + oW1) + i);
            const float4 vv = reinterpret_cast<const float4*>(a.v + oW1)[i];
            const float px[4] = {pv.x, pv.y, pv.z, pv.w}, gx[4] = {gv.x, gv.y, gv.z, gv.w}, vx[4] = {vv.x, vv.y, vv.z, vv.w};
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
            if (a.whi) {
                reinterpret_cast<float4*>(a.whi)[i] = make_float4(hi[0], hi[1], hi[2], hi[3]);
                reinterpret_cast<float4*>(a.wlo)[i] = make_float4(lo[0], lo[1], lo[2], lo[3]);
            }
        }
        return;
    }
    if (bid < a.n_w1_ctas + a.n_w2_ctas) {
        // ---- W2 / b2 rows [d0, d0 + UP_ROWS): SGD, then this slice's share of the Gram matrix
        const int ldg = H + 1 + C, HP = H + 1, LP = (ldg + 3) & ~3;
        float* sRow = up_smem;                                  // [UP_ROWS][LP] = [W2 | b2 | That] of the new weights
        const int d0 = (bid - a.n_w1_ctas) * UP_ROWS;
        // read-modify-write of the rows' parameters: all loads of the (up to UP_EPT) elements a thread owns are issued
        // before the first dependent use
        constexpr int UP_EPT = (UP_ROWS * 148 + UP_THREADS - 1) / UP_THREADS;      // LP <= 148 (H <= 128, C <= 16)
        float pv[UP_EPT], gv[UP_EPT], vv[UP_EPT];
#pragma unroll
        for (int i = 0; i < UP_EPT; ++i) {
            const int e = tid + i * UP_THREADS;
            pv[i] = 0.f; gv[i] = 0.f; vv[i] = 0.f;
            if (e < UP_ROWS * LP) {
                const int r = e / LP, k = e - r * LP, d = d0 + r;
                if (d < D && k < ldg) {
                    if (k <= H) {
                        const size_t fo = k < H ? (oW2 + (size_t)d * H + k) : (ob2 + d);
                        pv[i] = k < H ? a.W2[(size_t)d * H + k] : a.b2[d];
                        gv[i] = __ldcg(a.g + fo);
                        vv[i] = a.v[fo];
                    } else {
                        pv[i] = __ldg(a.That + (size_t)d * C + (k - H - 1));
                    }
                }
            }
        }
#pragma unroll
        for (int i = 0; i < UP_EPT; ++i) {
            const int e = tid + i * UP_THREADS;
            if (e < UP_ROWS * LP) {
                const int r = e / LP, k = e - r * LP, d = d0 + r;
                float out = pv[i];
                if (d < D && k <= H) {
                    const size_t fo = k < H ? (oW2 + (size_t)d * H + k) : (ob2 + d);
                    const float g = gv[i] + a.wd * pv[i];
                    const float vn = a.momentum * vv[i] + g;
                    a.v[fo] = vn;
                    out = pv[i] - lr * vn;
                    if (k < H) a.W2[(size_t)d * H + k] = out; else a.b2[d] = out;
                }
                sRow[e] = out;
            }
        }
        return;
    }
    // ---- last CTA: b1 / gamma / beta, BatchNorm running statistics, accumulator reset
    for (int e = tid; e < 3 * H; e += UP_THREADS) {
        const int seg = e / H, j = e - seg * H;
        float* pp = (seg == 0 ? a.b1 : (seg == 1 ? a.gamma : a.beta)) + j;
        const size_t fo = ob1 + e;                              // b1 | gamma | beta are contiguous in the flat layout
        const float pv = *pp;
        const float g = a.g[fo] + a.wd * pv;
        const float vn = a.momentum * a.v[fo] + g;
        a.v[fo] = vn;
        *pp = pv - lr * vn;
    }
    if (a.colsum) {
        for (int e = tid; e < a.nad * H; e += UP_THREADS) {
            const int ad = e / H, j = e - ad * H;
            const double m = fx_get<FX_COLSUM>(&a.colsum[((size_t)ad * 2 + 0) * H + j]) / (double)a.Bg;
            double var = fx_get<FX_COLSUM>(&a.colsum[((size_t)ad * 2 + 1) * H + j]) / (double)a.Bg - m * m;
            if (var < 0.0) var = 0.0;
            const float unbiased = (float)(var * (double)a.Bg / (double)(a.Bg - 1));
            a.rm[ad][j] = (1.0f - DBMM_BN_MOMENTUM) * a.rm[ad][j] + DBMM_BN_MOMENTUM * (float)m;
            a.rv[ad][j] = (1.0f - DBMM_BN_MOMENTUM) * a.rv[ad][j] + DBMM_BN_MOMENTUM * unbiased;
        }
        if (tid < a.nad) *a.nbt[tid] += 1;
    }
    if (a.zero_accum) {
        __syncthreads();
        for (int e = tid; e < a.nad * 2 * H; e += UP_THREADS) a.colsum[e].v = 0;
        for (int e = tid; e < 2 * H; e += UP_THREADS) a.dgb[e].v = 0;
    }
}

// BatchNorm running statistics of a train-mode forward that is NOT followed by the fused update (nn.Module forward under
// torch autograd, modules.py): running_mean / running_var (unbiased, momentum 0.1) and num_batches_tracked, as torch does.
struct BnRunningArgs { const fx64* colsum; int nad, H; int64_t Bg; float* rm[2]; float* rv[2]; long long* nbt[2]; };
__global__ void __launch_bounds__(256) k_bn_running(BnRunningArgs a) {
    const int H = a.H;
    for (int e = threadIdx.x; e < a.nad * H; e += 256) {
        const int ad = e / H, j = e - ad * H;
        const double m = fx_get<FX_COLSUM>(&a.colsum[((size_t)ad * 2 + 0) * H + j]) / (double)a.Bg;
        double var = fx_get<FX_COLSUM>(&a.colsum[((size_t)ad * 2 + 1) * H + j]) / (double)a.Bg - m * m;
        if (var < 0.0) var = 0.0;
        const float unbiased = (float)(var * (double)a.Bg / (double)(a.Bg - 1));
        a.rm[ad][j] = (1.0f - DBMM_BN_MOMENTUM) * a.rm[ad][j] + DBMM_BN_MOMENTUM * (float)m;
        a.rv[ad][j] = (1.0f - DBMM_BN_MOMENTUM) * a.rv[ad][j] + DBMM_BN_MOMENTUM * unbiased;
    }
    if (threadIdx.x < a.nad) *a.nbt[threadIdx.x] += 1;
}

static int launch_update(UpdateArgs a, cudaStream_t st) {
    const size_t smem = update_smem_bytes(a.H, a.C);
    DBMM_CHECK_SHAPE(a.H % 4 == 0 && a.D % 4 == 0 && a.H <= 8 * 16, "update kernel: unsupported H=%d D=%d", a.H, a.D);
    DBMM_CUDA(set_smem(k_update, smem));
    a.n_w1_ctas = 64;
    a.n_w2_ctas = ceil_div(a.D, UP_ROWS);
    DBMM_CUDA(launch_pdl(k_update, dim3(a.n_w1_ctas + a.n_w2_ctas + 1), dim3(UP_THREADS), smem, st, a));
    return DBMM_OK;
}

}  // namespace dbmm
