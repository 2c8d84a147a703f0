// dW1 on the 5th-generation tensor cores:   dW1[j][d] = sum_b da[b][j] * X[row(b)][d]        (H x D, K = batch)
//
//   * da (BatchNorm backward of the saved dL/d(ahat)) is formed on the fly by the producer warps from A / dahat and the
//     batch statistics -- it is never written to global memory.
//   * Both operands are "MN-major" for the tensor core: da is stored [b][j] (M = j contiguous), X is stored [b][d]
//     (N = d contiguous), K = b walks rows.  For 32-bit operands the only MN-major shared-memory layout the UMMA
//     descriptor accepts is SWIZZLE_128B_BASE32B: atoms of 4 K-rows x 128 bytes, the 32-byte chunk index XORed with
//     (row & 3).  One kind::tf32 MMA (K = 8) reads two such atoms per 32-float MN segment.
//   * da is split hi + lo (hi = top 19 bits) and two MMAs accumulate hi*x + lo*x in the fp32 TMEM accumulator; X is
//     used as stored (exact for fp16-valued embeddings, see DESIGN.md "Precision policy").
//   * The batch is split into chunks (grid.y); every CTA writes its partial 128 x 128 tile with plain vector stores
//     to part[chunk][j][d] -- no atomics, no zeroing, deterministic -- and k_finalize_grads sums the chunks.
#pragma once
#include "common.cuh"
#include "ptx_sm100.cuh"

namespace dbmm {

constexpr int WG_BK = 64;                      // batch rows per pipeline stage
constexpr int WG_TILE = 128;                   // hidden units (M) and embedding columns (N) per CTA
constexpr int WG_OP_BYTES = WG_BK * WG_TILE * 4;          // one operand tile: 64 rows x 128 floats = 32 KB
constexpr int WG_STAGE_BYTES = 3 * WG_OP_BYTES;           // da_hi | da_lo | X
constexpr int WG_STAGES = 2;
constexpr size_t WG_SMEM = (size_t)WG_STAGES * WG_STAGE_BYTES + 1024 + 256;
constexpr int WG_THREADS = 160;

struct WgradTcArgs {
    const float* X; int64_t ldx; const int32_t* idx;
    int B; int64_t Bg; int D, H;
    const float* A;        // [B][H] pre-BatchNorm activations of the trainable adapter
    const float* dahat;    // [B][H]
    const double* colsum;  // [2][H] sum a, sum a^2 (global batch)
    const double* dgb;     // [2][H] dgamma, dbeta (global batch)
    const float* gamma;
    float* part;           // [nchunk][H][D]
    int rows_per_chunk;    // multiple of WG_BK
};

// byte offset of the 16-byte chunk (4 floats) `c16` (0..31 along the 128-float MN extent) of K-row `row` (0..63)
// inside one operand tile laid out as [kstep 8 rows][MN atom 32 floats][K atom 4 rows][128 B], SW128_BASE32B.
__device__ __forceinline__ uint32_t wg_tile_offset(uint32_t row, uint32_t c16) {
    const uint32_t kstep = row >> 3, ka = (row >> 2) & 1u, r = row & 3u;
    const uint32_t atom = c16 >> 3, c32 = (c16 >> 1) & 3u, half = c16 & 1u;
    return kstep * 4096u + atom * 1024u + ka * 512u + r * 128u + ((c32 ^ r) << 5) + (half << 4);
}

// MN-major SWIZZLE_128B_BASE32B descriptor: LBO = bytes between 32-float MN atoms, SBO = bytes between 4-row K atoms.
__device__ __forceinline__ uint64_t wg_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((1024u >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((512u >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;      // descriptor version (Blackwell)
    d |= (uint64_t)1 << 61;      // SWIZZLE_128B_BASE32B
    return d;
}

__global__ void __launch_bounds__(WG_THREADS, 1) k_wgrad_tc(WgradTcArgs a) {
    extern __shared__ uint8_t wg_smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)wg_smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* full = (uint64_t*)(smem + (size_t)WG_STAGES * WG_STAGE_BYTES);
    uint64_t* empty = full + WG_STAGES;
    uint64_t* tmem_full = empty + WG_STAGES;
    uint32_t* tmem_ptr = (uint32_t*)(tmem_full + 1);
    __shared__ float sCst[4][WG_TILE];          // mu, rstd, mean(dahat), mean(dahat*ahat) per hidden unit

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int d0 = blockIdx.x * WG_TILE;
    const int chunk = blockIdx.y;
    const int b_begin = chunk * a.rows_per_chunk;
    const int b_end = min(a.B, b_begin + a.rows_per_chunk);
    const int n_sub = (b_end - b_begin + WG_BK - 1) / WG_BK;     // >= 1: the host never launches empty chunks
    const int H = a.H;

    if (tid < WG_TILE) {
        float mu = 0.f, rstd = 0.f, m1 = 0.f, m2 = 0.f;
        if (tid < H) {
            const double m = a.colsum[tid] / (double)a.Bg;
            double v = a.colsum[H + tid] / (double)a.Bg - m * m;
            if (v < 0.0) v = 0.0;
            mu = (float)m; rstd = 1.0f / sqrtf((float)v + DBMM_BN_EPS);
            const double gm = (double)a.gamma[tid];
            m1 = (float)(gm * a.dgb[H + tid] / (double)a.Bg);
            m2 = (float)(gm * a.dgb[tid] / (double)a.Bg);
        }
        sCst[0][tid] = mu; sCst[1][tid] = rstd; sCst[2][tid] = m1; sCst[3][tid] = m2;
    }
    if (tid == 0) {
        for (int s = 0; s < WG_STAGES; ++s) { ptx::mbar_init(&full[s], 128); ptx::mbar_init(&empty[s], 1); }
        ptx::mbar_init(tmem_full, 1);
        ptx::fence_mbar_init();
    }
    if (warp == 4) ptx::tmem_alloc<WG_TILE>(tmem_ptr);
    ptx::tc_fence_before_sync();
    __syncthreads();
    ptx::tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp < 4) {
        // ===================== producers =====================
        for (int sub = 0; sub < n_sub; ++sub) {
            const int s = sub % WG_STAGES;
            ptx::mbar_wait(&empty[s], ((sub / WG_STAGES) & 1) ^ 1);
            uint8_t* stage = smem + (size_t)s * WG_STAGE_BYTES;
            const uint32_t sX = ptx::smem_u32(stage) + 2 * WG_OP_BYTES;
            const int b0 = b_begin + sub * WG_BK;
            // X tile: 64 rows x 512 bytes, gathered through the batch's index list
#pragma unroll 4
            for (int id = tid; id < WG_BK * 32; id += 128) {
                const int row = id >> 5, c16 = id & 31;
                int b = b0 + row;
                if (b >= b_end) b = b_end - 1;                    // tail rows: any finite values (da is 0 there)
                const int64_t r = a.idx ? (int64_t)__ldg(a.idx + b) : (int64_t)b;
                ptx::cp_async16(sX + wg_tile_offset(row, c16), a.X + r * a.ldx + d0 + c16 * 4);
            }
            ptx::cp_async_commit();
            // da tile (hi / lo), computed while the X copies are in flight
#pragma unroll 2
            for (int id = tid; id < WG_BK * 32; id += 128) {
                const int row = id >> 5, c16 = id & 31;
                const int b = b0 + row, j = c16 * 4;
                float4 hi = make_float4(0.f, 0.f, 0.f, 0.f), lo = hi;
                if (b < b_end && j < H) {
                    const float4 av = __ldg(reinterpret_cast<const float4*>(a.A + (size_t)b * H + j));
                    const float4 dv = __ldg(reinterpret_cast<const float4*>(a.dahat + (size_t)b * H + j));
                    const float ax[4] = {av.x, av.y, av.z, av.w}, dx[4] = {dv.x, dv.y, dv.z, dv.w};
                    float h4[4], l4[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float rstd = sCst[1][j + q];
                        const float ah = (ax[q] - sCst[0][j + q]) * rstd;
                        const float da = (dx[q] - sCst[2][j + q] - ah * sCst[3][j + q]) * rstd;
                        h4[q] = __uint_as_float(__float_as_uint(da) & 0xffffe000u);
                        l4[q] = da - h4[q];
                    }
                    hi = make_float4(h4[0], h4[1], h4[2], h4[3]);
                    lo = make_float4(l4[0], l4[1], l4[2], l4[3]);
                }
                const uint32_t off = wg_tile_offset(row, c16);
                *reinterpret_cast<float4*>(stage + off) = hi;
                *reinterpret_cast<float4*>(stage + WG_OP_BYTES + off) = lo;
            }
            ptx::cp_async_wait<0>();
            ptx::fence_proxy_async_smem();
            ptx::mbar_arrive(&full[s]);
        }

        // ===================== epilogue: TMEM -> partial tile in global memory =====================
        ptx::mbar_wait(tmem_full, 0);
        ptx::tc_fence_after_sync();
        const int j = warp * 32 + lane;
        float* out = a.part + ((size_t)chunk * H + (j < H ? j : 0)) * a.D + d0;
#pragma unroll 1
        for (int ch = 0; ch < WG_TILE / 32; ++ch) {
            uint32_t r[32];
            ptx::tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(warp * 32) << 16) + ch * 32, r);
            ptx::tmem_ld_wait();
            if (j < H) {
#pragma unroll
                for (int q = 0; q < 32; q += 4)
                    *reinterpret_cast<float4*>(out + ch * 32 + q) =
                        make_float4(__uint_as_float(r[q]), __uint_as_float(r[q + 1]), __uint_as_float(r[q + 2]), __uint_as_float(r[q + 3]));
            }
        }
        ptx::tc_fence_before_sync();
    } else {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc = ptx::umma_idesc(/*tf32*/ 2, WG_TILE, WG_TILE, /*A MN-major*/ 1, /*B MN-major*/ 1);
            for (int sub = 0; sub < n_sub; ++sub) {
                const int s = sub % WG_STAGES;
                ptx::mbar_wait(&full[s], (sub / WG_STAGES) & 1);
                ptx::tc_fence_after_sync();
                const uint32_t base = ptx::smem_u32(smem + (size_t)s * WG_STAGE_BYTES);
#pragma unroll
                for (int ks = 0; ks < WG_BK / 8; ++ks) {
                    const uint64_t bdesc = wg_desc(base + 2 * WG_OP_BYTES + ks * 4096);
#pragma unroll
                    for (int t = 0; t < 2; ++t) {
                        const uint64_t adesc = wg_desc(base + t * WG_OP_BYTES + ks * 4096);
                        ptx::mma_tf32_ss(tmem_base, adesc, bdesc, idesc, (sub | ks | t) != 0 ? 1u : 0u);
                    }
                }
                ptx::mma_commit(&empty[s]);
            }
            ptx::mma_commit(tmem_full);
        }
        __syncwarp();
    }
    __syncthreads();
    if (warp == 4) ptx::tmem_dealloc<WG_TILE>(tmem_base);
}

static inline int wgrad_tc_chunks(int B, int* rows_per_chunk) {
    // ~16 batch chunks (128 CTAs at D = 1024), each a multiple of the 64-row stage
    int rpc = (B + 15) / 16;
    rpc = (rpc + WG_BK - 1) / WG_BK * WG_BK;
    *rows_per_chunk = rpc;
    return (B + rpc - 1) / rpc;
}

static int launch_wgrad_tc(const WgradTcArgs& a, int nchunk, cudaStream_t st) {
    DBMM_CHECK_SHAPE(a.D % WG_TILE == 0 && a.H <= WG_TILE && a.H % 4 == 0, "tensor-core dW1 needs D %% 128 == 0 and H <= 128 (D=%d H=%d)", a.D, a.H);
    DBMM_CUDA(cudaFuncSetAttribute(k_wgrad_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WG_SMEM));
    dim3 grid(a.D / WG_TILE, nchunk);
    k_wgrad_tc<<<grid, WG_THREADS, WG_SMEM, st>>>(a);
    DBMM_LAUNCH_CHECK();
    return DBMM_OK;
}

// ------------------------------------------------------------------------------------------------
// Gradient finalisation: sum the dW1 chunks, dW2a = [W2 | b2 | That] S, dgamma / dbeta / db1 -> flat gradient
// ------------------------------------------------------------------------------------------------
struct FinalizeArgs {
    const float* part; int nchunk;        // [nchunk][H][D], or nullptr when gW1 was produced directly
    const float* W2; const float* b2; const float* That; const float* S;   // S: [H+1+C][H+1]
    const double* dgb;
    float* gW1; float* gb1; float* ggamma; float* gbeta; float* gW2; float* gb2;
    int D, H, C;
    int n_w1_ctas;
};

constexpr int FIN_ROWS = 16;       // embedding rows of dW2a per CTA
constexpr int FIN_THREADS = 256;

static inline size_t finalize_smem_bytes(int H, int C) {
    return sizeof(float) * ((size_t)(H + 1 + C) * (H + 1) + (size_t)FIN_ROWS * (H + 1 + C)) + 16;
}

__global__ void __launch_bounds__(FIN_THREADS) k_finalize_grads(FinalizeArgs a) {
    extern __shared__ __align__(16) float fin_smem[];
    const int H = a.H, C = a.C, D = a.D, K = H + 1 + C, N = H + 1;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if ((int)blockIdx.x < a.n_w1_ctas) {
        // ---- dW1 = sum of the batch-chunk partial tiles (16-byte accesses; H * D is a multiple of 4)
        if (a.part) {
            const int64_t n4 = (int64_t)H * D / 4;
            for (int64_t i = (int64_t)blockIdx.x * FIN_THREADS + tid; i < n4; i += (int64_t)a.n_w1_ctas * FIN_THREADS) {
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int c = 0; c < a.nchunk; ++c) {
                    const float4 v = __ldcg(reinterpret_cast<const float4*>(a.part + (size_t)c * H * D) + i);
                    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
                }
                reinterpret_cast<float4*>(a.gW1)[i] = acc;
            }
        }
        if (blockIdx.x == 0) {
            for (int j = tid; j < H; j += FIN_THREADS) {
                a.ggamma[j] = (float)a.dgb[j];
                a.gbeta[j] = (float)a.dgb[H + j];
                // db1 = sum_B da vanishes identically (BatchNorm removes the bias); the reference's value is autograd
                // rounding noise (|db1| ~ 1e-9, tests/test_oracle_golden.py), so b1 moves by weight decay only.
                a.gb1[j] = 0.f;
            }
        }
        return;
    }
    // ---- dW2a rows [d0, d0 + FIN_ROWS):  out[d][n] = sum_k L[d][k] S[k][n],  L = [W2 | b2 | That]
    float* sS = fin_smem;                       // [K][N]
    float* sL = sS + (size_t)K * N;             // [FIN_ROWS][K]
    const int d0 = ((int)blockIdx.x - a.n_w1_ctas) * FIN_ROWS;
    for (int e = tid; e < K * N; e += FIN_THREADS) sS[e] = __ldcg(a.S + e);
    for (int e = tid; e < FIN_ROWS * K; e += FIN_THREADS) {
        const int r = e / K, k = e - r * K, d = d0 + r;
        float v = 0.f;
        if (d < D) v = k < H ? a.W2[(size_t)d * H + k] : (k == H ? a.b2[d] : a.That[(size_t)d * C + (k - H - 1)]);
        sL[e] = v;
    }
    __syncthreads();
    constexpr int RPW = FIN_ROWS / (FIN_THREADS / 32);      // rows per warp (2)
    constexpr int NS = 5;                                   // 32-wide output column slots (N <= 129 + padding)
    float acc[RPW][NS];
#pragma unroll
    for (int r = 0; r < RPW; ++r)
#pragma unroll
        for (int s = 0; s < NS; ++s) acc[r][s] = 0.f;
    const float* Lr = sL + (size_t)(warp * RPW) * K;
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
        float sv[NS];
#pragma unroll
        for (int s = 0; s < NS; ++s) { const int n = lane + 32 * s; sv[s] = n < N ? sS[(size_t)k * N + n] : 0.f; }
#pragma unroll
        for (int r = 0; r < RPW; ++r) {
            const float l = Lr[(size_t)r * K + k];
#pragma unroll
            for (int s = 0; s < NS; ++s) acc[r][s] = fmaf(l, sv[s], acc[r][s]);
        }
    }
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
        const int d = d0 + warp * RPW + r;
        if (d >= D) continue;
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            const int n = lane + 32 * s;
            if (n < H) a.gW2[(size_t)d * H + n] = acc[r][s];
            else if (n == H) a.gb2[d] = acc[r][s];
        }
    }
}

static int launch_finalize(FinalizeArgs a, cudaStream_t st) {
    const size_t smem = finalize_smem_bytes(a.H, a.C);
    DBMM_CHECK_SHAPE(smem <= 227 * 1024 && a.H + 1 <= 160, "finalize kernel: H=%d C=%d too large", a.H, a.C);
    DBMM_CUDA(cudaFuncSetAttribute(k_finalize_grads, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    a.n_w1_ctas = a.part ? 64 : 1;
    const int n_w2 = ceil_div(a.D, FIN_ROWS);
    k_finalize_grads<<<a.n_w1_ctas + n_w2, FIN_THREADS, smem, st>>>(a);
    DBMM_LAUNCH_CHECK();
    return DBMM_OK;
}

}  // namespace dbmm
