// Dense "NT" GEMM on the 5th-generation tensor cores, operands fed by TMA:
//
//     C[m][n] = scale * sum_k A[m][k] * B[n][k]          A: [M, K] row-major,  B: [N, K] row-major  (both K-major)
//
// used for the only truly dense contractions of the path: the contrastive regulariser's B x B similarity and its
// gradient GEMMs (config 3) and the zero-shot head against up to 1,000+ prompt columns (config 4).
//
//   * kind::tf32 MMA, M = N = 128 per CTA, K = 8 per instruction, fp32 accumulator in TMEM.
//   * "3xTF32": the B side always arrives pre-split (B = Bhi + Blo, Bhi tf32-exact); the A side either arrives
//     pre-split too (SPLIT_A: terms Ahi*Bhi + Alo*Bhi + Ahi*Blo, fp32-level accuracy for arbitrary data) or is used as
//     stored (terms A*Bhi + A*Blo: exact when A is tf32-representable, e.g. fp16-valued CLIP embeddings).
//   * persistent CTAs (one per SM) walk the output tiles; the fp32 accumulator is double-buffered in TMEM (2 x 128
//     columns) so the epilogue of tile i overlaps the MMAs of tile i+1, and set-up costs are paid once per SM;
//   * one elected thread issues 2-D tiled TMA loads (SWIZZLE_128B boxes of 32 floats x 128 rows) into a 3-4 stage
//     ring guarded by full/empty mbarriers; one thread issues the MMAs; four warps drain TMEM.
//   * epilogues: EPI_STORE writes the scaled tile; EPI_SOFTMAX_PART reduces the tile to per-row online-softmax
//     partials (max, sum-exp, argmax, target logit) so the logits of the zero-shot head are never materialised.
#pragma once
#include "common.cuh"
#include "ptx_sm100.cuh"

namespace dbmm {

constexpr int TG_BM = 128, TG_BN = 128, TG_BK = 32, TG_THREADS = 192;
constexpr int TG_TILE_BYTES = 128 * 128;                    // one 128-row x 32-float operand tile
enum { EPI_STORE = 0, EPI_SOFTMAX_PART = 1, EPI_HSPACE = 2, EPI_EVAL_H = 3 };

struct SoftmaxPart { float mx, se, ly; int am; };           // per (row, column tile): max, sum exp(l - max), target logit, argmax

struct TcGemmArgs {
    int M, N, K;
    float scale;                                            // acc * scale * (*scale_dev) * rowscale[m]
    const float* scale_dev; const float* rowscale;          // optional (null = 1)
    const float* col_bias;                                  // EPI_SOFTMAX_PART: optional additive bias per column (linear probe)
    float* C; int64_t ldc; int accumulate;                  // EPI_STORE: C = [C +] scaled tile
    const int32_t* y; const int32_t* idx;                   // EPI_SOFTMAX_PART: target column per DATASET row, row list (or null)
    int64_t pos0;
    SoftmaxPart* part;                                      // [number of column tiles][M]
    // EPI_HSPACE (eval forward, t = [h, 1] G with G^T as the B operand): column tile 0 reduces t[0:128] to
    // rowdot[m] = sum_j t_j h_j (h = Ahi + Alo, re-read from global), further tiles store t[128 + j] to tail[m][j]
    const float* hs_hi; const float* hs_lo; int64_t hs_ld; const float* hs_bias; float* rowdot; float* tail; int tail_ld;
    // EPI_EVAL_H (GEMM-1 of the eval forward): h = relu(BatchNorm_running(acc + b1)) split into tf32 hi + lo, [M][N] each
    const float* e_b1; const float* e_mean; const float* e_var; const float* e_gamma; const float* e_beta; float* e_hhi; float* e_hlo;
};

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// 2-D fp32 tensor map over a row-major [rows, cols] matrix (row stride ld floats), box = 32 floats x 128 rows, 128-byte swizzle.
static int make_tmap_2d(CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int64_t ld) {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        DBMM_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
        DBMM_CHECK_ARG(p != nullptr && q == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled is not available in this driver");
        fn = (PFN_encodeTiled)p;
    }
    DBMM_CHECK_ARG(((uintptr_t)base & 15) == 0 && ld % 4 == 0, "TMA operands need 16-byte aligned rows (ld=%lld)", (long long)ld);
    const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t gstr[1] = {(cuuint64_t)ld * sizeof(float)};
    const cuuint32_t box[2] = {TG_BK, 128};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    DBMM_CHECK_ARG(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with code %d", (int)r);
    return DBMM_OK;
}

template <bool SPLIT_A>
struct TgCfg {
    static constexpr int TILES = SPLIT_A ? 4 : 3;                       // A(hi) [A lo] B hi B lo
    static constexpr int STAGE_BYTES = TILES * TG_TILE_BYTES;
    static constexpr int STAGES = SPLIT_A ? 3 : 4;
    static constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + 1024 + 256;
};

template <bool SPLIT_A, int EPI>
__global__ void __launch_bounds__(TG_THREADS, 1)
k_tc_gemm_nt(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapAlo,
             const __grid_constant__ CUtensorMap mapBhi, const __grid_constant__ CUtensorMap mapBlo, TcGemmArgs a) {
    using Cfg = TgCfg<SPLIT_A>;
    constexpr int S = Cfg::STAGES;
    extern __shared__ uint8_t tg_smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)tg_smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* full = (uint64_t*)(smem + (size_t)S * Cfg::STAGE_BYTES);
    uint64_t* empty = full + S;
    uint64_t* tmem_full = empty + S;                                  // [2] accumulator stage ready for the epilogue
    uint64_t* tmem_empty = tmem_full + 2;                             // [2] accumulator stage drained by the epilogue
    uint32_t* tmem_ptr = (uint32_t*)(tmem_empty + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_ntiles = (a.N + TG_BN - 1) / TG_BN, n_mtiles = (a.M + TG_BM - 1) / TG_BM;
    const int total_tiles = n_ntiles * n_mtiles;                      // column tiles fastest: CTAs sharing an A row tile run together
    const int KB = (a.K + TG_BK - 1) / TG_BK;                         // the K tail is zero-filled by TMA

    if (tid == 0) {
        for (int s = 0; s < S; ++s) { ptx::mbar_init(&full[s], 1); ptx::mbar_init(&empty[s], 1); }
        for (int s = 0; s < 2; ++s) { ptx::mbar_init(&tmem_full[s], 1); ptx::mbar_init(&tmem_empty[s], 128); }
        ptx::fence_mbar_init();
        ptx::tma_prefetch_desc(&mapA); ptx::tma_prefetch_desc(&mapBhi); ptx::tma_prefetch_desc(&mapBlo);
        if (SPLIT_A) ptx::tma_prefetch_desc(&mapAlo);
    }
    if (warp == 0) ptx::tmem_alloc<2 * TG_BN>(tmem_ptr);
    ptx::tc_fence_before_sync();
    __syncthreads();
    ptx::tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        // ===================== TMA producer: one thread =====================
        if (lane == 0) {
            uint32_t g = 0;                                           // k-blocks issued so far (ring position across tiles)
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int m0 = (tile / n_ntiles) * TG_BM, n0 = (tile % n_ntiles) * TG_BN;
            for (int kb = 0; kb < KB; ++kb, ++g) {
                const int s = g % S;
                ptx::mbar_wait(&empty[s], ((g / S) & 1) ^ 1);
                uint8_t* st = smem + (size_t)s * Cfg::STAGE_BYTES;
                ptx::mbar_arrive_expect_tx(&full[s], Cfg::STAGE_BYTES);
                const int k0 = kb * TG_BK;
                int t = 0;
                ptx::tma_load_2d(&mapA, &full[s], st + (t++) * TG_TILE_BYTES, k0, m0);
                if (SPLIT_A) ptx::tma_load_2d(&mapAlo, &full[s], st + (t++) * TG_TILE_BYTES, k0, m0);
                ptx::tma_load_2d(&mapBhi, &full[s], st + (t++) * TG_TILE_BYTES, k0, n0);
                ptx::tma_load_2d(&mapBlo, &full[s], st + (t++) * TG_TILE_BYTES, k0, n0);
            }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer: one thread =====================
        if (lane == 0) {
            uint32_t g = 0, it = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
            const int n0 = (tile % n_ntiles) * TG_BN;
            int ncols = (a.N - n0 + 15) & ~15;                        // a narrow last column tile issues narrow MMAs
            if (ncols > TG_BN) ncols = TG_BN;
            const uint32_t idesc = ptx::umma_idesc(/*tf32*/ 2, TG_BM, ncols, 0, 0);
            const uint32_t as = it & 1u;
            const uint32_t tmem_acc = tmem_base + as * TG_BN;
            ptx::mbar_wait(&tmem_empty[as], ((it >> 1) & 1) ^ 1);     // the epilogue has drained this accumulator stage
            ptx::tc_fence_after_sync();
            for (int kb = 0; kb < KB; ++kb, ++g) {
                const int s = g % S;
                ptx::mbar_wait(&full[s], (g / S) & 1);
                ptx::tc_fence_after_sync();
                const uint32_t base = ptx::smem_u32(smem + (size_t)s * Cfg::STAGE_BYTES);
                const uint32_t sAhi = base, sAlo = base + TG_TILE_BYTES;
                const uint32_t sBhi = base + (SPLIT_A ? 2 : 1) * TG_TILE_BYTES, sBlo = sBhi + TG_TILE_BYTES;
#pragma unroll
                for (int kk = 0; kk < TG_BK / 8; ++kk) {
                    const uint64_t ahi = ptx::umma_smem_desc(sAhi + kk * 32, 0, 1024);
                    const uint64_t bhi = ptx::umma_smem_desc(sBhi + kk * 32, 0, 1024);
                    const uint64_t blo = ptx::umma_smem_desc(sBlo + kk * 32, 0, 1024);
                    ptx::mma_tf32_ss(tmem_acc, ahi, bhi, idesc, (kb | kk) != 0 ? 1u : 0u);
                    ptx::mma_tf32_ss(tmem_acc, ahi, blo, idesc, 1u);
                    if (SPLIT_A) {
                        const uint64_t alo = ptx::umma_smem_desc(sAlo + kk * 32, 0, 1024);
                        ptx::mma_tf32_ss(tmem_acc, alo, bhi, idesc, 1u);
                    }
                }
                ptx::mma_commit(&empty[s]);
            }
            ptx::mma_commit(&tmem_full[as]);
            }
        }
        __syncwarp();
    } else {
        // ===================== epilogue: warps 2..5, TMEM lane quarter = warp % 4 =====================
        const int q = warp & 3;
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const int nt = tile % n_ntiles, m0 = (tile / n_ntiles) * TG_BM, n0 = nt * TG_BN;
        const uint32_t as = it & 1u;
        const uint32_t tmem_acc = tmem_base + as * TG_BN;
        ptx::mbar_wait(&tmem_full[as], (it >> 1) & 1);
        ptx::tc_fence_after_sync();
        const int m = m0 + q * 32 + lane;
        const bool row_ok = m < a.M;
        float scale = a.scale;
        if (a.scale_dev) scale *= __ldg(a.scale_dev);
        if (a.rowscale && row_ok) scale *= __ldg(a.rowscale + m);
        if (EPI == EPI_STORE) {
            float* crow = a.C + (size_t)(row_ok ? m : 0) * a.ldc + n0;
#pragma unroll 1
            for (int ch = 0; ch < TG_BN / 32; ++ch) {
                uint32_t r[32];
                ptx::tmem_ld_32x32b_x32(tmem_acc + ((uint32_t)(q * 32) << 16) + ch * 32, r);
                ptx::tmem_ld_wait();
                if (!row_ok) continue;
                const int nb = n0 + ch * 32;
                if (nb + 32 <= a.N && (a.ldc & 3) == 0) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        float4* dst = reinterpret_cast<float4*>(crow + ch * 32 + j);
                        float4 v = make_float4(scale * __uint_as_float(r[j]), scale * __uint_as_float(r[j + 1]),
                                               scale * __uint_as_float(r[j + 2]), scale * __uint_as_float(r[j + 3]));
                        if (a.accumulate) { const float4 o = *dst; v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
                        *dst = v;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (nb + j < a.N) crow[ch * 32 + j] = scale * __uint_as_float(r[j]) + (a.accumulate ? crow[ch * 32 + j] : 0.f);
                }
            }
        } else if (EPI == EPI_EVAL_H) {
#pragma unroll 1
            for (int ch = 0; ch < TG_BN / 32; ++ch) {
                uint32_t r[32];
                ptx::tmem_ld_32x32b_x32(tmem_acc + ((uint32_t)(q * 32) << 16) + ch * 32, r);
                ptx::tmem_ld_wait();
                const int nb = n0 + ch * 32;
                if (!row_ok || nb >= a.N) continue;
                float hi[32], lo[32];
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4) {
                    const float4 b1 = __ldg(reinterpret_cast<const float4*>(a.e_b1 + nb) + j4);
                    const float4 mu = __ldg(reinterpret_cast<const float4*>(a.e_mean + nb) + j4);
                    const float4 va = __ldg(reinterpret_cast<const float4*>(a.e_var + nb) + j4);
                    const float4 ga = __ldg(reinterpret_cast<const float4*>(a.e_gamma + nb) + j4);
                    const float4 be = __ldg(reinterpret_cast<const float4*>(a.e_beta + nb) + j4);
                    const float b1x[4] = {b1.x, b1.y, b1.z, b1.w}, mux[4] = {mu.x, mu.y, mu.z, mu.w}, vax[4] = {va.x, va.y, va.z, va.w};
                    const float gax[4] = {ga.x, ga.y, ga.z, ga.w}, bex[4] = {be.x, be.y, be.z, be.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int j = 4 * j4 + e;
                        const float ah = (__uint_as_float(r[j]) + b1x[e] - mux[e]) * (1.0f / sqrtf(vax[e] + DBMM_BN_EPS));
                        const float h = fmaxf(fmaf(ah, gax[e], bex[e]), 0.f);
                        hi[j] = __uint_as_float(__float_as_uint(h) & 0xffffe000u);
                        lo[j] = h - hi[j];
                    }
                }
                const size_t off = (size_t)m * a.N + nb;
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    *reinterpret_cast<float4*>(a.e_hhi + off + j) = make_float4(hi[j], hi[j + 1], hi[j + 2], hi[j + 3]);
                    *reinterpret_cast<float4*>(a.e_hlo + off + j) = make_float4(lo[j], lo[j + 1], lo[j + 2], lo[j + 3]);
                }
            }
        } else if (EPI == EPI_HSPACE) {
            float dot = 0.f;
#pragma unroll 1
            for (int ch = 0; ch < TG_BN / 32; ++ch) {
                uint32_t r[32];
                ptx::tmem_ld_32x32b_x32(tmem_acc + ((uint32_t)(q * 32) << 16) + ch * 32, r);
                ptx::tmem_ld_wait();
                if (!row_ok) continue;
                const int nb = n0 + ch * 32;
                if (nb >= a.N) continue;
                if (nt == 0) {
                    const float4* hh = reinterpret_cast<const float4*>(a.hs_hi + (size_t)m * a.hs_ld + nb);
                    const float4* hl = reinterpret_cast<const float4*>(a.hs_lo + (size_t)m * a.hs_ld + nb);
                    const float4* gb = reinterpret_cast<const float4*>(a.hs_bias + nb);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 x = __ldcg(hh + j), y4 = __ldcg(hl + j), b = __ldg(gb + j);
                        dot = fmaf(__uint_as_float(r[4 * j + 0]) + b.x, x.x + y4.x, dot);
                        dot = fmaf(__uint_as_float(r[4 * j + 1]) + b.y, x.y + y4.y, dot);
                        dot = fmaf(__uint_as_float(r[4 * j + 2]) + b.z, x.z + y4.z, dot);
                        dot = fmaf(__uint_as_float(r[4 * j + 3]) + b.w, x.w + y4.w, dot);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (nb + j < a.N) a.tail[(size_t)m * a.tail_ld + (nb - TG_BN) + j] = __uint_as_float(r[j]) + __ldg(a.hs_bias + nb + j);
                }
            }
            if (row_ok && nt == 0) a.rowdot[m] = dot;
        } else {
            int yv = -1;
            if (row_ok && a.y) {
                const int64_t pos = a.pos0 + m;
                yv = a.y[a.idx ? (int64_t)a.idx[pos] : pos];
            }
            float mx = -INFINITY, se = 0.f, ly = -INFINITY; int am = 0;
#pragma unroll 1
            for (int ch = 0; ch < TG_BN / 32; ++ch) {
                uint32_t r[32];
                ptx::tmem_ld_32x32b_x32(tmem_acc + ((uint32_t)(q * 32) << 16) + ch * 32, r);
                ptx::tmem_ld_wait();
                const int nb = n0 + ch * 32;
                float cmx = -INFINITY; int cam = 0;
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float l = (nb + j < a.N) ? fmaf(scale, __uint_as_float(r[j]), a.col_bias ? __ldg(a.col_bias + nb + j) : 0.f) : -INFINITY;
                    r[j] = __float_as_uint(l);
                    if (l > cmx) { cmx = l; cam = nb + j; }           // strict >: first maximum wins, as torch.argmax
                    if (nb + j == yv) ly = l;
                }
                if (cmx > mx) { se *= expf(mx - cmx); mx = cmx; am = cam; }      // exp(-inf) = 0 on the first chunk
                if (mx > -INFINITY) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) se += expf(__uint_as_float(r[j]) - mx);
                }
            }
            if (row_ok) {
                SoftmaxPart p; p.mx = mx; p.se = se; p.ly = ly; p.am = am;
                a.part[(size_t)nt * a.M + m] = p;                      // [column tile][row]: coalesced here and in k_head_finish
            }
        }
        ptx::tc_fence_before_sync();
        ptx::mbar_arrive(&tmem_empty[as]);                            // 128 epilogue threads: accumulator stage reusable
        }
    }
    __syncthreads();
    if (warp == 0) ptx::tmem_dealloc<2 * TG_BN>(tmem_base);
}

template <bool SPLIT_A, int EPI>
static int launch_tc_gemm_nt(const float* A, const float* Alo, int64_t lda, const float* Bhi, const float* Blo, int64_t ldb,
                             const TcGemmArgs& a, cudaStream_t st) {
    using Cfg = TgCfg<SPLIT_A>;
    DBMM_CHECK_ARG(a.M >= 1 && a.N >= 1 && a.K >= 1, "empty GEMM %d x %d x %d", a.M, a.N, a.K);
    CUtensorMap mA, mAlo, mBhi, mBlo;
    if (int rc = make_tmap_2d(&mA, A, a.M, a.K, lda)) return rc;
    if (int rc = make_tmap_2d(&mAlo, SPLIT_A ? Alo : A, a.M, a.K, lda)) return rc;
    if (int rc = make_tmap_2d(&mBhi, Bhi, a.N, a.K, ldb)) return rc;
    if (int rc = make_tmap_2d(&mBlo, Blo, a.N, a.K, ldb)) return rc;
    auto kern = k_tc_gemm_nt<SPLIT_A, EPI>;
    DBMM_CUDA(set_smem(kern, Cfg::SMEM));
    int grid = ceil_div(a.N, TG_BN) * ceil_div(a.M, TG_BM);
    if (grid > 148) grid = 148;                                  // persistent: one CTA per SM walks the tiles
    kern<<<grid, TG_THREADS, Cfg::SMEM, st>>>(mA, mAlo, mBhi, mBlo, a);
    DBMM_LAUNCH_CHECK();
    return DBMM_OK;
}

// w = hi + lo with hi tf32-exact (top 19 bits); also the plain transpose used to make an operand K-major.
__global__ void __launch_bounds__(256) k_split_hi_lo(const float* __restrict__ w, int64_t ld_in, float* __restrict__ hi,
                                                     float* __restrict__ lo, int64_t rows, int64_t cols, int64_t ld_out) {
    const int64_t n = rows * cols;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / cols, c = i - r * cols;
        const float v = w[r * ld_in + c];
        const float h = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
        hi[r * ld_out + c] = h;
        lo[r * ld_out + c] = v - h;
    }
}
// out_hi/out_lo[c][r] = split(in[r][c])   (32 x 32 shared-memory tiles)
__global__ void __launch_bounds__(256) k_transpose_split(const float* __restrict__ in, int64_t ld_in, float* __restrict__ hi,
                                                         float* __restrict__ lo, int rows, int cols, int64_t ld_out) {
    __shared__ float t[32][33];
    const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int i = ty; i < 32; i += 8) {
        const int r = r0 + i, c = c0 + tx;
        t[i][tx] = (r < rows && c < cols) ? in[(size_t)r * ld_in + c] : 0.f;
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
        const int c = c0 + i, r = r0 + tx;
        if (c < cols && r < rows) {
            const float v = t[tx][i];
            const float h = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
            hi[(size_t)c * ld_out + r] = h;
            lo[(size_t)c * ld_out + r] = v - h;
        }
    }
}

}  // namespace dbmm
