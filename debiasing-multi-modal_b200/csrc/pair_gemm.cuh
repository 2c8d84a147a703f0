// Dense "NT" GEMM on kind::f16 tensor cores for ARBITRARY fp32 data, operands as fp16 PAIRS:
//
//     C[m][n] = scale * sum_k A[m][k] * B[n][k],      A = Ah + Al,  B = Bh + Bl   (both K-major)
//
// The caller scales each matrix by a power of two chosen from its largest magnitude (k_absmax_bits / k_head_bscale:
// max |x| 2^k in [2^14, 2^15)) and splits the scaled values into an UNSCALED fp16 pair (k_pair_split*): with the scaling the
// low half is a normal fp16 number for every element within 2^17 of the largest one, so a pair carries 22 significant
// bits -- the same as the tf32 hi + lo pair of tc_gemm.cuh.  Three products Ah*Bh + Ah*Bl + Al*Bh go into ONE fp32 TMEM
// accumulator (the missing Al*Bl is 2^-22 relative); the epilogue multiplies by the inverse powers of two (device scalars).
// Against the 3xTF32 GEMM: the same three MMAs per k-step, each at twice the rate and on half the bytes.
//
// Users: the contrastive regulariser (config 3): S = Z Z^T / tau and the gradient GEMMs dZ = G Z, G^T Z (one GEMM (G + G^T) Z when
// one rank holds the whole batch).  Measured, B = 8192, d = 768: 256 us per GEMM (1.2 PFLOP/s executed, L2 -> SM bound at 64 KB
// per 12 MMAs).  Tried and dropped: the masked log-sum-exp row pass inside the S GEMM's epilogue -- 128 columns of online
// softmax + label compares per thread outlast the tile's MMAs (S GEMM 256 -> 588 us); k_supcon_rows (144 us) stays a kernel.
// Persistent CTAs walk 128 x 128 tiles (column tiles fastest); TMA boxes of 64 halfs x 128 rows (SWIZZLE_128B), 3-stage ring of
// (Ah, Al, Bh, Bl), one TMA thread, one MMA thread, four epilogue warps, accumulator double-buffered in TMEM.
#pragma once
#include "head_f16.cuh"

namespace dbmm {

constexpr int PG_THREADS = 192, PG_BM = 128, PG_BN = 128, PG_BK = 64, PG_STAGES = 3;
constexpr int PG_STAGE_BYTES = 4 * EF_TILE_BYTES;
constexpr size_t PG_SMEM = (size_t)PG_STAGES * PG_STAGE_BYTES + 1024 + 256;

struct PairGemmArgs {
    int M, N, K;
    float scale; const float* sdev[3];        // C = acc * scale * prod(*sdev[i]) (null entries = 1)
    float* C; int64_t ldc; int accumulate;    // C = [C +] scaled tile
};

__global__ void __launch_bounds__(PG_THREADS, 1)
k_f16_pair_gemm(const __grid_constant__ CUtensorMap mapAh, const __grid_constant__ CUtensorMap mapAl,
                const __grid_constant__ CUtensorMap mapBh, const __grid_constant__ CUtensorMap mapBl, PairGemmArgs a) {
    constexpr int S = PG_STAGES;
    extern __shared__ uint8_t pg_smem_raw[];
    uint8_t* smem = pg_smem_raw + ((1024u - (ptx::smem_u32(pg_smem_raw) & 1023u)) & 1023u);
    uint64_t* full = (uint64_t*)(smem + (size_t)S * PG_STAGE_BYTES);
    uint64_t* empty = full + S;
    uint64_t* tmem_full = empty + S;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_ptr = (uint32_t*)(tmem_empty + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_ntiles = (a.N + PG_BN - 1) / PG_BN, n_mtiles = (a.M + PG_BM - 1) / PG_BM;
    const int total_tiles = n_ntiles * n_mtiles;
    const int KB = (a.K + PG_BK - 1) / PG_BK;                           // the K tail is zero-filled by TMA

    if (tid == 0) {
        for (int s = 0; s < S; ++s) { ptx::mbar_init(&full[s], 1); ptx::mbar_init(&empty[s], 1); }
        for (int s = 0; s < 2; ++s) { ptx::mbar_init(&tmem_full[s], 1); ptx::mbar_init(&tmem_empty[s], 128); }
        ptx::fence_mbar_init();
        ptx::tma_prefetch_desc(&mapAh); ptx::tma_prefetch_desc(&mapAl); ptx::tma_prefetch_desc(&mapBh); ptx::tma_prefetch_desc(&mapBl);
    }
    if (warp == 0) ptx::tmem_alloc<256>(tmem_ptr);
    ptx::tc_fence_before_sync();
    __syncthreads();
    ptx::tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t g = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int m0 = (tile / n_ntiles) * PG_BM, n0 = (tile % n_ntiles) * PG_BN;
                for (int kb = 0; kb < KB; ++kb, ++g) {
                    const int s = g % S;
                    ptx::mbar_wait(&empty[s], ((g / S) & 1) ^ 1);
                    uint8_t* st = smem + (size_t)s * PG_STAGE_BYTES;
                    ptx::mbar_arrive_expect_tx(&full[s], PG_STAGE_BYTES);
                    const int k0 = kb * PG_BK;
                    ptx::tma_load_2d(&mapAh, &full[s], st, k0, m0);
                    ptx::tma_load_2d(&mapAl, &full[s], st + EF_TILE_BYTES, k0, m0);
                    ptx::tma_load_2d(&mapBh, &full[s], st + 2 * EF_TILE_BYTES, k0, n0);
                    ptx::tma_load_2d(&mapBl, &full[s], st + 3 * EF_TILE_BYTES, k0, n0);
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {
            uint32_t g = 0, it = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
                const int n0 = (tile % n_ntiles) * PG_BN;
                int ncols = (a.N - n0 + 15) & ~15;                    // a narrow last column tile issues narrow MMAs
                if (ncols > PG_BN) ncols = PG_BN;
                const uint32_t idesc = ptx::umma_idesc(/*f16*/ 0, PG_BM, ncols, 0, 0);
                const uint32_t as = it & 1u;
                const uint32_t acc = tmem_base + as * PG_BN;
                ptx::mbar_wait(&tmem_empty[as], ((it >> 1) & 1) ^ 1);
                ptx::tc_fence_after_sync();
                for (int kb = 0; kb < KB; ++kb, ++g) {
                    const int s = g % S;
                    ptx::mbar_wait(&full[s], (g / S) & 1);
                    ptx::tc_fence_after_sync();
                    const uint32_t base = ptx::smem_u32(smem + (size_t)s * PG_STAGE_BYTES);
#pragma unroll
                    for (int kk = 0; kk < PG_BK / 16; ++kk) {
                        const uint64_t ah = ptx::umma_smem_desc(base + kk * 32, 0, 1024);
                        const uint64_t al = ptx::umma_smem_desc(base + EF_TILE_BYTES + kk * 32, 0, 1024);
                        const uint64_t bh = ptx::umma_smem_desc(base + 2 * EF_TILE_BYTES + kk * 32, 0, 1024);
                        const uint64_t bl = ptx::umma_smem_desc(base + 3 * EF_TILE_BYTES + kk * 32, 0, 1024);
                        ptx::mma_f16_ss(acc, ah, bh, idesc, (kb | kk) != 0 ? 1u : 0u);
                        ptx::mma_f16_ss(acc, ah, bl, idesc, 1u);
                        ptx::mma_f16_ss(acc, al, bh, idesc, 1u);
                    }
                    ptx::mma_commit(&empty[s]);
                }
                ptx::mma_commit(&tmem_full[as]);
            }
        }
        __syncwarp();
    } else {
        const int q = warp & 3;
        float scale = a.scale;
#pragma unroll
        for (int i = 0; i < 3; ++i) if (a.sdev[i]) scale *= __ldg(a.sdev[i]);
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
            const int m0 = (tile / n_ntiles) * PG_BM, n0 = (tile % n_ntiles) * PG_BN;
            const uint32_t as = it & 1u;
            const uint32_t acc = tmem_base + as * PG_BN + ((uint32_t)(q * 32) << 16);
            ptx::mbar_wait(&tmem_full[as], (it >> 1) & 1);
            ptx::tc_fence_after_sync();
            const int m = m0 + q * 32 + lane;
            const bool row_ok = m < a.M;
            float* crow = a.C + (size_t)(row_ok ? m : 0) * a.ldc + n0;
#pragma unroll 1
            for (int ch = 0; ch < PG_BN / 32; ++ch) {
                const int nb = n0 + ch * 32;
                if (nb >= a.N) break;
                uint32_t r[32];
                ptx::tmem_ld_32x32b_x32(acc + ch * 32, r);
                ptx::tmem_ld_wait();
                if (!row_ok) continue;
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
            ptx::tc_fence_before_sync();
            ptx::mbar_arrive(&tmem_empty[as]);
        }
    }
    __syncthreads();
    if (warp == 0) ptx::tmem_dealloc<256>(tmem_base);
}

static int launch_pair_gemm(const __half* Ah, const __half* Al, int64_t lda, const __half* Bh, const __half* Bl, int64_t ldb,
                            const PairGemmArgs& a, cudaStream_t st) {
    DBMM_CHECK_ARG(a.M >= 1 && a.N >= 1 && a.K >= 1, "empty GEMM %d x %d x %d", a.M, a.N, a.K);
    CUtensorMap mAh, mAl, mBh, mBl;
    if (int rc = make_tmap_2d_f16(&mAh, Ah, a.M, a.K, lda)) return rc;
    if (int rc = make_tmap_2d_f16(&mAl, Al, a.M, a.K, lda)) return rc;
    if (int rc = make_tmap_2d_f16(&mBh, Bh, a.N, a.K, ldb)) return rc;
    if (int rc = make_tmap_2d_f16(&mBl, Bl, a.N, a.K, ldb)) return rc;
    DBMM_CUDA(set_smem(k_f16_pair_gemm, PG_SMEM));
    int grid = ceil_div(a.N, PG_BN) * ceil_div(a.M, PG_BM);
    if (grid > 148) grid = 148;
    k_f16_pair_gemm<<<grid, PG_THREADS, PG_SMEM, st>>>(mAh, mAl, mBh, mBl, a);
    DBMM_LAUNCH_CHECK();
    return DBMM_OK;
}

// (hi, lo)[r][c] = unscaled fp16 pair of in[r][c] * sc[0];  ld_out in halfs (multiple of 8); columns [cols, ld_out) zero-filled
__global__ void __launch_bounds__(256) k_pair_split(const float* __restrict__ in, int64_t ld_in, __half* __restrict__ hi,
                                                    __half* __restrict__ lo, int64_t rows, int cols, int64_t ld_out,
                                                    const float* __restrict__ sc) {
    const float s2k = sc[0];
    const int64_t n = rows * ld_out;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / ld_out; const int c = (int)(i - r * ld_out);
        const float v = c < cols ? in[r * ld_in + c] * s2k : 0.f;
        const __half h = __float2half_rn(v);
        hi[i] = h;
        lo[i] = __float2half_rn(v - __half2float(h));
    }
}

// One pass over in [rows, cols]: the pair row-major (hi, lo: ld_out halfs; may be null) AND transposed (hiT, loT [cols][ldT]).
// 32 x 32 shared-memory tiles; padding columns of the outputs are left untouched (TMA never reads past the logical extent).
__global__ void __launch_bounds__(256) k_pair_split_both(const float* __restrict__ in, int64_t ld_in, __half* __restrict__ hi,
                                                         __half* __restrict__ lo, int64_t ld_out, __half* __restrict__ hiT,
                                                         __half* __restrict__ loT, int64_t ldT, int rows, int cols,
                                                         const float* __restrict__ sc) {
    __shared__ float t[32][33];
    const float s2k = sc[0];
    const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int i = ty; i < 32; i += 8) {
        const int r = r0 + i, c = c0 + tx;
        const float v = (r < rows && c < cols) ? in[(size_t)r * ld_in + c] * s2k : 0.f;
        t[i][tx] = v;
        if (hi && r < rows && c < cols) {
            const __half h = __float2half_rn(v);
            hi[(size_t)r * ld_out + c] = h;
            lo[(size_t)r * ld_out + c] = __float2half_rn(v - __half2float(h));
        }
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
        const int c = c0 + i, r = r0 + tx;
        if (c < cols && r < rows) {
            const float v = t[tx][i];
            const __half h = __float2half_rn(v);
            hiT[(size_t)c * ldT + r] = h;
            loT[(size_t)c * ldT + r] = __float2half_rn(v - __half2float(h));
        }
    }
}

// Whole-batch contrastive backward: W = G + G^T (anchor role + contrast role of every row in one matrix), as an unscaled fp16
// pair of W * sc[0], row-major [n][ld_out].  Tile (bi, bj) reads G tiles (bi, bj) and (bj, bi) (the latter through shared memory).
__global__ void __launch_bounds__(256) k_pair_split_sym(const float* __restrict__ G, int64_t ldg, __half* __restrict__ hi,
                                                        __half* __restrict__ lo, int64_t ld_out, int n, const float* __restrict__ sc) {
    // 64 x 64 tiles, a thread owns two adjacent columns: 256-byte row reads, 128-byte row writes per array (ldg, ld_out even)
    __shared__ float t[64][65];
    const float s2k = sc[0];
    const int r0 = blockIdx.y * 64, c0 = blockIdx.x * 64;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int i = ty; i < 64; i += 8) {                               // transposed partner tile: rows c0.., columns r0..
        const int r = c0 + i, c = r0 + 2 * tx;
        float2 v = make_float2(0.f, 0.f);
        if (r < n && c + 1 < n) v = *reinterpret_cast<const float2*>(G + (size_t)r * ldg + c);
        else if (r < n && c < n) v.x = G[(size_t)r * ldg + c];
        t[i][2 * tx] = v.x; t[i][2 * tx + 1] = v.y;
    }
    __syncthreads();
    for (int i = ty; i < 64; i += 8) {
        const int r = r0 + i, c = c0 + 2 * tx;
        if (r >= n || c >= n) continue;
        const bool two = c + 1 < n;
        float2 g = make_float2(0.f, 0.f);
        if (two) g = *reinterpret_cast<const float2*>(G + (size_t)r * ldg + c);
        else g.x = G[(size_t)r * ldg + c];
        const float v0 = (g.x + t[2 * tx][i]) * s2k, v1 = (g.y + t[2 * tx + 1][i]) * s2k;
        const __half h0 = __float2half_rn(v0), h1 = __float2half_rn(v1);
        const __half l0 = __float2half_rn(v0 - __half2float(h0)), l1 = __float2half_rn(v1 - __half2float(h1));
        if (two) {
            *reinterpret_cast<__half2*>(hi + (size_t)r * ld_out + c) = __halves2half2(h0, h1);
            *reinterpret_cast<__half2*>(lo + (size_t)r * ld_out + c) = __halves2half2(l0, l1);
        } else { hi[(size_t)r * ld_out + c] = h0; lo[(size_t)r * ld_out + c] = l0; }
    }
}

}  // namespace dbmm
