// Zero-shot head over an fp16-resident embedding matrix (validate_zs on raw embeddings, final_main.py:757-768; BASELINE
// config 4: up to 10 M rows against 1,000 prompt columns):
//
//     logits[m][c] = inv_tau * (x_m / |x_m|) . That_c  (+ bias_c)        -> CE, argmax, per-group counters; logits never stored
//
// CLIP embeddings are fp16-valued and the packed store keeps them as fp16 (pack.py), so the A operand is read as stored:
// 2 bytes per element from HBM, no split.  The prompt matrix enters as an fp16 PAIR That = Th + 2^-11 Tl (Tl pre-scaled by
// 2^11 so that it stays out of the fp16 subnormals: 22 significant bits, the same as the tf32 hi + lo pair of tc_gemm.cuh) and
// the two products accumulate in TWO fp32 TMEM accumulators, combined in the epilogue.  kind::f16 issues at twice the
// kind::tf32 rate and a stage moves half the bytes: 4x fewer tensor-pipe cycles per algorithmic flop than the tf32 head.
//
// Persistent CTAs (one per SM) walk 128 x 128 output tiles, column tiles fastest (the CTAs that share an A row tile run at
// the same time, so it comes from HBM once); TMA boxes of 64 halfs x 128 rows (SWIZZLE_128B), 4-stage ring, one TMA thread,
// one MMA thread, four epilogue warps; the accumulator pair is double-buffered (2 x 2 x 128 columns = all of TMEM) so the
// online-softmax epilogue of tile i overlaps the MMAs of tile i + 1.
#pragma once
#include "eval_f16.cuh"

namespace dbmm {

constexpr int HF_THREADS = 192, HF_BM = 128, HF_BN = 128, HF_BK = 64, HF_STAGES = 4;
constexpr int HF_STAGE_BYTES = 3 * EF_TILE_BYTES;                       // x, Th, Tl
constexpr size_t HF_SMEM = (size_t)HF_STAGES * HF_STAGE_BYTES + 1024 + 256;

struct HeadF16Args {
    int64_t M; int N, K;                      // rows, prompt columns, embedding width
    float scale;                              // inv_tau
    const float* rowscale;                    // optional [M]: 1 / |x_m|
    const float* col_bias;                    // optional [N]
    const int32_t* y; int64_t pos0;           // target column of row pos0 + m (or null)
    SoftmaxPart* part; int n_ntiles;          // [n_ntiles][M]
    const float* bscale_inv;                  // wide kernel: 2^-k of the prompt scaling (device scalar)
    // wide kernel, fused row norms (norm_out != nullptr): two extra warps square-sum the x tiles of every column-tile-0 pass out of
    // the TMA stages (no second pass over the matrix), write 1 / |x_m| to norm_out and raise norm_flag[row tile]; the epilogues of
    // the row tile's column tiles (other CTAs, the same round of the tile walk) wait for that flag.  rowscale is then ignored.
    float* norm_out; int* norm_flag;          // [M]; [ceil(M / 256)], zero before the launch
    const __half* X; int64_t ldx;             // the rows again: an epilogue thread whose flag does not arrive within 200 us (the CTA that
                                              // walks the row tile's first column tile is not resident: GPU shared with other work)
                                              // computes the norm of its own row from global memory instead of waiting any longer
};

__global__ void __launch_bounds__(HF_THREADS, 1)
k_f16_head(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapBhi,
           const __grid_constant__ CUtensorMap mapBlo, HeadF16Args a) {
    constexpr int S = HF_STAGES;
    extern __shared__ uint8_t hf_smem_raw[];
    uint8_t* smem = hf_smem_raw + ((1024u - (ptx::smem_u32(hf_smem_raw) & 1023u)) & 1023u);
    uint64_t* full = (uint64_t*)(smem + (size_t)S * HF_STAGE_BYTES);
    uint64_t* empty = full + S;
    uint64_t* tmem_full = empty + S;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_ptr = (uint32_t*)(tmem_empty + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_ntiles = a.n_ntiles;
    const int64_t total_tiles = ((a.M + HF_BM - 1) / HF_BM) * n_ntiles;
    const int KB = (a.K + HF_BK - 1) / HF_BK;                           // the K tail is zero-filled by TMA

    if (tid == 0) {
        for (int s = 0; s < S; ++s) { ptx::mbar_init(&full[s], 1); ptx::mbar_init(&empty[s], 1); }
        for (int s = 0; s < 2; ++s) { ptx::mbar_init(&tmem_full[s], 1); ptx::mbar_init(&tmem_empty[s], 128); }
        ptx::fence_mbar_init();
        ptx::tma_prefetch_desc(&mapA); ptx::tma_prefetch_desc(&mapBhi); ptx::tma_prefetch_desc(&mapBlo);
    }
    if (warp == 0) ptx::tmem_alloc<512>(tmem_ptr);
    ptx::tc_fence_before_sync();
    __syncthreads();
    ptx::tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        // ===================== TMA producer: one thread =====================
        if (lane == 0) {
            uint32_t g = 0;
            for (int64_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int m0 = (int)(tile / n_ntiles) * HF_BM, n0 = (int)(tile % n_ntiles) * HF_BN;
                for (int kb = 0; kb < KB; ++kb, ++g) {
                    const int s = g % S;
                    ptx::mbar_wait(&empty[s], ((g / S) & 1) ^ 1);
                    uint8_t* st = smem + (size_t)s * HF_STAGE_BYTES;
                    ptx::mbar_arrive_expect_tx(&full[s], HF_STAGE_BYTES);
                    const int k0 = kb * HF_BK;
                    ptx::tma_load_2d(&mapA, &full[s], st, k0, m0);
                    ptx::tma_load_2d(&mapBhi, &full[s], st + EF_TILE_BYTES, k0, n0);
                    ptx::tma_load_2d(&mapBlo, &full[s], st + 2 * EF_TILE_BYTES, k0, n0);
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer: one thread =====================
        if (lane == 0) {
            uint32_t g = 0, it = 0;
            for (int64_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
                const int n0 = (int)(tile % n_ntiles) * HF_BN;
                int ncols = (a.N - n0 + 15) & ~15;                    // a narrow last column tile issues narrow MMAs
                if (ncols > HF_BN) ncols = HF_BN;
                const uint32_t idesc = ptx::umma_idesc(/*f16*/ 0, HF_BM, ncols, 0, 0);
                const uint32_t as = it & 1u;
                const uint32_t acc0 = tmem_base + as * 256, acc1 = acc0 + 128;
                ptx::mbar_wait(&tmem_empty[as], ((it >> 1) & 1) ^ 1);
                ptx::tc_fence_after_sync();
                for (int kb = 0; kb < KB; ++kb, ++g) {
                    const int s = g % S;
                    ptx::mbar_wait(&full[s], (g / S) & 1);
                    ptx::tc_fence_after_sync();
                    const uint32_t base = ptx::smem_u32(smem + (size_t)s * HF_STAGE_BYTES);
#pragma unroll
                    for (int kk = 0; kk < HF_BK / 16; ++kk) {
                        const uint64_t ax = ptx::umma_smem_desc(base + kk * 32, 0, 1024);
                        const uint64_t bhi = ptx::umma_smem_desc(base + EF_TILE_BYTES + kk * 32, 0, 1024);
                        const uint64_t blo = ptx::umma_smem_desc(base + 2 * EF_TILE_BYTES + kk * 32, 0, 1024);
                        const uint32_t acc = (kb | kk) != 0 ? 1u : 0u;
                        ptx::mma_f16_ss(acc0, ax, bhi, idesc, acc);
                        ptx::mma_f16_ss(acc1, ax, blo, idesc, acc);
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
        for (int64_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
            const int nt = (int)(tile % n_ntiles), n0 = nt * HF_BN;
            const int64_t m = (tile / n_ntiles) * HF_BM + q * 32 + lane;
            const bool row_ok = m < a.M;
            const uint32_t as = it & 1u;
            const uint32_t acc0 = tmem_base + as * 256 + ((uint32_t)(q * 32) << 16), acc1 = acc0 + 128;
            float scale = a.scale;
            int yv = -1;
            if (row_ok) {
                if (a.rowscale) scale *= __ldg(a.rowscale + m);
                if (a.y) yv = a.y[a.pos0 + m];
            }
            ptx::mbar_wait(&tmem_full[as], (it >> 1) & 1);
            ptx::tc_fence_after_sync();
            float mx = -INFINITY, se = 0.f, ly = -INFINITY; int am = 0;
#pragma unroll 1
            for (int ch = 0; ch < HF_BN / 32; ++ch) {
                const int nb = n0 + ch * 32;
                if (nb >= a.N) break;                                 // (uniform over the CTA: columns past N were never computed)
                uint32_t r0[32], r1[32];
                ptx::tmem_ld_32x32b_x32(acc0 + ch * 32, r0);
                ptx::tmem_ld_32x32b_x32(acc1 + ch * 32, r1);
                ptx::tmem_ld_wait();
                float cmx = -INFINITY; int cam = 0;
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float dot = fmaf(__uint_as_float(r1[j]), EF_LO_INV, __uint_as_float(r0[j]));
                    const float l = (nb + j < a.N) ? fmaf(scale, dot, a.col_bias ? __ldg(a.col_bias + nb + j) : 0.f) : -INFINITY;
                    r0[j] = __float_as_uint(l);
                    if (l > cmx) { cmx = l; cam = nb + j; }           // strict >: first maximum wins, as torch.argmax
                    if (nb + j == yv) ly = l;
                }
                if (cmx > mx) { se *= expf(mx - cmx); mx = cmx; am = cam; }      // exp(-inf) = 0 on the first chunk
                if (mx > -INFINITY) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) se += expf(__uint_as_float(r0[j]) - mx);
                }
            }
            ptx::tc_fence_before_sync();
            ptx::mbar_arrive(&tmem_empty[as]);                        // 128 epilogue threads: accumulator pair reusable
            if (row_ok) {
                SoftmaxPart p; p.mx = mx; p.se = se; p.ly = ly; p.am = am;
                a.part[(size_t)nt * a.M + m] = p;                  // [column tile][row]: coalesced here and in k_head_finish
            }
        }
    }
    __syncthreads();
    if (warp == 0) ptx::tmem_dealloc<512>(tmem_base);
}

// ------------------------------------------------------------------------------------------------
// Wide variant: 256-row tiles, ONE accumulator per 128-row block.  The prompt matrix is scaled by a power of two 2^k chosen
// from its largest magnitude (max |T| 2^k in [2^14, 2^15): k_head_bscale) and split into an UNSCALED fp16 pair
// T 2^k = Th + Tl; with the scaling Tl is a normal fp16 number for every component within 2^17 of the largest one, so the
// pair still carries 22 significant bits, and both products x * Th and x * Tl add into the SAME fp32 TMEM accumulator (the
// epilogue multiplies by 2^-k).  (kind::f16 with a = f16, b = bf16 would avoid the scaling, but the tensor core rejects
// mixed a / b formats: illegal instruction.)  A stage then holds two x blocks and the (Th, Tl) pair (64 KB): the prompt
// tiles, which every CTA streams from L2, are read once per 256 rows instead of once per 128 -- the 128-row kernel above is
// L2 -> SM bandwidth bound (48 KB per 8 MMAs = 174 GB/s per SM at full tensor rate; this one needs 116).  Two accumulator
// blocks x double buffering = 512 TMEM columns; 8 epilogue warps (TMEM lane quarter = warp % 4, row block = (warp - 2) / 4);
// two more warps compute the row norms out of the TMA stages (HeadF16Args::norm_out).
// ------------------------------------------------------------------------------------------------
constexpr int HWD_THREADS = 384, HWD_BM = 256, HWD_STAGES = 3;          // TMA, MMA, 8 epilogue warps, 2 row-norm warps
constexpr int HWD_STAGE_BYTES = 4 * EF_TILE_BYTES;                       // x block 0, x block 1, Th, Tl
constexpr size_t HWD_SMEM = (size_t)HWD_STAGES * HWD_STAGE_BYTES + 1024 + 256;

__global__ void __launch_bounds__(HWD_THREADS, 1)
k_f16_head_wide(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapBhi,
                const __grid_constant__ CUtensorMap mapBlo, HeadF16Args a) {
    constexpr int S = HWD_STAGES;
    extern __shared__ uint8_t hf_smem_raw[];
    uint8_t* smem = hf_smem_raw + ((1024u - (ptx::smem_u32(hf_smem_raw) & 1023u)) & 1023u);
    uint64_t* full = (uint64_t*)(smem + (size_t)S * HWD_STAGE_BYTES);
    uint64_t* empty = full + S;
    uint64_t* tmem_full = empty + S;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_ptr = (uint32_t*)(tmem_empty + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_ntiles = a.n_ntiles;
    const int64_t total_tiles = ((a.M + HWD_BM - 1) / HWD_BM) * n_ntiles;
    const int KB = (a.K + HF_BK - 1) / HF_BK;

    if (tid == 0) {
        for (int s = 0; s < S; ++s) { ptx::mbar_init(&full[s], 1); ptx::mbar_init(&empty[s], 3); }      // MMA commit + the two norm warps
        for (int s = 0; s < 2; ++s) { ptx::mbar_init(&tmem_full[s], 1); ptx::mbar_init(&tmem_empty[s], 256); }
        ptx::fence_mbar_init();
        ptx::tma_prefetch_desc(&mapA); ptx::tma_prefetch_desc(&mapBhi); ptx::tma_prefetch_desc(&mapBlo);
    }
    if (warp == 0) ptx::tmem_alloc<512>(tmem_ptr);
    ptx::tc_fence_before_sync();
    __syncthreads();
    ptx::tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t g = 0;
            for (int64_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int64_t m0 = (tile / n_ntiles) * HWD_BM;
                const int n0 = (int)(tile % n_ntiles) * HF_BN;
                const int m1 = (int)(m0 + 128 < a.M ? m0 + 128 : m0);   // a tile whose second block lies past M re-reads the first
                for (int kb = 0; kb < KB; ++kb, ++g) {
                    const int s = g % S;
                    ptx::mbar_wait(&empty[s], ((g / S) & 1) ^ 1);
                    uint8_t* st = smem + (size_t)s * HWD_STAGE_BYTES;
                    ptx::mbar_arrive_expect_tx(&full[s], HWD_STAGE_BYTES);
                    const int k0 = kb * HF_BK;
                    ptx::tma_load_2d(&mapA, &full[s], st, k0, (int)m0);
                    ptx::tma_load_2d(&mapA, &full[s], st + EF_TILE_BYTES, k0, m1);
                    ptx::tma_load_2d(&mapBhi, &full[s], st + 2 * EF_TILE_BYTES, k0, n0);
                    ptx::tma_load_2d(&mapBlo, &full[s], st + 3 * EF_TILE_BYTES, k0, n0);
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {
            uint32_t g = 0, it = 0;
            for (int64_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
                const int n0 = (int)(tile % n_ntiles) * HF_BN;
                int ncols = (a.N - n0 + 15) & ~15;
                if (ncols > HF_BN) ncols = HF_BN;
                const uint32_t idesc = ptx::umma_idesc(/*f16*/ 0, 128, ncols, 0, 0);
                const uint32_t as = it & 1u;
                const uint32_t acc0 = tmem_base + as * 256, acc1 = acc0 + 128;
                ptx::mbar_wait(&tmem_empty[as], ((it >> 1) & 1) ^ 1);
                ptx::tc_fence_after_sync();
                for (int kb = 0; kb < KB; ++kb, ++g) {
                    const int s = g % S;
                    ptx::mbar_wait(&full[s], (g / S) & 1);
                    ptx::tc_fence_after_sync();
                    const uint32_t base = ptx::smem_u32(smem + (size_t)s * HWD_STAGE_BYTES);
#pragma unroll
                    for (int kk = 0; kk < HF_BK / 16; ++kk) {
                        const uint64_t a0 = ptx::umma_smem_desc(base + kk * 32, 0, 1024);
                        const uint64_t a1 = ptx::umma_smem_desc(base + EF_TILE_BYTES + kk * 32, 0, 1024);
                        const uint64_t bhi = ptx::umma_smem_desc(base + 2 * EF_TILE_BYTES + kk * 32, 0, 1024);
                        const uint64_t blo = ptx::umma_smem_desc(base + 3 * EF_TILE_BYTES + kk * 32, 0, 1024);
                        const uint32_t acc = (kb | kk) != 0 ? 1u : 0u;
                        ptx::mma_f16_ss(acc0, a0, bhi, idesc, acc);
                        ptx::mma_f16_ss(acc1, a1, bhi, idesc, acc);
                        ptx::mma_f16_ss(acc0, a0, blo, idesc, 1u);
                        ptx::mma_f16_ss(acc1, a1, blo, idesc, 1u);
                    }
                    ptx::mma_commit(&empty[s]);
                }
                ptx::mma_commit(&tmem_full[as]);
            }
        }
        __syncwarp();
    } else if (warp >= 10) {
        // ===================== row norms: warps 10, 11 (64 threads, 4 rows each: 2 per 128-row block) =====================
        const int t = tid - 320;
        uint32_t g = 0;
        for (int64_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int nt = (int)(tile % n_ntiles);
            const int64_t mt = tile / n_ntiles, m0 = mt * HWD_BM;
            const bool work = a.norm_out != nullptr && nt == 0;
            float ss[4] = {0.f, 0.f, 0.f, 0.f};
            for (int kb = 0; kb < KB; ++kb, ++g) {
                const int s = g % S;
                ptx::mbar_wait(&full[s], (g / S) & 1);
                if (work) {
                    const uint8_t* st = smem + (size_t)s * HWD_STAGE_BYTES;
#pragma unroll
                    for (int rr = 0; rr < 4; ++rr) {
                        const int row = (rr & 1) * 64 + t;                                   // row of the 128-row block rr / 2
                        const uint8_t* rp = st + (rr >> 1) * EF_TILE_BYTES + row * 128;
#pragma unroll
                        for (int c = 0; c < 8; ++c) {                                        // the 8 chunks of the row, rotated by the lane: no bank conflicts
                            const uint4 v = *reinterpret_cast<const uint4*>(rp + (((c + t) & 7) << 4));
                            const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w4[e]));
                                ss[rr] = fmaf(f.x, f.x, ss[rr]); ss[rr] = fmaf(f.y, f.y, ss[rr]);
                            }
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(&empty[s]);
            }
            if (work) {
#pragma unroll
                for (int rr = 0; rr < 4; ++rr) {
                    const int64_t m = m0 + (rr >> 1) * 128 + (rr & 1) * 64 + t;
                    if (m < a.M && !((rr >> 1) == 1 && m0 + 128 >= a.M)) a.norm_out[m] = 1.0f / sqrtf(ss[rr]);
                }
                __threadfence();
                __syncwarp();
                if (lane == 0) atomicAdd(a.norm_flag + mt, 1);
            }
        }
    } else {
        const int q = warp & 3, blk = (warp - 2) >> 2;
        uint32_t it = 0;
        for (int64_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
            const int nt = (int)(tile % n_ntiles), n0 = nt * HF_BN;
            const int64_t m = (tile / n_ntiles) * HWD_BM + blk * 128 + q * 32 + lane;
            const bool row_ok = m < a.M;
            const uint32_t as = it & 1u;
            const uint32_t acc = tmem_base + as * 256 + blk * 128 + ((uint32_t)(q * 32) << 16);
            float scale = a.scale * __ldg(a.bscale_inv);
            int yv = -1;
            if (row_ok) {
                if (a.rowscale && !a.norm_out) scale *= __ldg(a.rowscale + m);
                if (a.y) yv = a.y[a.pos0 + m];
            }
            ptx::mbar_wait(&tmem_full[as], (it >> 1) & 1);
            ptx::tc_fence_after_sync();
            if (a.norm_out) {              // the row tile's norms come from the CTA that walks its column tile 0 (this round or earlier)
                const int* fl = a.norm_flag + tile / n_ntiles;
                bool ready = false;
                unsigned long long t0 = 0;
                for (unsigned spin = 0;; ++spin) {
                    int v;
                    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(fl) : "memory");
                    if (v >= 2) { ready = true; break; }
                    if ((spin & 63u) == 63u) {                            // bounded by wall clock: never a hang, whatever the CTA scheduling
                        unsigned long long now;
                        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                        if (t0 == 0) t0 = now;
                        else if (now - t0 > 200000ull) break;
                    }
                }
                if (row_ok) {
                    if (ready) scale *= __ldcg(a.norm_out + m);
                    else {
                        const uint4* xp = reinterpret_cast<const uint4*>(a.X + m * a.ldx);
                        float ssq = 0.f;
                        for (int c = 0; c < a.K / 8; ++c) {
                            const uint4 v4 = __ldg(xp + c);
                            const uint32_t w4[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w4[e]));
                                ssq = fmaf(f.x, f.x, ssq); ssq = fmaf(f.y, f.y, ssq);
                            }
                        }
                        scale *= 1.0f / sqrtf(ssq);
                    }
                }
            }
            float mx = -INFINITY, se = 0.f, ly = -INFINITY; int am = 0;
#pragma unroll 1
            for (int ch = 0; ch < HF_BN / 32; ++ch) {
                const int nb = n0 + ch * 32;
                if (nb >= a.N) break;
                uint32_t r0[32];
                ptx::tmem_ld_32x32b_x32(acc + ch * 32, r0);
                ptx::tmem_ld_wait();
                float cmx = -INFINITY; int cam = 0;
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float l = (nb + j < a.N) ? fmaf(scale, __uint_as_float(r0[j]), a.col_bias ? __ldg(a.col_bias + nb + j) : 0.f) : -INFINITY;
                    r0[j] = __float_as_uint(l);
                    if (l > cmx) { cmx = l; cam = nb + j; }
                    if (nb + j == yv) ly = l;
                }
                if (cmx > mx) { se *= expf(mx - cmx); mx = cmx; am = cam; }
                if (mx > -INFINITY) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) se += expf(__uint_as_float(r0[j]) - mx);
                }
            }
            ptx::tc_fence_before_sync();
            ptx::mbar_arrive(&tmem_empty[as]);                        // 256 epilogue threads
            if (row_ok) {
                SoftmaxPart p; p.mx = mx; p.se = se; p.ly = ly; p.am = am;
                a.part[(size_t)nt * a.M + m] = p;                  // [column tile][row]: coalesced here and in k_head_finish
            }
        }
    }
    __syncthreads();
    if (warp == 0) ptx::tmem_dealloc<512>(tmem_base);
}

// largest magnitude of a matrix as the bit pattern of a non-negative float (atomicMax on the integer view); *out zeroed by the caller
__global__ void __launch_bounds__(256) k_absmax_bits(const float* __restrict__ in, int64_t n, unsigned* __restrict__ out) {
    float m = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float v = fabsf(in[i]);
        if (v > m && v < INFINITY) m = v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(out, __float_as_uint(m));
}
// sc[0] = 2^k with max |T| 2^k in [2^14, 2^15), sc[1] = 2^-k
__global__ void k_head_bscale(const unsigned* absmax_bits, float* sc, int headroom = 0) {      // headroom: extra bits kept free
    const float m = __uint_as_float(absmax_bits[0]);
    int e = 0;
    if (m > 0.f) e = ilogbf(m);
    int k = 14 - e - headroom;
    k = k > 100 ? 100 : (k < -100 ? -100 : k);
    sc[0] = ldexpf(1.0f, k); sc[1] = ldexpf(1.0f, -k);
}
// out_hi + out_lo [c][r] = unscaled fp16 pair of in[r][c] * sc[0]   (the wide kernel's B operands)
__global__ void __launch_bounds__(256) k_transpose_split_f16_scaled(const float* __restrict__ in, int64_t ld_in, __half* __restrict__ hi,
                                                                    __half* __restrict__ lo, int rows, int cols, int64_t ld_out,
                                                                    const float* __restrict__ sc) {
    __shared__ float t[32][33];
    const float s2k = sc[0];
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
            const float v = t[tx][i] * s2k;                           // exact: a power of two
            const __half h = __float2half_rn(v);
            hi[(size_t)c * ld_out + r] = h;
            lo[(size_t)c * ld_out + r] = __float2half_rn(v - __half2float(h));
        }
    }
}

// out_hi / out_lo [c][r] = fp16 pair of in[r][c]  (prompts [D, C] -> K-major [C, D]; 32 x 32 shared-memory tiles)
__global__ void __launch_bounds__(256) k_transpose_split_f16(const float* __restrict__ in, int64_t ld_in, __half* __restrict__ hi,
                                                             __half* __restrict__ lo, int rows, int cols, int64_t ld_out) {
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
            const __half h = __float2half_rn(v);
            hi[(size_t)c * ld_out + r] = h;
            lo[(size_t)c * ld_out + r] = __float2half_rn((v - __half2float(h)) * EF_LO_SCALE);
        }
    }
}

// inv_norm[r] = 1 / ||x_r||_2 over fp16 rows (one warp per row, 16-byte loads, fp32 accumulation)
__global__ void __launch_bounds__(256) k_row_inv_norm_f16(const __half* __restrict__ X, int64_t ldx, int64_t n, int D,
                                                          float* __restrict__ inv_norm) {
    const int lane = threadIdx.x & 31;
    const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp_global; r < n; r += nwarps) {
        const uint4* p = reinterpret_cast<const uint4*>(X + r * ldx);
        float s = 0.f;
        for (int c = lane; c < D / 8; c += 32) {
            const uint4 v = __ldg(p + c);
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[e]));
                s = fmaf(f.x, f.x, s); s = fmaf(f.y, f.y, s);
            }
        }
        s = warp_sum(s);
        if (lane == 0) inv_norm[r] = 1.0f / sqrtf(s);
    }
}

// wide: (Bhi, Blo) is the unscaled pair of the 2^k-scaled prompts (k_transpose_split_f16_scaled) and the 256-row
// single-accumulator kernel runs
static int launch_f16_head(const __half* A, int64_t lda, const __half* Bhi, const __half* Blo, int64_t ldb, const HeadF16Args& a,
                           bool wide, cudaStream_t st) {
    CUtensorMap mA, mBhi, mBlo;
    if (int rc = make_tmap_2d_f16(&mA, A, a.M, a.K, lda)) return rc;
    if (int rc = make_tmap_2d_f16(&mBhi, Bhi, a.N, a.K, ldb)) return rc;
    if (int rc = make_tmap_2d_f16(&mBlo, Blo, a.N, a.K, ldb)) return rc;
    if (wide) {
        DBMM_CUDA(set_smem(k_f16_head_wide, HWD_SMEM));
        const int64_t tiles = ((a.M + HWD_BM - 1) / HWD_BM) * a.n_ntiles;
        const int grid = tiles < 148 ? (int)tiles : 148;
        k_f16_head_wide<<<grid, HWD_THREADS, HWD_SMEM, st>>>(mA, mBhi, mBlo, a);
        DBMM_LAUNCH_CHECK();
        return DBMM_OK;
    }
    DBMM_CUDA(set_smem(k_f16_head, HF_SMEM));
    const int64_t tiles = ((a.M + HF_BM - 1) / HF_BM) * a.n_ntiles;
    const int grid = tiles < 148 ? (int)tiles : 148;
    k_f16_head<<<grid, HF_THREADS, HF_SMEM, st>>>(mA, mBhi, mBlo, a);
    DBMM_LAUNCH_CHECK();
    return DBMM_OK;
}

}  // namespace dbmm
