// W2 branch of the fused training step on the 5th-generation tensor cores (H-space formulation):
//
//     dW2a = [W2 | b2 | That] S  ->  SGD on W2 / b2 (torch.optim.SGD, demo/util.py:118-136)  ->  the rows' share of the next
//     Gram matrix  G = [W2 | b2]^T [W2 | b2 | That]
//
// One CTA owns HW_ROWS = 64 embedding rows d.  Same skeleton as hs_rows.cuh:
//   P0   B operand = S^T (the summed tiles of the row kernel, [j][l] as stored: K-major), A operand = the CTA's W2 rows, both
//        split hi + lo; the three-to-five K entries past H (b2 and the prompt columns of [W2 | b2 | That]) are added on the
//        CUDA cores in the epilogue.
//   MMA1 dW2a tile (3xTF32, 48 tcgen05.mma of M128 N144 K8; the upper 64 accumulator lanes are never read).
//   E1   one thread per (row d, column half): SGD, new W2 / b2 / momentum / gradient out, and the new rows written TRANSPOSED
//        into ONE K-major tile [144 n][64 d] that is both operands of the second contraction (A = its first 128 rows).
//   MMA2 Gram share  G_part[m][n] = sum_d W2'[d][m] [W2' | b2' | That][d][n]   (24 MMAs), stored to Gpart[tile].
// k_sum_gpart adds the tiles in tile order (deterministic) into the Gram matrix the row kernel of the next step reads; the
// row m = H of G (b2-weighted sums) is column H of the tiles by symmetry plus three scalars per tile.
#pragma once
#include "hs_rows.cuh"
#include "p2p.cuh"

namespace dbmm {

constexpr int HW_ROWS = 64;
constexpr int HW_TILE_KT_BYTES = HR_N * 128;                 // MMA2: one k-tile of the shared [144 n][32 d] operand tile
constexpr size_t HW_SMEM = HR_SMEM;

struct HsW2Args {
    float* W2; float* b2; float* g; float* v;     // trainable adapter's W2 [D][H], b2 [D]; flat gradient / momentum buffers
    size_t oW2, ob2;                              // offsets of W2 / b2 inside the flat buffers
    const float* lr_dev; float lr, momentum, wd;
    const float* That;                            // [D][C]
    const float* ST;                              // [H+1][HR_SP_LD]: S^T summed over the row tiles (k_sum_spart)
    float* Gpart;                                 // [tiles][H+1][HR_SP_LD]
    int D, H, C;
};

__device__ __forceinline__ void hs_w2_body(const HsW2Args& a) {
    extern __shared__ uint8_t hw_smem_raw[];
    uint8_t* smem = hw_smem_raw + ((1024u - (ptx::smem_u32(hw_smem_raw) & 1023u)) & 1023u);
    uint8_t* sB = smem;                                      // MMA1 B operand (S^T); MMA2: the shared operand tile
    uint8_t* sA = smem + HR_SB_BYTES;                        // MMA1 A operand (W2 rows)
    float* sTail = (float*)(smem + HR_SB_BYTES + HR_SA_BYTES);     // [HR_N][8]: per column j: S rows H (b2), H+1+c (That_c), padding
    float* sRs = sTail + 8 * HR_N;                           // [64][8]: per row d: b2, That[d][0..3]
    float* sRow = sRs + 64 * 8;                // [2][16]: per row warp: sum_d b2' * {b2', That_c}
    uint64_t* bars = (uint64_t*)(sRow + 2 * 16);             // [0]: MMA1 done; [1]: MMA2 operands ready; [2]: MMA2 done
    uint32_t* tmem_ptr = (uint32_t*)(bars + 3);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int H = HR_H, C = a.C, D = a.D;
    const int d0 = blockIdx.x * HW_ROWS;
    const int rows_here = min(HW_ROWS, D - d0);
    const float lr = a.lr_dev ? __ldg(a.lr_dev) : a.lr;

    if (tid == 0) {
        ptx::mbar_init(&bars[0], 1); ptx::mbar_init(&bars[1], HR_THREADS); ptx::mbar_init(&bars[2], 1);
        ptx::fence_mbar_init();
    }
    if (warp == 2) ptx::tmem_alloc<512>(tmem_ptr);
    HR_TICK(1, 0);

    // TMEM readers: warps with (warp & 2) == 0 (a warp reaches the lanes 32 * (warp % 4) .. +31 only): rows rw * 32 + lane,
    // accumulator columns of the quarter qt.  They only move the tile to shared memory; the update itself runs on all 16
    // warps with lane = column (per-column constants in registers, coalesced global traffic): warp w owns rows 4w .. 4w+3.
    const bool ep = (warp & 2) == 0;
    const int qt = warp >> 2, rw = warp & 1, j0 = qt * 32;
    const uint32_t tlane = ((uint32_t)(rw * 32)) << 16;
    // parameters and momentum of this thread's 4 rows x 4 columns (column = lane + 32 c) do not depend on the predecessor
    // kernel: requested before the dependency wait
    float pv[4][4], vv[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int d = d0 + warp * 4 + i;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            pv[i][c] = 0.f; vv[i][c] = 0.f;
            if (d < D) { pv[i][c] = __ldcg(a.W2 + (size_t)d * H + lane + 32 * c); vv[i][c] = __ldcg(a.v + a.oW2 + (size_t)d * H + lane + 32 * c); }
        }
    }
    // row scalars (thread tid < 64 <-> row tid): b2, its momentum, the prompt entries
    const int my_d = d0 + tid;
    const bool rvalid = tid < HW_ROWS && my_d < D;
    float b2v = 0.f, vb2 = 0.f, tv[HR_CT] = {0.f, 0.f, 0.f, 0.f};
    if (rvalid) {
        b2v = __ldcg(a.b2 + my_d); vb2 = __ldcg(a.v + a.ob2 + my_d);
#pragma unroll
        for (int c = 0; c < HR_CT; ++c) if (c < C) tv[c] = __ldg(a.That + (size_t)my_d * C + c);
    }
    if (tid < HW_ROWS) {
        sRs[tid * 8 + 0] = b2v;
#pragma unroll
        for (int c = 0; c < HR_CT; ++c) sRs[tid * 8 + 1 + c] = tv[c];
    }
    // A operand: the CTA's W2 rows, K = hidden unit.  64 x 32 tasks of 4 floats, 4 per thread.
    {
        float4 av[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int task = tid + u * HR_THREADS, row = task >> 5, k = (task & 31) * 4;
            av[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row < rows_here) av[u] = __ldcg(reinterpret_cast<const float4*>(a.W2 + (size_t)(d0 + row) * H + k));
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int task = tid + u * HR_THREADS, row = task >> 5, c16 = task & 31;
            const float ax[4] = {av[u].x, av[u].y, av[u].z, av[u].w};
            float hi[4], lo[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) hr_split(ax[q], hi[q], lo[q]);
            const uint32_t off = (uint32_t)(c16 >> 3) * HR_A_KT_BYTES + ptx::sw128_offset(row, c16 & 7);
            *reinterpret_cast<float4*>(sA + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<float4*>(sA + 4 * HR_A_KT_BYTES + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
        }
    }
    HR_TICK(1, 1);
    ptx::pdl_wait();                // S^T comes from k_sum_spart, the kernel in front of this one
    DBMM_TL_WAIT(TL_W2);
    ptx::pdl_launch();
    HR_TICK(1, 2);
    // B operand: B[n = j][k = l] = S^T[j][l], l < H;  rows j > H: zeros.   144 x 32 tasks, 9 per thread; every load of the
    // prologue is issued before the first one is consumed (one L2 round trip)
    {
        float4 x[9];
        float tl[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int u = 0; u < 9; ++u) {
            const int task = tid + u * HR_THREADS, n = task >> 5, k = (task & 31) * 4;
            x[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (n <= H) x[u] = __ldcg(reinterpret_cast<const float4*>(a.ST + (size_t)n * HR_SP_LD + k));
        }
#pragma unroll
        for (int u = 0; u < 3; ++u) {
            const int e = tid + u * HR_THREADS;
            if (e < 8 * HR_N) {
                const int j = e >> 3, t = e & 7;
                if (j <= H && t <= C) tl[u] = __ldcg(a.ST + (size_t)j * HR_SP_LD + H + t);
            }
        }
#pragma unroll
        for (int u = 0; u < 9; ++u) {
            const int task = tid + u * HR_THREADS, n = task >> 5, c16 = task & 31;
            const float xx[4] = {x[u].x, x[u].y, x[u].z, x[u].w};
            float hi[4], lo[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) hr_split(xx[q], hi[q], lo[q]);
            const uint32_t off = (uint32_t)(c16 >> 3) * HR_B_KT_BYTES + ptx::sw128_offset(n, c16 & 7);
            *reinterpret_cast<float4*>(sB + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<float4*>(sB + 4 * HR_B_KT_BYTES + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
        }
#pragma unroll
        for (int u = 0; u < 3; ++u) {
            const int e = tid + u * HR_THREADS;
            if (e < 8 * HR_N) sTail[e] = tl[u];
        }
    }
    ptx::fence_proxy_async_smem();
    ptx::tc_fence_before_sync();
    __syncthreads();
    ptx::tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_ptr;
    HR_TICK(1, 3);
    constexpr uint32_t TM_D1 = 0, TM_D2 = 160;
    constexpr uint32_t idesc = ptx::umma_idesc(/*tf32*/ 2, 128, HR_N, 0, 0);
    if (tid == 64) {
        const uint32_t bA = ptx::smem_u32(sA), bB = ptx::smem_u32(sB);
#pragma unroll
        for (int kt = 0; kt < 4; ++kt)
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const uint64_t ahi = ptx::umma_smem_desc(bA + kt * HR_A_KT_BYTES + kk * 32, 0, 1024);
                const uint64_t alo = ptx::umma_smem_desc(bA + 4 * HR_A_KT_BYTES + kt * HR_A_KT_BYTES + kk * 32, 0, 1024);
                const uint64_t bhi = ptx::umma_smem_desc(bB + kt * HR_B_KT_BYTES + kk * 32, 0, 1024);
                const uint64_t blo = ptx::umma_smem_desc(bB + 4 * HR_B_KT_BYTES + kt * HR_B_KT_BYTES + kk * 32, 0, 1024);
                ptx::mma_tf32_ss(tmem_base + TM_D1, alo, bhi, idesc, (kt | kk) != 0 ? 1u : 0u);
                ptx::mma_tf32_ss(tmem_base + TM_D1, ahi, blo, idesc, 1u);
                ptx::mma_tf32_ss(tmem_base + TM_D1, ahi, bhi, idesc, 1u);
            }
        ptx::mma_commit(&bars[0]);
    }
    ptx::mbar_wait(&bars[0], 0);
    ptx::tc_fence_after_sync();
    HR_TICK(1, 4);

    // MMA2 operand tile (the MMA1 tiles are dead): T2[n][k = d] = [W2' | b2' | That][d][n], hi | lo, 2 k-tiles of 144 rows
    uint8_t* sT2 = sB;
    constexpr int DLD = 132;                                             // row stride of the dumped accumulator tile (floats)
    float* sD = (float*)(sB + 2 * 2 * HW_TILE_KT_BYTES);                 // [64][DLD]: dW2a tile (MMA part), above the MMA2 tile
    auto put = [&](int k, int n, float x) {
        float hi, lo;
        hr_split(x, hi, lo);
        const uint32_t off = (uint32_t)(k >> 5) * HW_TILE_KT_BYTES + (uint32_t)n * 128u + (((((uint32_t)k & 31u) >> 2) ^ ((uint32_t)n & 7u)) << 4) + ((uint32_t)k & 3u) * 4u;
        *reinterpret_cast<float*>(sT2 + off) = hi;
        *reinterpret_cast<float*>(sT2 + 2 * HW_TILE_KT_BYTES + off) = lo;
    };
    if (ep) {                                                            // TMEM -> shared memory
        const uint32_t d = tmem_base + TM_D1 + tlane;
        uint32_t r[32];
        ptx::tmem_ld_32x32b_x32(d + j0, r);
        ptx::tmem_ld_wait();
        float* drow = sD + (size_t)(rw * 32 + lane) * DLD + j0;
#pragma unroll
        for (int q = 0; q < 32; q += 4)
            *reinterpret_cast<float4*>(drow + q) = make_float4(__uint_as_float(r[q]), __uint_as_float(r[q + 1]), __uint_as_float(r[q + 2]), __uint_as_float(r[q + 3]));
        if (qt == 0) {
            ptx::tmem_ld_32x32b_x32(d + 128, r);
            ptx::tmem_ld_wait();
            sD[(size_t)(rw * 32 + lane) * DLD + 128] = __uint_as_float(r[0]);
        }
    }
    __syncthreads();
    {   // SGD on W2: lane = column.  g = MMA part + b2[d] S[H][j] + sum_c That[d][c] S[H+1+c][j]
        float t0[4], t1[4], t2[4], t3[4], t4[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const float4 x = *reinterpret_cast<const float4*>(sTail + (lane + 32 * c) * 8);
            t0[c] = x.x; t1[c] = x.y; t2[c] = x.z; t3[c] = x.w; t4[c] = sTail[(lane + 32 * c) * 8 + 4];
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int k = warp * 4 + i, d = d0 + k;
            const float4 rs = *reinterpret_cast<const float4*>(sRs + k * 8);       // b2, That[d][0..2]
            const float rs4 = sRs[k * 8 + 4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int jcol = lane + 32 * c;
                float g = sD[(size_t)k * DLD + jcol] + rs.x * t0[c];
                g = fmaf(rs.y, t1[c], g); g = fmaf(rs.z, t2[c], g); g = fmaf(rs.w, t3[c], g); g = fmaf(rs4, t4[c], g);
                const float p0 = pv[i][c];
                const float vn = a.momentum * vv[i][c] + (g + a.wd * p0);
                const float pn = p0 - lr * vn;
                if (d < D) {
                    const size_t fo = (size_t)d * H + jcol;
                    a.W2[fo] = pn; a.v[a.oW2 + fo] = vn; a.g[a.oW2 + fo] = g;
                }
                put(k, jcol, d < D ? pn : 0.f);
            }
        }
    }
    if (tid < HW_ROWS) {                                  // column H: b2;  rows H+1+c of the tile: That;  rows >= H+1+C: zeros
        const int k = tid;
        float g = sD[(size_t)k * DLD + 128] + b2v * sTail[H * 8];
#pragma unroll
        for (int c = 0; c < HR_CT; ++c) g = fmaf(tv[c], sTail[H * 8 + 1 + c], g);
        const float vn = a.momentum * vb2 + (g + a.wd * b2v);
        const float bn = rvalid ? b2v - lr * vn : 0.f;
        if (rvalid) { a.b2[my_d] = bn; a.v[a.ob2 + my_d] = vn; a.g[a.ob2 + my_d] = g; }
        put(k, H, bn);
#pragma unroll 1
        for (int n = H + 1; n < HR_N; ++n) {
            float x = 0.f;
#pragma unroll
            for (int c = 0; c < HR_CT; ++c) if (n - H - 1 == c) x = tv[c];
            put(k, n, x);
        }
        // row H of the Gram matrix past the symmetric part: sum_d b2' * {b2', That_c}
        const float s0 = warp_sum(bn * bn);
        float sc[HR_CT];
#pragma unroll
        for (int c = 0; c < HR_CT; ++c) sc[c] = warp_sum(bn * tv[c]);
        if (lane < 16) {
            float x = 0.f;
            if (lane == 0) x = s0;
#pragma unroll
            for (int c = 0; c < HR_CT; ++c) if (lane == 1 + c) x = sc[c];
            sRow[warp * 16 + lane] = x;
        }
    }
    HR_TICK(1, 5);
    ptx::fence_proxy_async_smem();
    ptx::mbar_arrive(&bars[1]);
    if (tid == 64) {
        ptx::mbar_wait(&bars[1], 0);
        ptx::tc_fence_after_sync();
        const uint32_t bT = ptx::smem_u32(sT2);
#pragma unroll
        for (int kt = 0; kt < 2; ++kt)
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const uint64_t thi = ptx::umma_smem_desc(bT + kt * HW_TILE_KT_BYTES + kk * 32, 0, 1024);
                const uint64_t tlo = ptx::umma_smem_desc(bT + 2 * HW_TILE_KT_BYTES + kt * HW_TILE_KT_BYTES + kk * 32, 0, 1024);
                ptx::mma_tf32_ss(tmem_base + TM_D2, tlo, thi, idesc, (kt | kk) != 0 ? 1u : 0u);      // A = rows 0..127 of the same tile
                ptx::mma_tf32_ss(tmem_base + TM_D2, thi, tlo, idesc, 1u);
                ptx::mma_tf32_ss(tmem_base + TM_D2, thi, thi, idesc, 1u);
            }
        ptx::mma_commit(&bars[2]);
    }
    __syncthreads();
    HR_TICK(1, 6);
    float* tile = a.Gpart + (size_t)blockIdx.x * (H + 1) * HR_SP_LD;
    if (tid < 16) tile[(size_t)H * HR_SP_LD + H + tid] = sRow[tid] + sRow[16 + tid];
    {   // rows m < H from TMEM: warp w owns lanes 32 * (w % 4) .. +31 and the column chunk w / 4 (chunk 4, 16 wide: warps 0..3 again)
        ptx::mbar_wait(&bars[2], 0);
        ptx::tc_fence_after_sync();
        const int m = (warp & 3) * 32 + lane;
        float* out = tile + (size_t)m * HR_SP_LD;
        const uint32_t d2 = tmem_base + TM_D2 + (((uint32_t)((warp & 3) * 32)) << 16);
        float* scr2 = (float*)(sB + 2 * 2 * HW_TILE_KT_BYTES) + warp * (32 * 36);
#pragma unroll 1
        for (int ch = warp >> 2; ch < 5; ch += 4) {
            uint32_t r[32];
            ptx::tmem_ld_32x32b_x32(d2 + ch * 32, r);
            ptx::tmem_ld_wait();
            float v[32];
#pragma unroll
            for (int q = 0; q < 32; ++q) v[q] = __uint_as_float(r[q]);
            if (ch < 4) warp_store_block32(v, scr2, tile + (size_t)((warp & 3) * 32) * HR_SP_LD + ch * 32, HR_SP_LD, 32, lane);
            else {
#pragma unroll
                for (int q = 0; q < HR_N - 128; q += 4) *reinterpret_cast<float4*>(out + 128 + q) = make_float4(v[q], v[q + 1], v[q + 2], v[q + 3]);
            }
        }
        ptx::tc_fence_before_sync();
    }
    HR_TICK(1, 7);
    __syncthreads();
    if (warp == 2) ptx::tmem_dealloc<512>(tmem_base);
}

__global__ void __launch_bounds__(HR_THREADS, 1) k_hs_w2(HsW2Args a) { DBMM_TL_SCOPE(TL_W2); hs_w2_body(a); }

// G[m][n] (row stride H+1+C) = sum over the W2 tiles, in tile order: m < H: Gpart[t][m][n]; m = H: n < H: Gpart[t][n][H] (symmetry),
// n >= H: Gpart[t][H][n].  No early launch trigger: a row kernel joined from another stream may read G before its own wait.
struct SumGpartArgs { const float* Gpart; int tiles, H, C; float* G; };
__device__ __forceinline__ void sum_gpart_body(const SumGpartArgs& a) {
    const int H = a.H, ldg = H + 1 + a.C;
    ptx::pdl_wait();
    DBMM_TL_WAIT(TL_GSUM);
    const int e = blockIdx.x * 256 + threadIdx.x;
    if (e >= (H + 1) * ldg) return;
    const int m = e / ldg, n = e - m * ldg;
    const size_t ts = (size_t)(H + 1) * HR_SP_LD;
    const float* src = a.Gpart + (m < H ? (size_t)m * HR_SP_LD + n : (n < H ? (size_t)n * HR_SP_LD + H : (size_t)H * HR_SP_LD + n));
    float v = 0.f;
    int t = 0;
    for (; t + 4 <= a.tiles; t += 4) {
        const float p0 = __ldcg(src + (size_t)t * ts), p1 = __ldcg(src + (size_t)(t + 1) * ts);
        const float p2 = __ldcg(src + (size_t)(t + 2) * ts), p3 = __ldcg(src + (size_t)(t + 3) * ts);
        v += p0; v += p1; v += p2; v += p3;
    }
    for (; t < a.tiles; ++t) v += __ldcg(src + (size_t)t * ts);
    a.G[e] = v;
}
__global__ void __launch_bounds__(256) k_sum_gpart(SumGpartArgs a) { DBMM_TL_SCOPE(TL_GSUM); sum_gpart_body(a); }

// Data parallel: S^T summed over the ranks in place (peer memory, the S slots and flags of p2p.cuh).  CTA c pushes slice c of
// the rank's S^T to every rank, raises flag [c][rank] everywhere, waits for slice c of every rank and adds them in rank order:
// every rank ends with the same bits.  Off the critical path (W2 branch).
constexpr int PS_CTAS = 16, PS_THREADS = 256;
__global__ void __launch_bounds__(PS_THREADS) k_p2p_sum_st(float* ST, int n4, P2pArgs p) {
    ptx::pdl_wait();                // S^T comes from the kernel in front of this one (TN GEMM)
    ptx::pdl_launch();
    const unsigned inst = p2p_instance(p);
    const int parity = inst & 1u, c = blockIdx.x, tid = threadIdx.x;
    const int per = (n4 + PS_CTAS - 1) / PS_CTAS, lo4 = c * per, hi4 = min(n4, lo4 + per);
    for (int e = lo4 + tid; e < hi4; e += PS_THREADS) {
        const float4 v = __ldcg(reinterpret_cast<const float4*>(ST) + e);
        for (int r = 0; r < p.world; ++r) reinterpret_cast<float4*>(p2p_s_slot(p.peer[r], parity, p.rank))[e] = v;
    }
    __threadfence_system();
    __syncthreads();
    if (tid < p.world) {
        unsigned* f = p2p_s_flag(p.peer[tid], c, p.rank);
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f), "r"(inst + 1u) : "memory");
    }
    if (tid < p.world) {
        const unsigned* f = p2p_s_flag(p.peer[p.rank], c, tid);
        unsigned v = 0;
        unsigned long long t0 = 0;
        for (unsigned spin = 0;; ++spin) {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
            if ((int)(v - (inst + 1u)) >= 0 || (p.skip & 8) || p2p_expired(p, t0, spin)) break;
        }
    }
    __syncthreads();
    char* me = p.peer[p.rank];
    for (int e = lo4 + tid; e < hi4; e += PS_THREADS) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int r = 0; r < p.world; ++r) {
            const float4 v = __ldcg(reinterpret_cast<const float4*>(p2p_s_slot(me, parity, r)) + e);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        reinterpret_cast<float4*>(ST)[e] = acc;
    }
}
static int launch_p2p_sum_st(float* ST, int H, const P2pArgs& p, cudaStream_t st) {
    const int n4 = (H + 1) * HR_SP_LD / 4;
    DBMM_CHECK_ARG((size_t)(H + 1) * HR_SP_LD <= P2P_S_FLOATS && PS_CTAS <= P2P_S_CTAS, "S^T exceeds the peer-memory slot");
    DBMM_CUDA(launch_pdl(k_p2p_sum_st, dim3(PS_CTAS), dim3(PS_THREADS), 0, st, ST, n4, p));
    return DBMM_OK;
}

static inline size_t hs_gpart_floats(int D) { return (size_t)((D + HW_ROWS - 1) / HW_ROWS) * (HR_H + 1) * HR_SP_LD; }

static int launch_hs_w2(const HsW2Args& a, cudaStream_t st, bool pdl) {
    DBMM_CUDA(set_smem(k_hs_w2, HW_SMEM));
    if (pdl) DBMM_CUDA(launch_pdl(k_hs_w2, dim3(ceil_div(a.D, HW_ROWS)), dim3(HR_THREADS), HW_SMEM, st, a));
    else { k_hs_w2<<<ceil_div(a.D, HW_ROWS), HR_THREADS, HW_SMEM, st>>>(a); DBMM_LAUNCH_CHECK(); }
    return DBMM_OK;
}
static int launch_sum_gpart(const float* Gpart, int D, int H, int C, float* G, cudaStream_t st) {
    SumGpartArgs s; s.Gpart = Gpart; s.tiles = ceil_div(D, HW_ROWS); s.H = H; s.C = C; s.G = G;
    DBMM_CUDA(launch_pdl(k_sum_gpart, dim3(ceil_div((H + 1) * (H + 1 + C), 256)), dim3(256), 0, st, s));
    return DBMM_OK;
}

}  // namespace dbmm
