// Small "TN" GEMM of the H-space formulation, deterministic and atomics-free:
//
//     C[m][n] = sum_k A[k][m] * B[k][n]            m < M, n < N (both ~130), k < K (batch rows or embedding width, ~1024)
//
// Two per-step products have this shape:
//   * S    = [c*h | c | ds]^T [h | 1]              K = batch rows; operands written row by row by k_rows_train
//   * Gram = [W2 | b2]^T [W2 | b2 | That]          K = D; operands are the adapter's own tensors, column-concatenated
// Round 1 accumulated both with fp32 atomics from the kernels that produce the operands (17 K atomics per CTA on the
// critical path of the row kernel, and a run-to-run varying last bit).  Here every output tile is owned by ONE CTA that
// walks all of K in a fixed order: no atomics, no zeroing, bit-reproducible, and off the critical path (second branch of
// the epoch graph).  The contraction runs on the tensor cores as warp-level mma.sync m16n8k8 with 3xTF32 split operands
// (hi*hi + lo*hi + hi*lo, fp32 accumulate: ~2^-21 relative, DESIGN.md "Precision policy"): 15 CTAs of 8 warps, operands
// streamed through a 4-stage cp.async ring -- a 131 x 129 output does not warrant a tcgen05 / TMEM pipeline.
#pragma once
#include "common.cuh"
#include "ptx_sm100.cuh"

namespace dbmm {

__device__ __forceinline__ void tf32_split(float x, uint32_t& hi, uint32_t& lo) {
    hi = __float_as_uint(x) & 0xffffe000u;                 // tf32-exact head; the tensor core ignores the 13 low bits of lo
    lo = __float_as_uint(x - __uint_as_float(hi));         // (|lo| < 2^-10 |x|, truncated at 2^-21 |x|)
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void mma_3xtf32(float (&c)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4],
                                           const uint32_t (&bh)[2], const uint32_t (&bl)[2]) {
    mma_tf32(c, al, bh);
    mma_tf32(c, ah, bl);
    mma_tf32(c, ah, bh);
}

constexpr int TNG_MAX_TILES = 16, TNG_MAX_KSPLIT = 8;
constexpr int TNG_TM = 32, TNG_TN = 48, TNG_KT = 32, TNG_STAGES = 4, TNG_THREADS = 256;
constexpr int TNG_LDA = TNG_TM + 8, TNG_LDB = TNG_TN + 8;          // == 8 / 24 (mod 32): conflict-free mma fragment reads
constexpr int TNG_STAGE_FLOATS = TNG_KT * (TNG_LDA + TNG_LDB);
constexpr size_t TNG_SMEM = sizeof(float) * (size_t)TNG_STAGES * TNG_STAGE_FLOATS;

// [K rows] x [w0 + w1 + w2 columns]: column concatenation of up to three row-major sources (p[i] == nullptr: w[i] = 0).
struct CatMat { const float* p[3]; int ld[3]; int w[3]; };
static inline CatMat cat_mat(const float* p0, int ld0, int w0, const float* p1 = nullptr, int ld1 = 0, int w1 = 0,
                             const float* p2 = nullptr, int ld2 = 0, int w2 = 0) {
    CatMat m; m.p[0] = p0; m.ld[0] = ld0; m.w[0] = w0; m.p[1] = p1; m.ld[1] = ld1; m.w[1] = w1; m.p[2] = p2; m.ld[2] = ld2; m.w[2] = w2;
    return m;
}
struct TnGemmArgs {
    CatMat A, B; int M, N, K;
    float* C; int ldc; int n_store;       // columns [N, n_store) of every row are written as zeros (padding the consumer reads)
    int ksplit; float* part; int* ticket; // ksplit > 1: K is cut into `ksplit` slices (grid z); slice tiles go to part[ksplit][tiles][TM*TN] and
                                          // the LAST CTA of a tile (ticket counter, self-resetting) sums them in slice order: deterministic
    int no_early_trigger;                 // 1: no griddepcontrol.launch_dependents -- a consumer joined from ANOTHER stream with the
                                          // programmatic-launch attribute (k_rows_train after the W2 branch) gets a programmatic
                                          // edge from this kernel too and reads C before its own dependency wait
};

__device__ __forceinline__ void cp_async4(uint32_t dst_smem, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst_smem), "l"(src) : "memory");
}

// one 4-column chunk (columns c .. c+3, c % 4 == 0) of row k of a concatenated matrix -> shared memory
__device__ __forceinline__ void tng_load_chunk(float* dst, const CatMat& X, int k, int c, bool row_ok) {
    const int e0 = X.w[0], e1 = e0 + X.w[1], e2 = e1 + X.w[2];
    if (!row_ok || c >= e2) { *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f); return; }
    const int s = c < e0 ? 0 : (c < e1 ? 1 : 2);
    const int off = s == 0 ? 0 : (s == 1 ? e0 : e1), end = s == 0 ? e0 : (s == 1 ? e1 : e2);
    const float* src = X.p[s] + (size_t)k * X.ld[s] + (c - off);
    if (c + 4 <= end && ((X.ld[s] | (c - off)) & 3) == 0 && ((uintptr_t)X.p[s] & 15) == 0) {
        ptx::cp_async16(ptx::smem_u32(dst), src);
        return;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int cc = c + j;
        if (cc >= e2) { dst[j] = 0.f; continue; }
        const int sj = cc < e0 ? 0 : (cc < e1 ? 1 : 2);
        const int oj = sj == 0 ? 0 : (sj == 1 ? e0 : e1);
        cp_async4(ptx::smem_u32(dst + j), X.p[sj] + (size_t)k * X.ld[sj] + (cc - oj));
    }
}

__device__ __forceinline__ void tn_gemm_body(const TnGemmArgs& a, const int kz = 0) {
    extern __shared__ __align__(16) float tng_smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    const int m0 = blockIdx.x * TNG_TM, n0 = blockIdx.y * TNG_TN;
    const int KB_all = (a.K + TNG_KT - 1) / TNG_KT;
    const int ksplit = a.ksplit > 1 ? a.ksplit : 1;
    const int kb_per = (KB_all + ksplit - 1) / ksplit;
    const int kb_lo = kz * kb_per;
    const int KB = max(0, min(KB_all, kb_lo + kb_per) - kb_lo);
    ptx::pdl_wait();                // the operands come from the preceding kernel of this branch (row kernel / W2 update)
    if (a.ksplit > 1) DBMM_TL_WAIT(TL_TN);
    if (!a.no_early_trigger) ptx::pdl_launch();

    auto issue = [&](int kb) {
        float* As = tng_smem + (size_t)(kb % TNG_STAGES) * TNG_STAGE_FLOATS;
        float* Bs = As + TNG_KT * TNG_LDA;
        const int k0 = (kb_lo + kb) * TNG_KT;
        {
            const int r = tid >> 3, c4 = tid & 7;                                  // 32 rows x 8 chunks
            tng_load_chunk(As + r * TNG_LDA + c4 * 4, a.A, k0 + r, m0 + c4 * 4, k0 + r < a.K);
        }
        for (int id = tid; id < TNG_KT * (TNG_TN / 4); id += TNG_THREADS) {        // 32 rows x 12 chunks
            const int r = id / (TNG_TN / 4), c4 = id - r * (TNG_TN / 4);
            tng_load_chunk(Bs + r * TNG_LDB + c4 * 4, a.B, k0 + r, n0 + c4 * 4, k0 + r < a.K);
        }
        ptx::cp_async_commit();
    };
    for (int s = 0; s < TNG_STAGES - 1; ++s) { if (s < KB) issue(s); else ptx::cp_async_commit(); }

    // warp roles: k-half grp (k-steps 2*grp, 2*grp + 1 of every stage), m16 tile wm, three n8 tiles from wn * 3
    const int grp = warp >> 2, wm = warp & 1, wn = (warp >> 1) & 1;
    float acc[3][4];
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[j][q] = 0.f;
    for (int kb = 0; kb < KB; ++kb) {
        ptx::cp_async_wait<TNG_STAGES - 2>();
        __syncthreads();                                     // stage kb has landed; stage kb - 1 is free for everybody
        if (kb + TNG_STAGES - 1 < KB) issue(kb + TNG_STAGES - 1); else ptx::cp_async_commit();
        const float* As = tng_smem + (size_t)(kb % TNG_STAGES) * TNG_STAGE_FLOATS;
        const float* Bs = As + TNG_KT * TNG_LDA;
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
            const int k0 = (grp * 2 + kk) * 8;
            uint32_t ah[4], al[4];
            tf32_split(As[(k0 + t) * TNG_LDA + wm * 16 + g], ah[0], al[0]);
            tf32_split(As[(k0 + t) * TNG_LDA + wm * 16 + g + 8], ah[1], al[1]);
            tf32_split(As[(k0 + t + 4) * TNG_LDA + wm * 16 + g], ah[2], al[2]);
            tf32_split(As[(k0 + t + 4) * TNG_LDA + wm * 16 + g + 8], ah[3], al[3]);
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const int nb = (wn * 3 + j) * 8;
                uint32_t bh[2], bl[2];
                tf32_split(Bs[(k0 + t) * TNG_LDB + nb + g], bh[0], bl[0]);
                tf32_split(Bs[(k0 + t + 4) * TNG_LDB + nb + g], bh[1], bl[1]);
                mma_3xtf32(acc[j], ah, al, bh, bl);
            }
        }
    }
    ptx::cp_async_wait<0>();
    __syncthreads();                                         // the ring is free: the upper k-half parks its sums in it
    float* red = tng_smem;                                   // [4 warps][3][4][32]
    if (grp == 1) {
#pragma unroll
        for (int j = 0; j < 3; ++j)
#pragma unroll
            for (int q = 0; q < 4; ++q) red[(((warp & 3) * 3 + j) * 4 + q) * 32 + lane] = acc[j][q];
    }
    __syncthreads();
    if (ksplit == 1) {
        if (grp == 0) {
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const int n = n0 + (wn * 3 + j) * 8 + 2 * t;
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    const int m = m0 + wm * 16 + g + 8 * hh;
                    if (m >= a.M) continue;
                    const float v0 = acc[j][2 * hh] + red[((warp * 3 + j) * 4 + 2 * hh) * 32 + lane];
                    const float v1 = acc[j][2 * hh + 1] + red[((warp * 3 + j) * 4 + 2 * hh + 1) * 32 + lane];
                    if (n < a.n_store) a.C[(size_t)m * a.ldc + n] = n < a.N ? v0 : 0.f;
                    if (n + 1 < a.n_store) a.C[(size_t)m * a.ldc + n + 1] = n + 1 < a.N ? v1 : 0.f;
                }
            }
        }
        return;
    }
    // ---- K slices: park the slice tile, take a ticket; the last CTA of the tile adds the slices in slice order
    const int tile = blockIdx.y * gridDim.x + blockIdx.x, ntiles = gridDim.x * gridDim.y;
    constexpr int TE = TNG_TM * TNG_TN;
    float* mine = a.part + ((size_t)kz * ntiles + tile) * TE;
    if (grp == 0) {
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const int nl = (wn * 3 + j) * 8 + 2 * t;
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const int ml = wm * 16 + g + 8 * hh;
                const float v0 = acc[j][2 * hh] + red[((warp * 3 + j) * 4 + 2 * hh) * 32 + lane];
                const float v1 = acc[j][2 * hh + 1] + red[((warp * 3 + j) * 4 + 2 * hh + 1) * 32 + lane];
                __stcg(reinterpret_cast<float2*>(mine + ml * TNG_TN + nl), make_float2(v0, v1));
            }
        }
    }
    __threadfence();
    __syncthreads();
    __shared__ int s_last;
    if (tid == 0) {
        const int tk = atomicAdd(a.ticket + tile, 1);
        s_last = tk == ksplit - 1;
        if (s_last) a.ticket[tile] = 0;              // ready for the next launch on this stream
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    for (int e = tid; e < TE; e += TNG_THREADS) {
        const int ml = e / TNG_TN, nl = e - ml * TNG_TN, m = m0 + ml, n = n0 + nl;
        if (m >= a.M || n >= a.n_store) continue;
        float pz[TNG_MAX_KSPLIT];                    // all slices in flight, then added in slice order
#pragma unroll
        for (int z = 0; z < TNG_MAX_KSPLIT; ++z) pz[z] = z < ksplit ? __ldcg(a.part + ((size_t)z * ntiles + tile) * TE + e) : 0.f;
        float v = 0.f;
#pragma unroll
        for (int z = 0; z < TNG_MAX_KSPLIT; ++z) v += pz[z];
        a.C[(size_t)m * a.ldc + n] = n < a.N ? v : 0.f;
    }
}

__global__ void __launch_bounds__(TNG_THREADS) k_tn_gemm(TnGemmArgs a) { DBMM_TL_SCOPE(TL_TN); tn_gemm_body(a, (int)blockIdx.z); }

static inline size_t tn_gemm_part_floats() { return (size_t)TNG_MAX_KSPLIT * TNG_MAX_TILES * TNG_TM * TNG_TN; }

static int launch_tn_gemm(const TnGemmArgs& a, cudaStream_t st, bool pdl = true) {
    DBMM_CHECK_ARG(a.M >= 1 && a.N >= 1 && a.K >= 1 && a.C && a.n_store >= a.N && a.ldc >= a.n_store, "bad TN GEMM %d x %d x %d", a.M, a.N, a.K);
    DBMM_CUDA(set_smem(k_tn_gemm, TNG_SMEM));
    const int ks = a.ksplit > 1 ? a.ksplit : 1;
    dim3 grid(ceil_div(a.M, TNG_TM), ceil_div(a.n_store, TNG_TN), ks);
    DBMM_CHECK_ARG(ks == 1 || (a.part && a.ticket && ks <= TNG_MAX_KSPLIT && (int)(grid.x * grid.y) <= TNG_MAX_TILES), "bad TN GEMM K split");
    if (pdl) DBMM_CUDA(launch_pdl(k_tn_gemm, grid, dim3(TNG_THREADS), TNG_SMEM, st, a));
    else { k_tn_gemm<<<grid, TNG_THREADS, TNG_SMEM, st>>>(a); DBMM_LAUNCH_CHECK(); }
    return DBMM_OK;
}

}  // namespace dbmm
