// fp32 SIMT tile GEMM with functor operand loaders, used for the small / irregular contractions of the
// training step (Gram matrix, S = L^T R, dW2) where the shapes (129 x 133, K = 133 ...) do not map onto
// tensor-core tiles.
//   C[m][n] = sum_k A(m, k) * B(k, n)        m in [m0, m0+BM), n in [n0, n0+BN), k in [k0, k1)
// A(m,k) / B(k,n) are callables returning 0 outside the logical matrix; they must be branch-free enough for
// the compiler to batch their loads (select the pointer, then one load).  The k loop is software pipelined:
// the global loads of k-block i+1 are in flight while block i is multiplied out of shared memory.
#pragma once
#include "common.cuh"

namespace dbmm {

constexpr int GT_BM = 64, GT_BN = 64, GT_BK = 16, GT_TM = 4, GT_TN = 4;
constexpr int GT_THREADS = (GT_BM / GT_TM) * (GT_BN / GT_TN);   // 256
constexpr int GT_LDA = GT_BM + 4, GT_LDB = GT_BN + 4;
constexpr int GT_SMEM_FLOATS = GT_BK * GT_LDA + GT_BK * GT_LDB;
constexpr int GT_LD_A = GT_BM * GT_BK / GT_THREADS;             // elements of A each thread stages per k-block (4)
constexpr int GT_LD_B = GT_BN * GT_BK / GT_THREADS;

// A_KFAST: consecutive threads read consecutive k of A (A stored [m][k]); otherwise consecutive m ([k][m]).
// B_NFAST: consecutive threads read consecutive n of B (B stored [k][n]); otherwise consecutive k ([n][k]).
template <bool A_KFAST, bool B_NFAST, class FA, class FB>
__device__ __forceinline__ void simt_gemm_tile(float (&acc)[GT_TM][GT_TN], int m0, int n0, int k0, int k1,
                                               FA fa, FB fb, float* smem) {
    float* sA = smem;                       // [BK][LDA]
    float* sB = smem + GT_BK * GT_LDA;      // [BK][LDB]
    const int tid = threadIdx.x;
    const int tx = tid % (GT_BN / GT_TN);
    const int ty = tid / (GT_BN / GT_TN);
#pragma unroll
    for (int i = 0; i < GT_TM; ++i)
#pragma unroll
        for (int j = 0; j < GT_TN; ++j) acc[i][j] = 0.f;

    float ra[GT_LD_A], rb[GT_LD_B];
    auto fetch = [&](int kb) {
#pragma unroll
        for (int i = 0; i < GT_LD_A; ++i) {
            const int e = tid + i * GT_THREADS;
            int m, k;
            if (A_KFAST) { k = e % GT_BK; m = e / GT_BK; } else { m = e % GT_BM; k = e / GT_BM; }
            ra[i] = (kb + k < k1) ? fa(m0 + m, kb + k) : 0.f;
        }
#pragma unroll
        for (int i = 0; i < GT_LD_B; ++i) {
            const int e = tid + i * GT_THREADS;
            int n, k;
            if (B_NFAST) { n = e % GT_BN; k = e / GT_BN; } else { k = e % GT_BK; n = e / GT_BK; }
            rb[i] = (kb + k < k1) ? fb(kb + k, n0 + n) : 0.f;
        }
    };
    auto stash = [&]() {
#pragma unroll
        for (int i = 0; i < GT_LD_A; ++i) {
            const int e = tid + i * GT_THREADS;
            int m, k;
            if (A_KFAST) { k = e % GT_BK; m = e / GT_BK; } else { m = e % GT_BM; k = e / GT_BM; }
            sA[k * GT_LDA + m] = ra[i];
        }
#pragma unroll
        for (int i = 0; i < GT_LD_B; ++i) {
            const int e = tid + i * GT_THREADS;
            int n, k;
            if (B_NFAST) { n = e % GT_BN; k = e / GT_BN; } else { k = e % GT_BK; n = e / GT_BK; }
            sB[k * GT_LDB + n] = rb[i];
        }
    };

    if (k0 < k1) fetch(k0);
    for (int kb = k0; kb < k1; kb += GT_BK) {
        stash();
        __syncthreads();
        if (kb + GT_BK < k1) fetch(kb + GT_BK);          // next block's loads fly during the FMAs below
#pragma unroll
        for (int k = 0; k < GT_BK; ++k) {
            const float4 a4 = *reinterpret_cast<const float4*>(&sA[k * GT_LDA + ty * GT_TM]);
            const float4 b4 = *reinterpret_cast<const float4*>(&sB[k * GT_LDB + tx * GT_TN]);
            const float a[4] = {a4.x, a4.y, a4.z, a4.w};
            const float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int i = 0; i < GT_TM; ++i)
#pragma unroll
                for (int j = 0; j < GT_TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
}

// element (i, j) of the thread's micro-tile lives at row m0 + ty*TM + i, column n0 + tx*TN + j
__device__ __forceinline__ int gt_row(int m0, int i) { return m0 + (threadIdx.x / (GT_BN / GT_TN)) * GT_TM + i; }
__device__ __forceinline__ int gt_col(int n0, int j) { return n0 + (threadIdx.x % (GT_BN / GT_TN)) * GT_TN + j; }

// [k0, k1) of split `part` out of `parts` over K, in whole k-blocks
__device__ __forceinline__ void gt_split_k(int K, int parts, int part, int& k0, int& k1) {
    const int chunk = ((K + parts - 1) / parts + GT_BK - 1) / GT_BK * GT_BK;
    k0 = part * chunk;
    k1 = min(K, k0 + chunk);
}

}  // namespace dbmm
