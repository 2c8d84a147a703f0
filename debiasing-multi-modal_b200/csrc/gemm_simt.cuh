// fp32 SIMT tile GEMM with functor operand loaders.  Used for the small / irregular contractions of
// the training step (Gram matrices, S = L^T R, dW2) and, in round 1, for GEMM-1 and dW1 as well.
//   C[m][n] = sum_k A(m, k) * B(k, n)        m in [m0, m0+BM), n in [n0, n0+BN), k in [k0, k1)
// A(m,k) / B(k,n) are callables returning 0 outside the logical matrix.
#pragma once
#include "common.cuh"

namespace dbmm {

constexpr int GT_BM = 64, GT_BN = 64, GT_BK = 16, GT_TM = 4, GT_TN = 4;
constexpr int GT_THREADS = (GT_BM / GT_TM) * (GT_BN / GT_TN);   // 256
constexpr int GT_LDA = GT_BM + 4, GT_LDB = GT_BN + 4;
constexpr int GT_SMEM_FLOATS = GT_BK * GT_LDA + GT_BK * GT_LDB;

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

    for (int kb = k0; kb < k1; kb += GT_BK) {
        // stage A tile (BM x BK) and B tile (BK x BN)
#pragma unroll
        for (int e = tid; e < GT_BM * GT_BK; e += GT_THREADS) {
            int m, k;
            if (A_KFAST) { k = e % GT_BK; m = e / GT_BK; } else { m = e % GT_BM; k = e / GT_BM; }
            float v = (kb + k < k1) ? fa(m0 + m, kb + k) : 0.f;
            sA[k * GT_LDA + m] = v;
        }
#pragma unroll
        for (int e = tid; e < GT_BN * GT_BK; e += GT_THREADS) {
            int n, k;
            if (B_NFAST) { n = e % GT_BN; k = e / GT_BN; } else { k = e % GT_BK; n = e / GT_BK; }
            float v = (kb + k < k1) ? fb(kb + k, n0 + n) : 0.f;
            sB[k * GT_LDB + n] = v;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < GT_BK; ++k) {
            float4 a4 = *reinterpret_cast<const float4*>(&sA[k * GT_LDA + ty * GT_TM]);
            float4 b4 = *reinterpret_cast<const float4*>(&sB[k * GT_LDB + tx * GT_TN]);
            float a[4] = {a4.x, a4.y, a4.z, a4.w};
            float b[4] = {b4.x, b4.y, b4.z, b4.w};
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

}  // namespace dbmm
