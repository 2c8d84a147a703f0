// Batched-adapter training (BASELINE config 5; reference: the nested sequential sweep loops of
// run_multiple/final_main_iteration_wb.py:1129-1197 and run_multiple/final_main_iteration_ca.py:1179-1256).
//
// M members of a sweep (seeds x learning rates ...) train in LOCK STEP over the same resident embedding matrix: every kernel
// of the training step is launched ONCE per step for all members, with the member index in a grid dimension the kernel
// does not otherwise use.  A member is a MemberDev record in device memory (its batch order, adapters, optimizer state,
// statistics slots and carved workspace); the wrappers below patch the member's pointers into the by-value argument block
// (which carries everything the members share: X, labels, shapes, the step's position) and run the SAME kernel bodies as
// the single-run path.  One 1024-row step cannot fill 148 SMs (DESIGN.md "honest floor"); 64 of them can: GEMM-1 and dW1
// run un-split (one CTA per 128-row / 128-column tile walks all of K: no partial tiles, no reduction kernels), and the
// dependent-kernel latencies are shared by all members.
#pragma once
#include "gemm1_tc.cuh"
#include "rows_train.cuh"
#include "hs_rows.cuh"
#include "hs_w2.cuh"
#include "wgrad_tc.cuh"
#include "step_tail.cuh"
#include "tn_gemm.cuh"

namespace dbmm {

struct MemberDev {
    const int32_t* order;          // [n_rows] this member's batch order
    AdapterView ad[2];             // [0] frozen adapter (stage 2) or the trainable one again, [1] trainable adapter
    float* grads; float* mom;      // flat gradient / momentum
    const float* lr;               // [steps] learning rates of the running epoch
    double* loss_sum; int64_t* counts;
    TrainWs w;                     // the member's workspace, carved
};

__global__ void __launch_bounds__(256) k_zero_accum_b(const MemberDev* mem, int n_colsum, int n_dgb) {
    const MemberDev& m = mem[blockIdx.x];
    for (int e = threadIdx.x; e < n_colsum; e += 256) m.w.colsum[e].v = 0;
    for (int e = threadIdx.x; e < n_dgb; e += 256) m.w.dgb[e].v = 0;
}

template <int BN, int TERMS>
__global__ void __launch_bounds__(G1_THREADS, 1) k_gemm1_tc_b(Gemm1TcArgs a, const MemberDev* mem) {
    const MemberDev& m = mem[blockIdx.z];
    a.idx = m.order; a.A = m.w.A; a.colsum = m.w.colsum;
    const size_t hd = (size_t)a.H * a.D;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const size_t k = a.nad == 2 ? i : 0;
        a.Whi[i] = m.w.whi + k * hd; a.Wlo[i] = m.w.wlo + k * hd; a.b1[i] = m.ad[a.nad == 2 ? i : 1].b1;
    }
    gemm1_tc_body<BN, TERMS>(a, 0);
}

template <int NAD, int CT, int NW>
__global__ void __launch_bounds__(NW * 32) k_rows_train_b(RowsTrainArgs a, const MemberDev* mem, int64_t pos0) {
    const MemberDev& m = mem[blockIdx.y];
    a.idx = m.order + pos0; a.A = m.w.A; a.gram = m.w.gram; a.colsum = m.w.colsum;
    a.ad[0] = m.ad[0]; a.ad[1] = m.ad[1];
    a.loss_sum = m.loss_sum; a.counts = m.counts;
    a.dahat = m.w.dahat; a.dgb = m.w.dgb; a.Lrows = m.w.Lrows; a.Hrows = m.w.Hrows;
    rows_train_body<NAD, CT, NW>(a);
}

__global__ void __launch_bounds__(HR_THREADS, 1) k_hs_rows_b(HsRowsArgs a, const MemberDev* mem, int64_t pos0) {
    const MemberDev& m = mem[blockIdx.y];
    a.idx = m.order + pos0; a.A = m.w.A; a.gram = m.w.gram; a.colsum = m.w.colsum;
    a.ad[0] = m.ad[0]; a.ad[1] = m.ad[1];
    a.loss_sum = m.loss_sum; a.counts = m.counts;
    a.dahat = m.w.dahat; a.dgb = m.w.dgb; a.Spart = m.w.Spart;
    hs_rows_body(a);
}

__global__ void __launch_bounds__(256) k_sum_spart_b(SumSpartArgs a, const MemberDev* mem) {
    const MemberDev& m = mem[blockIdx.y];
    a.Spart = m.w.Spart; a.S = m.w.S; a.ST = m.w.ST;
    sum_spart_body(a);
}

__global__ void __launch_bounds__(HR_THREADS, 1) k_hs_w2_b(HsW2Args a, const MemberDev* mem, int step) {
    const MemberDev& m = mem[blockIdx.y];
    a.W2 = const_cast<float*>(m.ad[1].W2); a.b2 = const_cast<float*>(m.ad[1].b2); a.g = m.grads; a.v = m.mom;
    a.lr_dev = m.lr + step; a.ST = m.w.ST; a.Gpart = m.w.Gpart;
    hs_w2_body(a);
}

__global__ void __launch_bounds__(256) k_sum_gpart_b(SumGpartArgs a, const MemberDev* mem, int nad) {
    const MemberDev& m = mem[blockIdx.y];
    a.Gpart = m.w.Gpart; a.G = m.w.gram + (size_t)(nad - 1) * (a.H + 1) * (a.H + 1 + a.C);
    sum_gpart_body(a);
}

__global__ void __launch_bounds__(WG_THREADS, 1) k_wgrad_tc_b(WgradTcArgs a, const MemberDev* mem, int64_t pos0, int nad) {
    const MemberDev& m = mem[blockIdx.z];
    a.idx = m.order + pos0;
    a.A = m.w.A + (size_t)(nad - 1) * a.B * a.H; a.dahat = m.w.dahat;
    a.colsum = m.w.colsum + (size_t)(nad - 1) * 2 * a.H; a.dgb = m.w.dgb; a.dgb_wb = m.w.dgb;
    a.gamma = m.ad[1].gamma; a.part = m.w.part;
    wgrad_tc_body(a);
}

__device__ __forceinline__ void patch_tail(StepTailArgs& a, const MemberDev& m, int step) {
    a.W1 = const_cast<float*>(m.ad[1].W1); a.b1 = const_cast<float*>(m.ad[1].b1);
    a.gamma = const_cast<float*>(m.ad[1].gamma); a.beta = const_cast<float*>(m.ad[1].beta);
    a.W2 = const_cast<float*>(m.ad[1].W2); a.b2 = const_cast<float*>(m.ad[1].b2);
    a.g = m.grads; a.v = m.mom; a.lr_dev = m.lr + step;
    a.part = m.w.part;
    a.whi = m.w.whi + (size_t)(a.nad - 1) * a.H * a.D; a.wlo = m.w.wlo + (size_t)(a.nad - 1) * a.H * a.D;
    a.S = m.w.S; a.dgb = m.w.dgb; a.colsum = m.w.colsum;
    a.rm[0] = m.ad[0].running_mean; a.rv[0] = m.ad[0].running_var; a.nbt[0] = m.ad[0].nbt;
    a.rm[1] = m.ad[1].running_mean; a.rv[1] = m.ad[1].running_var; a.nbt[1] = m.ad[1].nbt;
}

__global__ void __launch_bounds__(ST_THREADS) k_tail_w1_b(StepTailArgs a, const MemberDev* mem, int step) {
    patch_tail(a, mem[blockIdx.y], step);
    tail_w1_body<false>(a);
}

template <int ST2_ROWS>
__global__ void __launch_bounds__(ST2_THREADS) k_tail_w2_b(StepTailArgs a, const MemberDev* mem, int step) {
    patch_tail(a, mem[blockIdx.y], step);
    tail_w2_body<false, ST2_ROWS>(a);
}

// which: 0 = S = L^T [h | 1] of the step's rows, 1 = Gram matrix of the trainable adapter's new W2 / b2
__global__ void __launch_bounds__(TNG_THREADS) k_tn_gemm_b(TnGemmArgs a, const MemberDev* mem, int which, int nad) {
    const MemberDev& m = mem[blockIdx.z];
    if (which == 0) { a.A.p[0] = m.w.Lrows; a.B.p[0] = m.w.Hrows; a.C = m.w.S; }
    else {
        a.A.p[0] = m.ad[1].W2; a.A.p[1] = m.ad[1].b2; a.B.p[0] = m.ad[1].W2; a.B.p[1] = m.ad[1].b2;
        a.C = m.w.gram + (size_t)(nad - 1) * a.M * a.N;
    }
    tn_gemm_body(a);
}

// dW1 batch chunks per member in the batched launch: fewer than the single run's 16 (the machine is already full; every
// extra chunk is another H x D partial tile through L2), more than 1 so that a CTA's serial walk over the batch stays short.
static inline int batched_wgrad_chunks(int B, int* rows_per_chunk) {
    static const int env = getenv("DBMM_BATCHED_CHUNKS") ? atoi(getenv("DBMM_BATCHED_CHUNKS")) : 0;
    int n = env >= 1 && env <= 16 ? env : 2;
    int rpc = (B + n - 1) / n;
    rpc = (rpc + WG_BK - 1) / WG_BK * WG_BK;
    *rows_per_chunk = rpc;
    return (B + rpc - 1) / rpc;
}

static bool batched_supported(int D, int H, int C) {
    return (D % G1_BK == 0) && (H % 32 == 0) && H == 128 && (D % WG_TILE == 0) && step_tail_supported(D, H, C);
}

// One lock-step training step of all members.  `dmem`: MemberDev[M] in device memory.
static int batched_step(const MemberDev* dmem, int M, int step, int64_t pos0, int B,
                        const float* X, int64_t ldx, const int32_t* y, const int32_t* grp, int D, int H, int C, int G, int nad,
                        float ebd_weight, const float* That, float inv_tau, float momentum, float wd, cudaStream_t st) {
    k_zero_accum_b<<<M, 256, 0, st>>>(dmem, nad * 2 * H, 2 * H);
    DBMM_LAUNCH_CHECK();
    {   // GEMM-1, un-split: A (+ bias) and the fixed-point column sums straight from the epilogue
        Gemm1TcArgs t;
        memset(&t, 0, sizeof(t));
        t.X = X; t.ldx = ldx; t.pos0 = pos0; t.B = B; t.D = D; t.H = H; t.nad = nad; t.ksplit = 1;
        using Cfg = G1Cfg<128, 2>;
        auto kern = k_gemm1_tc_b<128, 2>;
        DBMM_CUDA(set_smem(kern, Cfg::SMEM));
        const int kb = D / G1_BK;
        t.stages = kb < Cfg::STAGES ? kb : Cfg::STAGES;
        t.colsum = reinterpret_cast<fx64*>(1);          // non-null: the epilogue reduces the column sums (patched per member)
        kern<<<dim3(ceil_div(B, G1_BM), nad, M), G1_THREADS, Cfg::SMEM, st>>>(t, dmem);
        DBMM_LAUNCH_CHECK();
    }
    const bool tc_rows = hs_rows_supported(H, C) && !(getenv("DBMM_ROWS") && strcmp(getenv("DBMM_ROWS"), "simt") == 0);
    if (tc_rows) {
        HsRowsArgs ha;
        memset(&ha, 0, sizeof(ha));
        ha.B = B; ha.Bg = B; ha.y = y; ha.grp = grp; ha.H = H; ha.C = C; ha.G = G; ha.nad = nad; ha.strideA = (int64_t)B * H;
        ha.w_old = ebd_weight; ha.inv_tau = inv_tau; ha.inv_B = 1.0f / (float)B; ha.slot = step;
        DBMM_CUDA(set_smem(k_hs_rows_b, HR_SMEM));
        k_hs_rows_b<<<dim3(ceil_div(B, HR_ROWS), M), HR_THREADS, HR_SMEM, st>>>(ha, dmem, pos0);
        DBMM_LAUNCH_CHECK();
    } else
    {
        RowsTrainArgs ra;
        memset(&ra, 0, sizeof(ra));
        ra.B = B; ra.Bg = B; ra.y = y; ra.grp = grp; ra.H = H; ra.C = C; ra.G = G; ra.strideA = (int64_t)B * H;
        ra.w_old = ebd_weight; ra.inv_tau = inv_tau; ra.inv_B = 1.0f / (float)B; ra.slot = step;
        ra.loss_sum = reinterpret_cast<double*>(1); ra.counts = reinterpret_cast<int64_t*>(1);
        const int CT = C <= 4 ? 4 : 16, nw = nad == 1 ? 16 : 8;
        const size_t smem = rows_train_smem_bytes(H, C, nad, CT, nw);
        DBMM_CHECK_SHAPE(smem <= 227 * 1024, "train row kernel needs %zu bytes of shared memory", smem);
        int gx = ceil_div(B, RT_ROWS);
        if (gx > 148 * 2) gx = 148 * 2;
#define DBMM_RTB_CASE(NAD_, CT_, NW_)                                                        \
        do {                                                                                 \
            auto kern = k_rows_train_b<NAD_, CT_, NW_>;                                      \
            DBMM_CUDA(set_smem(kern, smem));                                                 \
            kern<<<dim3(gx, M), NW_ * 32, smem, st>>>(ra, dmem, pos0);                       \
        } while (0)
        if (nad == 1 && CT == 4) DBMM_RTB_CASE(1, 4, 16);
        else if (nad == 1) DBMM_RTB_CASE(1, 16, 16);
        else if (CT == 4) DBMM_RTB_CASE(2, 4, 8);
        else DBMM_RTB_CASE(2, 16, 8);
#undef DBMM_RTB_CASE
        DBMM_LAUNCH_CHECK();
    }
    int nchunk = 0;
    {
        WgradTcArgs t;
        memset(&t, 0, sizeof(t));
        t.X = X; t.ldx = ldx; t.B = B; t.Bg = B; t.D = D; t.H = H;
        nchunk = batched_wgrad_chunks(B, &t.rows_per_chunk);
        DBMM_CUDA(set_smem(k_wgrad_tc_b, WG_SMEM));
        const int n_sub_max = (t.rows_per_chunk + WG_BK - 1) / WG_BK;
        t.stages = n_sub_max < WG_STAGES ? n_sub_max : WG_STAGES;
        k_wgrad_tc_b<<<dim3(D / WG_TILE, nchunk, M), WG_THREADS, WG_SMEM, st>>>(t, dmem, pos0, nad);
        DBMM_LAUNCH_CHECK();
    }
    StepTailArgs ta;
    memset(&ta, 0, sizeof(ta));
    ta.momentum = momentum; ta.wd = wd; ta.That = That; ta.D = D; ta.H = H; ta.C = C; ta.nad = nad; ta.Bg = B; ta.nchunk = nchunk;
    ta.n_w1_ctas = ceil_div((int64_t)H * D / 4, ST_THREADS);
    if (ta.n_w1_ctas > 128) ta.n_w1_ctas = 128;
    k_tail_w1_b<<<dim3(ta.n_w1_ctas + 1, M), ST_THREADS, 0, st>>>(ta, dmem, step);
    DBMM_LAUNCH_CHECK();
    if (tc_rows) {   // S = sum of the row kernel's S^T tiles, in tile order
        SumSpartArgs sa;
        memset(&sa, 0, sizeof(sa));
        sa.tiles = ceil_div(B, HR_ROWS); sa.H = H; sa.C = C;
        const int n = (H + 1) * (H + 1 + C) + (H + 1 + C) * (s_stride(H) - (H + 1));
        k_sum_spart_b<<<dim3(ceil_div(n, 256), M), 256, 0, st>>>(sa, dmem);
        DBMM_LAUNCH_CHECK();
    } else
    {   // S = L^T [h | 1]
        TnGemmArgs g;
        memset(&g, 0, sizeof(g));
        g.A = cat_mat(nullptr, l_stride(H, C), H + 1 + C); g.B = cat_mat(nullptr, s_stride(H), H + 1);
        g.M = H + 1 + C; g.N = H + 1; g.K = B; g.ldc = s_stride(H); g.n_store = s_stride(H);
        DBMM_CUDA(set_smem(k_tn_gemm_b, TNG_SMEM));
        k_tn_gemm_b<<<dim3(ceil_div(g.M, TNG_TM), ceil_div(g.n_store, TNG_TN), M), TNG_THREADS, TNG_SMEM, st>>>(g, dmem, 0, nad);
        DBMM_LAUNCH_CHECK();
    }
    const bool tc_w2 = tc_rows && !(getenv("DBMM_W2") && strcmp(getenv("DBMM_W2"), "simt") == 0);
    if (tc_w2) {   // dW2a = [W2 | b2 | That] S -> SGD on W2 / b2 -> Gram tiles -> next step's Gram matrix
        HsW2Args wa;
        memset(&wa, 0, sizeof(wa));
        wa.oW2 = (size_t)H * D + 3 * (size_t)H; wa.ob2 = wa.oW2 + (size_t)D * H;
        wa.momentum = momentum; wa.wd = wd; wa.That = That; wa.D = D; wa.H = H; wa.C = C;
        DBMM_CUDA(set_smem(k_hs_w2_b, HW_SMEM));
        k_hs_w2_b<<<dim3(ceil_div(D, HW_ROWS), M), HR_THREADS, HW_SMEM, st>>>(wa, dmem, step);
        DBMM_LAUNCH_CHECK();
        SumGpartArgs sg;
        memset(&sg, 0, sizeof(sg));
        sg.tiles = ceil_div(D, HW_ROWS); sg.H = H; sg.C = C;
        k_sum_gpart_b<<<dim3(ceil_div((H + 1) * (H + 1 + C), 256), M), 256, 0, st>>>(sg, dmem, nad);
        DBMM_LAUNCH_CHECK();
        return DBMM_OK;
    }
    {   // dW2a = [W2 | b2 | That] S -> SGD on W2 / b2
        const int w2_rows = 64;
        ta.n_w2_ctas = ceil_div(D, w2_rows);
        const size_t smem = step_tail_w2_smem(w2_rows);
        DBMM_CUDA(set_smem(k_tail_w2_b<64>, smem));
        k_tail_w2_b<64><<<dim3(ta.n_w2_ctas, M), ST2_THREADS, smem, st>>>(ta, dmem, step);
        DBMM_LAUNCH_CHECK();
    }
    {   // next step's Gram matrix from the new W2 / b2
        TnGemmArgs g;
        memset(&g, 0, sizeof(g));
        g.A = cat_mat(nullptr, H, H, nullptr, 1, 1); g.B = cat_mat(nullptr, H, H, nullptr, 1, 1, That, C, C);
        g.M = H + 1; g.N = H + 1 + C; g.K = D; g.ldc = H + 1 + C; g.n_store = H + 1 + C;
        DBMM_CUDA(set_smem(k_tn_gemm_b, TNG_SMEM));
        k_tn_gemm_b<<<dim3(ceil_div(g.M, TNG_TM), ceil_div(g.n_store, TNG_TN), M), TNG_THREADS, TNG_SMEM, st>>>(g, dmem, 1, nad);
        DBMM_LAUNCH_CHECK();
    }
    return DBMM_OK;
}

}  // namespace dbmm
