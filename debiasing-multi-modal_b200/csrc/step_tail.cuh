// Fused tail of a training step: gradient finalisation AND optimizer step AND the next step's operands, as two kernels
// that the epoch graph launches on two branches so that the W2 half leaves the critical path.
//
//   k_tail_w1 (main branch, after k_wgrad_tc):  dW1 = sum of the batch-chunk partial tiles -> SGD on W1 -> tf32 split of
//       the new W1 (everything elementwise: one pass, 16 partial loads + p + v in flight per thread).  One more CTA
//       ("chores": k_wgrad_tc reads gamma, so not earlier): dgamma / dbeta / db1 -> SGD on b1 / gamma / beta, BatchNorm
//       running statistics (final_main.py:122,574: the frozen adapter drifts too).
//   k_tail_w2 (second branch: forked after k_rows_train, joined before the NEXT step's k_rows_train):
//       per CTA 64 / 32 / 16 embedding rows: dW2a = [W2 | b2 | That] S -> SGD on W2 / b2.  The same branch computes S before
//       it and the next step's Gram matrix G = [W2 | b2]^T [W2 | b2 | That] after it (k_tn_gemm: one CTA per output tile,
//       no atomics, bit-reproducible).
// The per-step accumulators are re-zeroed by the NEXT step's head kernels (column sums: k_gemm1_tc; dgamma / dbeta and S:
// k_reduce_stats), never here: k_wgrad_tc reads them while the W2 branch runs.  SGD as torch.optim.SGD
// (demo/util.py:118-136): g += wd*p; v = momentum*v + g; p -= lr*v (the caller zeroes v before the optimizer's first
// step, which makes v = g).  Data parallel (template P2P): the dW1 quads and S are summed over the ranks through peer
// memory inside these kernels (p2p.cuh) -- there is no gradient all-reduce kernel.
#pragma once
#include "common.cuh"
#include "ptx_sm100.cuh"
#include "p2p.cuh"
#include "tn_gemm.cuh"

namespace dbmm {

constexpr int ST_THREADS = 256, ST_MAXCHUNK = 16;                 // W1 role
constexpr int ST2_THREADS = 512, ST2_ROWS_MAX = 64;               // W2 role: 32 or 64 embedding rows per CTA (template)
constexpr int ST2_LP = 168, ST2_SP = 136, ST2_K8MAX = 152;        // smem row strides (== 8 mod 32: conflict-free mma fragments)
                                                                  // of the staged [W2 | b2 | That] rows / of S; max padded K

struct StepTailArgs {
    int roles;
    float* W1; float* b1; float* gamma; float* beta; float* W2; float* b2;    // trainable adapter (updated in place)
    float* g; float* v;                       // flat gradient (written: callers may read it back) / momentum
    const float* lr_dev; float lr; float momentum, wd;
    const float* part; int nchunk;            // [nchunk][H][D] batch-chunk partials of dW1
    float* whi; float* wlo;                   // tf32 split of the new W1
    const float* That; const float* S;        // [D][C]; [H+1+C][s_stride(H)] (k_tn_gemm's output)
    const fx64* dgb; const fx64* colsum;      // [2][H] (FX_DGB); [nad][2][H] (FX_COLSUM)
    int D, H, C, nad; int64_t Bg;
    float* rm[2]; float* rv[2]; long long* nbt[2];
    int n_w1_ctas, n_w2_ctas;
    P2pArgs p2p;                              // data parallel: dW1 slices and S are summed over the ranks through peer memory
};

static inline size_t step_tail_smem_bytes(int H, int C) {
    const size_t KP = (H + 1 + C + 3) & ~3, NP = (H + 1 + 3) & ~3;
    (void)KP; (void)NP;
    return sizeof(float) * ((size_t)ST2_K8MAX * ST2_SP + (size_t)ST2_ROWS_MAX * ST2_LP) + 16;
}

// ---- role bit 0: W1 CTAs + one chores CTA
template <bool P2P>
__device__ __forceinline__ void tail_w1_body(const StepTailArgs& a) {
    const int H = a.H, D = a.D;
    const int tid = threadIdx.x;
    const float lr = a.lr_dev ? __ldg(a.lr_dev) : a.lr;
    const size_t oW1 = 0, ob1 = (size_t)H * D;
    const int bid = blockIdx.x;
    if (bid < a.n_w1_ctas) {
        // chunk sum + SGD + tf32 split, 16-byte accesses (H * D is a multiple of 4)
        const int64_t n4 = (int64_t)H * D / 4;
        const size_t plane4 = (size_t)H * D / 4;
        // weights and momentum of the thread's first quad were written a whole step ago: requested before the wait
        const int64_t i0 = (int64_t)bid * ST_THREADS + tid;
        float4 pv0 = make_float4(0.f, 0.f, 0.f, 0.f), vv0 = pv0;
        if (i0 < n4) { pv0 = reinterpret_cast<const float4*>(a.W1)[i0]; vv0 = reinterpret_cast<const float4*>(a.v + oW1)[i0]; }
        ptx::pdl_wait();            // the chunk partials come from k_wgrad_tc
        DBMM_TL_WAIT(TL_TAIL);
        ptx::pdl_launch();
        unsigned inst = 0; int parity = 0;
        if constexpr (P2P) { inst = p2p_instance(a.p2p); parity = inst & 1u; }
        for (int64_t i = (int64_t)bid * ST_THREADS + tid; i < n4; i += (int64_t)a.n_w1_ctas * ST_THREADS) {
            float4 pt[ST_MAXCHUNK];
#pragma unroll
            for (int c = 0; c < ST_MAXCHUNK; ++c)
                pt[c] = c < a.nchunk ? __ldcg(reinterpret_cast<const float4*>(a.part) + c * plane4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
            if constexpr (P2P) {
                // Data parallel, reduce-scatter + all-gather by owner (LL words: data + instance tag in each 8-byte store, no
                // fence, no flag, no barrier).  A non-owner sends its chunk-summed quad to the owner and takes the owner's
                // UPDATED weights back; the owner adds the ranks' quads in rank order (own quad from registers), runs SGD
                // below and pushes the new weights to every rank: all replicas hold the same bits.
                float4 mine = pt[0];
#pragma unroll
                for (int c = 1; c < ST_MAXCHUNK; ++c) { mine.x += pt[c].x; mine.y += pt[c].y; mine.z += pt[c].z; mine.w += pt[c].w; }
                const int64_t per = (n4 + a.p2p.world - 1) / a.p2p.world;
                const int owner = (int)(i / per);
                char* me = a.p2p.peer[a.p2p.rank];
                if (owner != a.p2p.rank) {
                    p2p_g_store(p2p_g_ll(a.p2p.peer[owner], parity, a.p2p.rank), i, mine, inst + 1u);
                    const float4 w = p2p_g_load(a.p2p, p2p_w_ll(me, parity), i, inst + 1u, (a.p2p.skip & 4) != 0);
                    const float wx[4] = {w.x, w.y, w.z, w.w};
                    float hi[4], lo[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) { hi[q] = __uint_as_float(__float_as_uint(wx[q]) & 0xffffe000u); lo[q] = wx[q] - hi[q]; }
                    reinterpret_cast<float4*>(a.W1)[i] = w;
                    reinterpret_cast<float4*>(a.whi)[i] = make_float4(hi[0], hi[1], hi[2], hi[3]);
                    reinterpret_cast<float4*>(a.wlo)[i] = make_float4(lo[0], lo[1], lo[2], lo[3]);
                    continue;
                }
#pragma unroll
                for (int c = 0; c < ST_MAXCHUNK; ++c) pt[c] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int c = 0; c < P2P_MAX_WORLD; ++c)
                    if (c < a.p2p.world) pt[c] = c == a.p2p.rank ? mine : p2p_g_load(a.p2p, p2p_g_ll(me, parity, c), i, inst + 1u, (a.p2p.skip & 4) != 0);
            }
            const float4 pv = i == i0 ? pv0 : reinterpret_cast<const float4*>(a.W1)[i];
            const float4 vv = i == i0 ? vv0 : reinterpret_cast<const float4*>(a.v + oW1)[i];
            float4 gs = pt[0];
#pragma unroll
            for (int c = 1; c < ST_MAXCHUNK; ++c) { gs.x += pt[c].x; gs.y += pt[c].y; gs.z += pt[c].z; gs.w += pt[c].w; }
            const float px[4] = {pv.x, pv.y, pv.z, pv.w}, gx[4] = {gs.x, gs.y, gs.z, gs.w}, vx[4] = {vv.x, vv.y, vv.z, vv.w};
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
            reinterpret_cast<float4*>(a.whi)[i] = make_float4(hi[0], hi[1], hi[2], hi[3]);
            reinterpret_cast<float4*>(a.wlo)[i] = make_float4(lo[0], lo[1], lo[2], lo[3]);
            reinterpret_cast<float4*>(a.g + oW1)[i] = gs;
            if constexpr (P2P) {
                const float4 w = make_float4(po[0], po[1], po[2], po[3]);
                for (int r = 0; r < a.p2p.world; ++r)
                    if (r != a.p2p.rank) p2p_g_store(p2p_w_ll(a.p2p.peer[r], parity), i, w, inst + 1u);
            }
        }
        return;
    }
    ptx::pdl_wait();
    ptx::pdl_launch();
    // ---- chores CTA (this role because k_wgrad_tc reads gamma and must have finished): b1 / gamma / beta (db1 = sum_B da
    // vanishes identically under BatchNorm: b1 moves by weight decay only, see k_finalize_grads), BatchNorm running
    // statistics of every adapter in the forward
    for (int e = tid; e < 3 * H; e += ST_THREADS) {
        const int seg = e / H, j = e - seg * H;
        float* pp = (seg == 0 ? a.b1 : (seg == 1 ? a.gamma : a.beta)) + j;
        const size_t fo = ob1 + e;
        const float graw = seg == 0 ? 0.f : (float)fx_get<FX_DGB>(&a.dgb[(size_t)(seg - 1) * H + j]);
        const float pv = *pp;
        const float g = graw + a.wd * pv;
        const float vn = a.momentum * a.v[fo] + g;
        a.g[fo] = graw;
        a.v[fo] = vn;
        *pp = pv - lr * vn;
    }
    for (int e = tid; e < a.nad * H; e += ST_THREADS) {
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

template <bool P2P>
__global__ void __launch_bounds__(ST_THREADS) k_tail_w1(StepTailArgs a) { DBMM_TL_SCOPE(TL_TAIL); tail_w1_body<P2P>(a); }

// ---- role bit 1: W2 / b2 rows [d0, d0 + ST2_ROWS):  dW2a[d][n] = sum_k L[d][k] S[k][n],  L = [W2 | b2 | That], then SGD on
// those rows.  The contraction runs on the tensor cores as warp-level mma.sync m16n8k8 with 3xTF32 split operands
// (hi*hi + lo*hi + hi*lo, fp32 accumulate: ~2^-20 relative, the policy of DESIGN.md 3.2); 64-row tiles off the critical
// path do not warrant a tcgen05 / TMEM pipeline.
static inline size_t step_tail_w2_smem(int rows) { return sizeof(float) * ((size_t)ST2_K8MAX * ST2_SP + (size_t)rows * ST2_LP) + 16; }
// Embedding rows per W2-role CTA.  Measured on B200 (scripts/dp_time.py, scripts/train_only.py): one GPU runs best with 16
// fat CTAs on the ~20 SMs the 128-CTA kernels of the main branch leave free (41 us/step); under data parallelism the role
// also carries the S exchange and must be short, so it runs as 32 / 64 thin CTAs while the main branch's tensor-core
// kernels ask only for the stage ring they use and pack two to an SM (2 GPUs: 56 vs 62-69 us/step).  DBMM_W2_ROWS overrides.
static inline int step_tail_w2_rows(bool dp) {
    static const int env = getenv("DBMM_W2_ROWS") ? atoi(getenv("DBMM_W2_ROWS")) : 0;
    if (env == 16 || env == 32 || env == 64) return env;
    return dp ? 32 : 64;
}

template <bool P2P, int ST2_ROWS>
__device__ __forceinline__ void tail_w2_body(const StepTailArgs& a) {
    extern __shared__ __align__(16) float st_smem[];
    const int H = a.H, D = a.D, C = a.C;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    const float lr = a.lr_dev ? __ldg(a.lr_dev) : a.lr;
    const size_t oW2 = (size_t)H * D + 3 * (size_t)H, ob2 = oW2 + (size_t)D * H;
    const int K = H + 1 + C, N = H + 1, NPg = s_stride(H);            // NPg: row stride of S in global memory
    const int KT = (K + 7) >> 3, K8 = KT * 8;                          // k-steps of the dW2a contraction
    constexpr int LP = ST2_LP, SP = ST2_SP;
    const int w2 = blockIdx.x;
    float* sS = st_smem;                        // [K8][SP], zero padded
    float* sL = sS + (size_t)ST2_K8MAX * SP;    // [ST2_ROWS][LP], zero padded: the rows' [W2 | b2 | That]
    const int d0 = w2 * ST2_ROWS;
    const int n4row = NPg >> 2;                 // 16-byte chunks per row of S
    ptx::pdl_wait();        // S comes from k_tn_gemm, the kernel in front of this one on its stream (programmatic launch)
    ptx::pdl_launch();
    if constexpr (P2P) {
        // Data parallel: this CTA pushes its slice of the rank's S to every rank (off the critical path), raises S flag
        // [cta][rank] everywhere, then waits for every slice of every rank and sums the slots in rank order.
        const unsigned inst = p2p_instance(a.p2p);
        const int parity = inst & 1u;
        const int n4 = K * n4row, per = (n4 + (int)gridDim.x - 1) / (int)gridDim.x;
        const int lo4 = w2 * per, hi4 = min(n4, lo4 + per);
        for (int e = lo4 + tid; e < hi4; e += ST2_THREADS) {
            const float4 v = __ldcg(reinterpret_cast<const float4*>(a.S) + e);
            for (int r = 0; r < a.p2p.world; ++r) reinterpret_cast<float4*>(p2p_s_slot(a.p2p.peer[r], parity, a.p2p.rank))[e] = v;
        }
        __threadfence_system();
        __syncthreads();
        if (tid < a.p2p.world) {
            unsigned* f = p2p_s_flag(a.p2p.peer[tid], w2, a.p2p.rank);
            asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f), "r"(inst + 1u) : "memory");
        }
        if (tid < (int)gridDim.x * a.p2p.world) {
            const unsigned* f = p2p_s_flag(a.p2p.peer[a.p2p.rank], tid / a.p2p.world, tid % a.p2p.world);
            unsigned v = 0;
            unsigned long long t0 = 0;
            for (unsigned spin = 0;; ++spin) {
                asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
                if ((int)(v - (inst + 1u)) >= 0 || (a.p2p.skip & 8) || p2p_expired(a.p2p, t0, spin)) break;
            }
        }
        __syncthreads();
        char* me = a.p2p.peer[a.p2p.rank];
        for (int e = tid; e < n4; e += ST2_THREADS) {
            float4 acc4 = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int r = 0; r < a.p2p.world; ++r) {
                const float4 v = __ldcg(reinterpret_cast<const float4*>(p2p_s_slot(me, parity, r)) + e);
                acc4.x += v.x; acc4.y += v.y; acc4.z += v.z; acc4.w += v.w;
            }
            const int k = e / n4row, q = e - k * n4row;
            *reinterpret_cast<float4*>(sS + (size_t)k * SP + q * 4) = acc4;
        }
    } else {                // S rows: whole 16-byte chunks, everything in flight at once
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(sS);
        for (int e = tid; e < K * n4row; e += ST2_THREADS) {
            const int k = e / n4row, q = e - k * n4row;
            ptx::cp_async16(dst + ((size_t)k * SP + q * 4) * 4, a.S + (size_t)e * 4);
        }
    }
    {   // zero padding of sS: columns [NPg, SP) of every row, rows [K, K8)
        const int padc = SP - NPg;
        for (int e = tid; e < K * padc; e += ST2_THREADS) { const int k = e / padc; sS[(size_t)k * SP + NPg + (e - k * padc)] = 0.f; }
        for (int e = K * SP + tid; e < K8 * SP; e += ST2_THREADS) sS[e] = 0.f;
    }
    {   // W2 rows: 16-byte copies (H % 4 == 0); b2 / That / zero padding: scalars
        const int h4 = H >> 2;
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(sL);
        for (int e = tid; e < ST2_ROWS * h4; e += ST2_THREADS) {
            const int r = e / h4, q = e - r * h4, d = d0 + r;
            if (d < D) ptx::cp_async16(dst + ((size_t)r * LP + q * 4) * 4, a.W2 + (size_t)d * H + q * 4);
            else *reinterpret_cast<float4*>(sL + (size_t)r * LP + q * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        ptx::cp_async_commit();
        const int tail = LP - H;
        for (int e = tid; e < ST2_ROWS * tail; e += ST2_THREADS) {
            const int r = e / tail, k = H + (e - r * tail), d = d0 + r;
            float v = 0.f;
            if (d < D && k < K) v = k == H ? a.b2[d] : __ldg(a.That + (size_t)d * C + (k - H - 1));
            sL[(size_t)r * LP + k] = v;
        }
    }
    // ---- dW2a: warp w owns the 16-row tile mt = w % MTILES and the 8-column tiles nt = w / MTILES + NGRP j
    constexpr int MTILES = ST2_ROWS / 16, NGRP = (ST2_THREADS / 32) / MTILES;    // warps per 16-row tile
    constexpr int NJ = (17 + NGRP - 1) / NGRP;  // column tiles per warp at most (N <= 129 -> 17 tiles)
    const int mt = warp % MTILES, nt0 = warp / MTILES, NT = (N + 7) >> 3;
    const int r0 = mt * 16 + g;                  // this lane's rows: r0 and r0 + 8
    float2 vv[NJ][2];                            // momentum of the owned elements: requested before the contraction
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        const int n = (nt0 + NGRP * j) * 8 + 2 * t;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            const int d = d0 + r0 + 8 * hh;
            vv[j][hh] = make_float2(0.f, 0.f);
            if (nt0 + NGRP * j < NT && d < D) {
                if (n < H) vv[j][hh] = *reinterpret_cast<const float2*>(a.v + oW2 + (size_t)d * H + n);
                else if (n == H) vv[j][hh].x = a.v[ob2 + d];
            }
        }
    }
    ptx::cp_async_wait<0>();
    __syncthreads();
    float acc[NJ][4];
#pragma unroll
    for (int j = 0; j < NJ; ++j)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[j][q] = 0.f;
    for (int ks = 0; ks < KT; ++ks) {
        const int k0 = ks * 8;
        uint32_t ah[4], al[4];
        tf32_split(sL[(size_t)r0 * LP + k0 + t], ah[0], al[0]);
        tf32_split(sL[(size_t)(r0 + 8) * LP + k0 + t], ah[1], al[1]);
        tf32_split(sL[(size_t)r0 * LP + k0 + t + 4], ah[2], al[2]);
        tf32_split(sL[(size_t)(r0 + 8) * LP + k0 + t + 4], ah[3], al[3]);
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const int nt = nt0 + NGRP * j;
            if (nt < NT) {
                uint32_t bh[2], bl[2];
                tf32_split(sS[(size_t)(k0 + t) * SP + nt * 8 + g], bh[0], bl[0]);
                tf32_split(sS[(size_t)(k0 + t + 4) * SP + nt * 8 + g], bh[1], bl[1]);
                mma_3xtf32(acc[j], ah, al, bh, bl);
            }
        }
    }
    // SGD on the owned elements (c fragment: rows r0 / r0 + 8, columns n, n + 1); the new values replace the old in sL
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        const int nt = nt0 + NGRP * j, n = nt * 8 + 2 * t;
        if (nt >= NT) continue;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            const int r = r0 + 8 * hh, d = d0 + r;
            if (d >= D) continue;
            float* lrow = sL + (size_t)r * LP;
            const float g0 = acc[j][2 * hh], g1 = acc[j][2 * hh + 1];
            if (n < H) {
                const float p0 = lrow[n], p1 = lrow[n + 1];
                const float v0 = a.momentum * vv[j][hh].x + (g0 + a.wd * p0), v1 = a.momentum * vv[j][hh].y + (g1 + a.wd * p1);
                const float q0 = p0 - lr * v0, q1 = p1 - lr * v1;
                const size_t fo = (size_t)d * H + n;
                *reinterpret_cast<float2*>(a.W2 + fo) = make_float2(q0, q1);
                *reinterpret_cast<float2*>(a.v + oW2 + fo) = make_float2(v0, v1);
                *reinterpret_cast<float2*>(a.g + oW2 + fo) = make_float2(g0, g1);
            } else if (n == H) {
                const float p0 = lrow[H];
                const float v0 = a.momentum * vv[j][hh].x + (g0 + a.wd * p0);
                const float q0 = p0 - lr * v0;
                a.b2[d] = q0; a.v[ob2 + d] = v0; a.g[ob2 + d] = g0;
            }
        }
    }
}

template <bool P2P, int ST2_ROWS>
__global__ void __launch_bounds__(ST2_THREADS) k_tail_w2(StepTailArgs a) { tail_w2_body<P2P, ST2_ROWS>(a); }

static inline bool step_tail_supported(int D, int H, int C) {
    return H % 4 == 0 && D % 4 == 0 && H <= 128 && (H + 1 + C) <= ST2_K8MAX - 7 && s_stride(H) <= ST2_SP &&
           step_tail_smem_bytes(H, C) <= 227 * 1024;
}

static int launch_step_tail(StepTailArgs a, cudaStream_t st) {
    DBMM_CHECK_SHAPE(step_tail_supported(a.D, a.H, a.C), "step tail kernels: unsupported D=%d H=%d C=%d", a.D, a.H, a.C);
    DBMM_CHECK_ARG(a.nchunk <= ST_MAXCHUNK, "at most %d batch chunks (got %d)", ST_MAXCHUNK, a.nchunk);
#ifdef DBMM_EXPERIMENTS
    static const int skip_roles = getenv("DBMM_TAIL_SKIP") ? atoi(getenv("DBMM_TAIL_SKIP")) : 0;      // timing experiments only
    a.roles &= ~skip_roles;
#endif
    a.n_w1_ctas = ceil_div((int64_t)a.H * a.D / 4, ST_THREADS);
    if (a.n_w1_ctas > 128) a.n_w1_ctas = 128;
    const int w2_rows = step_tail_w2_rows(a.p2p.world > 1);
    a.n_w2_ctas = ceil_div(a.D, w2_rows);
    const bool p2p = a.p2p.world > 1;
    DBMM_CHECK_ARG(!p2p || (a.n_w2_ctas <= P2P_S_CTAS && (size_t)a.H * a.D <= P2P_G_FLOATS &&
                            (size_t)(a.H + 1 + a.C) * s_stride(a.H) <= P2P_S_FLOATS), "shape exceeds the peer-memory gradient slots");
    if (a.roles & 1) {
        // single GPU: no programmatic launch for the W1 role -- its 129 early-resident CTAs otherwise sit on the SMs the W2
        // branch's k_hs_w2 needs empty (42.9 vs 43.3 us / step); DBMM_NOPDL=none keeps the programmatic edge
        g_plain_next_launch = pdl_off_for("tail") || (!p2p && !pdl_off_for("none"));
        if (p2p) DBMM_CUDA(launch_pdl(k_tail_w1<true>, dim3(a.n_w1_ctas + 1), dim3(ST_THREADS), 0, st, a));
        else DBMM_CUDA(launch_pdl(k_tail_w1<false>, dim3(a.n_w1_ctas + 1), dim3(ST_THREADS), 0, st, a));
    }
    if (a.roles & 2) {
        const size_t smem = step_tail_w2_smem(w2_rows);
#define DBMM_W2_LAUNCH(P2P_, ROWS_)                                                         \
        do {                                                                                \
            DBMM_CUDA(set_smem(k_tail_w2<P2P_, ROWS_>, smem));                              \
            DBMM_CUDA(launch_pdl(k_tail_w2<P2P_, ROWS_>, dim3(a.n_w2_ctas), dim3(ST2_THREADS), smem, st, a));  \
        } while (0)
        if (p2p) { if (w2_rows == 64) DBMM_W2_LAUNCH(true, 64); else if (w2_rows == 32) DBMM_W2_LAUNCH(true, 32); else DBMM_W2_LAUNCH(true, 16); }
        else { if (w2_rows == 64) DBMM_W2_LAUNCH(false, 64); else if (w2_rows == 32) DBMM_W2_LAUNCH(false, 32); else DBMM_W2_LAUNCH(false, 16); }
#undef DBMM_W2_LAUNCH
        DBMM_LAUNCH_CHECK();
    }
    return DBMM_OK;
}

}  // namespace dbmm
