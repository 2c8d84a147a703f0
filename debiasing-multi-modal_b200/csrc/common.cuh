// Shared helpers for the dbmm kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <utility>

#include "../../include/dbmm.h"

#define DBMM_WARP 32
#define DBMM_BN_EPS 1e-5f
#define DBMM_BN_MOMENTUM 0.1f

namespace dbmm {

void set_error(const char* fmt, ...);

#define DBMM_CHECK_ARG(cond, ...)                 \
    do {                                          \
        if (!(cond)) {                            \
            dbmm::set_error(__VA_ARGS__);         \
            return DBMM_ERR_INVALID_ARG;          \
        }                                         \
    } while (0)

#define DBMM_CHECK_SHAPE(cond, ...)               \
    do {                                          \
        if (!(cond)) {                            \
            dbmm::set_error(__VA_ARGS__);         \
            return DBMM_ERR_UNSUPPORTED_SHAPE;    \
        }                                         \
    } while (0)

#define DBMM_CUDA(call)                                                                       \
    do {                                                                                      \
        cudaError_t e__ = (call);                                                             \
        if (e__ != cudaSuccess) {                                                             \
            dbmm::set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__,              \
                            cudaGetErrorString(e__));                                         \
            return DBMM_ERR_CUDA;                                                             \
        }                                                                                     \
    } while (0)

#define DBMM_LAUNCH_CHECK()                                                                   \
    do {                                                                                      \
        cudaError_t e__ = cudaGetLastError();                                                 \
        if (e__ != cudaSuccess) {                                                             \
            dbmm::set_error("kernel launch failed at %s:%d: %s", __FILE__, __LINE__,          \
                            cudaGetErrorString(e__));                                         \
            return DBMM_ERR_CUDA;                                                             \
        }                                                                                     \
    } while (0)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Every kernel of the training step asks for the same L1 / shared-memory split (maximum shared): consecutive kernels
// with different carve-outs force the SMs to drain and reconfigure between launches.
template <typename K>
static inline cudaError_t set_smem(K kern, size_t dyn_bytes) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_bytes);
}

// Kernel launch with the programmatic-dependent-launch attribute (see ptx::pdl_wait): the next kernel's set-up and its
// loads of data that do not depend on the predecessor overlap the predecessor's tail.  Measured on B200 with the fused
// step tail: 40.1 vs 42.5 us/step.  DBMM_PDL=0 turns it off.
static inline bool pdl_enabled() {
    static const bool on = !(getenv("DBMM_PDL") && strcmp(getenv("DBMM_PDL"), "0") == 0);
    return on;
}
static thread_local bool g_plain_next_launch = false;     // set by a call site to launch ITS next kernel without the programmatic attribute
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = (pdl_enabled() && !g_plain_next_launch) ? 1 : 0;
    g_plain_next_launch = false;
    return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}
// true when `tag` is listed in DBMM_NOPDL (comma-separated kernel tags: tuning switch): that kernel is then launched without the
// programmatic attribute, i.e. its CTAs do not become resident before its predecessor has completed
static inline bool pdl_off_for(const char* tag) {
    const char* e = getenv("DBMM_NOPDL");
    return e && strstr(e, tag) != nullptr;
}

// ---- step timeline (measurement builds only, -DDBMM_TIMELINE): %globaltimer stamps of every step kernel -- entry of its first
// CTA, the moment that CTA passed its dependency wait, and the latest exit of any of its CTAs -- in a ring indexed by the kernel's
// launch count; scripts/step_timeline.py turns them into the timeline of a step inside the running epoch graph (nsys is not
// available here).  Compiles to nothing in the product build.
enum { TL_GEMM1 = 0, TL_REDUCE, TL_ROWS, TL_WGRAD, TL_TAIL, TL_TN, TL_W2, TL_GSUM, TL_KERNELS };
#ifdef DBMM_TIMELINE
constexpr int TL_RING = 1024;
__device__ unsigned long long g_tl[TL_RING][TL_KERNELS][3];            // entry, after the wait, last exit
__device__ unsigned long long g_tl_entry[TL_KERNELS], g_tl_exit[TL_KERNELS];
__device__ unsigned g_tl_count[TL_KERNELS];
__device__ __forceinline__ unsigned long long tl_now() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
struct TlScope {
    int k; bool on;
    __device__ __forceinline__ TlScope(int k_) : k(k_), on(threadIdx.x == 0) {
        if (on && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) g_tl_entry[k] = tl_now();
    }
    __device__ __forceinline__ ~TlScope() { if (on) atomicMax(&g_tl_exit[k], tl_now()); }
};
__device__ __forceinline__ void tl_wait(int k) {
    if (threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) {
        const unsigned i = atomicAdd(&g_tl_count[k], 1u);
        if (i > 0) g_tl[(i - 1) % TL_RING][k][2] = atomicExch(&g_tl_exit[k], 0ull);      // the previous launch has fully retired
        g_tl[i % TL_RING][k][0] = g_tl_entry[k];
        g_tl[i % TL_RING][k][1] = tl_now();
    }
}
#define DBMM_TL_SCOPE(k) TlScope tl_scope_(k)
#define DBMM_TL_WAIT(k) tl_wait(k)
#else
#define DBMM_TL_SCOPE(k)
#define DBMM_TL_WAIT(k)
#endif

__host__ __device__ static inline int s_stride(int H) { return (H + 1 + 3) & ~3; }     // row stride of the S matrix and of the [h | 1] rows
__host__ __device__ static inline int l_stride(int H, int C) { return (H + 1 + C + 3) & ~3; }     // row stride of the [c*h | c | ds] rows

// ---- deterministic cross-CTA accumulators.  The small per-step batch reductions that several CTAs add into (BatchNorm
// column sums, dgamma / dbeta) are 64-bit FIXED-POINT integers: integer atomics commute, so the sum does not depend on the
// order in which the CTAs arrive and a training run is bit-reproducible (fp32 / fp64 atomics are not: the last bit follows
// the arrival order).  The large ones (S, the Gram matrix) are atomics-free GEMMs (tn_gemm.cuh).  The wrapper type keeps
// plain floating-point arithmetic on a slot from compiling.  LOG2 = binary point: resolution 2^-LOG2, range +-2^(63-LOG2).
struct fx64 { long long v; };
constexpr int FX_COLSUM = 20;     // sum a, sum a^2 over the global batch: range 8.8e12, resolution 9.5e-7 (fp32 partials carry more error)
constexpr int FX_DGB = 40;        // dgamma, dbeta: range 8.4e6, resolution 9.1e-13
template <int LOG2>
__device__ __forceinline__ unsigned long long fx_bits(double x) { return (unsigned long long)__double2ll_rn(x * (double)(1ll << LOG2)); }
template <int LOG2>
__device__ __forceinline__ void fx_add(fx64* p, double x) { atomicAdd(reinterpret_cast<unsigned long long*>(&p->v), fx_bits<LOG2>(x)); }
template <int LOG2>
__device__ __forceinline__ double fx_val(long long bits) { return (double)bits * (1.0 / (double)(1ll << LOG2)); }
template <int LOG2>
__device__ __forceinline__ double fx_get(const fx64* p) { return fx_val<LOG2>(__ldcg(&p->v)); }

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// Device-side view of one adapter's tensors (copied by value into kernel arguments).
struct AdapterView {
    const float* W1;
    const float* b1;
    const float* gamma;
    const float* beta;
    float* running_mean;
    float* running_var;
    long long* nbt;
    const float* W2;
    const float* b2;
};

static inline AdapterView view_of(const dbmm_adapter* a) {
    AdapterView v;
    v.W1 = a->W1; v.b1 = a->b1; v.gamma = a->gamma; v.beta = a->beta;
    v.running_mean = a->running_mean; v.running_var = a->running_var;
    v.nbt = (long long*)a->num_batches_tracked; v.W2 = a->W2; v.b2 = a->b2;
    return v;
}

// Train-step workspace carve-up (all offsets in bytes, 256-byte aligned).
// Shared memory a training launch of the tensor-core kernels asks for: 120 KB: two of these CTAs cannot share an SM (each
// keeps its SM's load bandwidth), but the 49 KB CTAs of the W2 branch's TN GEMM can sit beside one.  With the full 193 KB ring
// a TN-GEMM CTA that reached an SM first kept the next tensor-core CTA off it until it retired (measured, scripts/train_only.py:
// 193 KB 49.2 us / step, 120 KB 44.0; the step takes 36.1 us without its W2 branch).  DBMM_SOLO_KB overrides.
static inline size_t train_smem_bytes(size_t needed, size_t full, int pack) {
    static const int env = getenv("DBMM_SOLO_KB") ? atoi(getenv("DBMM_SOLO_KB")) : -1;      // tuning switch
    if (env >= 0) return needed > (size_t)env * 1024 ? needed : (size_t)env * 1024;
    (void)pack;         // data parallel used to ask for the used ring only (two CTAs per SM): 2 GPUs 56.5 us / step, 120 KB 55.0, 193 KB 58.8
    const size_t solo = (size_t)120 * 1024;
    return needed > solo ? needed : (full < solo ? full : solo);
}
constexpr int DBMM_LR_TABLE = 65536;          // learning rates (one per step) a single epoch graph can address
constexpr int DBMM_G1_PART_ROWS = 16384;      // ksplit * nad * ceil128(B) never exceeds this (see gemm1_ksplit)

struct TrainWs {
    fx64* colsum;     // [nad][2][H]   sum a, sum a^2         (fixed point; zero at the start of every step)
    fx64* dgb;        // [2][H]        dgamma, dbeta          (fixed point; zero at the start of every step)
    float* S;         // [H+1+C][SP]   batch reduction for dW2 = L^T [h | 1] (k_tn_gemm), SP = H+1 rounded up to 4
    float* gram;      // [nad][H+1][H+1+C]
    float* S2;        // second half of the S double buffer (fused step tail)
    float* gram2;     // second half of the Gram double buffer (fused step tail: step s reads half s & 1, fills the other)
    int* tn_ticket;   // [16]          tile tickets of the K-sliced TN GEMMs (zero between launches)
    float* tn_part;   // K-slice tiles of the TN GEMMs (S and the Gram matrix run one after the other and share it)
    float* Spart;     // [ceil(B/64)][H+1][144]  S^T tiles of the tensor-core row kernel (hs_rows.cuh)
    float* ST;        // [H+1][144]    S^T summed over the tiles (operand of the tensor-core W2 kernel, hs_w2.cuh)
    float* Gpart;     // [ceil(D/64)][H+1][144]  Gram tiles of the tensor-core W2 kernel
    float* Lrows;     // [B][l_stride]  rows [c*h | c | ds] of the trainable adapter (operand of S)
    float* Hrows;     // [B][s_stride]  rows [h | 1]
    float* A;         // [nad][B][H]   pre-BatchNorm activations
    float* dahat;     // [B][H]        dL/d(normalised activation)
    float* whi;       // [nad][H][D]   tf32-exact part of W1 (tensor-core GEMM-1 operand)
    float* wlo;       // [nad][H][D]   W1 - whi
    float* part;      // [16][H][D]    batch-chunk partial tiles of dW1 (tensor-core path)
    float* g1part;    // [ksplit][nad][B][H] D-slice partial tiles of GEMM-1
    float* lr;        // [DBMM_LR_TABLE] per-step learning rates of the running epoch
    size_t accum_bytes;  // bytes of the zeroed region at the start (colsum, dgb, S, gram)
    size_t total;
};

static inline TrainWs carve_train_ws(void* base, int64_t B, int D, int H, int C, int nad) {
    TrainWs w;
    char* p = (char*)base;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
    size_t o_colsum = take(sizeof(fx64) * nad * 2 * H);
    size_t o_dgb = take(sizeof(fx64) * 2 * H);
    size_t o_S = take(sizeof(float) * (size_t)(H + 1 + C) * s_stride(H));
    size_t o_gram = take(sizeof(float) * (size_t)nad * (H + 1) * (H + 1 + C));
    size_t o_S2 = take(sizeof(float) * (size_t)(H + 1 + C) * s_stride(H));
    size_t o_gram2 = take(sizeof(float) * (size_t)nad * (H + 1) * (H + 1 + C));
    size_t o_tk = take(sizeof(int) * 32);
    w.accum_bytes = off;
    size_t o_tp = take(sizeof(float) * (size_t)8 * 16 * 32 * 48);       // TNG_MAX_KSPLIT * TNG_MAX_TILES * TNG_TM * TNG_TN
    // S^T shares: 64-row tiles of the tensor-core row kernel, or one share per CTA (8 rows, at most 296) of the CUDA-core one
    const size_t sp_tiles = (size_t)((B + 63) / 64), sp_ctas = (size_t)((B + 7) / 8) <= 296 ? (size_t)((B + 7) / 8) : 0;
    size_t o_sp = take(sizeof(float) * (sp_tiles > sp_ctas ? sp_tiles : sp_ctas) * (H + 1) * 144);
    size_t o_st = take(sizeof(float) * (size_t)(H + 1) * 144);
    size_t o_gp = take(sizeof(float) * (size_t)((D + 63) / 64) * (H + 1) * 144);
    size_t o_L = take(sizeof(float) * (size_t)B * l_stride(H, C));
    size_t o_Hr = take(sizeof(float) * (size_t)B * s_stride(H));
    size_t o_A = take(sizeof(float) * (size_t)nad * B * H);
    size_t o_da = take(sizeof(float) * (size_t)B * H);
    size_t o_whi = take(sizeof(float) * (size_t)nad * H * D);
    size_t o_wlo = take(sizeof(float) * (size_t)nad * H * D);
    size_t o_part = take(sizeof(float) * (size_t)16 * H * D);
    size_t o_g1 = take(sizeof(float) * (size_t)(DBMM_G1_PART_ROWS + 2 * 128) * H);
    size_t o_lr = take(sizeof(float) * DBMM_LR_TABLE);
    w.total = off;
    w.colsum = (fx64*)(p + o_colsum); w.dgb = (fx64*)(p + o_dgb);
    w.Spart = (float*)(p + o_sp); w.ST = (float*)(p + o_st); w.Gpart = (float*)(p + o_gp);
    w.Lrows = (float*)(p + o_L); w.Hrows = (float*)(p + o_Hr);
    w.tn_ticket = (int*)(p + o_tk); w.tn_part = (float*)(p + o_tp);
    w.A = (float*)(p + o_A); w.dahat = (float*)(p + o_da);
    w.gram = (float*)(p + o_gram); w.gram2 = (float*)(p + o_gram2); w.S = (float*)(p + o_S); w.S2 = (float*)(p + o_S2);
    w.whi = (float*)(p + o_whi); w.wlo = (float*)(p + o_wlo); w.part = (float*)(p + o_part);
    w.g1part = (float*)(p + o_g1); w.lr = (float*)(p + o_lr);
    return w;
}

}  // namespace dbmm
