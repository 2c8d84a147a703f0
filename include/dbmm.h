/*
 * dbmm.h -- C ABI of the B200-native adapter hot path (libdbmm.so).
 *
 * The reference (Lainshower/debiasing-multi-modal) is pure Python/PyTorch and has no FFI of its own;
 * the entry points below are the operators its epoch loops call, one C function per operator,
 * each citing the reference code it replaces (paths are into the reference repository):
 *
 *   dbmm_normalize_text   final_main.py:77,136      text / text.norm(dim=0)
 *   dbmm_eval_fwd         final_main.py:655-719     validate(): eval-mode CustomCLIP/MultipleAdapter forward
 *                         final_main.py:66-92,121-158 + CE (302) + accuracy/update_dict (383-391)
 *   dbmm_train_step       final_main.py:455-466, 610-623   forward + CE + backward (+ optim step)
 *                         demo/util.py:118-136      optim.SGD (momentum 0.9, weight decay)
 *   dbmm_train_epoch      final_main.py:426-496, 571-653   the per-batch loop of one epoch
 *   dbmm_train_epoch_batched  run_multiple/final_main_iteration_wb.py:1129-1197   the sweep's members, one epoch in lock step
 *   dbmm_sgd_step         demo/util.py:118-136      SGD on a flat buffer (data-parallel path)
 *   dbmm_export_embeddings  demo/demo_visualization.ipynb:1117-1215  validate_adapter_with_return (features for the notebooks)
 *   dbmm_widen_f16        data/waterbirds_embeddings.py:69-78  embeddings -> float32 tensors (ingest, fp16 store)
 *   dbmm_group_counts     final_main.py:383-391     update_dict on given logits
 *   dbmm_logits_ce        final_main.py:757-759,768 raw-embedding cosine logits + CE (zero-shot head)
 *   dbmm_supcon_fwd/bwd   demo/visualizer_supcon.py:1532-1571  contrastive loss, all anchors at once
 *
 * Conventions: all pointers are DEVICE pointers unless the name ends in _host; fp32 storage,
 * int32 labels/indices, int64 counters.  No allocation, no exceptions: every function returns 0 on
 * success or a negative dbmm_status and records a message readable via dbmm_last_error()
 * (thread-local).  The caller owns every buffer including the workspace, whose size comes from
 * dbmm_workspace_bytes().  Work is enqueued on `stream` (a cudaStream_t passed as void*) and is
 * asynchronous with respect to the host.
 */
#ifndef DBMM_H
#define DBMM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DBMM_ABI_VERSION 1
#define DBMM_MAX_H 128      /* --adapter_feat_dim (final_main.py:241) supported up to 128 */
#define DBMM_MAX_C 16       /* prompts per head on the adapter path (2 class / 2 spurious / 4 group) */
#define DBMM_MAX_G 16       /* groups */

typedef enum dbmm_status {
    DBMM_OK = 0,
    DBMM_ERR_INVALID_ARG = -1,
    DBMM_ERR_UNSUPPORTED_SHAPE = -2,
    DBMM_ERR_WORKSPACE_TOO_SMALL = -3,
    DBMM_ERR_CUDA = -4
} dbmm_status;

/* One Adapter (final_main.py:160-174): Linear(D,H) -> BatchNorm1d(H) -> ReLU -> Linear(H,D).
 * Field order = state_dict order (layers.0.weight, layers.0.bias, layers.1.{weight,bias,
 * running_mean,running_var,num_batches_tracked}, layers.3.weight, layers.3.bias). */
typedef struct dbmm_adapter {
    float* W1;                  /* [H, D] */
    float* b1;                  /* [H]    */
    float* gamma;               /* [H]    */
    float* beta;                /* [H]    */
    float* running_mean;        /* [H]    */
    float* running_var;         /* [H]    */
    int64_t* num_batches_tracked; /* [1]  */
    float* W2;                  /* [D, H] */
    float* b2;                  /* [D]    */
} dbmm_adapter;

/* Trainable tensors in the flat order used by gradients / momentum:  W1 | b1 | gamma | beta | W2 | b2 */
static inline size_t dbmm_param_count(int D, int H) { return (size_t)2 * D * H + 3 * (size_t)H + D; }

/* phases of one training step (bit mask).  A single-GPU caller passes DBMM_PHASE_ALL; the
 * data-parallel caller interleaves all-reduces between the phases. */
#define DBMM_PHASE_GEMM1   1   /* a = x W1^T + b1, per-column sum / sum-of-squares            */
#define DBMM_PHASE_ROWS    2   /* BN, ReLU, logits, CE, row-wise backward, dgamma/dbeta sums  */
#define DBMM_PHASE_WGRAD   4   /* dW1, db1, dW2, db2, dgamma, dbeta -> flat gradient          */
#define DBMM_PHASE_UPDATE  8   /* SGD on the six tensors + BatchNorm running-stat update      */
#define DBMM_PHASE_ALL     15

/* workspace ops */
#define DBMM_OP_EVAL   1
#define DBMM_OP_TRAIN  2
#define DBMM_OP_HEAD   3
#define DBMM_OP_SUPCON 4

int dbmm_abi_version(void);
const char* dbmm_last_error(void);
const char* dbmm_build_info(void);

/* rows = rows per call (eval: chunk size used internally is bounded, pass N), n_adapters = 1 or 2 */
size_t dbmm_workspace_bytes(int op, int64_t rows, int D, int H, int C, int n_adapters);

/* Byte offsets (from the start of a DBMM_OP_TRAIN workspace) and element counts of the two accumulator blocks a
 * data-parallel caller must all-reduce between phases: column sums [n_adapters][2][H] after DBMM_PHASE_GEMM1 and
 * (dgamma, dbeta) [2][H] after DBMM_PHASE_ROWS.  Both are int64 FIXED-POINT sums (binary point at bit 20 / bit 40):
 * reduce them as int64 with SUM -- integer sums are order-independent, so the replicas stay bit-identical. */
int dbmm_train_accum_layout(int H, int n_adapters, size_t* colsum_offset, size_t* colsum_count,
                            size_t* dgb_offset, size_t* dgb_count);

/* That[:, c] = T[:, c] / ||T[:, c]||_2 ;  T, That: [D, C] row-major */
int dbmm_normalize_text(const float* T, float* That, int D, int C, void* stream);

/* Per-batch accumulators, one slot per batch (train step or eval batch):
 *   loss_sum[slot]        sum over the batch's rows of -log softmax(logits)[y]
 *   counts[slot][0][g]    #rows of group g whose argmax == y        (update_dict's `corr`)
 *   counts[slot][1][g]    #rows of group g                          (update_dict's `n`)
 * The kernels ADD into the slot; the caller zeroes the arrays before an epoch. */
typedef struct dbmm_batch_stats {
    double* loss_sum;           /* [n_slots]          */
    int64_t* counts;            /* [n_slots][2][G]    */
} dbmm_batch_stats;

/*
 * Eval-mode forward over N rows (validate / validate_zs, final_main.py:655-803).
 *   X[N_total, D] with row stride ldx (floats); idx == NULL -> rows 0..N-1, else rows idx[0..N-1].
 *   y, grp: labels per DATASET row (indexed like X), int32.  grp may be NULL (all rows count as group 0,
 *   pass G = 1); y may be NULL when no statistics are requested (logits / argmax only).
 *   old_ad == NULL: CustomCLIP(ad).  old_ad != NULL: MultipleAdapter(old_ad, ad) with ebd_weight.
 *   That: [D, C] column-normalised prompts; logits = (u . That) * inv_tau.
 *   batch_size: rows per stats slot (the reference's val/test loader batch size); slot = row / batch_size.
 *   logits_out [N, C] and pred_out [N] are optional (NULL to skip).
 */
int dbmm_eval_fwd(const float* X, int64_t ldx, const int32_t* idx, const int32_t* y, const int32_t* grp,
                  int64_t N, int D, int H, int C, int G,
                  const dbmm_adapter* old_ad, const dbmm_adapter* ad, float ebd_weight,
                  const float* That, float inv_tau, int64_t batch_size,
                  dbmm_batch_stats stats, float* logits_out, int32_t* pred_out,
                  void* ws, size_t ws_bytes, void* stream);

/*
 * The same eval forward over an fp16-RESIDENT embedding matrix (X16: IEEE half, row stride ldx halfs, rows contiguous in
 * dataset order: labels are indexed by row).  CLIP emits fp16 (clip/model.py:375-396 of the reference) and the packed store
 * keeps it losslessly, so this reads 2 bytes per element and runs kind::f16 tensor-core MMAs (csrc/eval_f16.cuh); W1, h and
 * the Gram block enter as scaled fp16 pairs (22 significant bits, as the tf32 hi + lo of dbmm_eval_fwd).  Needs H == 128,
 * D % 8 == 0, C <= 4 (dbmm_eval_f16_supported); callers fall back to dbmm_eval_fwd otherwise.  Replaces, like
 * dbmm_eval_fwd, validate / validate_zs of final_main.py:655-803.
 */
int dbmm_eval_f16_supported(int D, int H, int C);
size_t dbmm_eval_f16_workspace_bytes(int64_t rows, int D, int H, int C, int n_adapters);
int dbmm_eval_fwd_f16(const void* X16, int64_t ldx, const int32_t* y, const int32_t* grp,
                      int64_t N, int D, int H, int C, int G,
                      const dbmm_adapter* old_ad, const dbmm_adapter* ad, float ebd_weight,
                      const float* That, float inv_tau, int64_t batch_size,
                      dbmm_batch_stats stats, float* logits_out, int32_t* pred_out,
                      void* ws, size_t ws_bytes, void* stream);

/*
 * One training step of `--tl_method contrastive_adapter` (accepted by the reference's CLI, final_main.py:230, without a runnable
 * branch; formula and epoch: demo/visualizer_supcon.py:412-508, 1522-1587; forward_ca: workspace/jinsu/SupCon.ipynb:109-113):
 *     u = L2(adapter(L2(x)))  (pre_norm != 0: the input normalisation of forward_ca; head = identity),
 *     L = loss_weight * all-anchor supervised contrastive loss of the batch (labels[row], temperature 1 / inv_tau_cl),
 * forward, D-wide backward through the normalisation / GEMM-2 / ReLU / batch-stat BatchNorm / GEMM-1, SGD on the six adapter
 * tensors (torch.optim.SGD semantics) and the BatchNorm running statistics.  The B x B similarity and its gradients are the
 * tcgen05 GEMMs of dbmm_supcon_fwd / dbmm_supcon_bwd.  *loss_out += the weighted mean loss over the valid anchors.
 * grads: [dbmm_param_count] flat gradient (written), momentum_buf: flat momentum (zeroed when first_step).
 */
size_t dbmm_contrastive_workspace_bytes(int B, int D, int H);
/* The same step in three calls, for data parallelism with GLOBAL negatives (BASELINE config 3): forward of the rank's rows
 * (U_out [B, D] normalised embeddings), then -- on the caller's side -- all-gather, dbmm_supcon_fwd / _bwd with row0 / Bl,
 * reduce-scatter of the contrast-role gradient; backward from dL/du into the flat gradient (all-reduced by the caller);
 * SGD + BatchNorm running statistics.  The workspace carries the forward's state between the calls. */
int dbmm_contrastive_forward(const float* X, int64_t ldx, const int32_t* idx, const int32_t* labels, int B, int D, int H,
                             const dbmm_adapter* ad, int pre_norm, float* U_out, int32_t* labels_out, void* ws, size_t ws_bytes, void* stream);
int dbmm_contrastive_backward(const float* X, int64_t ldx, const int32_t* idx, int B, int D, int H, const dbmm_adapter* ad, int pre_norm,
                              float* dU, const float* dU2, float loss_weight, float* grads, void* ws, size_t ws_bytes, void* stream);
int dbmm_contrastive_apply(int B, int D, int H, const dbmm_adapter* ad, const float* grads, float* momentum_buf, float lr, float momentum,
                           float weight_decay, int first_step, void* ws, size_t ws_bytes, void* stream);
int dbmm_contrastive_step(const float* X, int64_t ldx, const int32_t* idx, const int32_t* labels, int B, int D, int H,
                          const dbmm_adapter* ad, int pre_norm, float inv_tau_cl, float loss_weight,
                          float* grads, float* momentum_buf, float lr, float momentum, float weight_decay, int first_step,
                          double* loss_out, int32_t* n_valid_out, void* ws, size_t ws_bytes, void* stream);

/* PCI bus id ("0000:1b:00.0") of a CUDA device: the host side binds each rank's staging threads / pinned buffers to the
 * GPU's NUMA node (no reference counterpart: the reference is single-GPU). */
int dbmm_device_pci_bus_id(int device, char* out, int len);

/*
 * One training step on B_local rows (train_one_epoch / train_reg_seq_one_epoch body).
 *   B_global: rows of the whole (possibly multi-GPU) batch -- BatchNorm statistics and the CE mean use it.
 *   old_ad != NULL selects the stage-2 MultipleAdapter step: both adapters run batch-stat BatchNorm and
 *   update their running stats, only `ad` receives gradients (final_main.py:121-140, demo/util.py:128).
 *   grads: [dbmm_param_count] flat gradient (written by WGRAD, read by UPDATE).
 *   momentum_buf: [dbmm_param_count]; first_step != 0 initialises it with the gradient (torch SGD).
 *   stats slot `slot` receives loss / group counts of this batch.
 */
int dbmm_train_step(int phases,
                    const float* X, int64_t ldx, const int32_t* idx, const int32_t* y, const int32_t* grp,
                    int B_local, int64_t B_global, int D, int H, int C, int G,
                    const dbmm_adapter* old_ad, const dbmm_adapter* ad, float ebd_weight,
                    const float* That, float inv_tau,
                    float* grads, float* momentum_buf, float lr, float momentum, float weight_decay, int first_step,
                    dbmm_batch_stats stats, int64_t slot,
                    void* ws, size_t ws_bytes, void* stream);

/*
 * The nn.Module boundary of the same step (final_main.py:66-80, 121-140: CustomCLIP / MultipleAdapter .forward in train()
 * mode; final_main.py:455-466: output = classifier(x); loss = criterion(output, y); loss.backward(); optimizer.step()):
 *   dbmm_train_forward   logits_out[B, C] with batch-statistics BatchNorm; update_running_stats != 0 also moves
 *                        running_mean / running_var / num_batches_tracked of every adapter in the forward, as torch does;
 *   dbmm_train_backward  flat gradient (dbmm_param_count floats, layout W1 | b1 | gamma | beta | W2 | b2) of the trainable
 *                        adapter from an arbitrary upstream dL/dlogits [B, C]; the forward is recomputed from X, so no
 *                        state has to survive between the two calls.
 * The caller's autograd engine and optimizer own the rest (modules.py wraps the pair in a torch.autograd.Function).
 */
int dbmm_train_forward(const float* X, int64_t ldx, const int32_t* idx, int B, int D, int H, int C,
                       const dbmm_adapter* old_ad, const dbmm_adapter* ad, float ebd_weight, const float* That, float inv_tau,
                       float* logits_out, int update_running_stats, void* ws, size_t ws_bytes, void* stream);
int dbmm_train_backward(const float* X, int64_t ldx, const int32_t* idx, int B, int D, int H, int C,
                        const dbmm_adapter* old_ad, const dbmm_adapter* ad, float ebd_weight, const float* That, float inv_tau,
                        const float* dlogits, float* grads, void* ws, size_t ws_bytes, void* stream);

/* As dbmm_train_step, for callers that chain the steps of an epoch themselves (the data-parallel epoch graph):
 *   fresh != 0   first step of a chain: accumulators are zeroed, Gram matrices and tf32 weight splits computed from
 *                scratch; fresh == 0 relies on the previous step's WGRAD / UPDATE phases having prepared them (same
 *                workspace, same adapters, no other dbmm call in between);
 *   lr_dev       if not NULL the learning rate is read from this device address when the update kernel runs (a
 *                captured CUDA graph can then be replayed with a new schedule). */
int dbmm_train_step_ex(int phases, int fresh,
                       const float* X, int64_t ldx, const int32_t* idx, const int32_t* y, const int32_t* grp,
                       int B_local, int64_t B_global, int D, int H, int C, int G,
                       const dbmm_adapter* old_ad, const dbmm_adapter* ad, float ebd_weight,
                       const float* That, float inv_tau,
                       float* grads, float* momentum_buf, float lr, const float* lr_dev, float momentum, float weight_decay,
                       int first_step, dbmm_batch_stats stats, int64_t slot,
                       void* ws, size_t ws_bytes, void* stream);

/*
 * A whole single-GPU epoch: ceil(n_rows / batch_size) steps over rows order[0..n_rows-1] (device int32,
 * the injected batch order), learning rate per step from lr_host[] (HOST array, one entry per step:
 * adjust_learning_rate + warmup_learning_rate[_reg], demo/util.py:70-115).  `first_step` applies to step 0.
 * Stats slot s receives step s.  A trailing batch of one row is rejected like torch's BatchNorm1d does.
 */
int dbmm_train_epoch(const float* X, int64_t ldx, const int32_t* order, int64_t n_rows, int batch_size,
                     const int32_t* y, const int32_t* grp, int D, int H, int C, int G,
                     const dbmm_adapter* old_ad, const dbmm_adapter* ad, float ebd_weight,
                     const float* That, float inv_tau,
                     float* grads, float* momentum_buf, const float* lr_host, float momentum, float weight_decay,
                     int first_step, dbmm_batch_stats stats,
                     void* ws, size_t ws_bytes, void* stream);

/*
 * Data-parallel epoch over the GPUs of one NVSwitch box, one process per GPU.  The library owns its NCCL communicator
 * (libnccl.so.2 is bound with dlopen, preferring the copy PyTorch has already loaded): rank 0 calls
 * dbmm_comm_unique_id, the host broadcasts the 128 bytes (torch.distributed), every rank calls dbmm_comm_init.
 * dbmm_train_epoch_dp is dbmm_train_epoch with, per step, the all-reduces of the BatchNorm column sums, the (dgamma,
 * dbeta) sums and the flat gradient between the phases, so that BatchNorm and the CE mean see the GLOBAL batch
 * (the reference is single-process); kernels and collectives are captured in one CUDA graph.
 * When the ranks can map each other's memory (CUDA IPC over NVLink, world <= 8) the two small fp64 vectors are NOT sent
 * through NCCL: the producing kernel's last CTA pushes them into every rank's symmetric buffer and raises flags, the
 * consuming kernel's prologue waits on its local flags and sums the slots in rank order (csrc/p2p.cuh); only the 1 MB
 * gradient goes through ncclAllReduce.  DBMM_P2P=0 forces NCCL for all three.
 *   local_batches == 0: `order` is the global batch order (identical on all ranks); each rank trains on its contiguous
 *                       shard of every batch -- same result as one GPU.
 *   local_batches != 0: `order` lists this rank's own rows (equal n_rows / batch_size on all ranks); the global batch is
 *                       the union, world x batch_size rows (weak scaling).
 *   reduce_stats  != 0: the per-batch loss sums / group counters are all-reduced at the end of the epoch.
 */
int dbmm_comm_unique_id(void* id_out_128_bytes);
int dbmm_comm_init(const void* id_128_bytes, int world, int rank, void** comm_out);
int dbmm_comm_has_p2p(void* comm);      /* 1 if the fused peer-memory all-reduce (below) is active, 0 if NCCL is used throughout */
/* Synchronises the device; DBMM_ERR_CUDA if a peer-memory wait of an earlier epoch timed out (a rank never arrived within
 * DBMM_P2P_TIMEOUT_S seconds, default 300): the ranks must issue the same epoch calls in the same order. */
int dbmm_comm_check(void* comm);
int dbmm_comm_destroy(void* comm);
int dbmm_train_epoch_dp(void* comm, int world, int rank, int local_batches,
                        const float* X, int64_t ldx, const int32_t* order, int64_t n_rows, int batch_size,
                        const int32_t* y, const int32_t* grp, int D, int H, int C, int G,
                        const dbmm_adapter* old_ad, const dbmm_adapter* ad, float ebd_weight,
                        const float* That, float inv_tau,
                        float* grads, float* momentum_buf, const float* lr_host, float momentum, float weight_decay,
                        int first_step, dbmm_batch_stats stats, int reduce_stats,
                        void* ws, size_t ws_bytes, void* stream);

/*
 * Batched-adapter training (BASELINE config 5): n_members members of a sweep -- the (seed, learning-rate, ...) grid the
 * reference walks one run after another, run_multiple/final_main_iteration_wb.py:1129-1197, _ca.py:1179-1256 -- train one
 * epoch in LOCK STEP over the same resident X / labels: every kernel of the step is launched once for all members (member
 * index in the grid), so a sweep fills the GPU that a single 1024-row step cannot.  Members share the data, the batch size
 * (hence the step count), the prompt matrix, momentum and weight decay, and are all in the same stage (old_ad NULL or not);
 * each has its own batch order, adapters, optimizer state, learning-rate schedule, statistics slots and workspace.  The
 * result of every member equals dbmm_train_epoch on that member alone (same arithmetic up to the summation order of the
 * un-split GEMMs: ~1e-6 relative).  `members` is a HOST array; lr_host is [n_members][steps] (host); bws is a device
 * buffer of dbmm_batched_workspace_bytes() that must stay untouched until the epoch has run.
 */
typedef struct dbmm_member {
    const int32_t* order;              /* device [n_rows]: this member's batch order */
    const dbmm_adapter* old_ad;        /* NULL (stage 1) or the frozen adapter (stage 2) */
    const dbmm_adapter* ad;            /* trainable adapter, updated in place */
    float* grads; float* momentum_buf; /* flat [dbmm_param_count] each */
    dbmm_batch_stats stats;            /* this member's per-batch slots */
    void* ws; size_t ws_bytes;         /* this member's DBMM_OP_TRAIN workspace */
} dbmm_member;
size_t dbmm_batched_workspace_bytes(int n_members, int64_t steps);
int dbmm_train_epoch_batched(int n_members, const dbmm_member* members,
                             const float* X, int64_t ldx, int64_t n_rows, int batch_size, const int32_t* y, const int32_t* grp,
                             int D, int H, int C, int G, float ebd_weight, const float* That, float inv_tau,
                             const float* lr_host, float momentum, float weight_decay, int first_step,
                             void* bws, size_t bws_bytes, void* stream);

/* Measurement aid: the same epoch as stream launches with CUDA events between the kernels of every step;
 * kernel_us_host[6] = mean device microseconds of GEMM-1, reduce/statistics, row kernel, dW1, gradient finalisation,
 * update, each timed inside the running step.  Synchronises the stream. */
int dbmm_train_epoch_profile(const float* X, int64_t ldx, const int32_t* order, int64_t n_rows, int batch_size,
                             const int32_t* y, const int32_t* grp, int D, int H, int C, int G,
                             const dbmm_adapter* old_ad, const dbmm_adapter* ad, float ebd_weight,
                             const float* That, float inv_tau, float* grads, float* momentum_buf, const float* lr_host,
                             float momentum, float weight_decay, dbmm_batch_stats stats,
                             void* ws, size_t ws_bytes, void* stream, float* kernel_us_host);

/* Adapted-embedding export for the visualisation notebooks (validate_adapter_with_return, demo/demo_visualization.ipynb:
 * 1117-1215): out[N, D] = the UN-normalised adapter output (single adapter; normalize_single != 0: L2-normalised rows) or
 * the MultipleAdapter mix w * u_old + (1 - w) * u_new; optionally the notebook's logits = out @ That / tau for two prompt
 * sets (class, spurious; That column-normalised [D, C]).  Eval mode (running statistics).  Workspace: DBMM_OP_EVAL. */
int dbmm_export_embeddings(const float* X, int64_t ldx, const int32_t* idx, int64_t N, int D, int H,
                           const dbmm_adapter* old_ad, const dbmm_adapter* ad, float ebd_weight, int normalize_single,
                           const float* That_a, int Ca, const float* That_b, int Cb, float inv_tau,
                           float* out, int64_t ld_out, float* logits_a, float* logits_b,
                           void* ws, size_t ws_bytes, void* stream);

/* How single-GPU epochs of this shape run the tail of a step: 0 = k_finalize_grads + k_update, 1 = fused step tail
 * (k_tail_w1, k_tail_w2) in line, 2 = fused with the W2 role on a second branch of the epoch graph (default when every
 * step takes the tensor-core kernels).  Names the kernels behind dbmm_train_epoch_profile's slots 4 and 5. */
int dbmm_train_tail_mode(int batch_size, int last_batch, int n_adapters, int D, int H, int C);

/* torch.optim.SGD on a flat buffer: g += wd*p; v = g (first_step) or momentum*v + g; p -= lr*v */
int dbmm_sgd_step(float* p, const float* g, float* v, int64_t n, float lr, float momentum, float weight_decay,
                  int first_step, void* stream);

/* Ingest: fp16 rows [n_rows, D] (row stride ld_src halves) -> fp32 rows (row stride ld_dst floats), exact.  CLIP emits
 * fp16 embeddings (clip_inference.py:169), so the packed store keeps them as fp16 when that is lossless and the host ->
 * device copy moves half the bytes; the widening runs on the device, on `stream` (the copy stream in the e2e path).
 * Replaces the per-item np.array(list) -> float32 tensor of data/waterbirds_embeddings.py:69-78. */
int dbmm_widen_f16(const void* src_f16, int64_t ld_src, float* dst, int64_t ld_dst, int64_t n_rows, int D, void* stream);

/* update_dict (final_main.py:383-391) on given logits [N, C]: adds into stats slot row/batch_size. */
int dbmm_group_counts(const float* logits, const int32_t* y, const int32_t* grp, int64_t N, int C, int G,
                      int64_t batch_size, dbmm_batch_stats stats, int32_t* pred_out, void* stream);

/*
 * Zero-shot head on raw embeddings (validate_zs with --tl_method linear_probing / feature-quality check,
 * final_main.py:757-768, and BASELINE config 4: N rows x C up to 1,000+ prompt columns):
 *   logits = (u / ||u||) . That * inv_tau  (normalize_rows != 0), CE against y, argmax, per-group counters.
 * The logits are never materialised: a tcgen05 GEMM reduces each 128-column tile to online-softmax partials.
 *   U[N_total, D] row stride ldu (multiple of 4), idx == NULL -> rows 0..N-1; That: [D, C] column-normalised prompts;
 *   col_bias: optional [C] additive bias (NULL for the cosine head).
 */
size_t dbmm_head_workspace_bytes(int64_t N, int D, int C, int gathered);
int dbmm_logits_ce(const float* U, int64_t ldu, const int32_t* idx, const int32_t* y, const int32_t* grp,
                   int64_t N, int D, int C, int G, const float* That, const float* col_bias, float inv_tau, int normalize_rows,
                   int64_t batch_size, dbmm_batch_stats stats, int32_t* pred_out, void* ws, size_t ws_bytes, void* stream);

/*
 * The same head over an fp16-RESIDENT embedding matrix (CLIP emits fp16, clip_inference.py:169; the packed store keeps it):
 * U16[N, D] halves, row stride ldu (multiple of 8), rows 0..N-1.  kind::f16 tcgen05 MMAs on the rows as stored; the prompts
 * enter as a scaled fp16 pair (22 significant bits), so the logits agree with dbmm_logits_ce to fp32 rounding.  D % 8 == 0,
 * D >= 64.  Replaces the same reference lines as dbmm_logits_ce (final_main.py:757-768).
 */
size_t dbmm_head_f16_workspace_bytes(int64_t N, int D, int C);
int dbmm_logits_ce_f16(const void* U16, int64_t ldu, const int32_t* y, const int32_t* grp,
                       int64_t N, int D, int C, int G, const float* That, const float* col_bias, float inv_tau, int normalize_rows,
                       int64_t batch_size, dbmm_batch_stats stats, int32_t* pred_out, void* ws, size_t ws_bytes, void* stream);

/*
 * Linear probing (--tl_method linear_probing: LinearClassifier, final_main.py:43-49, trained by train_one_epoch,
 * final_main.py:426-496): one epoch of logits = x W^T + b, CE, dW / db, SGD (momentum, weight decay) over rows
 * order[0..n_rows-1] in batches of batch_size; W [C, D], b [C] updated in place; grads / momentum_buf: flat [C*D + C].
 * Evaluation of a linear probe = dbmm_logits_ce with That = W^T ([D, C]), col_bias = b, inv_tau = 1, normalize_rows = 0.
 */
int dbmm_linear_train_epoch(const float* X, int64_t ldx, const int32_t* order, int64_t n_rows, int batch_size,
                            const int32_t* y, const int32_t* grp, int D, int C, int G, float* W, float* b,
                            float* grads, float* momentum_buf, const float* lr_host, float momentum, float weight_decay,
                            int first_step, dbmm_batch_stats stats, void* stream);

/*
 * Contrastive regulariser, all anchors of the batch at once (formula: SupervisedContrastiveLoss.forward,
 * demo/visualizer_supcon.py:1532-1571, which the reference evaluates for ONE anchor per call in a Python loop,
 * demo/visualizer_supcon.py:458-485).  Z_all: [Bg, d] L2-normalised rows of the GLOBAL batch (all-gathered under data
 * parallelism); this rank's anchors are rows [row0, row0 + Bl).  For anchor i with positives P_i (same label, not
 * itself) and negatives N_i:  loss_i = log sum_{j != i} exp(z_i.z_j / tau) - mean_{p in P_i} z_i.z_p / tau ; anchors
 * without a positive or without a negative are skipped; the loss is the mean over the valid anchors.
 *   fwd:  *loss_sum += sum of loss_i, *n_valid += #valid anchors (device scalars: all-reduce them across ranks);
 *         row_loss[Bl] optional.  The similarity gradient stays in the workspace for bwd.
 *   bwd:  dZ_local[Bl, d]  = anchor-role gradient of this rank's rows,
 *         dZ_all[Bg, d] (+)= contrast-role gradient of every row of the global batch from this rank's anchors
 *         (reduce-scatter it across ranks; on one GPU the gradient is dZ_local + dZ_all), both scaled by 1 / *n_valid_global.
 *         When this rank holds EVERY anchor (row0 == 0, Bl == Bg) only the SUM dZ_local + dZ_all is defined per row: the
 *         fp16-pair path folds both roles into one GEMM, dZ_local = (G + G^T) Z, and leaves dZ_all zero / unchanged.
 * GEMMs: kind::f16 tcgen05 on fp16 pairs of the power-of-two-scaled operands (22 significant bits, pair_gemm.cuh) when
 * d % 8 == 0, d >= 64 (then row0 % 8 == 0 is required); 3xTF32 otherwise (DBMM_SUPCON=tf32 forces it).
 */
size_t dbmm_supcon_workspace_bytes(int Bl, int Bg, int d);
int dbmm_supcon_fwd(const float* Z_all, int Bg, int d, int64_t row0, int Bl, const int32_t* labels, float inv_tau_cl,
                    double* loss_sum, int32_t* n_valid, float* row_loss, void* ws, size_t ws_bytes, void* stream);
int dbmm_supcon_bwd(const float* Z_all, int Bg, int d, int64_t row0, int Bl, float inv_tau_cl, const int32_t* n_valid_global,
                    float* dZ_local, float* dZ_all, int accumulate_all, void* ws, size_t ws_bytes, void* stream);

/* Measurement aid (-DDBMM_TIMELINE builds only; an error in the product build): %globaltimer stamps {entry of the first CTA,
 * after its dependency wait, last CTA exit} of every kernel of the training step, ring[1024][8 kernels][3], for
 * scripts/step_timeline.py (the timeline of a step inside the running epoch graph). */
int dbmm_timeline_dump(unsigned long long* out_host, unsigned* counts_host, int* ring, int* kernels);

#ifdef __cplusplus
}
#endif
#endif /* DBMM_H */
