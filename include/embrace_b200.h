/*
 * embrace_b200.h -- C ABI of libembrace_sm100.so, the B200 (sm_100a) engine for the
 * EmbraceNet training / inference hot path of the reference
 * (BIOINF_tesi/models/*, BIOINF_tesi/models/utils/training_models_multimodal.py).
 *
 * The reference has no FFI boundary of its own (it is pure Python on top of PyTorch);
 * the entry points below are what its Python surface binds through ctypes
 * (see INTEGRATION.md).  Each entry cites the reference interface it replaces.
 *
 * Conventions
 *   - plain C, no torch types; every pointer is a DEVICE pointer unless the name ends in
 *     `_host`; sizes in elements unless the name says bytes.
 *   - every function returns 0 on success or a negative EMB_E_* code; the message is
 *     available from emb_last_error() (thread local).
 *   - all work is enqueued on the `stream` argument (a cudaStream_t passed as void*);
 *     no function synchronises unless documented; nothing allocates after emb_bind().
 *   - activations/gradients are laid out channels-last ([B, L, C]); parameters and their
 *     gradients keep the reference's state_dict shapes (row-major) in flat fp32 arenas.
 *   - there is NO CPU fallback: if no sm_100 device is present every compute entry fails
 *     with EMB_E_NO_DEVICE.
 */
#ifndef EMBRACE_B200_H
#define EMBRACE_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EMB_ABI_VERSION 2

#define EMB_MAX_FFNN 4   /* FFNN_pre.py:19  suggest_int("FFNN_n_layers", 1, 4) */
#define EMB_MAX_CNN 4    /* CNN_pre.py:24   suggest_int("CNN_n_layers", 1, 4)  */
#define EMB_MAX_POST 3   /* EmbraceNetMultimodal.py:135 n_post_layers in 0..2; ConcatNetMultimodal.py:42 CONCATNET_n_post_layers in 1..3 */
#define EMB_SEQ_LEN 256  /* CNN_pre.py:21 */
#define EMB_POOL_K 10    /* CNN_pre.py:18 */
#define EMB_POOL_S 2     /* CNN_pre.py:20 */

enum { EMB_OK = 0, EMB_E_ARG = -1, EMB_E_NO_DEVICE = -2, EMB_E_CUDA = -3, EMB_E_STATE = -4, EMB_E_UNSUPPORTED = -5 };

/* model kind: EmbraceNetMultimodal.py:94 / FF_net.py:8 / CNN_net.py:10 / ConcatNetMultimodal.py:12 */
enum { EMB_KIND_EMBRACENET = 0, EMB_KIND_FFNN = 1, EMB_KIND_CNN = 2, EMB_KIND_CONCATNET = 3 };
/* arithmetic of activations and GEMM operands (accumulation, BN statistics, loss, master
 * weights and optimizer state are always fp32) */
enum { EMB_PREC_FP32 = 0, EMB_PREC_BF16 = 1 };
/* optimizer rules of training_models_multimodal.py:318-325 (+ north_star's AdamW) */
enum { EMB_OPT_ADAM = 0, EMB_OPT_ADAMW = 1, EMB_OPT_NADAM = 2, EMB_OPT_RMSPROP = 3 };

/* The architecture an Optuna trial (or a checkpoint's model_params dict) selects.
 * Replaces: the trial.suggest_* calls in FFNN_pre.py:19-39, CNN_pre.py:24-51,
 * EmbraceNetMultimodal.py:123-157 and their *_NoTrain twins. */
typedef struct EmbArchSpec {
    int32_t kind;
    int32_t in_features;                       /* in_features_FFNN */
    int32_t n_ffnn;
    int32_t ffnn_units[EMB_MAX_FFNN];
    float   ffnn_dropout[EMB_MAX_FFNN];
    int32_t n_cnn;
    int32_t cnn_channels[EMB_MAX_CNN];
    int32_t cnn_kernels[EMB_MAX_CNN];          /* odd; padding = (k-1)/2 */
    float   cnn_dropout[EMB_MAX_CNN];
    int32_t embracement_size;                  /* "c" of EmbraceNet */
    int32_t n_post;
    int32_t post_units[EMB_MAX_POST];
    float   post_dropout[EMB_MAX_POST];
    double  p_ffnn;                            /* selection_probabilities_FFNN */
    int32_t embracenet_dropout;                /* modality dropout during training (default 1) */
    int32_t reserved;
} EmbArchSpec;

/* Explicit random draws for one forward (replay mode; every pointer may be NULL, meaning
 * "draw with the engine's counter-based Philox generator").  Layouts are the REFERENCE's:
 * dropout uniforms are indexed like the tensor nn.Dropout sees ([B,units] / [B,C,Lpool]),
 * keep = (u >= p); embrace_u is [B, embracement_size] fp64, idx = (u > cum0)
 * (torch.multinomial's CPU path, EmbraceNetMultimodal.py:84). */
typedef struct EmbDraws {
    const float*  ffnn_drop[EMB_MAX_FFNN];
    const float*  cnn_drop[EMB_MAX_CNN];
    const float*  post_drop[EMB_MAX_POST];
    const double* embrace_u;
    const float*  modal_rows;       /* [B] per-row coin of EmbraceNetMultimodal.py:181 */
    float         modal_u0;         /* the rand(1) coin of EmbraceNetMultimodal.py:179 */
    int32_t       has_modal_u0;     /* 0: draw it with Philox */
} EmbDraws;

typedef struct EmbOptConfig {
    int32_t kind;                   /* EMB_OPT_* */
    float lr, weight_decay;
    float beta1, beta2, eps;        /* Adam/Nadam: 0.9, 0.999, 1e-8 */
    float alpha;                    /* RMSprop: 0.99 */
    float momentum_decay;           /* Nadam schedule_decay: 4e-3 */
} EmbOptConfig;

/* One row of the parameter table (state_dict key <-> arena offset). */
typedef struct EmbParamInfo {
    char    name[64];               /* e.g. "CNN.CNN_model.0.weight" */
    int64_t offset;                 /* element offset into the params/grads (or buffers) arena */
    int64_t numel;
    int32_t ndim;
    int32_t shape[3];
    int32_t is_buffer;              /* running_mean / running_var live in the buffers arena */
} EmbParamInfo;

/* Per-step results accumulated on the device, one record per step since emb_metrics_reset
 * (replaces loss.item() and AUPRC()/F1_precision_recall() per batch,
 * training_models_multimodal.py:160-162, utils.py:80-94). */
typedef struct EmbStepMetrics {
    float   loss;
    int32_t tp, fp, fn, tn;
} EmbStepMetrics;

typedef struct EmbEngine EmbEngine;

const char* emb_last_error(void);
int emb_abi_version(void);
/* number of sm_100 devices visible (0 = none: compute entry points will fail) */
int emb_device_count(void);

/* Process-wide tuning switches (A/B measurements, verification fall-backs).  Each is read once from the environment
 * variable of the same upper-case name with an EMB_ prefix at first use and can be changed here; none is consulted on a
 * launch path through getenv.  Names: k1_pairs, k1_lookup, no_onehot_wgrad_tc, epi_stats, no_tma_k2, prof_dump, conv_reuse,
 * conv_debug, conv_no_resident, conv_tps1, min_kiters, wgrad_taps, wgrad_ntile, wgrad_fuse_taps, deterministic, k2_wide
 * (csrc/common.cuh: struct Tuning documents each).  Changing one after a CUDA graph was captured does not alter that graph. */
int emb_set_option(const char* name, int32_t value);
int emb_get_option(const char* name, int32_t* value_out);

/* ---- lifetime ------------------------------------------------------------------------------
 * emb_create plans the layer shapes (utils.py:143-153 size_out_convolution chain), the parameter
 * table and the workspace for batches up to max_batch.  No device memory is touched. */
int emb_create(const EmbArchSpec* spec, int32_t max_batch, int32_t precision, EmbEngine** out);
void emb_destroy(EmbEngine* e);

int64_t emb_param_count(const EmbEngine* e);       /* trainable fp32 elements */
int64_t emb_buffer_count(const EmbEngine* e);      /* BatchNorm running stats, fp32 elements */
int64_t emb_workspace_bytes(const EmbEngine* e);
int32_t emb_num_tensors(const EmbEngine* e);
int emb_param_info(const EmbEngine* e, int32_t i, EmbParamInfo* out);
int32_t emb_output_size(const EmbEngine* e, int32_t which); /* 0: FFNN_pre_output_size, 1: CNN_pre_output_size */

/* Bind caller-owned device memory (PyTorch owns it in the Python host).  opt_m/opt_v may be
 * NULL for inference-only use.  Passing all NULL makes the engine cudaMalloc its own. */
int emb_bind(EmbEngine* e, float* params, float* grads, float* buffers, float* opt_m, float* opt_v,
             void* workspace, int64_t workspace_bytes);

/* seed of the counter-based generator used where EmbDraws leaves a draw NULL */
int emb_set_seed(EmbEngine* e, uint64_t seed);
/* data-parallel context: this rank's batch rows are global rows [row_offset, row_offset+B) of a
 * global batch of global_batch rows (Philox counters, loss weights and BatchNorm n use the
 * global view; the collectives themselves are issued by the host, see emb_bn_partial_*). */
int emb_set_shard(EmbEngine* e, int64_t row_offset, int64_t global_batch);
/* number of positive labels in the GLOBAL batch of the next step (get_loss_weights_from_labels, utils.py:121-140,
 * sees the whole batch; labels are known to the host before the step). -1: use the local count. */
int emb_set_global_positives(EmbEngine* e, int64_t n_pos_global);

/* ---- the hot path --------------------------------------------------------------------------*/
/* EmbraceNetMultimodal.forward(..., is_training=True) (EmbraceNetMultimodal.py:159-193), or
 * FFNN.forward / CNN.forward for the single-modality kinds.
 *   x_ffnn  [B, in_features] fp32 row-major (ignored for EMB_KIND_CNN)
 *   bases   [B, 256] uint8 codes 0..3 = a,c,g,t: the argmax of the one-hot [B,4,256] tensor the
 *           reference feeds its first Conv1d (data_pipe/utils.py:268-276)
 *   availabilities [B,2] fp32 or NULL (EmbraceNet.forward's argument)
 *   logits_out [B,2] fp32 */
int emb_forward_train(EmbEngine* e, const float* x_ffnn, const uint8_t* bases, const float* availabilities,
                      int32_t B, const EmbDraws* draws, float* logits_out, void* stream);
/* model.eval() forward (BatchNorm running stats, no dropout, multinomial still sampled).
 * probs_out (optional) [B] = softmax(logits)[:,1], the value the predict loop keeps
 * (EmbraceNetMultimodal_NoTrain.py:210-214, visual.py:290-293). */
int emb_forward_infer(EmbEngine* e, const float* x_ffnn, const uint8_t* bases, const float* availabilities,
                      int32_t B, const EmbDraws* draws, float* logits_out, float* probs_out, void* stream);
/* nn.CrossEntropyLoss(weight=[w_neg,w_pos]) with get_loss_weights_from_labels
 * (training_models_multimodal.py:140-154, utils.py:121-140) on fp32 logits; also the confusion
 * counts behind AUPRC()/F1_precision_recall() (utils.py:80-94).  Appends one EmbStepMetrics record.
 * dlogits_out [B,2] fp32 (may be NULL when only the loss/metrics are wanted). */
int emb_loss_ce_weighted(EmbEngine* e, const float* logits, const int32_t* labels, int32_t B,
                         float* dlogits_out, void* stream);
/* loss.backward(): gradients of every parameter into the grads arena (overwritten, not accumulated)
 * for the most recent emb_forward_train. */
int emb_backward(EmbEngine* e, const float* dlogits, void* stream);
/* optimizer.step() over the whole parameter arena (torch.optim.Adam / RMSprop, timm Nadam, all with
 * coupled L2 weight decay; EMB_OPT_ADAMW decoupled). */
int emb_opt_step(EmbEngine* e, const EmbOptConfig* cfg, void* stream);
/* The host-visible part of the optimizer state (torch.optim keeps `step` per parameter; the fused kernel keeps one):
 * number of steps taken and Nadam's running momentum-schedule product.  The moments live in opt_m / opt_v. */
int emb_opt_state_get(const EmbEngine* e, int64_t* step_out, double* nadam_mu_product_out);
int emb_opt_state_set(EmbEngine* e, int64_t step, double nadam_mu_product);
/* One iteration of the loop body at training_models_multimodal.py:132-162: forward, loss,
 * backward, optimizer step, metrics; nothing returns to the host. */
int emb_train_step(EmbEngine* e, const float* x_ffnn, const uint8_t* bases, const int32_t* labels, int32_t B,
                   const EmbDraws* draws, const EmbOptConfig* cfg, void* stream);
/* Same through HOST buffers: copies the batch host->device on `stream`, runs emb_train_step, and
 * copies this step's EmbStepMetrics back into *metrics_host (synchronises the stream). */
int emb_train_step_host(EmbEngine* e, const float* x_ffnn_host, const uint8_t* bases_host,
                        const int32_t* labels_host, int32_t B, const EmbOptConfig* cfg,
                        EmbStepMetrics* metrics_host, void* stream);

/* Software-pipelined form of emb_train_step_host for training loops (host buffers should be pinned): call i copies batch i
 * host->device on an engine-owned copy stream (overlapping the compute of step i-1), enqueues step i on `stream` and
 * returns the EmbStepMetrics of step i-1 (*have_metrics = 0 on the first call).  emb_train_step_host_flush returns the
 * last step's record.  Replaces the same loop body; the host never stalls the device. */
int emb_train_step_host_pipelined(EmbEngine* e, const float* x_ffnn_host, const uint8_t* bases_host, const int32_t* labels_host,
                                  int32_t B, const EmbOptConfig* cfg, EmbStepMetrics* metrics_prev_host, int32_t* have_metrics,
                                  void* stream);
int emb_train_step_host_flush(EmbEngine* e, EmbStepMetrics* metrics_host, int32_t* have_metrics, void* stream);
/* Predict through HOST buffers (probs_host [B]); synchronises the stream. */
int emb_predict_host(EmbEngine* e, const float* x_ffnn_host, const uint8_t* bases_host,
                     const float* availabilities_host, int32_t B, float* probs_host, void* stream);
/* Software-pipelined scoring loop over host batches (the batch-1 Python loop of Compare_Models_Result.get_model_predictions,
 * visual.py:263-295, as a stream of large batches): call i uploads batch i on an engine-owned copy stream while batch i-1 is
 * still being computed, enqueues its forward and the copy of its scores into `probs_host`, and returns when batch i-1 is complete
 * (*prev_done = 1: the buffer passed to the previous call holds its scores).  emb_predict_host_flush waits for the last batch.
 * The loop shares its two staging slots with emb_train_step_host_pipelined: flush one loop before starting the other. */
int emb_predict_host_pipelined(EmbEngine* e, const float* x_ffnn_host, const uint8_t* bases_host,
                               const float* availabilities_host, int32_t B, float* probs_host, int32_t* prev_done, void* stream);
int emb_predict_host_flush(EmbEngine* e, void* stream);

int emb_metrics_reset(EmbEngine* e, void* stream);
/* copies up to max_records records to host (synchronises); returns the number copied or <0 */
int emb_metrics_read(EmbEngine* e, EmbStepMetrics* out_host, int32_t max_records, void* stream);
/* the modality index map of the last forward: [B, embracement_size] uint8 */
int emb_last_selection(EmbEngine* e, uint8_t* idx_out, int32_t B, void* stream);
/* number of kernel launches issued by this engine since creation (bench.py's gpu_launches) */
int64_t emb_launch_count(const EmbEngine* e);
/* Time every launch of the GEMM kernel class (nn.Linear, docking, Conv1d implicit GEMMs, their dgrad and
 * wgrad) with CUDA event pairs on the launching stream; emb_profile_read synchronises the device and returns
 * the summed kernel time, the algorithmic FLOPs (2*M*N*K per launch) and the launch count since enable. */
int emb_profile_gemm(EmbEngine* e, int32_t enable);
int emb_profile_read(EmbEngine* e, double* ms_out, double* flops_out, int64_t* launches_out);
/* select GEMM back end: 0 = SIMT fp32-accumulate kernels only, 1 = tcgen05/TMEM/TMA where the shape
 * allows (bf16 precision only) */
int emb_set_tensor_core(EmbEngine* e, int32_t on);
/* on: emb_train_step (no replayed draws, no data-parallel hooks) captures the whole step -- forward, loss, backward and
 * the optimizer kernel -- into a CUDA graph the second time a batch size is seen and replays it afterwards (inputs are
 * staged into engine-owned buffers; step-dependent scalars live in device memory).  off: destroys the cached graphs.
 * on = 2 (data parallel, opt-in): the step is captured WITH its hooks -- the all-reduce callback (SyncBN sums) and the phase
 * hook are invoked during capture on the capture stream, so collectives the host enqueues there (NCCL) become graph nodes;
 * the phase hook is called a second time (phase = 2) after the backward pass, where the host completes its gradient
 * all-reduce before the captured optimizer kernel; the global positive count travels through a device slot. */
int emb_set_graph(EmbEngine* e, int32_t on);

/* ---- data-parallel hooks (SyncBN / global loss weights) --------------------------------------
 * When set, BatchNorm statistics are finished by the host's allreduce: the engine writes the local
 * partial sums into a device buffer and calls back between the stats and the finalize kernels. */
typedef int (*EmbAllreduceFn)(void* user, double* device_buf, int64_t count, void* stream);
int emb_set_allreduce(EmbEngine* e, EmbAllreduceFn fn, void* user);
/* Called from emb_backward / emb_train_step with phase = 1 once every gradient OUTSIDE the CNN stack is complete (head, post,
 * docking and FFNN layers), before the CNN backward is enqueued on `stream`: a data-parallel host starts the all-reduce of
 * those slices of the gradient arena there, overlapped with the rest of the backward pass.  Return 0 on success. */
typedef int (*EmbPhaseFn)(void* user, int32_t phase, void* stream);
int emb_set_phase_hook(EmbEngine* e, EmbPhaseFn fn, void* user);

/* ---- data parallelism over NVLink peer memory (csrc/dp_peer.cuh) ------------------------------
 * One process per GPU; rows of the global batch are partitioned (emb_set_shard).  After emb_dp_attach every exchange of
 * the train step is a kernel of this library: SyncBN sums and the global positive count travel as one-shot pushes into
 * every rank's communication block inside the finalize / loss kernels, and emb_opt_step (also inside emb_train_step and
 * its CUDA graph) becomes reduce-scatter(gradients, peer loads) -> optimizer on this rank's slice -> all-gather(parameters,
 * peer stores) in one kernel.  No NCCL call and no host callback remains on the step.  Replaces what
 * torch.nn.parallel.DistributedDataParallel + SyncBatchNorm would add around fit_multimodal's loop body
 * (training_models_multimodal.py:132-162); the reference itself is single-device.
 *   comm[q]   rank q's communication block (emb_dp_comm_bytes() bytes, 128-byte aligned, ZEROED before the first attach)
 *   params[q], grads[q]   rank q's parameter / gradient arenas (entry [rank] = the arenas this engine is bound to)
 * as pointers valid in THIS process (emb_ipc_export / emb_ipc_open map a peer process's memory; engines sharing a process
 * pass their pointers directly).  world <= 8.  Every rank must run the same sequence of training steps. */
int64_t emb_dp_comm_bytes(void);
int emb_dp_attach(EmbEngine* e, int32_t rank, int32_t world, void* const* comm, float* const* params, float* const* grads);
int emb_dp_detach(EmbEngine* e);
/* tests / the split emb_backward + emb_opt_step use: leaves the SUMMED gradient in every rank's gradient arena */
int emb_dp_allreduce_grads(EmbEngine* e, void* stream);
/* CUDA IPC plumbing: handle (64 bytes) + byte offset of dev_ptr inside its cudaMalloc allocation; emb_ipc_open maps it into
 * this process on the current device (peer access enabled lazily) and caches the mapping per handle. */
int emb_ipc_export(const void* dev_ptr, void* handle_out, int64_t* offset_out);
int emb_ipc_open(const void* handle, int64_t offset, void** dev_ptr_out);

/* ---- single-kernel entry points (unit tests and micro-benchmarks) ----------------------------*/
/* K1: Conv1d(4->C1,k) over one-hot input as a gather-sum (CNN_pre.py:39 with in_channels=4).
 * w [C1,4,k] fp32 (reference layout), y [B,256,C1] (fp32 or bf16 by `precision`). */
int emb_k_onehot_conv_fwd(const uint8_t* bases, const float* w, const float* bias, int32_t B, int32_t C1,
                          int32_t k, int32_t precision, void* y, void* stream);
int emb_k_onehot_conv_bwd(const uint8_t* bases, const void* dy, int32_t B, int32_t C1, int32_t k,
                          int32_t precision, float* dw, float* dbias, void* stream);
/* K1 on the tensor cores (what the bf16 train step and inference run): one-hot rows expanded in shared memory, read through a
 * Toeplitz descriptor; the fp32 weights enter as an exact hi/mid/lo bf16 split.  y bf16 [B,256,C1]; stats (nullable) [2][C1] doubles
 * accumulate sum / sum of squares of the rounded outputs (layer-0 BatchNorm statistics).  C1 % 8 == 0, C1 <= 64, odd k <= 15. */
int emb_k_onehot_conv_fwd_tc(const uint8_t* bases, const float* w, const float* bias, int32_t B, int32_t C1, int32_t k, void* y_bf16,
                             double* stats, void* stream);
/* The same weight gradient on the tensor cores (what the bf16 train step runs): the one-hot operand is expanded in shared
 * memory from the base codes and all taps come out of one tcgen05.mma per 16 positions (csrc/onehot_wgrad_tc.cuh).
 * dy bf16 [B,256,C1], C1 % 8 == 0, C1 <= 64, odd k <= 15, dw [C1,4,k] fp32 (zeroed by the call). */
int emb_k_onehot_conv_wgrad_tc(const uint8_t* bases, const void* dy_bf16, int32_t B, int32_t C1, int32_t k, float* dw, void* stream);
/* One GEMM-shaped op of the step on either back end (backend 0 = SIMT fp32-accumulate kernel, 1 = tcgen05/TMEM/TMA
 * kernel); fp32 in / fp32 out, inputs rounded to bf16 as the bf16 precision does.  kind:
 *   0 linear fwd   a[M,K] b[N,K] -> out[M,N]       3 conv fwd   a[B,L,Cin]  b=W[Cout,Cin,taps] -> out[B*L,Cout]
 *   1 linear dgrad a[M,K] b[K,N] -> out[M,N]       4 conv dgrad a[B,L,Cout] b=W[Cout,Cin,taps] -> out[B*L,Cin]
 *   2 linear wgrad a[K,M] b[K,N] -> out[M,N]       5 conv wgrad a[B,L,Cout] b=act[B,L,Cin]     -> out=dW[Cout,Cin,taps]
 * (nn.Linear / nn.Conv1d(stride 1, "same" padding) forward and their autograd formulas; all inner widths multiples
 * of 8).  Synchronises the stream. */
int emb_k_gemm(int32_t kind, int32_t backend, const float* a, const float* b, float* out, int32_t M, int32_t N, int32_t K,
               int32_t B, int32_t L, int32_t Cin, int32_t Cout, int32_t taps, void* stream);

/* the same GEMM launched `reps` times (first = warm-up); *ms_out = average device milliseconds per launch (kernel tuning) */
int emb_k_gemm_time(int32_t kind, int32_t backend, const float* a, const float* b, float* out, int32_t M, int32_t N, int32_t K,
                    int32_t B, int32_t L, int32_t Cin, int32_t Cout, int32_t taps, int32_t reps, float* ms_out, void* stream);

/* test-only hardware probe: UMMA shared-memory descriptors with a start address shifted by whole 128-byte rows
 * (csrc/probe.cuh); decides how multi-tap convolution tiles may be addressed. */
int emb_k_umma_shift_probe(int32_t mode, int32_t shift, int32_t use_base_offset, const float* a, const float* b, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* EMBRACE_B200_H */
