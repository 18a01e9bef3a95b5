/* libaudio2midi_b200 -- C ABI of the B200-native audio-to-midi hot path.
 *
 * Drop-in boundary for the reference's batched model forward
 *     jax.vmap(model.predict, in_axes=(None, 0, None))(state, samples[B,2,80000], rope_freqs)
 * (reference infer.py:40; OutputSequenceGenerator.__call__ / predict, model.py:740-773) and for the
 * Rust `modelutil` post-processing that consumes its output (rust-plugins/src/common.rs:13-144,
 * python.rs:423-447, and the reference's only existing C ABI, cbinds.rs:9-91).
 *
 * Conventions
 *   - plain C types only; `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - every function returning int returns 0 on success and a negative A2M_E* code on failure, in which
 *     case a2m_last_error(handle) describes it; nothing aborts or throws across the boundary;
 *   - the caller owns every buffer it passes; the library owns its weights arena and workspace;
 *   - one handle per device; a handle is not thread-safe, distinct handles are independent;
 *   - *_dev pointers are device memory on the handle's device, *_host pointers are host memory;
 *   - nothing here falls back to the CPU: without an sm_100 device a2m_create fails.
 */
#ifndef A2M_H_
#define A2M_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define A2M_OK 0
#define A2M_EINVAL (-1)    /* bad argument / shape / missing leaf */
#define A2M_ECUDA (-2)     /* CUDA runtime or driver error */
#define A2M_ENODEVICE (-3) /* no sm_100 device */
#define A2M_ESTATE (-4)    /* call order (e.g. forward before load_weights) */
#define A2M_ENCCL (-5)     /* NCCL not loadable / a collective failed */

#define A2M_F32 0          /* element types of the host-path buffers (a2m_submit_host_ex) */
#define A2M_F16 1          /* IEEE binary16 */

#define A2M_WINDOW_SAMPLES 80000 /* audio_to_midi_dataset.py:28,111: 5.0 s x 16 kHz */
#define A2M_FRAMES 250           /* model output frames per window */
#define A2M_VOCAB 90             /* audio_to_midi_dataset.py:26 MIDI_EVENT_VOCCAB_SIZE */
#define A2M_ROPE_DIM 32          /* rope.py:12-22 with dim = attention_size = 64 */
#define A2M_HOST_SLOTS 4         /* batches in flight on the pipelined host path (a2m_submit_host*) */

typedef struct A2mHandle A2mHandle;

/* One leaf of the reference parameter pytree (model.py:673-678 and below), fp32, C-contiguous, in the
 * reference's own layout (eqx.nn.Conv1d weight (out, in/groups, k), bias (out, 1); eqx.nn.Linear weight
 * (out, in); transformer leaves stacked on a leading axis of num_transformer_layers, model.py:646-647).
 * `path` is the dotted pytree key path, e.g. "layers.5.layers.3.point_conv_1.weight" or
 * "transformer.layers.local_attention.attention_block.self_attention.kv_down_proj.weight". */
typedef struct {
  const char* path;
  uint64_t offset_bytes; /* into the blob passed to a2m_load_weights */
  int32_t ndim;
  int64_t shape[4];
} A2mLeafDesc;

/* Creates a handle on CUDA device `device`.  Replaces OutputSequenceGenerator.__init__ (model.py:680-738)
 * for the default model_config (model.py:20-34); weights arrive through a2m_load_weights. */
int a2m_create(int device, A2mHandle** out);
/* The same with the architecture spelled out, i.e. OutputSequenceGenerator(conf, key) (model.py:680-738, model_config
 * model.py:20-34).  The kernels are specialised for the reference's default model_config: any other value of the
 * architecture fields is refused with A2M_EINVAL.  use_graph / use_pdl: 0 / 1, or -1 for the default. */
typedef struct {
  int32_t device;
  int32_t num_stages;                    /* len(dims) = 7 */
  int32_t dims[8];                       /* 4, 8, 16, 32, 64, 128, 256 */
  int32_t depths[8];                     /* 3, 3, 3, 3, 3, 21, 3 */
  int32_t cnn_hidden_expansion_x2;       /* cnn_hidden_expansion * 2 = 4 */
  int32_t num_transformer_layers;        /* 8 (each = local + global) */
  int32_t num_transformer_heads;         /* 4 */
  int32_t attention_size;                /* 64 */
  int32_t compressed_attention_kv_size;  /* 64 */
  int32_t transformer_intermediate;      /* dims[-1] * transformer_hidden_expansion = 512 */
  int32_t use_graph;
  int32_t use_pdl;
} A2mConfig;
int a2m_create_ex(const A2mConfig* config, A2mHandle** out);
void a2m_destroy(A2mHandle* h);
const char* a2m_last_error(const A2mHandle* h);

/* Loads (and re-packs for the kernels) a full parameter pytree from a host blob.  Replaces the
 * checkpoint-restore -> pytree hand-off of infer.py:172-236 / change_fp_precision (infer.py:27-32).
 * May be called again to swap weights. */
int a2m_load_weights(A2mHandle* h, const void* blob_host, size_t blob_bytes, const A2mLeafDesc* table,
                     int32_t n_leaves);

/* Bytes of device scratch a forward of `batch` windows needs.  The handle allocates and caches this
 * itself when a2m_forward is called with workspace_dev == NULL. */
size_t a2m_workspace_bytes(const A2mHandle* h, int32_t batch, int32_t training);

/* Batched forward; everything is enqueued on `stream`, no host synchronisation.
 *   audio_dev   [batch, 2, 80000] fp32           (samples of infer.py:40)
 *   rope_cos_dev / rope_sin_dev [rope_max_pos, 32] fp32, rope_max_pos >= 250   (RopeFreqs, rope.py:5-22)
 *   logits_dev, probs_dev [batch, 250, 90] fp32   (return value of model.predict, model.py:771-773)
 *   workspace_dev: NULL, or >= a2m_workspace_bytes(h, batch, 0) bytes, 1024-byte aligned, zero-initialised
 *                  once by the caller before its first use.
 * Calls on DIFFERENT streams may be in flight at the same time if each uses its own workspace_dev (one launch plan and CUDA
 * graph is kept per (batch, workspace)); that is how consecutive independent batches are overlapped (two lanes: 1.32 ms per
 * 64-window batch against 1.55 ms back to back).  workspace_dev = NULL is the handle's single default workspace. */
int a2m_forward(A2mHandle* h, const float* audio_dev, int32_t batch, const float* rope_cos_dev,
                const float* rope_sin_dev, int32_t rope_max_pos, float* logits_dev, float* probs_dev,
                void* workspace_dev, size_t workspace_bytes, void* stream);

/* Same call with HOST buffers: stages through pinned memory, copies in, runs, copies out and waits.
 * This is the end-to-end path a Python caller without device arrays uses. */
int a2m_forward_host(A2mHandle* h, const float* audio_host, int32_t batch, const float* rope_cos_host,
                     const float* rope_sin_host, int32_t rope_max_pos, float* logits_host, float* probs_host);

/* Pipelined host path: A2M_HOST_SLOTS slots (0 .. 3).  a2m_submit_host enqueues H2D copy -> forward -> D2H copy for one
 * batch and returns; a2m_collect_host waits for that slot's results.  Each slot has its own copy stream and device buffers;
 * the forwards run on two compute lanes (stream + workspace; slot & 1), so with four batches in flight two are computing --
 * overlapping each other -- while one uploads and one downloads.  Buffers allocated with a2m_host_alloc (page-locked) are
 * copied from / to directly; pageable buffers go through an internal staging copy.  The caller's buffers must
 * stay valid and untouched until the slot has been collected. */
int a2m_submit_host(A2mHandle* h, int32_t slot, const float* audio_host, int32_t batch, const float* rope_cos_host,
                    const float* rope_sin_host, int32_t rope_max_pos, float* logits_host, float* probs_host);
int a2m_collect_host(A2mHandle* h, int32_t slot);
/* The same with typed buffers.  audio_dtype A2M_F16: the windows arrive as IEEE binary16 -- lossless for audio that went
 * through load_full_audio, which rounds every sample to f16 (python.rs:235-264) -- and are widened by the stem kernel.
 * logits_host may be NULL (infer.py:41 keeps only the probabilities); out_dtype A2M_F16 returns the probabilities as
 * binary16.  f16 in, probabilities only, f16 out moves 22.9 MB per 64 windows instead of 52.5 MB. */
int a2m_submit_host_ex(A2mHandle* h, int32_t slot, const void* audio_host, int32_t audio_dtype, int32_t batch,
                       const float* rope_cos_host, const float* rope_sin_host, int32_t rope_max_pos, float* logits_host,
                       void* probs_host, int32_t out_dtype);
/* Page-locked host memory for the calls above (cudaMallocHost / cudaFreeHost). */
void* a2m_host_alloc(size_t bytes);
void a2m_host_free(void* p);

/* ---- the step before the path, on the device ------------------------------------------------ */
/* Windows a clip of n_samples per channel is cut into: ceil((N - overlap) / (80000 - overlap)), overlap in SECONDS at
 * 16 kHz (load_and_slice_full_audio, audio_to_midi_dataset.py:277-294). */
int64_t a2m_window_count(int64_t n_samples, double overlap_s);
/* clip_dev [2, n_samples] fp32 (raw decoded audio) -> windows_dev [a2m_window_count, 2, 80000] fp32: loudness
 * normalisation of the whole clip exactly as load_full_audio does (python.rs:235-264: untouched if the peak is <= 0.05,
 * else x / sqrt(mean square over both channels) in f64; rounded to f16) fused with the window slicing; the last window
 * is zero padded.  The result is the `samples` argument of a2m_forward. */
int a2m_prepare_windows(A2mHandle* h, const float* clip_dev, int64_t n_samples, double overlap_s, float* windows_dev,
                        int64_t max_windows, void* stream);

/* Per-window validation loss (testset_loss_function, train.py:99-102): losses_dev[b] = sum_{t,c} BCEWithLogits(logits, labels)
 * over the [250, 90] frame grid of window b.  logits_dev / labels_dev [batch, 250, 90] fp32. */
int a2m_window_losses(A2mHandle* h, const float* logits_dev, const float* labels_dev, int32_t batch, float* losses_dev, void* stream);

/* detailed_event_loss (infer.py:94-158) for a batch of windows in one launch: per window, eventize the
 * probabilities (extract_events, common.rs:47-144), rasterise the events (to_frame_events, python.rs:423-447) and compare with
 * the annotation.  probs_dev / expected_dev [batch, frames, 90] fp32; metrics_dev [batch, 5] = full_diff, phantom_notes_diff,
 * missed_notes_diff, notes_hit, hit_rate.  pred_frames_dev (optional) [batch, frames, 90] receives the rasterised prediction,
 * n_events_dev (optional) [batch, 90] the number of events per key.  One thread per (window, key). */
int a2m_event_metrics(A2mHandle* h, const float* probs_dev, const float* expected_dev, int32_t batch, int32_t frames,
                      float* metrics_dev, float* pred_frames_dev, int32_t* n_events_dev, void* stream);

/* Tensor-core operand format of THIS build of the library: "bf16" (libaudio2midi_b200.so: training and inference) or "f16"
 * (libaudio2midi_b200_f16.so, the inference variant: IEEE binary16 operands, 8x smaller operand rounding at the same tcgen05
 * rate; a2m_train_init is refused).  Accumulation, residual stream and LayerNorm / softmax statistics are fp32 in both.
 * This is the choice change_fp_precision makes in the reference (infer.py:27-32, 234). */
const char* a2m_operand_format(void);
/* test hook: host-side rounding of fp32 values to the operand format (what a2m_load_weights applies to the weights) */
int a2m_debug_round_operand(const float* in_host, uint16_t* out_host, int64_t n);

/* stitch_probs (common.rs:13-45) and extract_events (common.rs:47-144) on the device ("extract_events is
 * independent per key"), so that the probabilities of a long clip never travel to the host -- only its event list does.
 * a2m_stitch_probs_dev: probs_dev [windows, frames, cats] -> out_dev [out_frames, cats], bit-identical to a2m_stitch_probs;
 * returns out_frames (out_dev = NULL queries it), or A2M_EINVAL (also when overlap >= about half a window, where consecutive
 * cross-fades would chain: use the host function).  a2m_extract_events_dev: probs_dev [frames < 2^24, notes <= 96] ->
 * events_dev [cap] 64-bit words  attack << 32 | key << 24 | duration  (one per event, in no particular order: sorting the
 * words ascending IS the (attack, key, duration) order of common.rs:142; velocity is the constant 7) and *count_dev = the
 * number of events; a count above cap means the words beyond cap were dropped: call again with a larger buffer.  Two
 * launches: the comparisons of the state machine for every (frame, key) in parallel into three bit masks per key, then the
 * machine itself, one thread per key, jumping from set bit to set bit (its cost follows the events, not the frames). */
int64_t a2m_stitch_probs_dev(A2mHandle* h, const float* probs_dev, int64_t windows, int64_t frames, int64_t cats, double overlap,
                             double duration_per_frame, float* out_dev, int64_t out_capacity_frames, void* stream);
int a2m_extract_events_dev(A2mHandle* h, const float* probs_dev, int64_t frames, int64_t notes, uint64_t* events_dev, int64_t cap,
                           int32_t* count_dev, void* stream);

/* Number of kernels of this library launched by the last a2m_forward on this handle. */
int32_t a2m_last_launch_count(const A2mHandle* h);
/* Per-launch profile of the forward plan for `batch` windows: every step of the plan (all launches between
 * the stem and the decoder GEMM) is run `repeats` times back to back between two CUDA events on the
 * handle's own stream.  Returns the number of steps written (or, with out == NULL, the number of steps).
 * flops / bytes are the ALGORITHMIC counts of the launch (2 x MACs; operands in + results out). The
 * residual stream in the workspace is left in an arbitrary (finite) state. */
typedef struct {
  char kernel[32];
  float ms;
  double flops;
  double bytes;
} A2mStepProfile;
int32_t a2m_profile_steps(A2mHandle* h, int32_t batch, int32_t repeats, int32_t max_steps, A2mStepProfile* out);
/* Use a CUDA graph for the steady-state forward (default 1). */
int a2m_set_use_graph(A2mHandle* h, int32_t enable);

/* ---- test hooks (used by tests/ only) ------------------------------------------------------------ */
/* CTA-pair (tcgen05 cta_group::2, cluster of two CTAs) GEMM used to prove the recipe that the fused transformer kernels
 * will adopt: out32[M, N] = A[M, K] (bf16) x W[N, K]^T (bf16); M % 256 == 0, N in {128, 256}, K % 64 == 0. */
int a2m_debug_gemm_pair(A2mHandle* h, int32_t M, int32_t N, int32_t K, const void* A, const void* W, float* out32, void* stream);
/* Phase timeline (clock64 stamps of CTA 0) of the last ffn_fused_kernel launch; only in builds compiled with
 * -DA2M_FFN_TIMING (tools/ffn_timeline.py), returns -1 in the product build. */
int a2m_debug_read_timing(long long* out, int32_t n);
/* Runs the forward up to and including the step labelled `label` ("stage0".."stage6", "cnn_out",
 * "tl<i>_local", "tl<i>_global") and copies the fp32 residual stream at that point to out_dev:
 * stage taps are [batch * L_stage, C_stage]; the others are [batch * 256, 256] (rows 250..255 padding). */
int a2m_debug_forward_tap(A2mHandle* h, const float* audio_dev, int32_t batch, const float* rope_cos_dev,
                          const float* rope_sin_dev, int32_t rope_max_pos, const char* label, float* out_dev,
                          size_t out_elems, void* stream);
/* One tcgen05 GEMM: D[M,N] = A[M,K] (bf16, row stride lda) x W[N,K]^T (bf16), generic epilogue
 * (flags: 1 bias, 2 gelu, 4 gamma, 8 residual, 16 fp32 out, 32 bf16 out).  block_n in {64,128,256}. */
int a2m_debug_gemm(A2mHandle* h, int32_t block_n, int32_t M, int32_t N, int32_t K, const void* A_bf16_dev,
                   int32_t lda, const void* W_bf16_dev, uint32_t flags, const float* bias_dev,
                   const float* gamma_dev, const float* resid_dev, float* out32_dev, void* out16_dev, void* stream);

/* ---- training path (train.py:39-62 loss + value_and_grad; train.py:259-332 step) -------------------------- */
/* Master parameters (fp32, in the layout of `blob_host` / the leaf table, = the reference pytree leaves), AdamW moments
 * and the activation tape live in the handle.  Replaces the model / opt_state hand-off of train.py:246-264.  Also
 * loads the weights for a2m_forward. */
int a2m_train_init(A2mHandle* h, const void* blob_host, size_t blob_bytes, const A2mLeafDesc* table, int32_t n_leaves);
int64_t a2m_param_count(const A2mHandle* h);                          /* floats in the blob */
int a2m_get_params(A2mHandle* h, float* out_dev, void* stream);       /* current master parameters, blob layout */
/* Overwrites the master parameters and re-derives the kernels' packed images: together with a2m_get/set_opt_state this is
 * the snapshot / roll-back of the reference's non-finite recovery (train.py:334-382). */
int a2m_set_params(A2mHandle* h, const float* params_dev, void* stream);
/* AdamW first / second moments (optax ScaleByAdamState mu, nu), blob layout; either pointer may be NULL. */
int a2m_get_opt_state(A2mHandle* h, float* m_dev, float* v_dev, void* stream);
int a2m_set_opt_state(A2mHandle* h, const float* m_dev, const float* v_dev, void* stream);
/* One learning-rate multiplier per leaf (layer-wise decay of train.py:646-726); NULL resets to 1. */
int a2m_set_lr_multipliers(A2mHandle* h, const float* per_leaf_host, int32_t n_leaves);
/* Dropout of the following a2m_forward_train / a2m_backward pairs: `rate` = transformer_dropout_rate (model.py:30) applied
 * to the attention weights (model.py:254-255, per window for the local layers) and to the FeedForwardBlock output
 * (model.py:237), kept values scaled by 1 / (1 - rate); 0 disables it (the state after a2m_train_init).  `seed` plays
 * the role of the PRNG key (train.py:53): masks are a counter-based hash of (seed, site, element), regenerated in the
 * backward; the random stream is NOT jax's threefry. */
int a2m_set_dropout(A2mHandle* h, float rate, uint64_t seed);
/* Forward that records the tape; logits_dev / probs_dev [batch, 250, 90] may be NULL.  audio_dev must stay valid until
 * a2m_backward has run. */
int a2m_forward_train(A2mHandle* h, const float* audio_dev, int32_t batch, const float* rope_cos_dev,
                      const float* rope_sin_dev, int32_t rope_max_pos, float* logits_dev, float* probs_dev, void* stream);
/* grads_dev [param_count] += d/dparams of  mean_b( sum_{t,c} BCEWithLogits(z, y) * scale )  (compute_loss, train.py:50-62);
 * loss_dev[0] += that value.  labels_dev [batch, 250, 90] fp32.  Gradient accumulation over minibatches
 * (train.py:283-293) = repeated forward_train / backward calls on the same grads_dev. */
int a2m_backward(A2mHandle* h, const float* labels_dev, float scale, float* grads_dev, float* loss_dev, void* stream);
/* The same backward for an ARBITRARY cotangent of the logits (dlogits_dev [batch, 250, 90] fp32): grads_dev += J^T dlogits.
 * This is the `bwd` of a jax.custom_vjp around a2m_forward_train: the loss stays the caller's. */
int a2m_backward_dlogits(A2mHandle* h, const float* dlogits_dev, float* grads_dev, void* stream);
/* Gradient buckets for an all-reduce that overlaps the backward (train.py:238-244 shards the batch; the exchange jit
 * inserts there is one collective here).  Bucket 0 = final norm + transformer + decoder (the tail of the leaf order),
 * final once the transformer backward has run; bucket 1 = the CNN, final when a2m_backward's work completes.
 * a2m_grad_bucket_range: elements [*lo, *hi) of the gradient blob.  a2m_stream_wait_grad_bucket: `stream` waits until
 * that range is final for the most recent a2m_backward (call after it returned). */
int32_t a2m_grad_bucket_count(const A2mHandle* h);
int a2m_grad_bucket_range(const A2mHandle* h, int32_t bucket, size_t* lo, size_t* hi);
int a2m_stream_wait_grad_bucket(A2mHandle* h, int32_t bucket, void* stream);
/* Data-parallel gradient exchange (train.py:238-244 shards the batch over devices): mean over the ranks of
 * `nccl_comm` (an ncclComm_t; NULL = the handle's own, a2m_comm_init) of the gradient blob and loss of the most recent
 * a2m_backward.  Two buckets: bucket 0 is reduced on a communication stream owned by the handle as soon as the backward has
 * produced it, under the CNN backward still running on `stream`; bucket 1 and the loss follow on `stream`, which then waits
 * for the communication stream.  Collective: every rank calls it after the same a2m_backward.  libnccl is bound at run time
 * (dlopen of libnccl.so.2 -- the copy already in the process if there is one), it is not a link-time dependency. */
int a2m_allreduce_grads(A2mHandle* h, void* nccl_comm, void* stream);
/* Rendezvous helpers for callers that have no communicator of their own: rank 0 calls a2m_comm_unique_id (ncclGetUniqueId,
 * 128 bytes), ships the id to the other ranks by any side channel, every rank calls a2m_comm_init (ncclCommInitRank). */
int a2m_comm_unique_id(void* id128_out);
int a2m_comm_init(A2mHandle* h, const void* id128, int32_t nranks, int32_t rank);
void* a2m_comm_get(const A2mHandle* h);   /* the ncclComm_t, or NULL */
int a2m_comm_destroy(A2mHandle* h);
/* optax.adamw then clip_by_global_norm(clip_norm) on the updates, applied to the master parameters; gradients are
 * divided by grad_divisor first (train.py:314).  step counts from 1.  stats_dev (optional, 2 floats): squared norm of
 * the unclipped update, number of non-finite gradient entries (train.py:320-322 grads_valid == 0).  When that number is
 * not zero the step is a NO-OP on the device: neither the parameters nor the moments are touched (the reference rolls back
 * to a snapshot instead, train.py:369-377).  The data-parallel gradient all-reduce happens between a2m_backward and this
 * call (a2m_allreduce_grads, or the caller's own collective on grads_dev). */
int a2m_adamw_step(A2mHandle* h, const float* grads_dev, float lr, float b1, float b2, float eps, float weight_decay,
                   float grad_divisor, float clip_norm, int32_t step, float* stats_dev, void* stream);
/* Per-launch profile of the training plans (which = 0 forward-with-tape, 1 backward); contract of a2m_profile_steps. */
int32_t a2m_profile_train_steps(A2mHandle* h, int32_t which, int32_t repeats, int32_t max_steps, A2mStepProfile* out);
int32_t a2m_train_launch_count(const A2mHandle* h);   /* kernels of the last forward_train + backward */
/* test hook: dW[n_out, k_out] += dY[tokens, n_out]^T X[tokens, k_out] (bf16 operands) on the tcgen05 wgrad kernel */
int a2m_debug_wgrad(A2mHandle* h, int32_t tokens, int32_t n_out, int32_t k_out, const void* dY_bf16_dev, int32_t ldy,
                    const void* X_bf16_dev, int32_t ldx, float* dW_dev, void* stream);

/* ---- modelutil: post-processing of the probabilities (host code, like the reference's Rust) ------- */
/* stitch_probs (common.rs:13-45).  probs [windows, frames, cats] fp32 -> out [out_frames, cats];
 * returns out_frames (= windows*frames - trunc(overlap/dpf)*(windows-1)); out may be NULL to query. */
int64_t a2m_stitch_probs(const float* probs, int64_t windows, int64_t frames, int64_t cats, double overlap,
                         double duration_per_frame, float* out);

typedef struct { /* cbinds.rs:9-15 */
  uint64_t attack_time;
  uint8_t note;
  uint64_t duration;
  uint8_t velocity;
} MidiEvent;
typedef struct { /* cbinds.rs:17-22 */
  MidiEvent* ptr;
  size_t length;
  size_t _capacity;
} MidiEventList;
typedef struct { /* cbinds.rs:24-29, N = 3; strides in ELEMENTS, data is IEEE binary16 */
  uint64_t strides[3];
  uint64_t dims[3];
  const uint8_t* data;
} MLMultiArrayWrapper3;

/* extract_events (common.rs:47-144) over probs [frames, notes] fp32; caller frees with free_midi_events. */
MidiEventList* a2m_extract_events(const float* probs, int64_t frames, int64_t notes);
/* The reference's iOS entry points (cbinds.rs:51-91): f16 strided windows -> stitch -> extract. */
MidiEventList* extract_midi_events(MLMultiArrayWrapper3 data, double overlap, double duration_per_frame);
void free_midi_events(MidiEventList* ptr);
/* convert_to_frame_events (python.rs:423-447) as called by to_frame_events (python.rs:980-1005):
 * events -> out [frame_count, 90] fp32 (zero-filled first). */
int a2m_to_frame_events(const MidiEvent* events, int64_t n_events, int64_t frame_count, float* out);

#ifdef __cplusplus
}
#endif
#endif /* A2M_H_ */
