/* sedt_b200 — C ABI of the B200-native SEDT hot path.
 *
 * The reference (Anaesthesiaye/sound_event_detection_transformer) is pure
 * Python and has no FFI; its seam is the Python module API of package `sedt`
 * (SURVEY.md section 8b).  This header is the layer directly beneath that
 * seam: the Python mirror in sound_event_detection_transformer_b200/sedt/
 * keeps the reference's classes and binds these entry points with ctypes
 * (INTEGRATION.md shows the stub).  Each entry point names the reference
 * code it replaces (paths relative to the reference root).
 *
 * Conventions: every pointer is a raw CUDA device pointer unless marked
 * [host]; all memory is owned by the caller (PyTorch); `stream` is a
 * cudaStream_t passed as void*; functions return 0 on success or a negative
 * sedt_status and never throw, exit or synchronise the device unless stated;
 * sedt_last_error() returns the message for the calling thread.
 */
#ifndef SEDT_B200_H
#define SEDT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SEDT_ABI_VERSION 1
#if defined(__GNUC__)
#define SEDT_API __attribute__((visibility("default")))
#else
#define SEDT_API
#endif

enum sedt_status {
    SEDT_STATUS_OK = 0,
    SEDT_STATUS_INVALID = -1,
    SEDT_STATUS_CUDA = -2,
    SEDT_STATUS_WORKSPACE = -3,
    SEDT_STATUS_NUMERIC = -4,      /* NaN / -inf cost entries: scipy raises ValueError here */
    SEDT_STATUS_INFEASIBLE = -5,
    SEDT_STATUS_UNSUPPORTED = -6
};

enum sedt_dtype { SEDT_F32 = 0, SEDT_BF16 = 1 };

/* Model hyper-parameters `build_model(args)` reads (sedt/__init__.py:8-63,
 * sedt/transformer.py:409-420, sedt/backbone.py:135-141). */
typedef struct sedt_config {
    int32_t enc_layers, dec_layers;
    int32_t num_queries;          /* event queries, without the audio query */
    int32_t num_classes;
    int32_t hidden_dim, nheads, dim_feedforward;
    int32_t dec_at, pre_norm, dilation;
    int32_t self_sup, feature_recon, num_patches;
    int32_t aux_loss;
    int32_t precision;            /* 0: fp32 CUDA-core tier; 1: bf16 operands, fp32 accumulate, tcgen05 */
    int32_t use_tensor_cores;     /* precision 1 only; 0 routes every GEMM to the CUDA-core kernel */
} sedt_config;

typedef struct sedt_model sedt_model;

/* fp32 outputs of one forward; null members are skipped. */
typedef struct sedt_outputs {
    float* hs;            /* [D, B, Qall, 256]  decoder states (transformer.py:140-150); required */
    float* logits;        /* [D, B, Q, C+1]     class_embed on every decoder layer (sedt.py:90); required */
    float* boxes;         /* [D, B, Q, 2]       sigmoid(bbox_embed) = (center, width) (sedt.py:91); required */
    float* at;            /* [B, C]             sigmoid(weak_class_embed(hs[-1,:,0])) (sedt.py:92); dec_at only */
    float* memory;        /* [B, S, 256]        encoder output, optional */
    float* pred_feature;  /* [D, B, Q, 2048]    feature_align(hs) (spsedt.py:80), SP-SEDT only, optional */
    float* gt_feature;    /* [B*P, 2048]        avgpool(backbone(patches)) (spsedt.py:50), SP-SEDT only, optional */
    float* feat;          /* [B, H, W, 2048]    layer4 output, fp32 tier only, optional (parity tests) */
} sedt_outputs;

SEDT_API const char* sedt_last_error(void);
SEDT_API int sedt_abi_version(void);
/* number of kernels this library has launched since load (bench.py reports the delta) */
SEDT_API unsigned long long sedt_launch_count(void);
/* eager launches per kernel kind (the size-dependent choices of the dispatcher: "conv_tc2", "conv_tc3_2sm", "conv_tc4_ws",
 * "ffn_fused", "enc_attn_fused", "bottleneck_fused", ...): lets a parity test assert which kernels produced the output it checked */
SEDT_API int sedt_kernel_kinds(void);
SEDT_API const char* sedt_kernel_kind_name(int kind);
SEDT_API unsigned long long sedt_kernel_kind_count(int kind);

/* Per-kernel-class device timing for roofline evidence (bench.py).  While enabled every launch is
 * bracketed by CUDA events on its stream; sedt_profile_read synchronises the device and returns
 * the summed milliseconds and launch counts per class since the last read, indexed by
 * sedt_kernel_class ([host] arrays of SEDT_KC_COUNT entries). */
enum sedt_kernel_class { SEDT_KC_GEMM_TCGEN05 = 0, SEDT_KC_GEMM_CUDA_CORE, SEDT_KC_STEM, SEDT_KC_ATTENTION, SEDT_KC_NORM,
                         SEDT_KC_MATCHER, SEDT_KC_OTHER, SEDT_KC_COUNT };
SEDT_API int sedt_profile_enable(int on);
SEDT_API int sedt_profile_read(double* ms_per_class, long long* launches_per_class);

/* ---- model lifecycle: replaces SEDT.__init__/SPSEDT.__init__ state (sedt/sedt.py:20-61) ---- */
SEDT_API int sedt_model_create(const sedt_config* cfg, sedt_model** out);
SEDT_API void sedt_model_destroy(sedt_model* m);
/* The weight table: slot i is the reference state_dict entry sedt_model_weight_name(m, i)
 * (fp32, contiguous, reference shape).  SURVEY.md section 8b lists the names. */
SEDT_API int sedt_model_num_weights(const sedt_model* m);
SEDT_API const char* sedt_model_weight_name(const sedt_model* m, int i);
SEDT_API int64_t sedt_model_weight_numel(const sedt_model* m, int i);
SEDT_API int64_t sedt_model_packed_bytes(const sedt_model* m);
/* Snapshot the weights into `packed` (256-byte aligned): OIHW -> O(HW)I repack in the tier's
 * dtype, FrozenBatchNorm2d fold (sedt/backbone.py:43-53), conv0-into-conv1 fold.  Call again
 * whenever a parameter changes.  `weights` is a [host] array of device pointers. */
SEDT_API int sedt_model_pack(sedt_model* m, const void* const* weights, void* packed, int64_t packed_bytes, void* stream);

/* Spatial size of the layer4 feature map for a [T, F] clip (backbone.py + resnet strides). */
SEDT_API int sedt_feature_shape(int T, int F, int dilation, int* H, int* W);
/* Bytes of scratch sedt_forward needs for this shape (P = PT = 0 unless SP-SEDT). */
SEDT_API int64_t sedt_workspace_bytes(sedt_model* m, int B, int T, int F, int P, int PT);

/* The forward hot path: replaces SEDT.forward (sedt/sedt.py:64-123) and, with patches,
 * SPSEDT.forward's eval branch (sedt/spsedt.py:34-91).
 *   x       [B, 1, T, F] fp32 log-mel clips (already zero-padded to the batch maximum)
 *   mask    [B, T, F] uint8, 1 = padding (utilities/utils.py:470-492), or null if nothing is padded
 *   patches [B, P, 1, PT, F] fp32 or null */
SEDT_API int sedt_forward(sedt_model* m, const float* x, const uint8_t* mask, int B, int T, int F,
                 const float* patches, int P, int PT, void* workspace, int64_t workspace_bytes,
                 const sedt_outputs* out, void* stream);

/* ---- Training step (bf16 tier, pre-norm supervised model, dropout = 0) ----------------------------------
 * sedt_forward_train is sedt_forward in train mode (sedt/sedt.py:64-123 under model.train()): same outputs, and
 * every activation the backward pass needs is kept in `tape` (sedt_train_tape_bytes, caller-owned).
 * sedt_backward replaces loss.backward() through the model (engine.py:70-74): from the gradients of
 * pred_logits [D,B,Q,C+1], pred_boxes [D,B,Q,2] (all decoder layers, aux_outputs included) and at [B,C]
 * (any may be NULL = zero) it writes the gradient of every trainable state_dict entry, fp32, reference
 * layout, at grads + sedt_grad_offset(slot) (one flat buffer of sedt_grad_numel floats = the data-parallel
 * all-reduce bucket; frozen entries - conv1, layer1, FrozenBN buffers - stay zero).  `weights` is the
 * array given to sedt_model_pack; train_backbone = 0 stops at input_proj (lr_backbone = 0,
 * sedt/backbone.py:135-141).  dropout = args.dropout (transformer.py: attention weights, after out_proj, FFN hidden,
 * after linear2): masks are counter-based (Philox keyed by `seed`, the dropout site and a per-tape step counter that
 * sedt_forward_train advances), so the backward pass regenerates exactly the forward's masks; pass the same
 * dropout to both calls. */
/* Data-parallel training (train_spsedt.py:157-158 wraps the model in DistributedDataParallel): sedt_backward records
 * `cuda_event` (a cudaEvent_t, or NULL to stop) on its stream as soon as every gradient outside the backbone (transformer, heads,
 * input_proj, query_embed: slots below the first "backbone." entry and from query_embed on) is final, so that the caller can
 * all-reduce that bucket on a side stream while the backbone backward is still running. */
SEDT_API int sedt_model_set_bucket_event(sedt_model* m, void* cuda_event);
SEDT_API int64_t sedt_train_tape_bytes(sedt_model* m, int B, int T, int F, int has_mask);
SEDT_API int64_t sedt_backward_workspace_bytes(sedt_model* m, int B, int T, int F);
SEDT_API int64_t sedt_grad_numel(const sedt_model* m);
SEDT_API int64_t sedt_grad_offset(const sedt_model* m, int slot);
/* SP-SEDT pretraining step (sedt/spsedt.py:34-91 in train mode; the backbone is frozen, train_spsedt.py:50): the training
 * branch of SPSEDT.forward -- query positions 2 * query_embed + query_keep * patch2query(avgpool(backbone(patch))) with
 * query_keep [B, Q] uint8 = (torch.rand(Q, bs, 1) > mask_ratio) drawn by the caller (spsedt.py:65), block-diagonal decoder mask,
 * dropout -- and its backward: gradients of input_proj, the transformer, the heads (class_embed, bbox_embed, feature_align),
 * query_embed and patch2query from d(pred_logits), d(pred_boxes), d(pred_feature [D,B,Q,2048]).  P must equal num_patches. */
SEDT_API int64_t sedt_train_tape_bytes_sp(sedt_model* m, int B, int T, int F, int has_mask, int P, int PT);
SEDT_API int64_t sedt_backward_workspace_bytes_sp(sedt_model* m, int B, int T, int F, int P, int PT);
SEDT_API int sedt_forward_train_sp(sedt_model* m, const float* x, const uint8_t* mask, int B, int T, int F, const float* patches, int P,
                                   int PT, const uint8_t* query_keep, void* tape, int64_t tape_bytes, const sedt_outputs* out,
                                   float dropout, uint64_t seed, void* stream);
SEDT_API int sedt_backward_sp(sedt_model* m, const void* const* weights, const float* x, const uint8_t* mask, int B, int T, int F, int P,
                              int PT, void* tape, int64_t tape_bytes, void* workspace, int64_t workspace_bytes, const float* d_logits,
                              const float* d_boxes, const float* d_pred_feature, float* grads, float dropout, void* stream);
SEDT_API int sedt_forward_train(sedt_model* m, const float* x, const uint8_t* mask, int B, int T, int F, void* tape,
                                int64_t tape_bytes, const sedt_outputs* out, float dropout, uint64_t seed, void* stream);
SEDT_API int sedt_backward(sedt_model* m, const void* const* weights, const float* x, const uint8_t* mask, int B, int T, int F,
                           void* tape, int64_t tape_bytes, void* workspace, int64_t workspace_bytes, const float* d_logits,
                           const float* d_boxes, const float* d_at, float* grads, int train_backbone, float dropout,
                           void* stream);

/* ---- HungarianMatcher.forward default path (sedt/matcher.py:41-97; utilities/box_ops.py:9-56)
 *   logits [B, Q, C1] fp32, boxes [B, Q, 2] fp32 (center, width)
 *   tgt_labels [sumK] int64, tgt_boxes [sumK, 2] fp32, offsets [B+1] int32 (clip b owns targets
 *   offsets[b] .. offsets[b+1]); kmax >= max_b K_b.
 *   rows, cols [B, Q] int64 (first counts[b] entries valid, rows ascending; rest -1), counts [B] int32,
 *   status [1] int32 (must be zeroed by the caller; set to a negative sedt_status on NaN/-inf costs).
 *   cost_out: optional [B, Q, ld_cost] fp32 copy of the cost blocks. */
SEDT_API int sedt_matcher(const float* logits, const float* boxes, const int64_t* tgt_labels, const float* tgt_boxes,
                 const int32_t* offsets, int B, int Q, int C1, int kmax,
                 float cost_class, float cost_bbox, float cost_giou,
                 float* cost_out, int ld_cost, int64_t* rows, int64_t* cols, int32_t* counts, int32_t* status,
                 void* stream);
/* The matcher's other branches (sedt/matcher.py:77-82, 99-106): fl != 0 replaces the softmax class cost by the focal
 * cost on sigmoid probabilities (alpha_fl / gamma_fl = config.py:71-72); lmin / largmin [B, Q] (optional, together) receive
 * each query's smallest location cost cost_bbox * L1 + cost_giou * (-GIoU) and the target attaining it (first minimum; -1
 * when the clip has no target) -- the quantities HungarianMatcher's fine_tune relaxation thresholds with epsilon. */
SEDT_API int sedt_matcher_ex(const float* logits, const float* boxes, const int64_t* tgt_labels, const float* tgt_boxes,
                    const int32_t* offsets, int B, int Q, int C1, int kmax, float cost_class, float cost_bbox, float cost_giou,
                    int fl, float alpha_fl, float gamma_fl, float* cost_out, int ld_cost, int64_t* rows, int64_t* cols,
                    int32_t* counts, int32_t* status, float* lmin, int64_t* largmin, void* stream);
/* Only the per-clip assignment (scipy.optimize.linear_sum_assignment, sedt/matcher.py:95) on caller
 * supplied fp32 cost blocks cost[B, Q, ld_cost]; clip b uses its first offsets[b+1]-offsets[b] columns. */
SEDT_API int sedt_lsap(const float* cost, int ld_cost, const int32_t* offsets, int B, int Q, int kmax,
              int64_t* rows, int64_t* cols, int32_t* counts, int32_t* status, void* stream);

/* ---- SetCriterion.forward, supervised default path (sedt/sedt.py:309-352 with fine_tune = normalize = fl = False and no
 * mixup 'ratio'; losses labels / boxes / cardinality / weak): for every decoder layer l (aux_outputs 0..L-2, then the top
 * layer) and clip b one warp runs the matcher (as sedt_matcher), loss_labels (:188-221), loss_boxes (:238-261),
 * loss_cardinality (:223-236) AND their gradients; a second launch reduces the per-clip partials in a fixed order and
 * evaluates loss_weak (:161-186) with its gradient.
 *   logits [L, B, Q, C1], boxes [L, B, Q, 2] fp32; the first Bs clips are strong_mask (matched), all B count for cardinality
 *   at [Bw, C1-1] or null; wl_labels / wl_offsets [Bw+1]: labels of the first Bw clips (targets[i]['labels'], :169-174)
 *   tgt_labels / tgt_boxes / offsets [Bs+1]: as sedt_matcher, for the strong clips; n_tgt [B] = len(targets[b]['labels'])
 *   num_boxes: sum of the Coef the matcher returns = sum_b min(Q, K_b) (sedt.py:323-324)
 *   rows, cols [L, Bs, Q] int64 (-1 padded), status [1] zeroed by the caller, partials [L, B, 8] scratch
 *   losses [L, 8]: loss_ce, loss_bbox, loss_giou, class_error, cardinality_error, loss_weak (row L-1 only), 0, 0
 *   g_logits [L, B, Q, C1] = d loss_ce[l] / d logits; g_l1, g_giou [L, B, Q, 2] = d loss_bbox[l], d loss_giou[l] / d boxes;
 *   g_at [Bw, C1-1] = d loss_weak / d at */
SEDT_API int sedt_set_criterion(const float* logits, const float* boxes, const float* at, const int64_t* tgt_labels,
                                const float* tgt_boxes, const int32_t* offsets, const float* n_tgt, const int64_t* wl_labels,
                                const int32_t* wl_offsets, int L, int B, int Bs, int Bw, int Q, int C1, int kmax,
                                float cost_class, float cost_bbox, float cost_giou, float eos_coef, float num_boxes,
                                int64_t* rows, int64_t* cols, int32_t* status, float* partials, float* losses,
                                float* g_logits, float* g_l1, float* g_giou, float* g_at, void* stream);

/* ---- The consumer right after the eval forward (engine.py:277-291): PostProcess.forward (sedt/sedt.py:359-396) and
 * BoxEncoder.decode_strong with del_overlap (utilities/BoxEncoder.py:179-226), one warp per clip.
 *   logits [B, Q, C1], boxes [B, Q, 2] (center, width), target_sizes [B] seconds (ignored when is_semi)
 *   audio_tags [B, C1-1] fp32 0/1 or null; at_m 1/2/3 = `fusion_strategy` (train_sedt.py:71); fuse_threshold = PostProcess's
 *   `threshold` (0.5); score_threshold = decode_strong's threshold; min_duration = 0.2 s (BoxEncoder.py:205)
 *   scores [B, Q], labels [B, Q] int64, boxes_se [B, Q, 2]: the PostProcess result
 *   ev_class / ev_onset / ev_offset / ev_score [B, Q], ev_count [B]: decoded events per clip in the reference's order
 *   (classes by first appearance among the kept queries, events of a class by onset); pass ev_count = null to skip decoding. */
SEDT_API int sedt_decode_events(const float* logits, const float* boxes, const float* target_sizes, const float* audio_tags,
                                int B, int Q, int C1, int at_m, float fuse_threshold, int is_semi, float score_threshold,
                                float min_duration, float* scores, int64_t* labels, float* boxes_se, int32_t* ev_class,
                                float* ev_onset, float* ev_offset, float* ev_score, int32_t* ev_count, void* stream);

/* get_pseudo_labels (engine.py:300-348): teacher outputs -> pseudo targets for the unlabelled clips.  PostProcess with
 * at_m = 1 and is_semi (audio_tags [B, C1-1] fp32 0/1 = at >= classwise_threshold, or null), keep queries with
 * score >= class_threshold[label] and width > min_width (= 0.2 / clip seconds), then (del_overlap) greedy same-class overlap
 * suppression in descending score order.  labels / scores [B, Q], boxes_out [B, Q, 2] (center, width): the first counts[b]
 * entries are the kept queries in the reference's order. */
SEDT_API int sedt_pseudo_labels(const float* logits, const float* boxes, const float* audio_tags, const float* class_threshold,
                                int B, int Q, int C1, float min_width, int del_overlap, int64_t* labels, float* boxes_out,
                                float* scores, int32_t* counts, void* stream);

/* ---- The deterministic (evaluation) input pipeline of utilities/BoxTransforms.py:454-490: ApplyLog (librosa.amplitude_to_db:
 * max(10 log10(max(1e-10, S^2)), clip maximum - 80)) -> PadOrTrunc(frames) (zero rows in the dB domain) -> ToTensor ->
 * Normalize ((x - mean_[f]) / std_[f] in float64, utilities/Scaler.py:102-108).
 *   raw [sum T_b, F] fp32 mel amplitudes, clip b = rows offsets[b] .. offsets[b+1] (int64 [B+1]); mean / std [F] float64 or
 *   both null; out [B, 1, frames, F] fp32 = the tensor sedt_forward reads; apply_log = 0 skips the dB conversion. */
SEDT_API int sedt_prepare_clips(const float* raw, const int64_t* offsets, const double* mean, const double* std, float* out,
                                int B, int frames, int F, int apply_log, void* stream);

/* ---- training-time input transforms (csrc/augment.cu) ------------------------------------------------
 * sedt_augment_clips: TimeMask -> FreqMask -> FreqShift (utilities/BoxTransforms.py:363-452) in place on x [B, T, F] fp32 (the
 * padded log-mel clips BEFORE Normalize).  One record per clip, the integer bands as the reference computes them on the host
 * (int(fraction * n)); tm_t == 0 / fm_mode == 0 / fs_shift == 0 skip a transform.  fm_mode: 1 = constant fill (fm_const),
 * 2 = mean of the band (np.mean in float32, reproduced bit for bit).  scratch: B * T floats. */
typedef struct sedt_augment_params {
    int32_t tm_t0, tm_t;          /* TimeMask: rows [tm_t0, tm_t0 + tm_t) are multiplied by 0 */
    int32_t fm_f0, fm_f, fm_mode; /* FreqMask: bins [fm_f0, fm_f0 + fm_f) */
    float fm_const;
    int32_t fs_shift, reserved;   /* FreqShift: np.roll by fs_shift bins, wrapped bins zeroed */
} sedt_augment_params;
SEDT_API int sedt_augment_clips(float* x, const sedt_augment_params* params, int B, int T, int F, float* scratch, void* stream);
/* mixup's data path (utilities/mixup.py:35): out[k] = a * x[i1] + b * x[i2] over rows of row_elems fp32 (b == 0: copy of x[i1]) */
typedef struct sedt_mix_row { int32_t i1, i2; float a, b; } sedt_mix_row;
SEDT_API int sedt_mix_rows(const float* x, float* out, const sedt_mix_row* rows, int n_out, int64_t row_elems, void* stream);
/* SP-SEDT patch crop + resize (utilities/BoxTransforms.py:315-360, Query): x [B, 1, T, F] fp32, bounds [B*P][2] = (s_idx, e_idx)
 * frame range of every patch, out [B, P, 1, 128, F]; Pillow's 8-bit antialiased bilinear resample, bit-exact */
SEDT_API int sedt_query_patches(const float* x, const int32_t* bounds, float* out, int B, int P, int T, int F, int fixed_patch_size,
                                void* stream);

/* ---- clip_grad_norm_ + AdamW: the optimizer half of the training step (engine.py:76-80; AdamW with two lr groups,
 * train_sedt.py:234-240,269-270; torch/optim/adamw.py _single_tensor_adamw arithmetic, amsgrad = maximize = False).
 * The caller keeps a device table of tensors and a device table of (tensor index, chunk index) pairs that splits
 * every tensor into chunks of sedt_optim_chunk_elems() elements (one CTA each).
 *   sedt_grad_norm:  norm_out[0] = || all grads ||_2 (fixed-order reduction; partials: nchunks floats of scratch)
 *   sedt_clip_grads: grads *= min(1, max_norm / (norm[0] + 1e-6))        (torch.nn.utils.clip_grad_norm_)
 *   sedt_adamw_step: one fused pass; when `norm` is non-null and max_norm > 0 the clip coefficient is applied to the
 *                    gradients on the fly (the stored grads are left unscaled) -- no host synchronisation anywhere. */
typedef struct sedt_optim_tensor {
    float* param; float* grad; float* exp_avg; float* exp_avg_sq;
    int64_t numel; int32_t group; int32_t reserved;
} sedt_optim_tensor;
/* per parameter group, computed in double on the host and rounded once (as torch does with Python scalars):
 * decay = 1 - lr*weight_decay, w1 = 1 - beta1, w2 = 1 - beta2, bc2_sqrt = sqrt(1 - beta2^step),
 * neg_step = -lr / (1 - beta1^step) */
typedef struct sedt_adamw_group { float decay, w1, beta2, w2, bc2_sqrt, eps, neg_step, reserved; } sedt_adamw_group;
SEDT_API int sedt_optim_chunk_elems(void);
SEDT_API int sedt_grad_norm(const sedt_optim_tensor* tensors, const int32_t* chunks, int nchunks, float* partials,
                            float* norm_out, void* stream);
SEDT_API int sedt_clip_grads(const sedt_optim_tensor* tensors, const int32_t* chunks, int nchunks, const float* norm,
                             float max_norm, void* stream);
SEDT_API int sedt_adamw_step(const sedt_optim_tensor* tensors, const int32_t* chunks, int nchunks,
                             const sedt_adamw_group* groups /* [host], 1..8 */, int ngroups, const float* norm, float max_norm,
                             void* stream);

/* ---- single operators, exported so the parity tests can pin each kernel against the oracle ---- */
typedef struct sedt_conv_desc {
    const void* in; const void* w; const float* scale; const float* bias; const void* residual; void* out;
    int32_t in_dtype, out_dtype;
    int32_t B, H, W, Cin, lda, Ho, Wo, Cout, ldc, ld_res, R, S, stride, dil, pad, relu;
} sedt_conv_desc;
/* Conv/linear + FrozenBN scale/bias + residual + ReLU as implicit GEMM (NHWC in, [Cout][R][S][Cin] weights).
 * engine: 0 = CUDA-core kernel, 1 = TMA + tcgen05 kernel (bf16 in; picks the 1-SM or 2-SM variant),
 *         2 = force the cta_group::2 variant (Cout % 256 == 0), 3 = force the weight-stationary variant (K <= 256,
 *         Cout % 128 == 0). */
SEDT_API int sedt_op_conv(const sedt_conv_desc* d, int engine, void* stream);
/* Weight gradient of the same layer (autograd's conv2d / addmm weight backward for sedt/backbone.py,
 * sedt/transformer.py): dw [Cout][k*k*Cin] fp32 += sum over output pixels of dy[m, co] * x[m shifted by tap, ci].
 * x NHWC bf16 [B,H,W,Cin], dy NHWC bf16 [B,Ho,Wo,Cout]; dw must be zeroed (or hold a running sum) by the caller. */
/* Non-GEMM backward operators (bf16 activations / gradients, fp32 parameter gradients accumulated atomically).
 * Data gradient of a conv / linear layer = sedt_op_conv on dy with the weights from sedt_op_repack_dgrad
 * (wd[ci][r'][s'][co] = scale[co] * w[co][ci][R-1-r'][S-1-s']), after sedt_op_upsample2 when the layer had stride 2;
 * relu = 2 in the conv descriptor turns the residual input into a ReLU mask (out = residual > 0 ? acc : 0). */
SEDT_API int sedt_op_repack_dgrad(const float* w_oihw, const float* scale, void* out, int dtype, int Cout, int Cin, int R, int S,
                                  void* stream);
SEDT_API int sedt_op_upsample2(const void* dy, void* u, int B, int H, int W, int Ho, int Wo, int C, void* stream);
SEDT_API int sedt_op_relu_mask(const void* act, const void* g1, const void* g2, void* out, int64_t n, void* stream);
SEDT_API int sedt_op_colsum(const void* in, int dtype, int64_t ld, float* out, int64_t M, int N, void* stream);
/* torch.nn.LayerNorm backward over 256 features (sedt/transformer.py norm1-3): g1, g2 bf16 and g3 fp32 are gradients of the
 * forward's y / y+pos / fp32 outputs (NULL = none), dres is added to dx. */
SEDT_API int sedt_op_layernorm_bwd(const float* x, const float* gamma, const void* g1, const void* g2, const float* g3,
                                   const float* dres, float* dx, float* dgamma, float* dbeta, int64_t rows, void* stream);
/* backward of the attention core of nn.MultiheadAttention (softmax(QK^T * scale + masks) V), bf16, head_dim 32;
 * engine 0 = CUDA-core kernel, 1 = tcgen05 kernel */
SEDT_API int sedt_op_attention_bwd(const void* Q, int ldq, const void* K, int ldk, const void* V, int ldv, const void* dO, int ldo,
                                   void* dQ, int lddq, void* dK, int lddk, void* dV, int lddv, const uint8_t* kpm,
                                   const float* amask, int B, int nheads, int Lq, int Lk, float scale, int engine, void* stream);
/* keep flags (1 = kept) of elements [0, n) of dropout site `site` at training step `step`: exactly the masks the
 * training kernels draw (sites: encoder layer l -> 8l + {0 attention weights, 1 after out_proj, 2 FFN hidden, 3 after
 * linear2}; decoder layer l -> 1024 + 8l + {0 self-attn weights, 1 after its out_proj, 2 cross-attn weights, 3 after its
 * out_proj, 4 FFN hidden, 5 after linear2}; element index: row-major [B*L, C] for activations,
 * ((clip*heads + head)*128 + query)*128 + key for attention weights).  Test hook. */
SEDT_API int sedt_op_dropout_mask(uint8_t* out, int64_t n, uint64_t seed, uint64_t step, uint32_t site, float p, void* stream);
SEDT_API int sedt_op_conv_wgrad(const void* x, const void* dy, float* dw, int B, int H, int W, int Cin, int Cout, int k,
                                int stride, int dil, int pad, void* stream);
SEDT_API int sedt_op_conv_tc_supported(const sedt_conv_desc* d);
/* Fused transformer FFN of the eval forward (sedt/transformer.py:202-203): out[M,256] fp32 = residual + relu(x W1^T + b1) W2^T + b2,
 * x [M,256] bf16, W1 [ff,256] / W2 [256,ff] bf16 (nn.Linear layout), ff % 256 == 0; the [M,ff] hidden activation never leaves the SM. */
/* fused tail of a layer1 bottleneck (csrc/bneck_fused.cu; torchvision resnet.py:150-161): out[B,H,W,256] (bf16 NHWC) =
 * relu(conv1x1(relu(conv3x3(h1, w2) + bias2), w3) + bias3 + residual); h1 [B,H,W,64] bf16, w2 [64][3][3][64] / w3 [256][64] bf16
 * with the FrozenBN scale folded in, fp32 biases, residual [B,H,W,256] bf16.  Exported for its parity test. */
SEDT_API int sedt_op_bneck_tail(const void* h1, const void* w2, const float* bias2, const void* w3, const float* bias3,
                                const void* residual, void* out, int B, int H, int W, void* stream);
/* fused encoder self-attention block (csrc/enc_attn_fused.cu; sedt/transformer.py:192-198): in place
 * x[B*S,256] (fp32) += out_proj(MHA(q = k = nap, v = na)), na / nap [B*S,256] bf16 = LN(x) / LN(x)+pos, w_in [768,256] and
 * w_out [256,256] bf16 as nn.MultiheadAttention stores them, b_in [768] / b_out [256] fp32, kpm [B,S] uint8 (1 = padded key)
 * or NULL; S <= 128 tokens per clip, 8 heads of 32.  ln_out (bf16 [B*S,256], optional with ln_g / ln_b [256]): LayerNorm of the
 * updated rows (the layer's norm2, sedt/transformer.py:199) from the same launch */
SEDT_API int sedt_op_enc_attn(const void* na, const void* nap, const void* w_in, const float* b_in, const void* w_out,
                              const float* b_out, const uint8_t* kpm, float* x, int B, int S, const float* ln_g, const float* ln_b,
                              void* ln_out, void* stream);

SEDT_API int sedt_op_ffn(const void* x, const void* w1, const float* b1, const void* w2, const float* b2, const float* residual,
                         float* out, int64_t M, int ff, void* stream);
/* OIHW fp32 -> O(HW)I in `dtype` */
SEDT_API int sedt_op_repack_conv(const float* w_oihw, void* out, int dtype, int Cout, int Cin, int R, int S, void* stream);
SEDT_API int sedt_op_cast(const float* in, void* out, int dtype, int64_t n, void* stream);
/* conv0 + conv1 + bn1 + relu + maxpool (sedt/backbone.py:102 + resnet stem); out NHWC [B, Hp, 16, 64] */
SEDT_API int sedt_op_stem(const float* x, const float* conv0_w, const float* conv0_b, const float* conv1_w,
                 const float* bn_w, const float* bn_b, const float* bn_mean, const float* bn_var,
                 void* scratch /* >= 32 KiB */, void* out, int out_dtype, int B, int T, int F, void* stream);
/* the same stem on tcgen05 tensor cores (bf16 output only) */
SEDT_API int sedt_op_stem_tc(const float* x, const float* conv0_w, const float* conv0_b, const float* conv1_w,
                    const float* bn_w, const float* bn_b, const float* bn_mean, const float* bn_var,
                    void* scratch /* >= 64 KiB */, void* out, int B, int T, int F, void* stream);
SEDT_API int sedt_op_layernorm(const float* x, const float* gamma, const float* beta, const float* pos, int64_t pos_rows,
                      void* y, void* ypos, float* y32, int dtype, int64_t rows, void* stream);
SEDT_API int sedt_op_attention(const void* Q, int ldq, const void* K, int ldk, const void* V, int ldv, void* O, int ldo,
                      int dtype, const uint8_t* key_padding_mask, const float* attn_mask,
                      int B, int nheads, int Lq, int Lk, float scale, void* stream);
SEDT_API int sedt_op_pos_table(const uint8_t* mask /* [B,T,F] or null */, uint8_t* mask_ds /* [B,H*W] scratch or null */,
                      float* pos, int B, int T, int F, int H, int W, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SEDT_B200_H */
