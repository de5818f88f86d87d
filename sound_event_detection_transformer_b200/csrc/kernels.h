// Internal launcher declarations shared by the kernel translation units and
// the host runtime (model.cu / api.cu).  Nothing here is exported.
#pragma once
#include "common.cuh"

namespace sedt {

enum DType : int { DT_F32 = 0, DT_BF16 = 1 };
static inline size_t dtype_size(int dt) { return dt == DT_F32 ? 4 : 2; }

// One convolution / linear layer as an implicit GEMM over NHWC activations:
//   out[m, n] = act( (sum_k A[m,k] * W[n,k]) * scale[n] + bias[n] + residual[m,n] )
// with m = (b, ho, wo), k = (r, s, c), W stored [Cout][R][S][Cin].
// A linear layer on [rows, K] is B=rows, H=W=Ho=Wo=1, R=S=1.
struct ConvGemm {
    const void* in = nullptr;      // activation, dtype in_dt, pixel stride lda elements
    const void* w = nullptr;       // weights, dtype in_dt
    const float* scale = nullptr;  // per-Cout (FrozenBN fold) or nullptr (=1)
    const float* bias = nullptr;   // per-Cout or nullptr (=0)
    const void* residual = nullptr;  // dtype out_dt, row stride ld_res, or nullptr
    void* out = nullptr;           // dtype out_dt, row stride ldc
    int in_dt = DT_F32, out_dt = DT_F32;
    int B = 0, H = 1, W = 1, Cin = 0, lda = 0;
    int Ho = 1, Wo = 1, Cout = 0, ldc = 0, ld_res = 0;
    int R = 1, S = 1, stride = 1, dil = 1, pad = 0;
    int relu = 0;
};

// ---- conv_simt.cu : fp32-accumulate CUDA-core implicit GEMM (precise tier + small shapes)
int launch_conv_simt(const ConvGemm& g, cudaStream_t stream);

// ---- gemm_tc.cu : TMA + tcgen05/TMEM implicit GEMM (bf16 tier)
bool conv_tc_supported(const ConvGemm& g);
int launch_conv_tc(const ConvGemm& g, cudaStream_t stream);      // gemm_tc2.cu: persistent kernel (or v1 if SEDT_TC_V1=1)
int launch_conv_tc_v1(const ConvGemm& g, cudaStream_t stream);
// gemm_tc4.cu: weight-stationary variant for K <= 256
bool conv_tc_ws_supported(const ConvGemm& g);
int launch_conv_tc_ws(const ConvGemm& g, cudaStream_t stream);
// gemm_tc3.cu: cta_group::2 (two SMs per 256 x 256 tile) for bf16-out layers with Cout % 256 == 0
bool conv_tc_2sm_supported(const ConvGemm& g);
bool conv_tc_2sm_preferred(const ConvGemm& g);    // ... and enough tiles to fill the SM pairs
int launch_conv_tc_2sm(const ConvGemm& g, cudaStream_t stream);
int tc_init();     // resolves cuTensorMapEncodeTiled once; safe without a GPU

// ---- gemm_wgrad.cu : weight gradient dW[co,(r,s),ci] += sum_pixels dY[m,co] * X[shifted m, ci] on tcgen05 (MN-major operands)
struct WgradGemm {
    const void* x = nullptr;       // forward input, NHWC bf16, pixel stride lda
    const void* dy = nullptr;      // gradient of the forward output, NHWC bf16, pixel stride ldy
    float* dw = nullptr;           // fp32 [Cout][R*S*Cin], accumulated atomically (caller zeroes)
    const float* row_scale = nullptr;   // optional [Cout]: every contribution to row co is multiplied by row_scale[co] (folded BN scale)
    int B = 0, H = 1, W = 1, Cin = 0, lda = 0;
    int Ho = 1, Wo = 1, Cout = 0, ldy = 0;
    int R = 1, S = 1, stride = 1, dil = 1, pad = 0;
};
bool conv_wgrad_tc_supported(const WgradGemm& g);
int launch_conv_wgrad_tc(const WgradGemm& g, cudaStream_t stream);

// ---- stem.cu : conv0(1x1,bias) + conv1(7x7 s2 p3) + FrozenBN + ReLU + maxpool(3x3 s2 p1), F == 64
struct StemWeights {
    const float* weff;   // [49][64]   sum_c conv1[o][c][tap] * conv0.w[c]
    const float* sat;    // [8][8][64] inclusive 2-D prefix sums of sum_c conv1[o][c][tap] * conv0.b[c]
    const float* scale;  // [64] bn1 fold
    const float* bias;   // [64]
};
int launch_stem_pack(const float* conv0_w, const float* conv0_b, const float* conv1_w, float* weff, float* sat,
                     cudaStream_t stream);
int launch_stem(const float* x, const StemWeights& w, void* out, int out_dt, int B, int T, int F, cudaStream_t stream);

// stem_tc.cu: the same stem on tcgen05 (bf16 output).  wtc: 16 KiB pre-swizzled weight image, bn_scale folded in
int launch_stem_tc_pack(const float* conv0_w, const float* conv0_b, const float* conv1_w, const float* bn_scale, void* wtc,
                        cudaStream_t stream);
// amax (optional, training): [B, Hp, 16, 64] uint8 arg-max of every pooling window (0..8 = dr*3+dc, 9 = dead ReLU)
// sat: the 8x8x64 summed-area table of the conv0-bias taps (launch_stem_pack), bn_scale / bn_bias: folded FrozenBN
int launch_stem_tc(const float* x, const void* wtc, const float* bn_bias, const float* bn_scale, const float* sat, void* out, int B,
                   int T, int F, cudaStream_t stream, uint8_t* amax = nullptr);

// ---- pack.cu
int launch_bn_fold(const float* w, const float* b, const float* mean, const float* var, float* scale, float* bias,
                   int n, cudaStream_t stream);
// OIHW fp32 -> O(HW)I in dtype dt; optional per-output-channel scale folded into the weights
int launch_repack_conv(const float* w_oihw, const float* scale, void* out, int dt, int Cout, int Cin, int R, int S,
                       cudaStream_t stream);
int launch_cast(const float* in, void* out, int dt, int64_t n, cudaStream_t stream);

// ---- backward.cu : the non-GEMM kernels of the backward pass (bf16 tier)
// dX = conv(dY, Wd):  Wd[ci][r'][s'][co] = scale[co] * W[co][ci][R-1-r'][S-1-s'] from the OIHW fp32 weights
// (the Cout axis, which the data-gradient GEMM reduces over, is zero-padded to Cout_pad)
int launch_repack_dgrad(const float* w_oihw, const float* scale, void* out, int dt, int Cout, int Cout_pad, int Cin, int R, int S,
                        cudaStream_t stream);
// grad[co][ci][r][s] = scale[co] * dw[co][(r,s)][ci]: gemm_wgrad.cu's result back in the reference's OIHW layout
int launch_unpack_wgrad(const float* dw, const float* scale, float* grad, int Cout, int Cin, int RS, cudaStream_t stream);
// gradients of pred_logits / pred_boxes / at -> zero-padded [rows, 128] bf16 GEMM operands (sigmoid backward included)
int launch_heads_bwd_prepare(const float* d_logits, const float* d_boxes, const float* d_at, const float* boxes, const float* at,
                             void* dcls, void* dbox, void* dweak, int D_, int B, int Qall, int start, int C1, int C,
                             cudaStream_t stream);
// zero insertion for stride-2 transposed convolutions: u[b,2ho,2wo,:] = dy[b,ho,wo,:] (bf16, C % 8 == 0)
int launch_upsample2(const void* dy, void* u, int B, int H, int W, int Ho, int Wo, int C, cudaStream_t stream);
// out = act > 0 ? g1 + g2 : 0 (bf16; g2 may be null; out may alias g1)
int launch_relu_mask(const void* act, const void* g1, const void* g2, void* out, int64_t n, cudaStream_t stream);
// out[n] += sum_m in[m*ld + n]
int launch_colsum(const void* in, int dt, int64_t ld, float* out, int64_t M, int N, cudaStream_t stream);
// LayerNorm (D = 256) backward; g1/g2 bf16 and g3 fp32 are the gradients of the forward's y / ypos / y32 outputs (any may
// be null), dres an fp32 gradient added to dx (the residual branch); dgamma / dbeta accumulate atomically
int launch_layernorm_bwd(const float* x, const float* gamma, const void* g1, const void* g2, const float* g3, const float* dres,
                         float* dx, float* dgamma, float* dbeta, int64_t rows, cudaStream_t stream);
// gradients of conv0.weight[3] / conv0.bias[3] from G = d/d(stem output) (bf16 [B,Hp,16,64], ReLU mask applied) and the
// pooling arg-max the training forward recorded (launch_stem_tc amax); scratch: 2*49*64 floats
int launch_stem_bwd(const float* x, const float* conv1_w, const float* bn_scale, const void* G, const uint8_t* amax,
                    float* scratch, float* g_conv0_w, float* g_conv0_b, int B, int T, int F, cudaStream_t stream);
// attention core backward (bf16, head_dim 32, Lq, Lk <= 128): recomputes P from Q, K and the masks
int launch_attention_bwd(const void* Q, int ldq, const void* K, int ldk, const void* V, int ldv, const void* dO, int ldo,
                         void* dQ, int lddq, void* dK, int lddk, void* dV, int lddv, const uint8_t* kpm, const float* amask,
                         int B, int nheads, int Lq, int Lk, float scale, cudaStream_t stream);

// attention_bwd_tc.cu: the same on tcgen05 / TMEM (production path; the SIMT kernel above is the plain reference)
bool attention_bwd_tc_supported(const void* Q, int ldq, const void* K, int ldk, const void* V, int ldv, const void* dO, int ldo,
                                const void* dQ, int lddq, const void* dK, int lddk, const void* dV, int lddv, int Lq, int Lk);
int launch_attention_bwd_tc(const void* Q, int ldq, const void* K, int ldk, const void* V, int ldv, const void* dO, int ldo,
                            void* dQ, int lddq, void* dK, int lddk, void* dV, int lddv, const uint8_t* kpm, const float* amask,
                            int B, int nheads, int Lq, int Lk, float scale, const struct DropSite* drop, cudaStream_t stream);

// ---- dropout.cu / dropout.cuh : counter-based (Philox) dropout of the training step; state = {seed, step} on the device
struct DropSite;
int launch_fill_value(float* p, float v, int n, cudaStream_t stream);
int launch_dropout_mask(unsigned char* out, int64_t n, unsigned long long seed, unsigned long long step, uint32_t site, float p,
                        cudaStream_t stream);
int launch_rng_init(unsigned long long* state, unsigned long long seed, cudaStream_t stream);
int launch_rng_step(unsigned long long* state, cudaStream_t stream);
int launch_dropout_add(const float* y, const float* resid, float* out, int64_t n, const DropSite& d, cudaStream_t stream);
int launch_dropout_bf16(void* h, int64_t n, const DropSite& d, cudaStream_t stream);
int launch_cast_dropout(const float* g, void* out16, int64_t n, const DropSite& d, cudaStream_t stream);
// attention_tc.cu: training forward with dropout on the attention weights
int launch_attention_tc_drop(const void* Q, int ldq, const void* K, int ldk, const void* V, int ldv, void* O, int ldo,
                             const uint8_t* kpm, const float* amask, int B, int nheads, int Lq, int Lk, float scale,
                             const DropSite& drop, cudaStream_t stream);

// ---- transformer.cu
// LayerNorm over D=256, eps 1e-5.  Any of y / ypos / y32 may be null.
//   y    = LN(x)                      (dtype dt)
//   ypos = LN(x) + pos[row % pos_rows] (dtype dt)
//   y32  = LN(x)                      (fp32)
int launch_layernorm(const float* x, const float* gamma, const float* beta, const float* pos, int64_t pos_rows,
                     void* y, void* ypos, float* y32, int dt, int64_t rows, cudaStream_t stream);
// y = x (cast), ypos = x + pos[row % pos_rows]  (post-norm path and memory+pos)
int launch_cast_addpos(const float* x, const float* pos, int64_t pos_rows, void* y, void* ypos, int dt, int64_t rows,
                       cudaStream_t stream);
// Multi-head attention core, head_dim 32: O = softmax(Q K^T * scale + amask + kpm) V
int launch_attention(const void* Q, int ldq, const void* K, int ldk, const void* V, int ldv, void* O, int ldo, int dt,
                     const uint8_t* kpm, const float* amask, int B, int nheads, int Lq, int Lk, float scale,
                     cudaStream_t stream);
// attention_tc.cu: tcgen05 version for bf16, Lq, Lk <= 128 (SEDT_ATT_SIMT=1 disables it)
bool attention_tc_supported(const void* Q, int ldq, const void* K, int ldk, const void* V, int ldv, const void* O, int ldo,
                            int dt, int nheads, int Lq, int Lk);
int launch_attention_tc(const void* Q, int ldq, const void* K, int ldk, const void* V, int ldv, void* O, int ldo,
                        const uint8_t* kpm, const float* amask, int B, int nheads, int Lq, int Lk, float scale,
                        cudaStream_t stream);
int launch_mask_downsample(const uint8_t* mask, uint8_t* out, int B, int T, int F, int H, int W, cudaStream_t stream);
// sine position table: [nb][H*W][256] fp32; mask_ds null => unpadded (nb must be 1)
int launch_pos_table(const uint8_t* mask_ds, float* pos, int nb, int H, int W, cudaStream_t stream);
int launch_fill_zero(void* p, size_t bytes, cudaStream_t stream);
// slice decoder slots and apply sigmoid: see model.cu
// class_embed / last bbox_embed layer (+ sigmoid) / weak_class_embed (+ sigmoid) in one pass over the decoder states:
// hs [D*B*Qall, 256] fp32, h2 = second box-MLP hidden layer [D*B*Qall, 256] fp32, fp32 weights [out, 256]
int launch_heads_out(const float* hs, const float* h2, const float* wc, const float* bc, const float* wb, const float* bb,
                     const float* ww, const float* bw, float* logits, float* boxes, float* at, int D_, int B, int Qall, int start,
                     int C1, int C, cudaStream_t stream);
int launch_heads_finalize(const float* cls_raw, const float* box_raw, const float* weak_raw, float* logits, float* boxes,
                          float* at, int D, int B, int Qall, int start, int C1, int C, cudaStream_t stream);
// [N, HW, C] -> [N, C] mean (SP-SEDT avgpool), fp32 out
int launch_avgpool(const void* x, int dt, float* out, int N, int HW, int C, cudaStream_t stream);
// query_pos[b, q, :] = pq[b, q / qpp, :] + query_embed[start + q, :]
// training branch (spsedt.py:63-67): keep [B, P*qpp] 1 = add the patch feature; qe_scale = 2 (decoder_input += ... + decoder_input)
int launch_patch_query(const float* pq, const float* query_embed, float* out, int B, int P, int qpp, int start,
                       cudaStream_t stream, const uint8_t* keep = nullptr, float qe_scale = 1.f);
// acc[i] += g[i] (bf16 -> fp32), and the backward of launch_patch_query's training branch:
//   d_qe[q, :] += qe_scale * sum_b dq[b, q, :],   d_pq[b, p, :] = sum_{q in patch p} keep[b, q] * dq[b, q, :]
int launch_accum_bf16(const void* g, float* acc, int64_t n, cudaStream_t stream);
int launch_patch_query_bwd(const float* dq, const uint8_t* keep, float* d_qe, float* d_pq, int B, int P, int qpp, float qe_scale,
                           cudaStream_t stream);

// additive block-diagonal 0/-inf decoder mask [Q,Q] (sedt/spsedt.py:27-32)
int launch_blockdiag_mask(float* m, int Q, int qpp, cudaStream_t stream);

int prof_read(double* ms, long long* counts);   // model.cu

// every data-gradient weight re-layout (launch_repack_dgrad with R == S, bf16 output) of a backward pass in one launch
struct DgradJob { const float* w; const float* scale; void* out; int32_t Cout, Cout_pad, Cin, RS, tile0, pad_; };
constexpr int kDgradMaxJobs = 256;
struct DgradJobs { DgradJob j[kDgradMaxJobs]; int32_t n; };
int launch_repack_dgrad_batched(const DgradJob* jobs, int n, cudaStream_t stream);

// ---- pack.cu: all casts / repacks / BN folds of one pack call batched into one launch (jobs in the kernel parameters)
enum : int { PACK_CAST = 0, PACK_REPACK = 1, PACK_BNFOLD = 2 };
struct PackJob {
    const float* src; void* dst; float* dst2; const float* bn_w; const float* bn_mean; const float* bn_var;
    int32_t kind, dt, a, b, c, chunk0;
};
constexpr int kPackMaxJobs = 400;                       // 400 * 72 B = 28.8 KB of kernel parameters (limit 32 KB)
struct PackJobs { PackJob j[kPackMaxJobs]; int32_t n; };
class PackBatch {
public:
    explicit PackBatch(cudaStream_t s) : stream_(s) { cur_.n = 0; }
    int cast(const float* in, void* out, int dt, int64_t n);
    // out[o][tap][c] = w[o][c][tap] * (bn_w ? bn_w[o] * rsqrt(bn_var[o] + 1e-5) : 1)
    int repack_conv(const float* w_oihw, const float* bn_w, const float* bn_var, void* out, int dt, int Cout, int Cin, int RS);
    int bn_fold(const float* w, const float* b, const float* mean, const float* var, float* scale, float* bias, int n);
    int flush();
private:
    int add(const PackJob& job, int64_t elems);
    PackJobs cur_;
    int chunks_ = 0;
    cudaStream_t stream_;
};

// ---- ffn_fused.cu: out = residual + relu(x W1^T + b1) W2^T + b2 with the hidden activation kept on chip (eval forward)
bool ffn_fused_supported(int d, int ff, int64_t M, const void* x, const void* w1, const void* w2, const float* residual, const float* out,
                         int ld_res, int ldo);
bool ffn_fused_preferred(int64_t M);  // true where the fused kernel beats the two GEMMs (whole rounds of tiles over the SMs)
int launch_ffn_fused(const void* x, const void* w1, const float* b1, const void* w2, const float* b2, const float* residual, int ld_res,
                     float* out, int ldo, int64_t M, int ff, cudaStream_t stream);

// ---- bneck_fused.cu: conv2 (3x3, 64 -> 64) + conv3 (1x1, 64 -> 256) + residual + ReLU of a layer1 bottleneck in one launch
bool bneck_tail_enabled();            // SEDT_BNECK_FUSED (default on)
bool bneck_tail_supported(const ConvGemm& g2, const ConvGemm& g3);
int launch_bneck_tail(const ConvGemm& g2, const ConvGemm& g3, cudaStream_t stream);

// ---- enc_attn_fused.cu: x += out_proj(MHA(q = k = nap, v = na)) for one <= 128-token clip per tile, everything on chip
bool enc_attn_fused_supported(int d, int nheads, int S, const void* na, const void* nap, const void* w_in, const void* w_out,
                              const float* x, int dt);
bool enc_attn_fused_enabled();        // SEDT_ENC_ATTN_FUSED (default on)
// ln_out (optional, bf16 [B*S, 256]): LayerNorm(updated x; ln_g, ln_b) written by the same launch (the layer's norm2)
int launch_enc_attn_fused(const void* na, const void* nap, const void* w_in, const float* b_in, const void* w_out, const float* b_out,
                          const uint8_t* kpm, float* x, int B, int S, float scale, cudaStream_t stream, const float* ln_g = nullptr,
                          const float* ln_b = nullptr, void* ln_out = nullptr);

// ---- matcher.cu
int launch_matcher(const float* logits, const float* boxes, const int64_t* tgt_labels, const float* tgt_boxes,
                   const int32_t* offsets, int B, int Q, int C1, int Kmax, float w_class, float w_bbox, float w_giou,
                   const float* cost_in, int ld_in, float* cost_out, int ld_out,
                   int64_t* rows, int64_t* cols, int32_t* counts, int32_t* status, int solve, cudaStream_t stream,
                   int fl = 0, float alpha_fl = 0.f, float gamma_fl = 0.f, float* lmin = nullptr, int64_t* largmin = nullptr);

// matcher + loss_labels / loss_boxes / loss_cardinality / loss_weak and their gradients (sedt/sedt.py:309-352)
int launch_set_criterion(const float* logits, const float* boxes, const float* at, const int64_t* tgt_labels, const float* tgt_boxes,
                         const int32_t* offsets, const float* n_tgt, const int64_t* wl_labels, const int32_t* wl_offsets,
                         int L, int B, int Bs, int Bw, int Q, int C1, int Kmax, float w_class, float w_bbox, float w_giou,
                         float eos_coef, float num_boxes, int64_t* rows, int64_t* cols, int32_t* status, float* partials,
                         float* losses, float* g_logits, float* g_l1, float* g_giou, float* g_at, cudaStream_t stream);

// ---- decode.cu: PostProcess.forward + BoxEncoder.decode_strong (sedt/sedt.py:359-396, utilities/BoxEncoder.py:179-226)
int launch_decode_events(const float* logits, const float* boxes, const float* sizes, const float* tags, int B, int Q, int C1,
                         int at_m, float fuse_threshold, int is_semi, float score_threshold, float min_duration,
                         float* out_scores, int64_t* out_labels, float* out_boxes, int32_t* ev_class, float* ev_onset,
                         float* ev_offset, float* ev_score, int32_t* ev_count, cudaStream_t stream);

int launch_pseudo_labels(const float* logits, const float* boxes, const float* tags, const float* class_thr, int B, int Q, int C1,
                         float min_width, int del_overlap, int64_t* out_labels, float* out_boxes, float* out_scores,
                         int32_t* out_count, cudaStream_t stream);

// ---- prepare.cu: ApplyLog -> PadOrTrunc -> Normalize of the evaluation input pipeline (utilities/BoxTransforms.py:454-490)
int launch_prepare_clips(const float* raw, const int64_t* offsets, const double* mean, const double* stdv, float* out, int B,
                         int frames, int F, int apply_log, cudaStream_t stream);

// ---- augment.cu: TimeMask / FreqMask / FreqShift, mixup rows, SP-SEDT patch crop + resize (utilities/BoxTransforms.py:315-452,
// utilities/mixup.py:13-127); the structs mirror include/sedt_b200.h
struct AugmentParams { int32_t tm_t0, tm_t, fm_f0, fm_f, fm_mode; float fm_const; int32_t fs_shift, reserved; };
struct MixRow { int32_t i1, i2; float a, b; };
int launch_augment_clips(float* x, const AugmentParams* params, int B, int T, int F, float* row_sums, cudaStream_t stream);
int launch_mix_rows(const float* x, float* out, const MixRow* rows, int n_out, int64_t row_elems, cudaStream_t stream);
int launch_query_patches(const float* x, const int32_t* bounds, float* out, int B, int P, int T, int F, int fixed, cudaStream_t stream);

// ---- optim.cu: clip_grad_norm_ + AdamW over a (tensor, chunk) table (engine.py:76-80)
int optim_chunk_elems();
int launch_grad_norm(const void* tensors, const int32_t* chunks, int nchunks, float* partials, float* norm_out, cudaStream_t stream);
int launch_grad_scale(const void* tensors, const int32_t* chunks, int nchunks, const float* norm, float max_norm, cudaStream_t stream);
int launch_adamw(const void* tensors, const int32_t* chunks, int nchunks, const float* groups_host, int ngroups,
                 const float* norm, float max_norm, cudaStream_t stream);

}  // namespace sedt
