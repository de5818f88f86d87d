// extern "C" surface declared in include/sedt_b200.h.
#include "../../include/sedt_b200.h"
#include "model.h"

using namespace sedt;

struct sedt_model { Model* impl; };

static_assert(sizeof(sedt_config) == sizeof(Config), "sedt_config and sedt::Config must stay in lock-step");

static ConvGemm to_gemm(const sedt_conv_desc* d)
{
    ConvGemm g;
    g.in = d->in; g.w = d->w; g.scale = d->scale; g.bias = d->bias; g.residual = d->residual; g.out = d->out;
    g.in_dt = d->in_dtype; g.out_dt = d->out_dtype;
    g.B = d->B; g.H = d->H; g.W = d->W; g.Cin = d->Cin; g.lda = d->lda; g.Ho = d->Ho; g.Wo = d->Wo; g.Cout = d->Cout;
    g.ldc = d->ldc; g.ld_res = d->ld_res; g.R = d->R; g.S = d->S; g.stride = d->stride; g.dil = d->dil; g.pad = d->pad;
    g.relu = d->relu;
    return g;
}

extern "C" {

const char* sedt_last_error(void) { return get_error(); }
int sedt_abi_version(void) { return SEDT_ABI_VERSION; }
unsigned long long sedt_launch_count(void) { return g_launch_count; }
int sedt_kernel_kinds(void) { return KK_NKINDS; }
const char* sedt_kernel_kind_name(int kind) { return kernel_kind_name(kind); }
unsigned long long sedt_kernel_kind_count(int kind) { return kind >= 0 && kind < KK_NKINDS ? g_kind_count[kind] : 0ull; }

int sedt_profile_enable(int on) { g_prof_on = on != 0; return SEDT_OK; }
int sedt_profile_read(double* ms_per_class, long long* launches_per_class)
{
    SEDT_REQUIRE(ms_per_class != nullptr && launches_per_class != nullptr, "profile_read: null argument");
    return prof_read(ms_per_class, launches_per_class);
}

int sedt_model_create(const sedt_config* cfg, sedt_model** out)
{
    SEDT_REQUIRE(cfg != nullptr && out != nullptr, "model_create: null argument");
    SEDT_REQUIRE(cfg->hidden_dim == 256 && cfg->nheads == 8, "model_create: kernels are built for hidden_dim=256, nheads=8 "
                 "(train_sedt.py:86,95 defaults); got %d / %d", cfg->hidden_dim, cfg->nheads);
    SEDT_REQUIRE(cfg->dim_feedforward % 64 == 0 && cfg->dim_feedforward >= 64, "model_create: dim_feedforward=%d", cfg->dim_feedforward);
    SEDT_REQUIRE(cfg->enc_layers >= 0 && cfg->dec_layers >= 1 && cfg->num_queries >= 1 && cfg->num_classes >= 1,
                 "model_create: bad layer/query/class counts");
    SEDT_REQUIRE(cfg->precision == 0 || cfg->precision == 1, "model_create: precision must be 0 (fp32) or 1 (bf16)");
    SEDT_REQUIRE(!(cfg->self_sup && cfg->dec_at), "model_create: SP-SEDT has no audio query");
    SEDT_REQUIRE(!cfg->self_sup || (cfg->num_patches >= 1 && cfg->num_queries % cfg->num_patches == 0),
                 "model_create: num_queries must be a multiple of num_patches (sedt/spsedt.py:27)");
    Config c;
    memcpy(&c, cfg, sizeof(c));
    sedt_model* m = new sedt_model;
    m->impl = new Model(c);
    *out = m;      // the TMA driver entry point is resolved lazily at the first tensor-core launch
    return SEDT_OK;
}

void sedt_model_destroy(sedt_model* m)
{
    if (m == nullptr) return;
    delete m->impl;
    delete m;
}

int sedt_model_num_weights(const sedt_model* m) { return m ? (int)m->impl->slots().size() : 0; }
const char* sedt_model_weight_name(const sedt_model* m, int i)
{
    if (m == nullptr || i < 0 || i >= (int)m->impl->slots().size()) return nullptr;
    return m->impl->slots()[i].name.c_str();
}
int64_t sedt_model_weight_numel(const sedt_model* m, int i)
{
    if (m == nullptr || i < 0 || i >= (int)m->impl->slots().size()) return -1;
    return m->impl->slots()[i].numel;
}
int64_t sedt_model_packed_bytes(const sedt_model* m) { return m ? (int64_t)m->impl->packed_bytes() : -1; }

int sedt_model_pack(sedt_model* m, const void* const* weights, void* packed, int64_t packed_bytes, void* stream)
{
    SEDT_REQUIRE(m != nullptr && weights != nullptr && packed != nullptr, "model_pack: null argument");
    return m->impl->pack(weights, packed, (size_t)packed_bytes, (cudaStream_t)stream);
}

int sedt_feature_shape(int T, int F, int dilation, int* H, int* W)
{
    SEDT_REQUIRE(H != nullptr && W != nullptr && T >= 1 && F >= 1, "feature_shape: bad argument");
    Model::feature_shape(T, F, dilation != 0, H, W);
    return SEDT_OK;
}

int64_t sedt_workspace_bytes(sedt_model* m, int B, int T, int F, int P, int PT)
{
    if (m == nullptr) { set_error("workspace_bytes: null model"); return SEDT_ERR_INVALID; }
    Arena a(nullptr, 0);
    ForwardOut o{};
    // sized for the padded case (per-clip position table), an upper bound for the unpadded one
    static const uint8_t kAnyMask = 0;
    int rc = m->impl->forward(nullptr, &kAnyMask, B, T, F, nullptr, P, PT, a, o, nullptr, true);
    if (rc != 0) return rc;
    return (int64_t)((a.peak + 255) & ~(size_t)255);
}

int sedt_forward(sedt_model* m, const float* x, const uint8_t* mask, int B, int T, int F, const float* patches, int P,
                 int PT, void* workspace, int64_t workspace_bytes, const sedt_outputs* out, void* stream)
{
    SEDT_REQUIRE(m != nullptr && x != nullptr && out != nullptr, "forward: null argument");
    SEDT_REQUIRE(out->hs != nullptr && out->logits != nullptr && out->boxes != nullptr, "forward: hs/logits/boxes outputs are required");
    SEDT_REQUIRE(!m->impl->cfg().dec_at || out->at != nullptr, "forward: dec_at model needs the `at` output");
    SEDT_REQUIRE(((uintptr_t)workspace & 255) == 0, "forward: workspace must be 256-byte aligned");
    const int64_t need = sedt_workspace_bytes(m, B, T, F, P, PT);
    if (need < 0) return (int)need;
    if (workspace == nullptr || workspace_bytes < need) {
        set_error("forward: workspace too small (%lld bytes needed, %lld given)", (long long)need, (long long)workspace_bytes);
        return SEDT_ERR_WORKSPACE;
    }
    Arena a(workspace, (size_t)workspace_bytes);
    ForwardOut o{out->hs, out->logits, out->boxes, out->at, out->memory, out->pred_feature, out->gt_feature, out->feat};
    return m->impl->forward(x, mask, B, T, F, patches, P, PT, a, o, (cudaStream_t)stream, false);
}

// ---- training step ---------------------------------------------------------------------------------------
int64_t sedt_train_tape_bytes(sedt_model* m, int B, int T, int F, int has_mask) { return sedt_train_tape_bytes_sp(m, B, T, F, has_mask, 0, 0); }
int64_t sedt_backward_workspace_bytes(sedt_model* m, int B, int T, int F) { return sedt_backward_workspace_bytes_sp(m, B, T, F, 0, 0); }

int64_t sedt_train_tape_bytes_sp(sedt_model* m, int B, int T, int F, int has_mask, int P, int PT)
{
    if (m == nullptr) { set_error("train_tape_bytes: null model"); return SEDT_ERR_INVALID; }
    return m->impl->tape_bytes(B, T, F, has_mask != 0, P, PT);
}

int64_t sedt_backward_workspace_bytes_sp(sedt_model* m, int B, int T, int F, int P, int PT)
{
    if (m == nullptr) { set_error("backward_workspace_bytes: null model"); return SEDT_ERR_INVALID; }
    return m->impl->backward_workspace_bytes(B, T, F, P, PT);
}

int sedt_forward_train_sp(sedt_model* m, const float* x, const uint8_t* mask, int B, int T, int F, const float* patches, int P, int PT,
                          const uint8_t* query_keep, void* tape, int64_t tape_bytes, const sedt_outputs* out, float dropout,
                          uint64_t seed, void* stream)
{
    SEDT_REQUIRE(m != nullptr && x != nullptr && out != nullptr && tape != nullptr && patches != nullptr && query_keep != nullptr,
                 "forward_train_sp: null argument");
    SEDT_REQUIRE(out->hs != nullptr && out->logits != nullptr && out->boxes != nullptr, "forward_train_sp: hs/logits/boxes outputs are required");
    SEDT_REQUIRE(((uintptr_t)tape & 255) == 0, "forward_train_sp: tape must be 256-byte aligned");
    ForwardOut o{out->hs, out->logits, out->boxes, nullptr, out->memory, out->pred_feature, out->gt_feature, nullptr};
    Model::SpTrain sp;
    sp.patches = patches; sp.P = P; sp.PT = PT; sp.query_keep = query_keep;
    return m->impl->forward_train(x, mask, B, T, F, tape, (size_t)tape_bytes, o, dropout, (unsigned long long)seed, (cudaStream_t)stream, &sp);
}

int sedt_backward_sp(sedt_model* m, const void* const* weights, const float* x, const uint8_t* mask, int B, int T, int F, int P, int PT,
                     void* tape, int64_t tape_bytes, void* workspace, int64_t workspace_bytes, const float* d_logits,
                     const float* d_boxes, const float* d_pred_feature, float* grads, float dropout, void* stream)
{
    SEDT_REQUIRE(m != nullptr && weights != nullptr && x != nullptr && tape != nullptr && workspace != nullptr && grads != nullptr,
                 "backward_sp: null argument");
    SEDT_REQUIRE(((uintptr_t)tape & 255) == 0 && ((uintptr_t)workspace & 255) == 0 && ((uintptr_t)grads & 255) == 0,
                 "backward_sp: tape, workspace and grads must be 256-byte aligned");
    Model::SpTrain sp;
    sp.P = P; sp.PT = PT; sp.d_pred_feature = d_pred_feature;
    return m->impl->backward(weights, x, mask, B, T, F, tape, (size_t)tape_bytes, workspace, (size_t)workspace_bytes, d_logits,
                             d_boxes, nullptr, grads, 0, dropout, (cudaStream_t)stream, &sp);
}

int sedt_model_set_bucket_event(sedt_model* m, void* cuda_event)
{
    SEDT_REQUIRE(m != nullptr, "set_bucket_event: null model");
    m->impl->set_bucket_event((cudaEvent_t)cuda_event);
    return SEDT_OK;
}

int64_t sedt_grad_numel(const sedt_model* m) { return m == nullptr ? (int64_t)SEDT_ERR_INVALID : m->impl->grad_numel(); }

int64_t sedt_grad_offset(const sedt_model* m, int slot)
{
    if (m == nullptr || slot < 0 || slot >= (int)m->impl->slots().size()) { set_error("grad_offset: bad argument"); return SEDT_ERR_INVALID; }
    return m->impl->grad_offset(slot);
}

int sedt_forward_train(sedt_model* m, const float* x, const uint8_t* mask, int B, int T, int F, void* tape, int64_t tape_bytes,
                       const sedt_outputs* out, float dropout, uint64_t seed, void* stream)
{
    SEDT_REQUIRE(m != nullptr && x != nullptr && out != nullptr && tape != nullptr, "forward_train: null argument");
    SEDT_REQUIRE(out->hs != nullptr && out->logits != nullptr && out->boxes != nullptr, "forward_train: hs/logits/boxes outputs are required");
    SEDT_REQUIRE(!m->impl->cfg().dec_at || out->at != nullptr, "forward_train: dec_at model needs the `at` output");
    SEDT_REQUIRE(((uintptr_t)tape & 255) == 0, "forward_train: tape must be 256-byte aligned");
    ForwardOut o{out->hs, out->logits, out->boxes, out->at, out->memory, nullptr, nullptr, nullptr};
    return m->impl->forward_train(x, mask, B, T, F, tape, (size_t)tape_bytes, o, dropout, (unsigned long long)seed, (cudaStream_t)stream);
}

int sedt_backward(sedt_model* m, const void* const* weights, const float* x, const uint8_t* mask, int B, int T, int F, void* tape,
                  int64_t tape_bytes, void* workspace, int64_t workspace_bytes, const float* d_logits, const float* d_boxes,
                  const float* d_at, float* grads, int train_backbone, float dropout, void* stream)
{
    SEDT_REQUIRE(m != nullptr && weights != nullptr && x != nullptr && tape != nullptr && workspace != nullptr && grads != nullptr,
                 "backward: null argument");
    SEDT_REQUIRE(((uintptr_t)tape & 255) == 0 && ((uintptr_t)workspace & 255) == 0 && ((uintptr_t)grads & 255) == 0,
                 "backward: tape, workspace and grads must be 256-byte aligned");
    return m->impl->backward(weights, x, mask, B, T, F, tape, (size_t)tape_bytes, workspace, (size_t)workspace_bytes, d_logits,
                             d_boxes, d_at, grads, train_backbone, dropout, (cudaStream_t)stream);
}

int sedt_matcher(const float* logits, const float* boxes, const int64_t* tgt_labels, const float* tgt_boxes,
                 const int32_t* offsets, int B, int Q, int C1, int kmax, float cost_class, float cost_bbox, float cost_giou,
                 float* cost_out, int ld_cost, int64_t* rows, int64_t* cols, int32_t* counts, int32_t* status, void* stream)
{
    SEDT_REQUIRE(B == 0 || (logits && boxes && offsets && rows && cols && counts && status), "matcher: null argument");
    SEDT_REQUIRE(cost_out == nullptr || ld_cost >= kmax, "matcher: ld_cost=%d < kmax=%d", ld_cost, kmax);
    return launch_matcher(logits, boxes, tgt_labels, tgt_boxes, offsets, B, Q, C1, kmax, cost_class, cost_bbox, cost_giou,
                          nullptr, 0, cost_out, ld_cost, rows, cols, counts, status, 1, (cudaStream_t)stream);
}

int sedt_matcher_ex(const float* logits, const float* boxes, const int64_t* tgt_labels, const float* tgt_boxes,
                    const int32_t* offsets, int B, int Q, int C1, int kmax, float cost_class, float cost_bbox, float cost_giou,
                    int fl, float alpha_fl, float gamma_fl, float* cost_out, int ld_cost, int64_t* rows, int64_t* cols,
                    int32_t* counts, int32_t* status, float* lmin, int64_t* largmin, void* stream)
{
    SEDT_REQUIRE(B == 0 || (logits && boxes && offsets && rows && cols && counts && status), "matcher: null argument");
    SEDT_REQUIRE(cost_out == nullptr || ld_cost >= kmax, "matcher: ld_cost=%d < kmax=%d", ld_cost, kmax);
    return launch_matcher(logits, boxes, tgt_labels, tgt_boxes, offsets, B, Q, C1, kmax, cost_class, cost_bbox, cost_giou,
                          nullptr, 0, cost_out, ld_cost, rows, cols, counts, status, 1, (cudaStream_t)stream, fl, alpha_fl, gamma_fl,
                          lmin, largmin);
}

int sedt_lsap(const float* cost, int ld_cost, const int32_t* offsets, int B, int Q, int kmax, int64_t* rows, int64_t* cols,
              int32_t* counts, int32_t* status, void* stream)
{
    SEDT_REQUIRE(B == 0 || (cost && offsets && rows && cols && counts && status), "lsap: null argument");
    SEDT_REQUIRE(ld_cost >= kmax, "lsap: ld_cost=%d < kmax=%d", ld_cost, kmax);
    return launch_matcher(nullptr, nullptr, nullptr, nullptr, offsets, B, Q, 1, kmax, 0.f, 0.f, 0.f, cost, ld_cost, nullptr, 0,
                          rows, cols, counts, status, 1, (cudaStream_t)stream);
}

int sedt_set_criterion(const float* logits, const float* boxes, const float* at, const int64_t* tgt_labels, const float* tgt_boxes,
                       const int32_t* offsets, const float* n_tgt, const int64_t* wl_labels, const int32_t* wl_offsets,
                       int L, int B, int Bs, int Bw, int Q, int C1, int kmax, float cost_class, float cost_bbox, float cost_giou,
                       float eos_coef, float num_boxes, int64_t* rows, int64_t* cols, int32_t* status, float* partials,
                       float* losses, float* g_logits, float* g_l1, float* g_giou, float* g_at, void* stream)
{
    return launch_set_criterion(logits, boxes, at, tgt_labels, tgt_boxes, offsets, n_tgt, wl_labels, wl_offsets, L, B, Bs, Bw, Q, C1,
                                kmax, cost_class, cost_bbox, cost_giou, eos_coef, num_boxes, rows, cols, status, partials, losses,
                                g_logits, g_l1, g_giou, g_at, (cudaStream_t)stream);
}

int sedt_decode_events(const float* logits, const float* boxes, const float* target_sizes, const float* audio_tags, int B, int Q,
                       int C1, int at_m, float fuse_threshold, int is_semi, float score_threshold, float min_duration,
                       float* scores, int64_t* labels, float* boxes_se, int32_t* ev_class, float* ev_onset, float* ev_offset,
                       float* ev_score, int32_t* ev_count, void* stream)
{
    return launch_decode_events(logits, boxes, target_sizes, audio_tags, B, Q, C1, at_m, fuse_threshold, is_semi, score_threshold,
                                min_duration, scores, labels, boxes_se, ev_class, ev_onset, ev_offset, ev_score, ev_count,
                                (cudaStream_t)stream);
}

int sedt_pseudo_labels(const float* logits, const float* boxes, const float* audio_tags, const float* class_threshold, int B, int Q,
                       int C1, float min_width, int del_overlap, int64_t* labels, float* boxes_out, float* scores,
                       int32_t* counts, void* stream)
{
    return launch_pseudo_labels(logits, boxes, audio_tags, class_threshold, B, Q, C1, min_width, del_overlap, labels, boxes_out,
                                scores, counts, (cudaStream_t)stream);
}

int sedt_op_ffn(const void* x, const void* w1, const float* b1, const void* w2, const float* b2, const float* residual, float* out,
                int64_t M, int ff, void* stream)
{
    SEDT_REQUIRE(x && w1 && b1 && w2 && b2 && residual && out, "op_ffn: null argument");
    if (!ffn_fused_supported(256, ff, M, x, w1, w2, residual, out, 256, 256)) { set_error("op_ffn: unsupported shape"); return SEDT_ERR_UNSUPPORTED; }
    return launch_ffn_fused(x, w1, b1, w2, b2, residual, 256, out, 256, M, ff, (cudaStream_t)stream);
}

int sedt_op_bneck_tail(const void* h1, const void* w2, const float* bias2, const void* w3, const float* bias3, const void* residual,
                       void* out, int B, int H, int W, void* stream)
{
    SEDT_REQUIRE(h1 && w2 && bias2 && w3 && bias3 && residual && out, "op_bneck_tail: null argument");
    ConvGemm g2, g3;
    g2.in = h1; g2.w = w2; g2.bias = bias2; g2.out = nullptr; g2.in_dt = g2.out_dt = DT_BF16;
    g2.B = B; g2.H = g2.Ho = H; g2.W = g2.Wo = W; g2.Cin = g2.lda = 64; g2.Cout = g2.ldc = g2.ld_res = 64;
    g2.R = g2.S = 3; g2.stride = 1; g2.dil = 1; g2.pad = 1; g2.relu = 1;
    g2.out = out;                                   // (placeholder: alignment checks only, never written)
    g3.in = out; g3.w = w3; g3.bias = bias3; g3.residual = residual; g3.out = out; g3.in_dt = g3.out_dt = DT_BF16;
    g3.B = B; g3.H = g3.Ho = H; g3.W = g3.Wo = W; g3.Cin = g3.lda = 64; g3.Cout = g3.ldc = g3.ld_res = 256; g3.relu = 1;
    SEDT_REQUIRE(bneck_tail_supported(g2, g3), "op_bneck_tail: unsupported shape (H x W = %d x %d)", H, W);
    return launch_bneck_tail(g2, g3, (cudaStream_t)stream);
}

int sedt_op_enc_attn(const void* na, const void* nap, const void* w_in, const float* b_in, const void* w_out, const float* b_out,
                     const uint8_t* kpm, float* x, int B, int S, const float* ln_g, const float* ln_b, void* ln_out, void* stream)
{
    SEDT_REQUIRE(na && nap && w_in && b_in && w_out && b_out && x, "op_enc_attn: null argument");
    return launch_enc_attn_fused(na, nap, w_in, b_in, w_out, b_out, kpm, x, B, S, 0.17677669529663687f, (cudaStream_t)stream, ln_g, ln_b,
                                 ln_out);
}

int sedt_prepare_clips(const float* raw, const int64_t* offsets, const double* mean, const double* std, float* out, int B, int frames,
                       int F, int apply_log, void* stream)
{
    return launch_prepare_clips(raw, offsets, mean, std, out, B, frames, F, apply_log, (cudaStream_t)stream);
}

static_assert(sizeof(sedt_augment_params) == sizeof(AugmentParams) && sizeof(sedt_mix_row) == sizeof(MixRow), "augment structs out of step");

int sedt_augment_clips(float* x, const sedt_augment_params* params, int B, int T, int F, float* scratch, void* stream)
{
    return launch_augment_clips(x, (const AugmentParams*)params, B, T, F, scratch, (cudaStream_t)stream);
}

int sedt_mix_rows(const float* x, float* out, const sedt_mix_row* rows, int n_out, int64_t row_elems, void* stream)
{
    return launch_mix_rows(x, out, (const MixRow*)rows, n_out, row_elems, (cudaStream_t)stream);
}

int sedt_query_patches(const float* x, const int32_t* bounds, float* out, int B, int P, int T, int F, int fixed_patch_size, void* stream)
{
    return launch_query_patches(x, bounds, out, B, P, T, F, fixed_patch_size, (cudaStream_t)stream);
}

int sedt_optim_chunk_elems(void) { return optim_chunk_elems(); }

int sedt_grad_norm(const sedt_optim_tensor* tensors, const int32_t* chunks, int nchunks, float* partials, float* norm_out, void* stream)
{
    return launch_grad_norm(tensors, chunks, nchunks, partials, norm_out, (cudaStream_t)stream);
}

int sedt_clip_grads(const sedt_optim_tensor* tensors, const int32_t* chunks, int nchunks, const float* norm, float max_norm,
                    void* stream)
{
    return launch_grad_scale(tensors, chunks, nchunks, norm, max_norm, (cudaStream_t)stream);
}

int sedt_adamw_step(const sedt_optim_tensor* tensors, const int32_t* chunks, int nchunks, const sedt_adamw_group* groups,
                    int ngroups, const float* norm, float max_norm, void* stream)
{
    static_assert(sizeof(sedt_adamw_group) == 32 && sizeof(sedt_optim_tensor) == 48, "optimizer table layout");
    return launch_adamw(tensors, chunks, nchunks, (const float*)groups, ngroups, norm, max_norm, (cudaStream_t)stream);
}

int sedt_op_conv(const sedt_conv_desc* d, int engine, void* stream)
{
    SEDT_REQUIRE(d != nullptr, "op_conv: null descriptor");
    ConvGemm g = to_gemm(d);
    if (engine == 0) return launch_conv_simt(g, (cudaStream_t)stream);
    SEDT_TRY(tc_init());
    if (engine == 3) {
        if (!conv_tc_ws_supported(g)) { set_error("op_conv: shape not supported by the weight-stationary kernel"); return SEDT_ERR_UNSUPPORTED; }
        return launch_conv_tc_ws(g, (cudaStream_t)stream);
    }
    if (engine == 2) {
        if (!conv_tc_2sm_supported(g)) { set_error("op_conv: shape not supported by the cta_group::2 kernel"); return SEDT_ERR_UNSUPPORTED; }
        return launch_conv_tc_2sm(g, (cudaStream_t)stream);
    }
    if (!conv_tc_supported(g)) {
        set_error("op_conv: shape not supported by the tcgen05 kernel");
        return SEDT_ERR_UNSUPPORTED;
    }
    return launch_conv_tc(g, (cudaStream_t)stream);
}

int sedt_op_conv_wgrad(const void* x, const void* dy, float* dw, int B, int H, int W, int Cin, int Cout, int k, int stride,
                       int dil, int pad, void* stream)
{
    SEDT_REQUIRE(x != nullptr && dy != nullptr && dw != nullptr, "op_conv_wgrad: null argument");
    WgradGemm g;
    g.x = x; g.dy = dy; g.dw = dw;
    g.B = B; g.H = H; g.W = W; g.Cin = Cin; g.lda = Cin; g.Cout = Cout; g.ldy = Cout;
    g.R = g.S = k; g.stride = stride; g.dil = dil; g.pad = pad;
    g.Ho = (H + 2 * pad - dil * (k - 1) - 1) / stride + 1;
    g.Wo = (W + 2 * pad - dil * (k - 1) - 1) / stride + 1;
    if (!conv_wgrad_tc_supported(g)) { set_error("op_conv_wgrad: shape not supported by the tcgen05 kernel"); return SEDT_ERR_UNSUPPORTED; }
    return launch_conv_wgrad_tc(g, (cudaStream_t)stream);
}

int sedt_op_repack_dgrad(const float* w_oihw, const float* scale, void* out, int dtype, int Cout, int Cin, int R, int S, void* stream)
{
    return launch_repack_dgrad(w_oihw, scale, out, dtype, Cout, Cout, Cin, R, S, (cudaStream_t)stream);
}

int sedt_op_upsample2(const void* dy, void* u, int B, int H, int W, int Ho, int Wo, int C, void* stream)
{
    return launch_upsample2(dy, u, B, H, W, Ho, Wo, C, (cudaStream_t)stream);
}

int sedt_op_relu_mask(const void* act, const void* g1, const void* g2, void* out, int64_t n, void* stream)
{
    return launch_relu_mask(act, g1, g2, out, n, (cudaStream_t)stream);
}

int sedt_op_colsum(const void* in, int dtype, int64_t ld, float* out, int64_t M, int N, void* stream)
{
    return launch_colsum(in, dtype, ld, out, M, N, (cudaStream_t)stream);
}

int sedt_op_layernorm_bwd(const float* x, const float* gamma, const void* g1, const void* g2, const float* g3, const float* dres,
                          float* dx, float* dgamma, float* dbeta, int64_t rows, void* stream)
{
    return launch_layernorm_bwd(x, gamma, g1, g2, g3, dres, dx, dgamma, dbeta, rows, (cudaStream_t)stream);
}

int sedt_op_attention_bwd(const void* Q, int ldq, const void* K, int ldk, const void* V, int ldv, const void* dO, int ldo,
                          void* dQ, int lddq, void* dK, int lddk, void* dV, int lddv, const uint8_t* kpm, const float* amask,
                          int B, int nheads, int Lq, int Lk, float scale, int engine, void* stream)
{
    if (engine == 0)
        return launch_attention_bwd(Q, ldq, K, ldk, V, ldv, dO, ldo, dQ, lddq, dK, lddk, dV, lddv, kpm, amask, B, nheads, Lq, Lk,
                                    scale, (cudaStream_t)stream);
    if (!attention_bwd_tc_supported(Q, ldq, K, ldk, V, ldv, dO, ldo, dQ, lddq, dK, lddk, dV, lddv, Lq, Lk)) {
        set_error("op_attention_bwd: shape / alignment not supported by the tcgen05 kernel");
        return SEDT_ERR_UNSUPPORTED;
    }
    return launch_attention_bwd_tc(Q, ldq, K, ldk, V, ldv, dO, ldo, dQ, lddq, dK, lddk, dV, lddv, kpm, amask, B, nheads, Lq, Lk,
                                   scale, nullptr, (cudaStream_t)stream);
}

int sedt_op_dropout_mask(uint8_t* out, int64_t n, uint64_t seed, uint64_t step, uint32_t site, float p, void* stream)
{
    SEDT_REQUIRE(out != nullptr || n == 0, "op_dropout_mask: null output");
    return launch_dropout_mask(out, n, (unsigned long long)seed, (unsigned long long)step, site, p, (cudaStream_t)stream);
}

int sedt_op_conv_tc_supported(const sedt_conv_desc* d) { return d != nullptr && conv_tc_supported(to_gemm(d)) ? 1 : 0; }

int sedt_op_repack_conv(const float* w_oihw, void* out, int dtype, int Cout, int Cin, int R, int S, void* stream)
{
    return launch_repack_conv(w_oihw, nullptr, out, dtype, Cout, Cin, R, S, (cudaStream_t)stream);
}

int sedt_op_cast(const float* in, void* out, int dtype, int64_t n, void* stream)
{
    return launch_cast(in, out, dtype, n, (cudaStream_t)stream);
}

int sedt_op_stem(const float* x, const float* conv0_w, const float* conv0_b, const float* conv1_w, const float* bn_w,
                 const float* bn_b, const float* bn_mean, const float* bn_var, void* scratch, void* out, int out_dtype, int B,
                 int T, int F, void* stream)
{
    SEDT_REQUIRE(scratch != nullptr && ((uintptr_t)scratch & 255) == 0, "op_stem: scratch must be 256-byte aligned");
    float* weff = (float*)scratch;              // 49*64
    float* sat = weff + 49 * 64 + 64;           // 64*64 (offset keeps 16-byte alignment)
    float* scale = sat + 64 * 64;
    float* bias = scale + 64;
    cudaStream_t s = (cudaStream_t)stream;
    SEDT_TRY(launch_stem_pack(conv0_w, conv0_b, conv1_w, weff, sat, s));
    SEDT_TRY(launch_bn_fold(bn_w, bn_b, bn_mean, bn_var, scale, bias, 64, s));
    StemWeights w{weff, sat, scale, bias};
    return launch_stem(x, w, out, out_dtype, B, T, F, s);
}

int sedt_op_stem_tc(const float* x, const float* conv0_w, const float* conv0_b, const float* conv1_w, const float* bn_w,
                    const float* bn_b, const float* bn_mean, const float* bn_var, void* scratch, void* out, int B, int T, int F,
                    void* stream)
{
    SEDT_REQUIRE(scratch != nullptr && ((uintptr_t)scratch & 255) == 0, "op_stem_tc: scratch must be 256-byte aligned");
    uint8_t* wtc = (uint8_t*)scratch;                 // 16 KiB
    float* scale = (float*)(wtc + 16384);
    float* bias = scale + 64;
    float* weff = bias + 64;                          // 49 x 64
    float* sat = weff + 49 * 64;                      // 8 x 8 x 64
    cudaStream_t s = (cudaStream_t)stream;
    SEDT_TRY(launch_bn_fold(bn_w, bn_b, bn_mean, bn_var, scale, bias, 64, s));
    SEDT_TRY(launch_stem_pack(conv0_w, conv0_b, conv1_w, weff, sat, s));
    SEDT_TRY(launch_stem_tc_pack(conv0_w, conv0_b, conv1_w, scale, wtc, s));
    return launch_stem_tc(x, wtc, bias, scale, sat, out, B, T, F, s);
}

int sedt_op_layernorm(const float* x, const float* gamma, const float* beta, const float* pos, int64_t pos_rows, void* y,
                      void* ypos, float* y32, int dtype, int64_t rows, void* stream)
{
    return launch_layernorm(x, gamma, beta, pos, pos_rows < 1 ? 1 : pos_rows, y, ypos, y32, dtype, rows, (cudaStream_t)stream);
}

int sedt_op_attention(const void* Q, int ldq, const void* K, int ldk, const void* V, int ldv, void* O, int ldo, int dtype,
                      const uint8_t* key_padding_mask, const float* attn_mask, int B, int nheads, int Lq, int Lk, float scale,
                      void* stream)
{
    return launch_attention(Q, ldq, K, ldk, V, ldv, O, ldo, dtype, key_padding_mask, attn_mask, B, nheads, Lq, Lk, scale,
                            (cudaStream_t)stream);
}

int sedt_op_pos_table(const uint8_t* mask, uint8_t* mask_ds, float* pos, int B, int T, int F, int H, int W, void* stream)
{
    cudaStream_t s = (cudaStream_t)stream;
    if (mask == nullptr) return launch_pos_table(nullptr, pos, 1, H, W, s);
    SEDT_REQUIRE(mask_ds != nullptr, "op_pos_table: mask_ds scratch required with a mask");
    SEDT_TRY(launch_mask_downsample(mask, mask_ds, B, T, F, H, W, s));
    return launch_pos_table(mask_ds, pos, B, H, W, s);
}

}  // extern "C"
