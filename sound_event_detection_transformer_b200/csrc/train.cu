// Training step of the SEDT hot path (bf16 tier, pre-norm, supervised model): a forward that keeps every
// activation the backward pass needs (the "tape"), and the backward pass itself.
//
//   forward_train : the launch sequence of Model::forward with one buffer per activation instead of the
//                   ping-pong arena (sedt/sedt.py:64-123 in train mode with dropout = 0).
//   backward      : gradients of every trainable parameter from d(pred_logits), d(pred_boxes), d(at) of all
//                   decoder layers (the set loss itself, sedt/sedt.py:309-352, stays a torch expression on
//                   these small tensors).  Data gradients run through the forward implicit-GEMM kernels with
//                   re-laid-out weights, weight gradients through gemm_wgrad.cu, the rest through backward.cu.
//                   Replaces autograd's traversal of sedt/backbone.py, sedt/transformer.py, sedt/sedt.py.
//
// Gradients are written in the reference's parameter layout into ONE flat fp32 buffer (offset of a
// state_dict entry = Model::grad_offset(slot)), which is also the bucket a data-parallel run all-reduces.
#include "model.h"
#include "dropout.cuh"
#include <algorithm>
#include <cmath>
#include <cstdlib>

namespace sedt {

namespace {
inline int conv_out_dim(int n, int k, int stride, int pad, int dil) { return (n + 2 * pad - dil * (k - 1) - 1) / stride + 1; }
}

struct Model::BlockTape { void *h1, *h2, *ds, *out; int H, W, Ho, Wo; };
struct Model::EncTape { float* x_in; void *na, *nap, *qk, *v, *ao; float* x_mid; void *n2, *h; float* x_out; };
struct Model::DecTape {
    float* t_in; void *da, *dap, *qk, *v, *ao1; float* t_mid1; void *dap2, *qb, *ao2; float* t_mid2; void *da3, *h; float* t_out;
};
struct Model::Tape {
    int H0, W0, H, W, S;
    unsigned long long* rng;    // {seed, step} of the dropout masks
    float* tmp32;               // GEMM output before dropout + residual add
    void* stem_out;
    uint8_t* stem_amax;         // [B, H0, W0, 64] arg-max of the stem's pooling windows (9 = dead ReLU)
    std::vector<BlockTape> blk;
    uint8_t* mask_ds; float* pos; int64_t pos_rows;
    float* x0;
    std::vector<EncTape> enc;
    void *mem, *mempos, *ck, *cv;
    std::vector<DecTape> dec;
    void* hs_t;                 // [D, B*Qall, d] bf16 decoder states after decoder.norm
    void *hh1; float* hh2;      // box MLP hidden layers (bf16 / fp32)
    float *cls_raw, *box_raw, *weak_raw, *boxes, *at;
    // SP-SEDT pretraining (frozen backbone: no block tape, the clip / patch backbones run through the eval launch sequence
    // in `scratch`; only the layer4 feature map `feat` is kept for input_proj's weight gradient)
    void* scratch; size_t scratch_bytes; const void* feat;
    float *gt, *pq, *qpos, *amask; uint8_t* keep; void* fh1;
    int Q;                      // queries per clip of this model (qall_, or num_queries for SP-SEDT)
};

struct Model::BwdBufs {
    void* wd;            // re-laid-out weights of the layer being differentiated (bf16): scratch for the unplanned sites
    char* wd_all;        // one slot per data-gradient site of the pass (filled by the batched re-layout)
    size_t wd_all_bytes;
    float* dw;           // weight gradient of a conv layer in the GEMM layout, before the OIHW permute
    float* vscale;       // [dim_feedforward] = 1 / (1 - dropout)
    // heads
    void *dcls, *dbox, *dweak, *dh2, *dh1, *hh2b; float* dhs32;
    // transformer (sized for max(rows, qrows))
    float *gA, *gB; void *g16, *dh, *dn_a, *dn_b, *dao, *dqk, *dv, *dq;
    void *dck, *dcv;
    // backbone
    void *G0, *G1, *gh2, *gh1, *up;
    // SP-SEDT: d(pred_feature) as a bf16 GEMM operand, d(feature_align hidden), per-clip d(query_pos), d(patch feature), bf16 copies
    void *dfeat, *dfh1, *dpq16, *gt16; float *dqpos32, *dpq32;
};

// Every buffer of the tape, in a fixed order: forward_train and backward derive the same pointers from the
// same base address.
void Model::tape_layout(int B, int T, int F, bool has_mask, int P, int PT, Arena& a, Tape& tp)
{
    const int d = cfg_.hidden_dim, ff = cfg_.dim_feedforward;
    const size_t es = 2;
    const bool sp = cfg_.self_sup != 0;
    tp.H0 = conv_out_dim(conv_out_dim(T, 7, 2, 3, 1), 3, 2, 1, 1);
    tp.W0 = conv_out_dim(conv_out_dim(F, 7, 2, 3, 1), 3, 2, 1, 1);
    tp.rng = (unsigned long long*)a.alloc(256);
    tp.blk.clear();
    tp.scratch = nullptr; tp.scratch_bytes = 0; tp.feat = nullptr;
    tp.gt = tp.pq = tp.qpos = tp.amask = nullptr; tp.keep = nullptr; tp.fh1 = nullptr;
    tp.Q = sp ? cfg_.num_queries : qall_;
    int h = tp.H0, w = tp.W0;
    if (!sp) {
        tp.stem_out = a.alloc((size_t)B * tp.H0 * tp.W0 * 64 * es);
        tp.stem_amax = (uint8_t*)a.alloc((size_t)B * tp.H0 * tp.W0 * 64);
        for (const Block& b : blocks_) {
            BlockTape bt{};
            bt.H = h; bt.W = w;
            bt.Ho = conv_out_dim(h, 3, b.c2.stride, b.c2.pad, b.c2.dil); bt.Wo = conv_out_dim(w, 3, b.c2.stride, b.c2.pad, b.c2.dil);
            bt.h1 = a.alloc((size_t)B * h * w * b.c1.cout * es);
            bt.h2 = a.alloc((size_t)B * bt.Ho * bt.Wo * b.c2.cout * es);
            bt.ds = b.has_ds ? a.alloc((size_t)B * bt.Ho * bt.Wo * b.ds.cout * es) : nullptr;
            bt.out = a.alloc((size_t)B * bt.Ho * bt.Wo * b.c3.cout * es);
            tp.blk.push_back(bt);
            h = bt.Ho; w = bt.Wo;
        }
    } else {
        tp.stem_out = nullptr; tp.stem_amax = nullptr;
        feature_shape(T, F, cfg_.dilation != 0, &h, &w);
        // scratch = the larger of the two eval backbones' workspaces; the clip backbone runs last, so its layer4 output stays valid
        Arena dry1(nullptr, 0), dry2(nullptr, 0);
        void* f = nullptr; int fh = 0, fw = 0;
        (void)backbone(nullptr, B * P, PT, F, dry1, &f, &fh, &fw, nullptr, true);
        (void)backbone(nullptr, B, T, F, dry2, &f, &fh, &fw, nullptr, true);
        tp.scratch_bytes = std::max(dry1.peak, dry2.peak) + 256;
        tp.scratch = a.alloc(tp.scratch_bytes);
        tp.gt = (float*)a.alloc((size_t)B * P * 2048 * 4);
        tp.pq = (float*)a.alloc((size_t)B * P * d * 4);
        tp.qpos = (float*)a.alloc((size_t)B * tp.Q * d * 4);
        tp.keep = (uint8_t*)a.alloc((size_t)B * tp.Q);
        tp.amask = (float*)a.alloc((size_t)tp.Q * tp.Q * 4);
    }
    tp.H = h; tp.W = w; tp.S = h * w;
    const int64_t rows = (int64_t)B * tp.S;
    tp.mask_ds = has_mask ? (uint8_t*)a.alloc((size_t)rows) : nullptr;
    tp.pos_rows = has_mask ? rows : tp.S;
    tp.pos = (float*)a.alloc((size_t)tp.pos_rows * d * 4);
    tp.x0 = (float*)a.alloc((size_t)rows * d * 4);
    tp.enc.clear();
    float* x = tp.x0;
    for (size_t l = 0; l < enc_.size(); ++l) {
        EncTape e{};
        e.x_in = x;
        e.na = a.alloc((size_t)rows * d * es); e.nap = a.alloc((size_t)rows * d * es);
        e.qk = a.alloc((size_t)rows * 2 * d * es); e.v = a.alloc((size_t)rows * d * es); e.ao = a.alloc((size_t)rows * d * es);
        e.x_mid = (float*)a.alloc((size_t)rows * d * 4);
        e.n2 = a.alloc((size_t)rows * d * es); e.h = a.alloc((size_t)rows * ff * es);
        e.x_out = (float*)a.alloc((size_t)rows * d * 4);
        tp.enc.push_back(e);
        x = e.x_out;
    }
    const size_t Dn = dec_.size();
    tp.mem = a.alloc((size_t)rows * d * es); tp.mempos = a.alloc((size_t)rows * d * es);
    tp.ck = a.alloc((size_t)rows * Dn * d * es); tp.cv = a.alloc((size_t)rows * Dn * d * es);
    const int64_t qrows = (int64_t)B * tp.Q;
    tp.dec.clear();
    float* t = (float*)a.alloc((size_t)qrows * d * 4);          // tgt = 0
    for (size_t l = 0; l < Dn; ++l) {
        DecTape e{};
        e.t_in = t;
        e.da = a.alloc((size_t)qrows * d * es); e.dap = a.alloc((size_t)qrows * d * es);
        e.qk = a.alloc((size_t)qrows * 2 * d * es); e.v = a.alloc((size_t)qrows * d * es); e.ao1 = a.alloc((size_t)qrows * d * es);
        e.t_mid1 = (float*)a.alloc((size_t)qrows * d * 4);
        e.dap2 = a.alloc((size_t)qrows * d * es); e.qb = a.alloc((size_t)qrows * d * es); e.ao2 = a.alloc((size_t)qrows * d * es);
        e.t_mid2 = (float*)a.alloc((size_t)qrows * d * 4);
        e.da3 = a.alloc((size_t)qrows * d * es); e.h = a.alloc((size_t)qrows * ff * es);
        e.t_out = (float*)a.alloc((size_t)qrows * d * 4);
        tp.dec.push_back(e);
        t = e.t_out;
    }
    const int64_t hrows = (int64_t)Dn * qrows;
    const int ncls = sp ? 1 : cfg_.num_classes, C1 = ncls + 1;
    tp.hs_t = a.alloc((size_t)hrows * d * es);
    tp.hh1 = a.alloc((size_t)hrows * d * es);
    tp.hh2 = (float*)a.alloc((size_t)hrows * d * 4);
    tp.cls_raw = (float*)a.alloc((size_t)hrows * C1 * 4);
    tp.box_raw = (float*)a.alloc((size_t)hrows * 2 * 4);
    tp.weak_raw = cfg_.dec_at ? (float*)a.alloc((size_t)B * ncls * 4) : nullptr;
    tp.tmp32 = (float*)a.alloc((size_t)std::max(rows, qrows) * d * 4);
    tp.boxes = (float*)a.alloc((size_t)Dn * B * cfg_.num_queries * 2 * 4);
    tp.at = cfg_.dec_at ? (float*)a.alloc((size_t)B * ncls * 4) : nullptr;
    if (sp && cfg_.feature_recon) tp.fh1 = a.alloc((size_t)hrows * d * es);
}

int Model::check_train_config() const
{
    SEDT_REQUIRE(cfg_.precision == 1 && cfg_.use_tensor_cores, "training kernels exist for the bf16 tcgen05 tier only");
    SEDT_REQUIRE(cfg_.pre_norm, "training kernels implement the pre-norm layers only (transformer.py:192-204, :263-284)");
    SEDT_REQUIRE(!(cfg_.self_sup && cfg_.dec_at), "SP-SEDT has no audio query (sedt/spsedt.py:59,72)");
    SEDT_REQUIRE(cfg_.hidden_dim == 256 && cfg_.nheads == 8, "kernels are built for hidden_dim 256 / 8 heads");
    SEDT_REQUIRE(cfg_.num_classes + 1 <= 128, "head gradients are padded to 128 columns: num_classes <= 127");
    return SEDT_OK;
}

int64_t Model::tape_bytes(int B, int T, int F, bool has_mask, int P, int PT)
{
    Arena a(nullptr, 0);
    Tape tp;
    tape_layout(B, T, F, has_mask, P, PT, a, tp);
    return (int64_t)a.peak + 256;
}

// dropout sites: encoder layer l -> 8*l + {0 attention weights, 1 after out_proj, 2 FFN hidden, 3 after linear2};
// decoder layer l -> 1024 + 8*l + {0 self-attention weights, 1 after its out_proj, 2 cross-attention weights, 3 after its
// out_proj, 4 FFN hidden, 5 after linear2}   (transformer.py:160-175, :220-240)
static inline uint32_t enc_site(int l, int k) { return (uint32_t)(8 * l + k); }
static inline uint32_t dec_site(int l, int k) { return (uint32_t)(1024 + 8 * l + k); }

int Model::forward_train(const float* x, const uint8_t* mask, int B, int T, int F, void* tape, size_t tape_bytes_,
                         const ForwardOut& out, float dropout, unsigned long long seed, cudaStream_t s, const SpTrain* sp)
{
    const bool spm = cfg_.self_sup != 0;
    const SpTrain none;
    if (sp == nullptr) sp = &none;
    SEDT_REQUIRE(!spm || (sp->patches != nullptr && sp->query_keep != nullptr), "forward_train: SP-SEDT needs patches and the query-drop mask");
    SEDT_REQUIRE(!spm || sp->P == cfg_.num_patches, "forward_train: the training branch uses exactly num_patches = %d patches per clip (sedt/spsedt.py:63-69), got %d",
                 cfg_.num_patches, sp->P);
    SEDT_REQUIRE(!spm || !cfg_.feature_recon || (out.pred_feature != nullptr && out.gt_feature != nullptr), "forward_train: pred_feature / gt_feature outputs are required");
    SEDT_REQUIRE(dropout >= 0.f && dropout < 1.f, "forward_train: dropout=%f", dropout);
    SEDT_TRY(check_train_config());
    SEDT_REQUIRE(packed_ != nullptr, "forward_train: sedt_model_pack has not been called");
    SEDT_REQUIRE(F == 64, "stem: the fused stem kernel needs 64 mel bins, got F=%d", F);
    Arena a(tape, tape_bytes_);
    Tape tp;
    tape_layout(B, T, F, mask != nullptr, sp->P, sp->PT, a, tp);
    if (a.overflow) { set_error("forward_train: tape too small (%zu bytes needed, %zu given)", a.peak, a.cap); return SEDT_ERR_WORKSPACE; }
    const int d = cfg_.hidden_dim, ff = cfg_.dim_feedforward, dt = DT_BF16;
    const size_t es = 2;
    auto P_ = [&](size_t off) { return (const float*)(packed_ + off); };
    auto LN = [&](const Norm& n, const float* xin, const float* pos, int64_t pos_rows, void* y, void* ypos, float* y32, int64_t rows) {
        return launch_layernorm(xin, P_(n.off_g), P_(n.off_b), pos, pos_rows, y, ypos, y32, dt, rows, s);
    };
    const float scale = (float)std::sqrt(1.0 / (double)(d / cfg_.nheads));
    const bool drop = dropout > 0.f;
    if (drop) {
        // the {seed, step} pair lives in the tape: initialised once per tape (outside any graph capture, which only
        // records the per-step increment), so that every replayed step draws fresh masks
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        SEDT_CHECK_CUDA(cudaStreamIsCapturing(s, &cap));
        // several tapes may be alive at once (two forwards before one backward, engine.py:134-170): each keeps its own state
        auto it = rng_tapes_.find(tape);
        if (cap == cudaStreamCaptureStatusNone && (it == rng_tapes_.end() || it->second != seed)) {
            SEDT_TRY(launch_rng_init(tp.rng, seed, s));
            rng_tapes_[tape] = seed;
            it = rng_tapes_.find(tape);
        }
        SEDT_REQUIRE(it != rng_tapes_.end(), "forward_train: dropout RNG state of this tape was not initialised before graph capture");
        SEDT_TRY(launch_rng_step(tp.rng, s));
    }
    auto site = [&](uint32_t id) { return make_drop_site(tp.rng, id, dropout); };
    // y = resid + dropout(x W^T + b)
    auto linear_drop_add = [&](const Linear& L, const void* in, int lda, int64_t R, const float* resid, float* y, uint32_t id) -> int {
        if (!drop) return linear(L, 0, L.out, in, dt, lda, R, resid, y, DT_F32, L.out, 0, s, false);
        SEDT_TRY(linear(L, 0, L.out, in, dt, lda, R, nullptr, tp.tmp32, DT_F32, L.out, 0, s, false));
        return launch_dropout_add(tp.tmp32, resid, y, R * L.out, site(id), s);
    };
    auto attention = [&](const void* Qp, int ldq, const void* Kp, int ldk, const void* Vp, int ldv, void* Op, const uint8_t* kpm,
                         int Lq, int Lk, uint32_t id, const float* am = nullptr) -> int {
        if (!drop) return launch_attention(Qp, ldq, Kp, ldk, Vp, ldv, Op, d, dt, kpm, am, B, cfg_.nheads, Lq, Lk, scale, s);
        return launch_attention_tc_drop(Qp, ldq, Kp, ldk, Vp, ldv, Op, d, kpm, am, B, cfg_.nheads, Lq, Lk, scale, site(id), s);
    };

    // ---- backbone
    const void* feat = nullptr;
    if (!spm) {
        SEDT_TRY(launch_stem_tc(x, packed_ + off_stem_wtc, P_(off_stem_bias), P_(off_stem_scale), P_(off_sat), tp.stem_out, B, T, F, s,
                                tp.stem_amax));
        const void* cur = tp.stem_out;
        for (size_t i = 0; i < blocks_.size(); ++i) {
            const Block& b = blocks_[i];
            const BlockTape& bt = tp.blk[i];
            int ho, wo;
            SEDT_TRY(conv(b.c1, cur, B, bt.H, bt.W, nullptr, bt.h1, &ho, &wo, s, false));
            SEDT_TRY(conv(b.c2, bt.h1, B, bt.H, bt.W, nullptr, bt.h2, &ho, &wo, s, false));
            const void* idn = cur;
            if (b.has_ds) { SEDT_TRY(conv(b.ds, cur, B, bt.H, bt.W, nullptr, bt.ds, &ho, &wo, s, false)); idn = bt.ds; }
            SEDT_TRY(conv(b.c3, bt.h2, B, bt.Ho, bt.Wo, idn, bt.out, &ho, &wo, s, false));
            cur = bt.out;
        }
        feat = cur;
    } else {
        // frozen backbone (train_spsedt.py:50): patches first (avgpool -> gt_feature -> patch2query), then the clips, whose
        // layer4 output stays in the scratch arena until backward (spsedt.py:40-57)
        const int qpp = cfg_.num_queries / cfg_.num_patches;
        {
            Arena ws(tp.scratch, tp.scratch_bytes);
            void* pfeat = nullptr; int ph = 0, pw = 0;
            SEDT_TRY(backbone(sp->patches, B * sp->P, sp->PT, F, ws, &pfeat, &ph, &pw, s, false));
            SEDT_REQUIRE(!ws.overflow, "forward_train: patch backbone scratch too small");
            SEDT_TRY(launch_avgpool(pfeat, dt, tp.gt, B * sp->P, ph * pw, 2048, s));
            SEDT_TRY(linear(patch2query_, 0, d, tp.gt, DT_F32, 2048, (int64_t)B * sp->P, nullptr, tp.pq, DT_F32, d, 0, s, false));
        }
        SEDT_CHECK_CUDA(cudaMemcpyAsync(tp.keep, sp->query_keep, (size_t)B * tp.Q, cudaMemcpyDeviceToDevice, s));
        SEDT_TRY(launch_patch_query(tp.pq, P_(off_query_embed), tp.qpos, B, sp->P, qpp, 0, s, tp.keep, 2.f));
        SEDT_TRY(launch_blockdiag_mask(tp.amask, tp.Q, qpp, s));
        if (out.gt_feature != nullptr)
            SEDT_CHECK_CUDA(cudaMemcpyAsync(out.gt_feature, tp.gt, (size_t)B * sp->P * 2048 * 4, cudaMemcpyDeviceToDevice, s));
        Arena ws(tp.scratch, tp.scratch_bytes);
        void* cfeat = nullptr; int ch = 0, cw = 0;
        SEDT_TRY(backbone(x, B, T, F, ws, &cfeat, &ch, &cw, s, false));
        SEDT_REQUIRE(!ws.overflow && ch == tp.H && cw == tp.W, "forward_train: clip backbone scratch / shape mismatch");
        feat = cfeat;
    }
    const int S = tp.S;
    const int64_t rows = (int64_t)B * S;

    // ---- mask, position table, input_proj
    if (mask != nullptr) {
        SEDT_TRY(launch_mask_downsample(mask, tp.mask_ds, B, T, F, tp.H, tp.W, s));
        SEDT_TRY(launch_pos_table(tp.mask_ds, tp.pos, B, tp.H, tp.W, s));
    } else {
        SEDT_TRY(launch_pos_table(nullptr, tp.pos, 1, tp.H, tp.W, s));
    }
    SEDT_TRY(linear(input_proj_, 0, d, feat, dt, 2048, rows, nullptr, tp.x0, DT_F32, d, 0, s, false));

    // ---- encoder
    for (size_t l = 0; l < enc_.size(); ++l) {
        const EncLayer& e = enc_[l];
        const EncTape& t = tp.enc[l];
        SEDT_TRY(LN(e.n1, t.x_in, tp.pos, tp.pos_rows, t.na, t.nap, nullptr, rows));
        SEDT_TRY(linear(e.attn.in_proj, 2 * d, d, t.na, dt, d, rows, nullptr, t.v, dt, d, 0, s, false));
        SEDT_TRY(linear(e.attn.in_proj, 0, 2 * d, t.nap, dt, d, rows, nullptr, t.qk, dt, 2 * d, 0, s, false));
        SEDT_TRY(attention(t.qk, 2 * d, (const char*)t.qk + d * es, 2 * d, t.v, d, t.ao, tp.mask_ds, S, S, enc_site((int)l, 0)));
        SEDT_TRY(linear_drop_add(e.attn.out_proj, t.ao, d, rows, t.x_in, t.x_mid, enc_site((int)l, 1)));
        SEDT_TRY(LN(e.n2, t.x_mid, nullptr, 1, t.n2, nullptr, nullptr, rows));
        SEDT_TRY(linear(e.lin1, 0, ff, t.n2, dt, d, rows, nullptr, t.h, dt, ff, 1, s, false));
        if (drop) SEDT_TRY(launch_dropout_bf16(t.h, rows * ff, site(enc_site((int)l, 2)), s));
        SEDT_TRY(linear_drop_add(e.lin2, t.h, ff, rows, t.x_mid, t.x_out, enc_site((int)l, 3)));
    }
    const float* xe = enc_.empty() ? tp.x0 : tp.enc.back().x_out;
    SEDT_TRY(LN(enc_norm_, xe, tp.pos, tp.pos_rows, tp.mem, tp.mempos, out.memory, rows));

    // ---- cross-attention keys / values of every decoder layer
    const int Dn = (int)dec_.size();
    {
        ConvGemm g;
        g.in_dt = g.out_dt = dt; g.B = (int)rows; g.H = g.W = g.Ho = g.Wo = 1; g.Cin = d; g.lda = d;
        g.Cout = Dn * d; g.ldc = Dn * d; g.ld_res = Dn * d;
        g.in = tp.mempos; g.w = packed_ + off_ck_w; g.bias = P_(off_ck_b); g.out = tp.ck;
        SEDT_TRY(gemm(g, s, false));
        g.in = tp.mem; g.w = packed_ + off_cv_w; g.bias = P_(off_cv_b); g.out = tp.cv;
        SEDT_TRY(gemm(g, s, false));
    }

    // ---- decoder
    const int Qall = tp.Q;
    const int64_t qrows = (int64_t)B * Qall;
    const float* qpos = spm ? tp.qpos : P_(off_query_embed);
    const int64_t qpos_rows = spm ? qrows : Qall;
    const float* amask = spm ? tp.amask : nullptr;                       // block-diagonal decoder mask (spsedt.py:27-32)
    SEDT_TRY(launch_fill_zero(tp.dec[0].t_in, (size_t)qrows * d * 4, s));
    for (int l = 0; l < Dn; ++l) {
        const DecLayer& e = dec_[l];
        const DecTape& t = tp.dec[l];
        SEDT_TRY(LN(e.n1, t.t_in, qpos, qpos_rows, t.da, t.dap, nullptr, qrows));
        SEDT_TRY(linear(e.self_attn.in_proj, 2 * d, d, t.da, dt, d, qrows, nullptr, t.v, dt, d, 0, s, false));
        SEDT_TRY(linear(e.self_attn.in_proj, 0, 2 * d, t.dap, dt, d, qrows, nullptr, t.qk, dt, 2 * d, 0, s, false));
        SEDT_TRY(attention(t.qk, 2 * d, (const char*)t.qk + d * es, 2 * d, t.v, d, t.ao1, nullptr, Qall, Qall, dec_site(l, 0), amask));
        SEDT_TRY(linear_drop_add(e.self_attn.out_proj, t.ao1, d, qrows, t.t_in, t.t_mid1, dec_site(l, 1)));
        SEDT_TRY(LN(e.n2, t.t_mid1, qpos, qpos_rows, nullptr, t.dap2, nullptr, qrows));
        SEDT_TRY(linear(e.cross_attn.in_proj, 0, d, t.dap2, dt, d, qrows, nullptr, t.qb, dt, d, 0, s, false));
        SEDT_TRY(attention(t.qb, d, (const char*)tp.ck + (size_t)l * d * es, Dn * d, (const char*)tp.cv + (size_t)l * d * es, Dn * d,
                           t.ao2, tp.mask_ds, Qall, S, dec_site(l, 2)));
        SEDT_TRY(linear_drop_add(e.cross_attn.out_proj, t.ao2, d, qrows, t.t_mid1, t.t_mid2, dec_site(l, 3)));
        SEDT_TRY(LN(e.n3, t.t_mid2, nullptr, 1, t.da3, nullptr, nullptr, qrows));
        SEDT_TRY(linear(e.lin1, 0, ff, t.da3, dt, d, qrows, nullptr, t.h, dt, ff, 1, s, false));
        if (drop) SEDT_TRY(launch_dropout_bf16(t.h, qrows * ff, site(dec_site(l, 4)), s));
        SEDT_TRY(linear_drop_add(e.lin2, t.h, ff, qrows, t.t_mid2, t.t_out, dec_site(l, 5)));
        SEDT_TRY(LN(dec_norm_, t.t_out, nullptr, 1, (char*)tp.hs_t + (size_t)l * qrows * d * es, nullptr,
                    out.hs + (size_t)l * qrows * d, qrows));
    }

    // ---- heads
    const int64_t hrows = (int64_t)Dn * qrows;
    const int ncls = spm ? 1 : cfg_.num_classes, C1 = ncls + 1, start = cfg_.dec_at ? 1 : 0;
    SEDT_TRY(linear(bbox0_, 0, d, tp.hs_t, dt, d, hrows, nullptr, tp.hh1, dt, d, 1, s, false));
    SEDT_TRY(linear(bbox1_, 0, d, tp.hh1, dt, d, hrows, nullptr, tp.hh2, DT_F32, d, 1, s, false));
    SEDT_TRY(launch_heads_out(out.hs, tp.hh2, P_(class_embed_.off_w), P_(class_embed_.off_b), P_(bbox2_.off_w), P_(bbox2_.off_b),
                              cfg_.dec_at ? P_(weak_.off_w) : nullptr, cfg_.dec_at ? P_(weak_.off_b) : nullptr, out.logits, out.boxes,
                              cfg_.dec_at ? out.at : nullptr, Dn, B, Qall, start, C1, ncls, s));
    if (spm && cfg_.feature_recon) {                                     // feature_align MLP (spsedt.py:80)
        SEDT_TRY(linear(falign0_, 0, d, tp.hs_t, dt, d, hrows, nullptr, tp.fh1, dt, d, 1, s, false));
        SEDT_TRY(linear(falign1_, 0, 2048, tp.fh1, dt, d, hrows, nullptr, out.pred_feature, DT_F32, 2048, 0, s, false));
    }
    // sigmoid outputs for the backward pass (the caller's tensors may be modified by then)
    SEDT_CHECK_CUDA(cudaMemcpyAsync(tp.boxes, out.boxes, (size_t)Dn * B * cfg_.num_queries * 2 * 4, cudaMemcpyDeviceToDevice, s));
    if (cfg_.dec_at) SEDT_CHECK_CUDA(cudaMemcpyAsync(tp.at, out.at, (size_t)B * ncls * 4, cudaMemcpyDeviceToDevice, s));
    return SEDT_OK;
}

// ---- gradient buffer layout -----------------------------------------------------------------------------
int64_t Model::grad_offset(int slot) const
{
    int64_t off = 0;
    for (int i = 0; i < slot; ++i) off += (slots_[i].numel + 63) & ~(int64_t)63;       // 256-byte aligned segments
    return off;
}
int64_t Model::grad_numel() const { return grad_offset((int)slots_.size()); }

int64_t Model::backward_workspace_bytes(int B, int T, int F, int P, int PT)
{
    Arena a(nullptr, 0);
    BwdBufs bb;
    Arena ta(nullptr, 0);
    Tape tp;
    tape_layout(B, T, F, false, P, PT, ta, tp);
    bwd_layout(B, tp, a, bb);
    return (int64_t)a.peak + 256;
}


void Model::bwd_layout(int B, const Tape& tp, Arena& a, BwdBufs& bb) const
{
    const int d = cfg_.hidden_dim, ff = cfg_.dim_feedforward;
    const size_t es = 2;
    size_t wmax = (size_t)std::max(ff * d, (int)dec_.size() * d * d), dwmax = (size_t)128 * d;
    size_t g_max = 0, h2_max = 0, h1_max = 0, up_max = 0;
    for (size_t i = 0; i < tp.blk.size(); ++i) {              // (empty for SP-SEDT: frozen backbone, no block tape)
        const Block& b = blocks_[i];
        const BlockTape& bt = tp.blk[i];
        for (const ConvLayer* L : {&b.c1, &b.c2, &b.c3, b.has_ds ? &b.ds : nullptr}) {
            if (L == nullptr) continue;
            wmax = std::max(wmax, (size_t)L->k * L->k * L->cin * L->cout);
            dwmax = std::max(dwmax, (size_t)L->k * L->k * L->cin * L->cout);
        }
        g_max = std::max(g_max, (size_t)B * bt.H * bt.W * b.c1.cin);
        g_max = std::max(g_max, (size_t)B * bt.Ho * bt.Wo * b.c3.cout);
        h2_max = std::max(h2_max, (size_t)B * bt.Ho * bt.Wo * b.c2.cout);
        h1_max = std::max(h1_max, (size_t)B * bt.H * bt.W * b.c1.cout);
        if (b.c2.stride == 2) up_max = std::max(up_max, (size_t)B * bt.H * bt.W * std::max(b.c2.cout, b.has_ds ? b.ds.cout : 0));
    }
    wmax = std::max(wmax, (size_t)2048 * d);
    dwmax = std::max(dwmax, (size_t)2048 * d);
    bb.wd = a.alloc(wmax * es);
    // every parameter appears in at most one data-gradient GEMM; padded head matrices and 256-byte slot alignment are
    // covered by the slack
    bb.wd_all_bytes = ((size_t)grad_numel() + (size_t)4 * 1024 * 1024) * es;
    for (const Block& b : blocks_)            // frozen layers carry no gradient slot but their data gradient still flows
        for (const ConvLayer* L : {&b.c1, &b.c2, &b.c3, b.has_ds ? &b.ds : nullptr})
            if (L != nullptr) bb.wd_all_bytes += (size_t)L->k * L->k * L->cin * L->cout * es + 256;
    bb.wd_all = (char*)a.alloc(bb.wd_all_bytes);
    bb.dw = (float*)a.alloc(dwmax * 4);
    bb.vscale = (float*)a.alloc((size_t)ff * 4);
    const int64_t rows = (int64_t)B * tp.S, qrows = (int64_t)B * tp.Q, hrows = (int64_t)dec_.size() * qrows;
    bb.dcls = a.alloc((size_t)hrows * 128 * es); bb.dbox = a.alloc((size_t)hrows * 128 * es);
    bb.dweak = a.alloc((size_t)std::max(B, 1) * 128 * es);
    bb.dh2 = a.alloc((size_t)hrows * d * es); bb.dh1 = a.alloc((size_t)hrows * d * es); bb.hh2b = a.alloc((size_t)hrows * d * es);
    bb.dhs32 = (float*)a.alloc((size_t)hrows * d * 4);
    const size_t r = (size_t)std::max(rows, qrows);
    bb.gA = (float*)a.alloc(r * d * 4); bb.gB = (float*)a.alloc(r * d * 4);
    bb.g16 = a.alloc(r * d * es); bb.dh = a.alloc(r * ff * es);
    bb.dn_a = a.alloc(r * d * es); bb.dn_b = a.alloc(r * d * es); bb.dao = a.alloc(r * d * es);
    bb.dqk = a.alloc(r * 2 * d * es); bb.dv = a.alloc(r * d * es); bb.dq = a.alloc(r * d * es);
    bb.dck = a.alloc((size_t)rows * dec_.size() * d * es); bb.dcv = a.alloc((size_t)rows * dec_.size() * d * es);
    bb.G0 = a.alloc(std::max(g_max, (size_t)rows * 2048) * es); bb.G1 = a.alloc(std::max<size_t>(g_max, 8) * es);
    bb.gh2 = a.alloc(std::max<size_t>(h2_max, 8) * es); bb.gh1 = a.alloc(std::max<size_t>(h1_max, 8) * es);
    bb.up = a.alloc(std::max<size_t>(up_max, 8) * es);
    bb.dfeat = bb.dfh1 = bb.dpq16 = bb.gt16 = nullptr; bb.dqpos32 = bb.dpq32 = nullptr;
    if (cfg_.self_sup) {
        const size_t np = (size_t)B * cfg_.num_patches;
        bb.dfeat = a.alloc((size_t)hrows * 2048 * es); bb.dfh1 = a.alloc((size_t)hrows * d * es);
        bb.dqpos32 = (float*)a.alloc((size_t)qrows * d * 4); bb.dpq32 = (float*)a.alloc(np * d * 4);
        bb.dpq16 = a.alloc(np * d * es); bb.gt16 = a.alloc(np * 2048 * es);
    }
}

int Model::backward(const void* const* weights, const float* x, const uint8_t* mask, int B, int T, int F, void* tape, size_t tape_bytes_,
                    void* workspace, size_t ws_bytes, const float* d_logits, const float* d_boxes, const float* d_at,
                    float* grads, int train_backbone, float dropout, cudaStream_t s, const SpTrain* sp)
{
    const bool spm = cfg_.self_sup != 0;
    const SpTrain none;
    if (sp == nullptr) sp = &none;
    SEDT_REQUIRE(!spm || !train_backbone, "backward: SP-SEDT pretraining keeps the backbone frozen (train_spsedt.py:50)");
    SEDT_REQUIRE(!spm || sp->P == cfg_.num_patches, "backward: SP-SEDT needs the patch configuration of the forward");
    SEDT_TRY(check_train_config());
    SEDT_REQUIRE(dropout >= 0.f && dropout < 1.f, "backward: dropout=%f", dropout);
    SEDT_REQUIRE(packed_ != nullptr, "backward: sedt_model_pack has not been called");
    SEDT_TRY(tc_init());
    Arena ta(tape, tape_bytes_);
    Tape tp;
    tape_layout(B, T, F, mask != nullptr, sp->P, sp->PT, ta, tp);
    if (ta.overflow) { set_error("backward: tape too small (%zu bytes needed, %zu given)", ta.peak, ta.cap); return SEDT_ERR_WORKSPACE; }
    Arena wa(workspace, ws_bytes);
    BwdBufs bb;
    bwd_layout(B, tp, wa, bb);
    if (wa.overflow) { set_error("backward: workspace too small (%zu bytes needed, %zu given)", wa.peak, wa.cap); return SEDT_ERR_WORKSPACE; }

    const int d = cfg_.hidden_dim, ff = cfg_.dim_feedforward, dt = DT_BF16;
    const size_t es = 2;
    const int S = tp.S, Qall = tp.Q, Dn = (int)dec_.size();
    const int64_t rows = (int64_t)B * S, qrows = (int64_t)B * Qall, hrows = (int64_t)Dn * qrows;
    const float scale = (float)std::sqrt(1.0 / (double)(d / cfg_.nheads));
    auto Wp = [&](int slot) { return (const float*)weights[slot]; };
    auto Gp = [&](int slot) { return grads + grad_offset(slot); };
    auto P_ = [&](size_t off) { return (const float*)(packed_ + off); };

    SEDT_TRY(launch_fill_zero(grads, (size_t)grad_numel() * 4, s));

    // ---- data-gradient weights: wd[ci][r'][s'][co] = scale[co] * w[co][ci][R-1-r'][S-1-s'] for every layer.  The first pass
    // over given weight / workspace addresses re-lays them out site by site and records the job list; later passes (the
    // CUDA-graph capture included) run the whole list as ONE launch up front (SEDT_DGRAD_BATCH=0: always site by site).
    static const bool dgrad_batch = [] { const char* e = getenv("SEDT_DGRAD_BATCH"); return e == nullptr || atoi(e) != 0; }();
    unsigned long long plan_key = 1469598103934665603ull;
    {
        auto mix = [&](unsigned long long v) { plan_key = (plan_key ^ v) * 1099511628211ull; };
        for (size_t i = 0; i < slots_.size(); ++i) mix((unsigned long long)(uintptr_t)weights[i]);
        mix((unsigned long long)(uintptr_t)bb.wd_all); mix((unsigned long long)(uintptr_t)packed_); mix((unsigned long long)train_backbone);
    }
    const bool use_plan = dgrad_batch && dgrad_plan_.ready && dgrad_plan_.key == plan_key;
    if (!use_plan) { dgrad_plan_.jobs.clear(); dgrad_plan_.ready = false; }
    else SEDT_TRY(launch_repack_dgrad_batched(dgrad_plan_.jobs.data(), (int)dgrad_plan_.jobs.size(), s));
    size_t wd_site = 0, wd_bump = 0;
    auto dgrad_weights = [&](const float* W, const float* scale, int Cout, int Cout_pad, int Cin, int k, const void** out) -> int {
        const size_t bytes = align_up((size_t)Cout_pad * Cin * k * k * 2, 256);
        SEDT_REQUIRE(wd_bump + bytes <= bb.wd_all_bytes, "backward: data-gradient weight region too small");
        void* dst = bb.wd_all + wd_bump;
        wd_bump += bytes;
        *out = dst;
        if (use_plan) {
            SEDT_REQUIRE(wd_site < dgrad_plan_.jobs.size(), "backward: data-gradient plan is out of step");
            const DgradJob& J = dgrad_plan_.jobs[wd_site++];
            SEDT_REQUIRE(J.w == W && J.out == dst && J.Cout == Cout && J.Cout_pad == Cout_pad && J.Cin == Cin && J.RS == k * k,
                         "backward: data-gradient plan does not match this pass");
            return SEDT_OK;
        }
        DgradJob J{};
        J.w = W; J.scale = scale; J.out = dst; J.Cout = Cout; J.Cout_pad = Cout_pad; J.Cin = Cin; J.RS = k * k;
        dgrad_plan_.jobs.push_back(J);
        return launch_repack_dgrad(W, scale, dst, DT_BF16, Cout, Cout_pad, Cin, k, k, s);
    };
    auto finish_plan = [&]() -> int {
        if (use_plan) SEDT_REQUIRE(wd_site == dgrad_plan_.jobs.size(), "backward: data-gradient plan has unused entries");
        else if (dgrad_batch) { dgrad_plan_.key = plan_key; dgrad_plan_.ready = true; }
        return SEDT_OK;
    };
    const bool drop = dropout > 0.f;
    SEDT_REQUIRE(!drop || rng_tapes_.count(tape) != 0, "backward: forward_train with dropout has not run on this tape");
    auto site = [&](uint32_t id) { return make_drop_site(tp.rng, id, dropout); };
    if (drop) SEDT_TRY(launch_fill_value(bb.vscale, 1.f / (1.f - dropout), ff, s));
    // bf16 copy of a gradient that enters a GEMM whose forward output went through dropout site `id`
    auto to16_drop = [&](const float* src, void* dst, int64_t n, uint32_t id) -> int {
        if (!drop) return launch_cast(src, dst, dt, n, s);
        return launch_cast_dropout(src, dst, n, site(id), s);
    };

    // dx[M, in_f] = dy[M, out_pad] * W[out_f, in_f]  (+ residual, or masked by `aux` when relu_mode == 2)
    auto dgrad_lin = [&](const float* W, int out_f, int out_pad, int in_f, const void* dy, int ldy, int64_t M, const void* aux,
                         int ld_aux, int relu_mode, void* dx, int dx_dt, int lddx, const float* oscale = nullptr,
                         bool scratch_weights = false) -> int {
        const void* wd = bb.wd;
        if (scratch_weights) SEDT_TRY(launch_repack_dgrad(W, nullptr, bb.wd, dt, out_f, out_pad, in_f, 1, 1, s));   // W is itself a scratch
        else SEDT_TRY(dgrad_weights(W, nullptr, out_f, out_pad, in_f, 1, &wd));
        ConvGemm g;
        g.in = dy; g.w = wd; g.residual = aux; g.out = dx; g.scale = oscale;
        g.in_dt = dt; g.out_dt = dx_dt;
        g.B = (int)M; g.H = g.W = g.Ho = g.Wo = 1; g.Cin = out_pad; g.lda = ldy;
        g.Cout = in_f; g.ldc = lddx; g.ld_res = ld_aux; g.relu = relu_mode;
        SEDT_REQUIRE(conv_tc_supported(g), "backward: data-gradient GEMM %d x %d not supported by the tcgen05 kernel", out_pad, in_f);
        return launch_conv_tc(g, s);
    };
    // dw[out_f, in_f] += dy^T x
    auto wgrad_lin = [&](const void* x, int lda, int in_f, const void* dy, int ldy, int out_f, int64_t M, float* dw) -> int {
        WgradGemm g;
        g.x = x; g.dy = dy; g.dw = dw; g.B = (int)M; g.Cin = in_f; g.lda = lda; g.Cout = out_f; g.ldy = ldy;
        SEDT_REQUIRE(conv_wgrad_tc_supported(g), "backward: weight-gradient GEMM %d x %d not supported", out_f, in_f);
        return launch_conv_wgrad_tc(g, s);
    };
    auto to16 = [&](const float* src, void* dst, int64_t n) { return launch_cast(src, dst, dt, n, s); };
    // one linear layer y = x W^T + b, W = weights[w_slot] rows [r0, r0 + out_f):  parameter gradients
    auto lin_param_grads = [&](const Linear& L, int r0, int out_f, const void* x, int lda, const void* dy, int ldy, int64_t M) -> int {
        SEDT_TRY(wgrad_lin(x, lda, L.in, dy, ldy, out_f, M, Gp(L.w_slot) + (size_t)r0 * L.in));
        return launch_colsum(dy, dt, ldy, Gp(L.b_slot) + r0, M, out_f, s);
    };

    // ================= heads (sedt/sedt.py:89-95) =====================================================
    const int ncls = spm ? 1 : cfg_.num_classes, C1 = ncls + 1, start = cfg_.dec_at ? 1 : 0;
    SEDT_TRY(launch_heads_bwd_prepare(d_logits, d_boxes, d_at, tp.boxes, tp.at, bb.dcls, bb.dbox, cfg_.dec_at ? bb.dweak : nullptr, Dn,
                                      B, Qall, start, C1, ncls, s));
    auto padded_param_grads = [&](const Linear& L, const void* x, int lda, const void* dy_pad, int64_t M) -> int {
        SEDT_TRY(launch_fill_zero(bb.dw, (size_t)128 * L.in * 4, s));
        SEDT_TRY(wgrad_lin(x, lda, L.in, dy_pad, 128, 128, M, bb.dw));
        SEDT_CHECK_CUDA(cudaMemcpyAsync(Gp(L.w_slot), bb.dw, (size_t)L.out * L.in * 4, cudaMemcpyDeviceToDevice, s));
        return launch_colsum(dy_pad, dt, 128, Gp(L.b_slot), M, L.out, s);
    };
    // class_embed
    SEDT_TRY(padded_param_grads(class_embed_, tp.hs_t, d, bb.dcls, hrows));
    SEDT_TRY(dgrad_lin(Wp(class_embed_.w_slot), C1, 128, d, bb.dcls, 128, hrows, nullptr, 0, 0, bb.dhs32, DT_F32, d));
    // bbox MLP: raw = W2 relu(W1 relu(W0 hs + b0) + b1) + b2
    SEDT_TRY(to16(tp.hh2, bb.hh2b, hrows * d));
    SEDT_TRY(padded_param_grads(bbox2_, bb.hh2b, d, bb.dbox, hrows));
    SEDT_TRY(dgrad_lin(Wp(bbox2_.w_slot), 2, 128, d, bb.dbox, 128, hrows, bb.hh2b, d, 2, bb.dh2, dt, d));
    SEDT_TRY(lin_param_grads(bbox1_, 0, d, tp.hh1, d, bb.dh2, d, hrows));
    SEDT_TRY(dgrad_lin(Wp(bbox1_.w_slot), d, d, d, bb.dh2, d, hrows, tp.hh1, d, 2, bb.dh1, dt, d));
    SEDT_TRY(lin_param_grads(bbox0_, 0, d, tp.hs_t, d, bb.dh1, d, hrows));
    SEDT_TRY(dgrad_lin(Wp(bbox0_.w_slot), d, d, d, bb.dh1, d, hrows, bb.dhs32, d, 0, bb.dhs32, DT_F32, d));
    if (cfg_.dec_at) {      // slot 0 of the last layer (sedt/sedt.py:92)
        const void* xw = (const char*)tp.hs_t + (size_t)(Dn - 1) * qrows * d * es;
        float* dst = bb.dhs32 + (size_t)(Dn - 1) * qrows * d;
        SEDT_TRY(padded_param_grads(weak_, xw, Qall * d, bb.dweak, B));
        SEDT_TRY(dgrad_lin(Wp(weak_.w_slot), ncls, 128, d, bb.dweak, 128, B, dst, Qall * d, 0, dst, DT_F32, Qall * d));
    }

    if (spm && cfg_.feature_recon && sp->d_pred_feature != nullptr) {        // feature_align: pred_feature = W1 relu(W0 hs + b0) + b1
        SEDT_TRY(to16(sp->d_pred_feature, bb.dfeat, hrows * 2048));
        SEDT_TRY(lin_param_grads(falign1_, 0, 2048, tp.fh1, d, bb.dfeat, 2048, hrows));
        SEDT_TRY(dgrad_lin(Wp(falign1_.w_slot), 2048, 2048, d, bb.dfeat, 2048, hrows, tp.fh1, d, 2, bb.dfh1, dt, d));
        SEDT_TRY(lin_param_grads(falign0_, 0, d, tp.hs_t, d, bb.dfh1, d, hrows));
        SEDT_TRY(dgrad_lin(Wp(falign0_.w_slot), d, d, d, bb.dfh1, d, hrows, bb.dhs32, d, 0, bb.dhs32, DT_F32, d));
    }

    // attention core backward: tcgen05 kernel (SEDT_ATT_BWD_SIMT=1 selects the CUDA-core one)
    static const bool att_simt = [] { const char* e = getenv("SEDT_ATT_BWD_SIMT"); return e != nullptr && e[0] == '1'; }();
    auto attn_bwd = [&](const void* Qp, int ldq, const void* Kp, int ldk, const void* Vp, int ldv, const void* dOp, int ldo, void* dQp,
                        int lddq, void* dKp, int lddk, void* dVp, int lddv, const uint8_t* kpm, const float* am, int Bn, int nh, int Lq,
                        int Lk, float sc, uint32_t site_p, cudaStream_t st) -> int {
        const DropSite ds = site(site_p);
        if ((!att_simt || drop) && attention_bwd_tc_supported(Qp, ldq, Kp, ldk, Vp, ldv, dOp, ldo, dQp, lddq, dKp, lddk, dVp, lddv, Lq, Lk))
            return launch_attention_bwd_tc(Qp, ldq, Kp, ldk, Vp, ldv, dOp, ldo, dQp, lddq, dKp, lddk, dVp, lddv, kpm, am, Bn, nh, Lq, Lk, sc,
                                           drop ? &ds : nullptr, st);
        SEDT_REQUIRE(!drop, "backward: attention dropout needs the tcgen05 kernel");
        return launch_attention_bwd(Qp, ldq, Kp, ldk, Vp, ldv, dOp, ldo, dQp, lddq, dKp, lddk, dVp, lddv, kpm, am, Bn, nh, Lq, Lk, sc, st);
    };

    // ================= decoder (transformer.py:263-284), last layer first ===============================
    // self-attention + its LayerNorm, shared by encoder and decoder layers.
    //   in : gx = d(loss)/d(x_mid) fp32 (x_mid = x_in + out_proj(attn))      out: gout = d(loss)/d(x_in) fp32
    // query-position gradient: summed over the batch into query_embed's slot (SEDT), kept per clip for SP-SEDT (the positions
    // are batch dependent there: 2 * query_embed + keep * patch2query(patch feature))
    if (spm) SEDT_TRY(launch_fill_zero(bb.dqpos32, (size_t)qrows * d * 4, s));
    auto add_dqpos = [&](const void* g16, float* dqpos_acc, int L) -> int {
        if (spm) return launch_accum_bf16(g16, bb.dqpos32, (int64_t)B * L * d, s);
        return launch_colsum(g16, dt, (int64_t)L * d, dqpos_acc, B, L * d, s);
    };
    auto self_attn_bwd = [&](const Mha& A, const Norm& n1, const float* x_in, const void* na, const void* nap, const void* qk,
                             const void* v, const void* ao, const uint8_t* kpm, int L, int64_t R, const float* gx, float* gout,
                             float* dqpos_acc, uint32_t site_p, uint32_t site_o, const float* am = nullptr) -> int {
        SEDT_TRY(to16_drop(gx, bb.g16, R * d, site_o));
        SEDT_TRY(lin_param_grads(A.out_proj, 0, d, ao, d, bb.g16, d, R));
        SEDT_TRY(dgrad_lin(Wp(A.out_proj.w_slot), d, d, d, bb.g16, d, R, nullptr, 0, 0, bb.dao, dt, d));
        SEDT_TRY(attn_bwd(qk, 2 * d, (const char*)qk + d * es, 2 * d, v, d, bb.dao, d, bb.dqk, 2 * d,
                          (char*)bb.dqk + d * es, 2 * d, bb.dv, d, kpm, am, B, cfg_.nheads, L, L, scale, site_p, s));
        SEDT_TRY(lin_param_grads(A.in_proj, 0, 2 * d, nap, d, bb.dqk, 2 * d, R));
        SEDT_TRY(lin_param_grads(A.in_proj, 2 * d, d, na, d, bb.dv, d, R));
        SEDT_TRY(dgrad_lin(Wp(A.in_proj.w_slot), 2 * d, 2 * d, d, bb.dqk, 2 * d, R, nullptr, 0, 0, bb.dn_b, dt, d));     // d(LN + pos)
        SEDT_TRY(dgrad_lin(Wp(A.in_proj.w_slot) + (size_t)2 * d * d, d, d, d, bb.dv, d, R, nullptr, 0, 0, bb.dn_a, dt, d));   // d(LN)
        if (dqpos_acc != nullptr) SEDT_TRY(add_dqpos(bb.dn_b, dqpos_acc, L));
        return launch_layernorm_bwd(x_in, P_(n1.off_g), bb.dn_a, bb.dn_b, nullptr, gx, gout, Gp(n1.w_slot), Gp(n1.b_slot), R, s);
    };
    // FFN + its LayerNorm:  in: gy = d/d(x_out) fp32, x_out = x_mid + lin2(relu(lin1(LN(x_mid))))   out: gout = d/d(x_mid)
    auto ffn_bwd = [&](const Linear& lin1, const Linear& lin2, const Norm& nn, const float* x_mid, const void* n_out, const void* h,
                       int64_t R, const float* gy, float* gout, uint32_t site_h, uint32_t site_o) -> int {
        SEDT_TRY(to16_drop(gy, bb.g16, R * d, site_o));
        SEDT_TRY(lin_param_grads(lin2, 0, d, h, ff, bb.g16, d, R));
        // h is stored after dropout: h > 0 is the ReLU AND the keep mask; the 1/(1-p) factor rides on the GEMM's output scale
        (void)site_h;
        SEDT_TRY(dgrad_lin(Wp(lin2.w_slot), d, d, ff, bb.g16, d, R, h, ff, 2, bb.dh, dt, ff, drop ? bb.vscale : nullptr));
        SEDT_TRY(lin_param_grads(lin1, 0, ff, n_out, d, bb.dh, ff, R));
        SEDT_TRY(dgrad_lin(Wp(lin1.w_slot), ff, ff, d, bb.dh, ff, R, nullptr, 0, 0, bb.dn_a, dt, d));
        return launch_layernorm_bwd(x_mid, P_(nn.off_g), bb.dn_a, nullptr, nullptr, gy, gout, Gp(nn.w_slot), Gp(nn.b_slot), R, s);
    };

    float* dqpos = Gp(s_query_embed);
    float* gcur = bb.gA;            // gradient w.r.t. the residual stream at the current point
    float* gnext = bb.gB;
    for (int l = Dn - 1; l >= 0; --l) {
        const DecLayer& e = dec_[l];
        const DecTape& t = tp.dec[l];
        // decoder.norm on this layer's output (+ what flows in from the layer above)
        SEDT_TRY(launch_layernorm_bwd(t.t_out, P_(dec_norm_.off_g), nullptr, nullptr, bb.dhs32 + (size_t)l * qrows * d,
                                      l == Dn - 1 ? nullptr : gcur, gnext, Gp(dec_norm_.w_slot), Gp(dec_norm_.b_slot), qrows, s));
        std::swap(gcur, gnext);
        SEDT_TRY(ffn_bwd(e.lin1, e.lin2, e.n3, t.t_mid2, t.da3, t.h, qrows, gcur, gnext, dec_site(l, 4), dec_site(l, 5)));
        std::swap(gcur, gnext);
        // cross attention: t_mid2 = t_mid1 + out_proj(attn(q = Wq(LN2(t_mid1) + qpos), K_l, V_l))
        SEDT_TRY(to16_drop(gcur, bb.g16, qrows * d, dec_site(l, 3)));
        SEDT_TRY(lin_param_grads(e.cross_attn.out_proj, 0, d, t.ao2, d, bb.g16, d, qrows));
        SEDT_TRY(dgrad_lin(Wp(e.cross_attn.out_proj.w_slot), d, d, d, bb.g16, d, qrows, nullptr, 0, 0, bb.dao, dt, d));
        SEDT_TRY(attn_bwd(t.qb, d, (const char*)tp.ck + (size_t)l * d * es, Dn * d, (const char*)tp.cv + (size_t)l * d * es,
                                      Dn * d, bb.dao, d, bb.dq, d, (char*)bb.dck + (size_t)l * d * es, Dn * d,
                                      (char*)bb.dcv + (size_t)l * d * es, Dn * d, tp.mask_ds, nullptr, B, cfg_.nheads, Qall, S, scale,
                                      dec_site(l, 2), s));
        SEDT_TRY(lin_param_grads(e.cross_attn.in_proj, 0, d, t.dap2, d, bb.dq, d, qrows));
        SEDT_TRY(dgrad_lin(Wp(e.cross_attn.in_proj.w_slot), d, d, d, bb.dq, d, qrows, nullptr, 0, 0, bb.dn_b, dt, d));
        SEDT_TRY(add_dqpos(bb.dn_b, dqpos, Qall));
        SEDT_TRY(launch_layernorm_bwd(t.t_mid1, P_(e.n2.off_g), nullptr, bb.dn_b, nullptr, gcur, gnext, Gp(e.n2.w_slot), Gp(e.n2.b_slot),
                                      qrows, s));
        std::swap(gcur, gnext);
        // K_l = Wk (memory + pos) + bk, V_l = Wv memory + bv: parameter gradients per layer
        SEDT_TRY(lin_param_grads(e.cross_attn.in_proj, d, d, tp.mempos, d, (const char*)bb.dck + (size_t)l * d * es, Dn * d, rows));
        SEDT_TRY(lin_param_grads(e.cross_attn.in_proj, 2 * d, d, tp.mem, d, (const char*)bb.dcv + (size_t)l * d * es, Dn * d, rows));
        // self attention
        SEDT_TRY(self_attn_bwd(e.self_attn, e.n1, t.t_in, t.da, t.dap, t.qk, t.v, t.ao1, nullptr, Qall, qrows, gcur, gnext, dqpos,
                               dec_site(l, 0), dec_site(l, 1), spm ? tp.amask : nullptr));
        std::swap(gcur, gnext);
    }
    if (spm) {
        // query positions = 2 * query_embed + keep * patch2query(gt_feature) (spsedt.py:63-67): query_embed gets twice the batch sum,
        // patch2query the kept queries' gradient summed per patch; gt_feature itself only leads into the frozen backbone
        const int qpp = cfg_.num_queries / cfg_.num_patches;
        const int64_t np = (int64_t)B * sp->P;
        SEDT_TRY(launch_patch_query_bwd(bb.dqpos32, tp.keep, dqpos, bb.dpq32, B, sp->P, qpp, 2.f, s));
        SEDT_TRY(to16(bb.dpq32, bb.dpq16, np * d));
        SEDT_TRY(to16(tp.gt, bb.gt16, np * 2048));
        SEDT_TRY(wgrad_lin(bb.gt16, 2048, 2048, bb.dpq16, d, d, np, Gp(patch2query_.w_slot)));
        SEDT_TRY(launch_colsum(bb.dpq32, DT_F32, d, Gp(patch2query_.b_slot), np, d, s));
    }
    // memory: d(mem + pos) = sum_l dK_l Wk_l, d(mem) = sum_l dV_l Wv_l  -> encoder.norm backward
    {
        // [Wk_0; Wk_1; ...] as one [D*d, d] matrix: gather the fp32 slices into the scratch, then one GEMM with K = D*d
        float* wcat = bb.dw;
        for (int l = 0; l < Dn; ++l)
            SEDT_CHECK_CUDA(cudaMemcpyAsync(wcat + (size_t)l * d * d, Wp(dec_[l].cross_attn.in_proj.w_slot) + (size_t)d * d,
                                            (size_t)d * d * 4, cudaMemcpyDeviceToDevice, s));
        SEDT_TRY(dgrad_lin(wcat, Dn * d, Dn * d, d, bb.dck, Dn * d, rows, nullptr, 0, 0, bb.dn_b, dt, d, nullptr, true));
        for (int l = 0; l < Dn; ++l)
            SEDT_CHECK_CUDA(cudaMemcpyAsync(wcat + (size_t)l * d * d, Wp(dec_[l].cross_attn.in_proj.w_slot) + (size_t)2 * d * d,
                                            (size_t)d * d * 4, cudaMemcpyDeviceToDevice, s));
        SEDT_TRY(dgrad_lin(wcat, Dn * d, Dn * d, d, bb.dcv, Dn * d, rows, nullptr, 0, 0, bb.dn_a, dt, d, nullptr, true));
    }
    const float* xe = enc_.empty() ? tp.x0 : tp.enc.back().x_out;
    SEDT_TRY(launch_layernorm_bwd(xe, P_(enc_norm_.off_g), bb.dn_a, bb.dn_b, nullptr, nullptr, bb.gA, Gp(enc_norm_.w_slot),
                                  Gp(enc_norm_.b_slot), rows, s));
    gcur = bb.gA; gnext = bb.gB;

    // ================= encoder (transformer.py:192-204) ==================================================
    for (int l = (int)enc_.size() - 1; l >= 0; --l) {
        const EncLayer& e = enc_[l];
        const EncTape& t = tp.enc[l];
        SEDT_TRY(ffn_bwd(e.lin1, e.lin2, e.n2, t.x_mid, t.n2, t.h, rows, gcur, gnext, enc_site(l, 2), enc_site(l, 3)));
        std::swap(gcur, gnext);
        SEDT_TRY(self_attn_bwd(e.attn, e.n1, t.x_in, t.na, t.nap, t.qk, t.v, t.ao, tp.mask_ds, S, rows, gcur, gnext, nullptr,
                               enc_site(l, 0), enc_site(l, 1)));
        std::swap(gcur, gnext);
    }

    // ================= input_proj (sedt/sedt.py:36,88) ===================================================
    const void* feat = nullptr;
    if (spm) {                                        // same arena, same order as forward_train: the clip backbone's layer4 output
        Arena ws(tp.scratch, tp.scratch_bytes);
        void* cfeat = nullptr; int ch = 0, cw = 0;
        SEDT_TRY(backbone(nullptr, B, T, F, ws, &cfeat, &ch, &cw, s, true));
        feat = cfeat;
    } else {
        feat = tp.blk.back().out;
    }
    SEDT_TRY(to16(gcur, bb.g16, rows * d));
    SEDT_TRY(lin_param_grads(input_proj_, 0, d, feat, 2048, bb.g16, d, rows));
    // everything outside the backbone is final: the first all-reduce bucket may go (see Model::set_bucket_event)
    if (bucket_event_ != nullptr) SEDT_CHECK_CUDA(cudaEventRecord(bucket_event_, s));
    if (!train_backbone) return finish_plan();
    // gradient w.r.t. the layer4 output, already masked by its ReLU
    SEDT_TRY(dgrad_lin(Wp(input_proj_.w_slot), d, d, 2048, bb.g16, d, rows, feat, 2048, 2, bb.G0, dt, 2048));

    // ================= backbone (torchvision resnet.py:143-163), last block first =========================
    void* G = bb.G0;            // d(loss)/d(block output), ReLU mask applied
    void* Gn = bb.G1;
    auto conv_dgrad = [&](const ConvLayer& L, const void* dy, int Hin, int Win, int Ho, int Wo, const void* aux, int relu_mode,
                          void* dx) -> int {
        const void* wd = nullptr;
        SEDT_TRY(dgrad_weights(Wp(L.w_slot), P_(L.off_scale), L.cout, L.cout, L.cin, L.k, &wd));
        const void* gin = dy;
        if (L.stride == 2) { SEDT_TRY(launch_upsample2(dy, bb.up, B, Hin, Win, Ho, Wo, L.cout, s)); gin = bb.up; }
        ConvGemm g;
        g.in = gin; g.w = wd; g.residual = aux; g.out = dx;
        g.in_dt = g.out_dt = dt;
        g.B = B; g.H = Hin; g.W = Win; g.Ho = Hin; g.Wo = Win; g.Cin = L.cout; g.lda = L.cout;
        g.Cout = L.cin; g.ldc = L.cin; g.ld_res = L.cin;
        g.R = g.S = L.k; g.stride = 1; g.dil = L.dil; g.pad = L.k == 3 ? L.dil : 0; g.relu = relu_mode;
        SEDT_REQUIRE(conv_tc_supported(g), "backward: conv data gradient %dx%d k%d not supported", L.cout, L.cin, L.k);
        return launch_conv_tc(g, s);
    };
    auto conv_wgrad = [&](const ConvLayer& L, const void* x, int Hin, int Win, int Ho, int Wo, const void* dy) -> int {
        const size_t n = (size_t)L.cout * L.k * L.k * L.cin;
        // 1x1: the GEMM layout [co][ci] IS the reference's OIHW layout, so the kernel accumulates straight into the (already
        // zeroed) gradient slot with the folded BN scale applied per row; 3x3 goes through the scratch + tap permutation
        const bool direct = L.k == 1;
        if (!direct) SEDT_TRY(launch_fill_zero(bb.dw, n * 4, s));
        WgradGemm g;
        g.x = x; g.dy = dy; g.dw = direct ? Gp(L.w_slot) : bb.dw; g.row_scale = direct ? P_(L.off_scale) : nullptr;
        g.B = B; g.H = Hin; g.W = Win; g.Cin = L.cin; g.lda = L.cin; g.Ho = Ho; g.Wo = Wo; g.Cout = L.cout; g.ldy = L.cout;
        g.R = g.S = L.k; g.stride = L.stride; g.dil = L.dil; g.pad = L.pad;
        SEDT_REQUIRE(conv_wgrad_tc_supported(g), "backward: conv weight gradient %dx%d k%d not supported", L.cout, L.cin, L.k);
        SEDT_TRY(launch_conv_wgrad_tc(g, s));
        if (direct) return SEDT_OK;
        return launch_unpack_wgrad(bb.dw, P_(L.off_scale), Gp(L.w_slot), L.cout, L.cin, L.k * L.k, s);
    };
    const size_t first_trainable = 3;                 // layer1 (blocks 0-2) is frozen (sedt/backbone.py:60-62)
    for (int i = (int)blocks_.size() - 1; i >= 0; --i) {
        const Block& b = blocks_[i];
        const BlockTape& bt = tp.blk[i];
        const void* x_in = i == 0 ? tp.stem_out : tp.blk[i - 1].out;
        const bool trainable = (size_t)i >= first_trainable;
        // conv3 (1x1): d(h2) = G W3, masked by h2 > 0
        if (trainable) SEDT_TRY(conv_wgrad(b.c3, bt.h2, bt.Ho, bt.Wo, bt.Ho, bt.Wo, G));
        SEDT_TRY(conv_dgrad(b.c3, G, bt.Ho, bt.Wo, bt.Ho, bt.Wo, bt.h2, 2, bb.gh2));
        // conv2 (3x3, maybe stride 2 / dilated)
        if (trainable) SEDT_TRY(conv_wgrad(b.c2, bt.h1, bt.H, bt.W, bt.Ho, bt.Wo, bb.gh2));
        SEDT_TRY(conv_dgrad(b.c2, bb.gh2, bt.H, bt.W, bt.Ho, bt.Wo, bt.h1, 2, bb.gh1));
        // conv1 (1x1) + the identity / downsample branch
        if (trainable) SEDT_TRY(conv_wgrad(b.c1, x_in, bt.H, bt.W, bt.H, bt.W, bb.gh1));
        if (b.has_ds) {
            if (trainable) SEDT_TRY(conv_wgrad(b.ds, x_in, bt.H, bt.W, bt.Ho, bt.Wo, G));
            SEDT_TRY(conv_dgrad(b.c1, bb.gh1, bt.H, bt.W, bt.H, bt.W, nullptr, 0, Gn));
            SEDT_TRY(conv_dgrad(b.ds, G, bt.H, bt.W, bt.Ho, bt.Wo, Gn, 0, Gn));
        } else {
            SEDT_TRY(conv_dgrad(b.c1, bb.gh1, bt.H, bt.W, bt.H, bt.W, G, 0, Gn));
        }
        // ReLU of the block below (or of the stem)
        SEDT_TRY(launch_relu_mask(x_in, Gn, nullptr, Gn, (int64_t)B * bt.H * bt.W * b.c1.cin, s));
        std::swap(G, Gn);
    }
    // G = d(loss)/d(stem output): conv0 is the only trainable parameter below (sedt/backbone.py:60,102)
    SEDT_TRY(launch_stem_bwd(x, Wp(s_conv1_w), P_(off_stem_scale), G, tp.stem_amax, bb.dw, Gp(s_conv0_w), Gp(s_conv0_b), B, T, F, s));
    return finish_plan();
}

}  // namespace sedt
