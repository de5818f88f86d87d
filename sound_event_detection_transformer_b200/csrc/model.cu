// Host runtime: layer table, weight packing and the forward launch sequence of
// SEDT / SP-SEDT (sedt/sedt.py:64-131, sedt/spsedt.py:34-91,
// sedt/transformer.py:48-86, sedt/backbone.py:71-113).  See model.h.
#include "model.h"
#include <cmath>
#include <cstdarg>
#include <cstring>
#include <cstdlib>

namespace sedt {

// ---- error string / launch counter ------------------------------------------------
static thread_local char g_err[1024] = "";
unsigned long long g_launch_count = 0;
unsigned long long g_kind_count[KK_NKINDS] = {};
const char* kernel_kind_name(int kind)
{
    static const char* names[KK_NKINDS] = {"conv_tc2", "conv_tc3_2sm", "conv_tc4_ws", "conv_tc_v1", "ffn_fused", "attention_tc", "stem_tc",
                                           "wgrad_tc", "attention_bwd_tc", "enc_attn_fused", "bottleneck_fused", "dec_layer_fused"};
    return kind >= 0 && kind < KK_NKINDS ? names[kind] : nullptr;
}

bool pdl_enabled()
{
    // off by default: measured on B200 inside the CUDA graph, 4.85 ms (off) vs 5.00 ms (on) per 256-clip forward and no
    // change for the training step -- graph replay already hides the launch latency and the early CTAs only add contention
    static const bool on = [] { const char* e = getenv("SEDT_PDL"); return e != nullptr && atoi(e) != 0; }();
    return on;
}

void set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char* get_error() { return g_err; }

// ---- per-class kernel timing -----------------------------------------------------------
bool g_prof_on = false;
namespace {
struct ProfRec { int cls; cudaEvent_t a, b; };
std::vector<ProfRec> g_prof_recs;
std::vector<cudaEvent_t> g_prof_pool;
cudaEvent_t prof_event()
{
    if (!g_prof_pool.empty()) { cudaEvent_t e = g_prof_pool.back(); g_prof_pool.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
}
}  // namespace
void prof_begin(int cls, cudaStream_t s)
{
    ProfRec r{cls, prof_event(), prof_event()};
    cudaEventRecord(r.a, s);
    g_prof_recs.push_back(r);
}
void prof_end(cudaStream_t s) { cudaEventRecord(g_prof_recs.back().b, s); }
int prof_read(double* ms, long long* counts)
{
    for (int i = 0; i < PROF_NCLASS; ++i) { ms[i] = 0.0; counts[i] = 0; }
    SEDT_CHECK_CUDA(cudaDeviceSynchronize());
    for (auto& r : g_prof_recs) {
        float t = 0.f;
        SEDT_CHECK_CUDA(cudaEventElapsedTime(&t, r.a, r.b));
        ms[r.cls] += t; counts[r.cls] += 1;
        g_prof_pool.push_back(r.a); g_prof_pool.push_back(r.b);
    }
    g_prof_recs.clear();
    return SEDT_OK;
}

// ---- layer table ------------------------------------------------------------------
static const char* kBody = "backbone.0.body.";

int Model::add_slot(const std::string& name, int64_t numel)
{
    slots_.push_back({name, numel});
    return (int)slots_.size() - 1;
}

size_t Model::reserve(size_t bytes)
{
    size_t o = (packed_bytes_ + 255) & ~(size_t)255;
    packed_bytes_ = o + bytes;
    return o;
}

ConvLayer Model::make_conv(const std::string& conv, const std::string& bn, int cin, int cout, int k, int stride, int dil,
                           int pad, int relu)
{
    ConvLayer L{};
    L.cin = cin; L.cout = cout; L.k = k; L.stride = stride; L.dil = dil; L.pad = pad; L.relu = relu;
    L.w_slot = add_slot(conv + ".weight", (int64_t)cout * cin * k * k);
    L.bn_slot = add_slot(bn + ".weight", cout);
    add_slot(bn + ".bias", cout);
    add_slot(bn + ".running_mean", cout);
    add_slot(bn + ".running_var", cout);
    L.off_w = reserve((size_t)cout * cin * k * k * dtype_size(act_dt()));
    L.off_scale = reserve((size_t)cout * 4);
    L.off_bias = reserve((size_t)cout * 4);
    return L;
}

Linear Model::make_linear_slots(int w_slot, int b_slot, int in, int out, bool f32_only)
{
    Linear L{};
    L.in = in; L.out = out; L.w_slot = w_slot; L.b_slot = b_slot; L.f32_only = f32_only;
    L.off_w = reserve((size_t)in * out * (f32_only ? 4 : dtype_size(act_dt())));
    L.off_b = reserve((size_t)out * 4);
    return L;
}

Linear Model::make_linear(const std::string& name, int in, int out, bool f32_only)
{
    int w = add_slot(name + ".weight", (int64_t)in * out);
    int b = add_slot(name + ".bias", out);
    return make_linear_slots(w, b, in, out, f32_only);
}

Norm Model::make_norm(const std::string& name)
{
    Norm n{};
    n.w_slot = add_slot(name + ".weight", cfg_.hidden_dim);
    n.b_slot = add_slot(name + ".bias", cfg_.hidden_dim);
    n.off_g = reserve((size_t)cfg_.hidden_dim * 4);
    n.off_b = reserve((size_t)cfg_.hidden_dim * 4);
    return n;
}

Mha Model::make_mha(const std::string& name)
{
    const int d = cfg_.hidden_dim;
    Mha m{};
    int w = add_slot(name + ".in_proj_weight", (int64_t)3 * d * d);
    int b = add_slot(name + ".in_proj_bias", 3 * d);
    m.in_proj = make_linear_slots(w, b, d, 3 * d, false);
    m.out_proj = make_linear(name + ".out_proj", d, d, false);
    return m;
}

Model::Model(const Config& c) : cfg_(c)
{
    const int d = cfg_.hidden_dim, ff = cfg_.dim_feedforward;
    // slot order follows the reference's state_dict order (transformer, heads, input_proj, backbone, embeddings)
    for (int l = 0; l < cfg_.enc_layers; ++l) {
        std::string p = "transformer.encoder.layers." + std::to_string(l) + ".";
        EncLayer e{};
        e.attn = make_mha(p + "self_attn");
        e.lin1 = make_linear(p + "linear1", d, ff, false);
        e.lin2 = make_linear(p + "linear2", ff, d, false);
        e.n1 = make_norm(p + "norm1");
        e.n2 = make_norm(p + "norm2");
        enc_.push_back(e);
    }
    if (cfg_.pre_norm) enc_norm_ = make_norm("transformer.encoder.norm");
    for (int l = 0; l < cfg_.dec_layers; ++l) {
        std::string p = "transformer.decoder.layers." + std::to_string(l) + ".";
        DecLayer e{};
        e.self_attn = make_mha(p + "self_attn");
        e.cross_attn = make_mha(p + "multihead_attn");
        e.lin1 = make_linear(p + "linear1", d, ff, false);
        e.lin2 = make_linear(p + "linear2", ff, d, false);
        e.n1 = make_norm(p + "norm1");
        e.n2 = make_norm(p + "norm2");
        e.n3 = make_norm(p + "norm3");
        dec_.push_back(e);
    }
    dec_norm_ = make_norm("transformer.decoder.norm");
    const int ncls = cfg_.self_sup ? 1 : cfg_.num_classes;
    class_embed_ = make_linear("class_embed", d, ncls + 1, true);
    // the 256x256 layers of the box MLP (and feature_align) run on the tensor cores in the bf16 tier; the narrow
    // output layers (C+1, 2, C columns) stay fp32 CUDA-core GEMMs
    bbox0_ = make_linear("bbox_embed.layers.0", d, d, false);
    bbox1_ = make_linear("bbox_embed.layers.1", d, d, false);
    bbox2_ = make_linear("bbox_embed.layers.2", d, 2, true);
    input_proj_ = make_linear("input_proj", 2048, d, false);

    s_conv0_w = add_slot(std::string(kBody) + "conv0.weight", 3);
    s_conv0_b = add_slot(std::string(kBody) + "conv0.bias", 3);
    s_conv1_w = add_slot(std::string(kBody) + "conv1.weight", 64 * 3 * 49);
    s_bn1 = add_slot(std::string(kBody) + "bn1.weight", 64);
    add_slot(std::string(kBody) + "bn1.bias", 64);
    add_slot(std::string(kBody) + "bn1.running_mean", 64);
    add_slot(std::string(kBody) + "bn1.running_var", 64);
    off_weff = reserve(49 * 64 * 4);
    off_sat = reserve(64 * 64 * 4);
    off_stem_scale = reserve(64 * 4);
    off_stem_bias = reserve(64 * 4);
    off_stem_wtc = reserve(16384);

    // torchvision resnet50 (resnet.py:225-262) with replace_stride_with_dilation=[F,F,dilation]
    static const int planes_[4] = {64, 128, 256, 512}, nblk_[4] = {3, 4, 6, 3}, stride_[4] = {1, 2, 2, 2};
    int inplanes = 64, cur_dil = 1;
    for (int li = 0; li < 4; ++li) {
        int stride = stride_[li];
        const int prev_dil = cur_dil;
        if (li == 3 && cfg_.dilation) { cur_dil *= stride; stride = 1; }
        for (int bi = 0; bi < nblk_[li]; ++bi) {
            std::string p = std::string(kBody) + "layer" + std::to_string(li + 1) + "." + std::to_string(bi) + ".";
            const int s = bi == 0 ? stride : 1, dl = bi == 0 ? prev_dil : cur_dil, pl = planes_[li];
            Block b{};
            b.c1 = make_conv(p + "conv1", p + "bn1", inplanes, pl, 1, 1, 1, 0, 1);
            b.c2 = make_conv(p + "conv2", p + "bn2", pl, pl, 3, s, dl, dl, 1);
            b.c3 = make_conv(p + "conv3", p + "bn3", pl, pl * 4, 1, 1, 1, 0, 1);   // relu after the residual add
            b.has_ds = bi == 0 && (s != 1 || inplanes != pl * 4);
            if (b.has_ds) b.ds = make_conv(p + "downsample.0", p + "downsample.1", inplanes, pl * 4, 1, s, 1, 0, 0);
            blocks_.push_back(b);
            inplanes = pl * 4;
        }
    }
    qall_ = cfg_.num_queries + (cfg_.dec_at ? 1 : 0);
    s_query_embed = add_slot("query_embed.weight", (int64_t)qall_ * d);
    off_query_embed = reserve((size_t)qall_ * d * 4);
    {
        const size_t nd = (size_t)cfg_.dec_layers;
        off_ck_w = reserve(nd * d * d * dtype_size(act_dt())); off_ck_b = reserve(nd * d * 4);
        off_cv_w = reserve(nd * d * d * dtype_size(act_dt())); off_cv_b = reserve(nd * d * 4);
    }
    if (cfg_.dec_at) weak_ = make_linear("weak_class_embed", d, ncls, true);
    if (cfg_.self_sup) {
        patch2query_ = make_linear("patch2query", 2048, d, true);
        if (cfg_.feature_recon) {
            falign0_ = make_linear("feature_align.layers.0", d, d, false);
            falign1_ = make_linear("feature_align.layers.1", d, 2048, false);
        }
    }
    packed_bytes_ = (packed_bytes_ + 255) & ~(size_t)255;
}

// ---- packing ------------------------------------------------------------------------
int Model::pack(const void* const* weights, void* packed, size_t bytes, cudaStream_t s)
{
    SEDT_REQUIRE(bytes >= packed_bytes_, "pack: packed buffer has %zu bytes, need %zu", bytes, packed_bytes_);
    SEDT_REQUIRE(((uintptr_t)packed & 255) == 0, "pack: packed buffer must be 256-byte aligned");
    for (size_t i = 0; i < slots_.size(); ++i)
        SEDT_REQUIRE(weights[i] != nullptr, "pack: weight '%s' is null", slots_[i].name.c_str());
    packed_ = (char*)packed;
    auto W = [&](int slot) { return (const float*)weights[slot]; };
    auto P = [&](size_t off) { return (void*)(packed_ + off); };
    const int dt = act_dt();

    // everything below is recorded into one batched launch (pack.cu: pack_jobs_kernel); SEDT_PACK_BATCH=0 keeps the
    // one-launch-per-tensor path
    static const bool batched = [] { const char* e = getenv("SEDT_PACK_BATCH"); return e == nullptr || atoi(e) != 0; }();
    PackBatch pb(s);
    auto do_cast = [&](const float* in, void* out, int odt, int64_t n) -> int {
        return batched ? pb.cast(in, out, odt, n) : launch_cast(in, out, odt, n, s);
    };
    auto do_bn_fold = [&](int bn_slot, float* scale, float* bias, int n) -> int {
        return batched ? pb.bn_fold(W(bn_slot), W(bn_slot + 1), W(bn_slot + 2), W(bn_slot + 3), scale, bias, n)
                       : launch_bn_fold(W(bn_slot), W(bn_slot + 1), W(bn_slot + 2), W(bn_slot + 3), scale, bias, n, s);
    };
    auto pack_conv = [&](const ConvLayer& L) -> int {
        SEDT_TRY(do_bn_fold(L.bn_slot, (float*)P(L.off_scale), (float*)P(L.off_bias), L.cout));
        // bf16 tier: the FrozenBN scale is folded into the weights before rounding (one rounding instead of
        // two, and the epilogue only adds the bias); the fp32 tier keeps x*scale+bias like the reference
        if (batched)
            return pb.repack_conv(W(L.w_slot), fold_scale() ? W(L.bn_slot) : nullptr, fold_scale() ? W(L.bn_slot + 3) : nullptr,
                                  P(L.off_w), dt, L.cout, L.cin, L.k * L.k);
        return launch_repack_conv(W(L.w_slot), fold_scale() ? (const float*)P(L.off_scale) : nullptr, P(L.off_w), dt, L.cout,
                                  L.cin, L.k, L.k, s);
    };
    auto pack_linear = [&](const Linear& L) -> int {
        SEDT_TRY(do_cast(W(L.w_slot), P(L.off_w), L.f32_only ? DT_F32 : dt, (int64_t)L.in * L.out));
        return do_cast(W(L.b_slot), P(L.off_b), DT_F32, L.out);
    };
    auto pack_norm = [&](const Norm& n) -> int {
        SEDT_TRY(do_cast(W(n.w_slot), P(n.off_g), DT_F32, cfg_.hidden_dim));
        return do_cast(W(n.b_slot), P(n.off_b), DT_F32, cfg_.hidden_dim);
    };
    for (auto& e : enc_) {
        SEDT_TRY(pack_linear(e.attn.in_proj)); SEDT_TRY(pack_linear(e.attn.out_proj));
        SEDT_TRY(pack_linear(e.lin1)); SEDT_TRY(pack_linear(e.lin2));
        SEDT_TRY(pack_norm(e.n1)); SEDT_TRY(pack_norm(e.n2));
    }
    if (cfg_.pre_norm) SEDT_TRY(pack_norm(enc_norm_));
    for (auto& e : dec_) {
        SEDT_TRY(pack_linear(e.self_attn.in_proj)); SEDT_TRY(pack_linear(e.self_attn.out_proj));
        SEDT_TRY(pack_linear(e.cross_attn.in_proj)); SEDT_TRY(pack_linear(e.cross_attn.out_proj));
        SEDT_TRY(pack_linear(e.lin1)); SEDT_TRY(pack_linear(e.lin2));
        SEDT_TRY(pack_norm(e.n1)); SEDT_TRY(pack_norm(e.n2)); SEDT_TRY(pack_norm(e.n3));
    }
    SEDT_TRY(pack_norm(dec_norm_));
    for (size_t l = 0; l < dec_.size(); ++l) {          // [Wk_0; Wk_1; ...] and [Wv_0; Wv_1; ...]
        const int d = cfg_.hidden_dim;
        const size_t es = dtype_size(dt);
        const float* w = W(dec_[l].cross_attn.in_proj.w_slot);
        const float* b = W(dec_[l].cross_attn.in_proj.b_slot);
        SEDT_TRY(do_cast(w + (size_t)d * d, (char*)P(off_ck_w) + l * d * d * es, dt, (int64_t)d * d));
        SEDT_TRY(do_cast(w + (size_t)2 * d * d, (char*)P(off_cv_w) + l * d * d * es, dt, (int64_t)d * d));
        SEDT_TRY(do_cast(b + d, (float*)P(off_ck_b) + l * d, DT_F32, d));
        SEDT_TRY(do_cast(b + 2 * d, (float*)P(off_cv_b) + l * d, DT_F32, d));
    }
    SEDT_TRY(pack_linear(class_embed_)); SEDT_TRY(pack_linear(bbox0_)); SEDT_TRY(pack_linear(bbox1_));
    SEDT_TRY(pack_linear(bbox2_)); SEDT_TRY(pack_linear(input_proj_));
    SEDT_TRY(launch_stem_pack(W(s_conv0_w), W(s_conv0_b), W(s_conv1_w), (float*)P(off_weff), (float*)P(off_sat), s));
    // the stem's scale / bias feed the stem_tc_pack launch right below: folded by their own launch, not by the batch
    SEDT_TRY(launch_bn_fold(W(s_bn1), W(s_bn1 + 1), W(s_bn1 + 2), W(s_bn1 + 3), (float*)P(off_stem_scale),
                            (float*)P(off_stem_bias), 64, s));
    if (cfg_.precision == 1)
        SEDT_TRY(launch_stem_tc_pack(W(s_conv0_w), W(s_conv0_b), W(s_conv1_w), (const float*)P(off_stem_scale), P(off_stem_wtc), s));
    for (auto& b : blocks_) {
        SEDT_TRY(pack_conv(b.c1)); SEDT_TRY(pack_conv(b.c2)); SEDT_TRY(pack_conv(b.c3));
        if (b.has_ds) SEDT_TRY(pack_conv(b.ds));
    }
    SEDT_TRY(do_cast(W(s_query_embed), P(off_query_embed), DT_F32, (int64_t)qall_ * cfg_.hidden_dim));
    if (cfg_.dec_at) SEDT_TRY(pack_linear(weak_));
    if (cfg_.self_sup) {
        SEDT_TRY(pack_linear(patch2query_));
        if (cfg_.feature_recon) { SEDT_TRY(pack_linear(falign0_)); SEDT_TRY(pack_linear(falign1_)); }
    }
    return pb.flush();
}

// ---- launch helpers -------------------------------------------------------------------
int Model::gemm(const ConvGemm& g, cudaStream_t s, bool dry)
{
    if (dry) return SEDT_OK;
    if (cfg_.precision == 1 && cfg_.use_tensor_cores && conv_tc_supported(g)) return launch_conv_tc(g, s);
    return launch_conv_simt(g, s);
}

static inline int conv_out_dim(int n, int k, int stride, int pad, int dil) { return (n + 2 * pad - dil * (k - 1) - 1) / stride + 1; }

ConvGemm Model::conv_gemm(const ConvLayer& L, const void* in, int N, int H, int W, const void* residual, void* out) const
{
    ConvGemm g;
    g.in = in; g.w = packed_ + L.off_w; g.scale = fold_scale() ? nullptr : (const float*)(packed_ + L.off_scale);
    g.bias = (const float*)(packed_ + L.off_bias); g.residual = residual; g.out = out;
    g.in_dt = g.out_dt = act_dt();
    g.B = N; g.H = H; g.W = W; g.Cin = L.cin; g.lda = L.cin;
    g.Ho = conv_out_dim(H, L.k, L.stride, L.pad, L.dil); g.Wo = conv_out_dim(W, L.k, L.stride, L.pad, L.dil);
    g.Cout = L.cout; g.ldc = L.cout; g.ld_res = L.cout;
    g.R = g.S = L.k; g.stride = L.stride; g.dil = L.dil; g.pad = L.pad; g.relu = L.relu;
    return g;
}

int Model::conv(const ConvLayer& L, const void* in, int N, int H, int W, const void* residual, void* out, int* Ho, int* Wo,
                cudaStream_t s, bool dry)
{
    const ConvGemm g = conv_gemm(L, in, N, H, W, residual, out);
    *Ho = g.Ho; *Wo = g.Wo;
    return gemm(g, s, dry);
}

int Model::linear(const Linear& L, int row0, int nrows, const void* in, int in_dt, int lda, int64_t rows,
                  const void* residual, void* out, int out_dt, int ldc, int relu, cudaStream_t s, bool dry)
{
    const int w_dt = L.f32_only ? DT_F32 : act_dt();
    SEDT_REQUIRE(in_dt == w_dt, "linear: input dtype %d does not match packed weight dtype %d", in_dt, w_dt);
    ConvGemm g;
    g.in = in; g.w = packed_ + L.off_w + (size_t)row0 * L.in * dtype_size(w_dt);
    g.scale = nullptr; g.bias = (const float*)(packed_ + L.off_b) + row0; g.residual = residual; g.out = out;
    g.in_dt = in_dt; g.out_dt = out_dt;
    g.B = (int)rows; g.H = g.W = g.Ho = g.Wo = 1; g.Cin = L.in; g.lda = lda;
    g.Cout = nrows; g.ldc = ldc; g.ld_res = ldc; g.relu = relu;
    return gemm(g, s, dry);
}

void Model::feature_shape(int T, int F, bool dilation, int* H, int* W)
{
    int h = conv_out_dim(T, 7, 2, 3, 1), w = conv_out_dim(F, 7, 2, 3, 1);
    h = conv_out_dim(h, 3, 2, 1, 1); w = conv_out_dim(w, 3, 2, 1, 1);
    const int nstride2 = dilation ? 2 : 3;
    for (int i = 0; i < nstride2; ++i) { h = conv_out_dim(h, 3, 2, 1, 1); w = conv_out_dim(w, 3, 2, 1, 1); }
    *H = h; *W = w;
}

// Blocks [b0, b1) of the trunk on N images of size H x W held in `in`.  The last block writes to
// `final_out` when given (otherwise to one of the ping-pong buffers); *out receives that pointer.
int Model::run_blocks(size_t b0, size_t b1, const void* in, int N, int* H, int* W, const BlockBufs& bb, void* final_out,
                      void** out, cudaStream_t s, bool dry)
{
    const void* cur = in;
    void* ping = bb.ping; void* pong = bb.pong;
    for (size_t i = b0; i < b1; ++i) {
        const Block& b = blocks_[i];
        int h1, w1, ho, wo, h3, w3;
        void* dst = (i + 1 == b1 && final_out != nullptr) ? final_out : ping;
        SEDT_TRY(conv(b.c1, cur, N, *H, *W, nullptr, bb.t1, &h1, &w1, s, dry));
        const void* idn = cur;
        if (b.has_ds) {
            int hd, wd;
            SEDT_TRY(conv(b.ds, cur, N, *H, *W, nullptr, bb.ds, &hd, &wd, s, dry));
            idn = bb.ds;
        }
        // layer1: conv2 + conv3 + residual as one launch, the 64-channel h2 tile never leaves the SM (bneck_fused.cu)
        const ConvGemm g2 = conv_gemm(b.c2, bb.t1, N, *H, *W, nullptr, bb.t2);
        const ConvGemm g3 = conv_gemm(b.c3, bb.t2, N, g2.Ho, g2.Wo, idn, dst);
        if (!dry && cfg_.precision == 1 && cfg_.use_tensor_cores && bneck_tail_enabled() && bneck_tail_supported(g2, g3)) {
            SEDT_TRY(launch_bneck_tail(g2, g3, s));
            ho = g2.Ho; wo = g2.Wo;
        } else {
            SEDT_TRY(conv(b.c2, bb.t1, N, *H, *W, nullptr, bb.t2, &ho, &wo, s, dry));
            SEDT_TRY(conv(b.c3, bb.t2, N, ho, wo, idn, dst, &h3, &w3, s, dry));
        }
        cur = dst;
        std::swap(ping, pong);
        *H = ho; *W = wo;
    }
    *out = const_cast<void*>(cur);
    return SEDT_OK;
}

// scratch for blocks [b0, b1) on N images entering at h x w (in_elems: the stage input if it must live in ping/pong)
Model::BlockBufs Model::alloc_block_bufs(size_t b0, size_t b1, int N, int h, int w, size_t extra_out_elems, Arena& ws,
                                         int* Hout, int* Wout, size_t* last_elems)
{
    const size_t es = dtype_size(act_dt());
    size_t max_out = extra_out_elems, max_t1 = 0, max_t2 = 0, max_ds = 0, last = 0;
    for (size_t i = b0; i < b1; ++i) {
        const Block& b = blocks_[i];
        const int ho = conv_out_dim(h, 3, b.c2.stride, b.c2.pad, b.c2.dil), wo = conv_out_dim(w, 3, b.c2.stride, b.c2.pad, b.c2.dil);
        max_t1 = std::max(max_t1, (size_t)N * h * w * b.c1.cout);
        max_t2 = std::max(max_t2, (size_t)N * ho * wo * b.c2.cout);
        last = (size_t)N * ho * wo * b.c3.cout;
        max_out = std::max(max_out, last);
        if (b.has_ds) max_ds = std::max(max_ds, (size_t)N * ho * wo * b.ds.cout);
        h = ho; w = wo;
    }
    BlockBufs bb;
    bb.ping = ws.alloc(max_out * es);
    bb.pong = ws.alloc(max_out * es);
    bb.t1 = ws.alloc(max_t1 * es);
    bb.t2 = ws.alloc(max_t2 * es);
    bb.ds = ws.alloc(max_ds * es);
    *Hout = h; *Wout = w; *last_elems = last;
    return bb;
}

// Backbone: stem -> layer1..layer4.  The bandwidth-bound front (stem, layer1, layer2: 9-10 MB of bf16
// activations per clip between convolutions) runs in chunks of clips small enough for the tensors
// handed from one convolution to the next to stay in the 126 MB L2; layer3/4 (compute-bound) run
// on the whole batch.
int Model::backbone(const float* x, int N, int T, int F, Arena& ws, void** feat, int* Hout, int* Wout, cudaStream_t s, bool dry)
{
    const int dt = act_dt();
    const size_t es = dtype_size(dt);
    SEDT_REQUIRE(F == 64, "stem: the fused stem kernel needs 64 mel bins (config.py n_mels), got F=%d", F);
    const int H0 = conv_out_dim(conv_out_dim(T, 7, 2, 3, 1), 3, 2, 1, 1), W0 = conv_out_dim(conv_out_dim(F, 7, 2, 3, 1), 3, 2, 1, 1);
    static const int env_chunk = [] { const char* e = getenv("SEDT_CHUNK_MB"); return e ? atoi(e) : 0; }();   // 0 = off: measured slower on B200
    // clips per chunk: keep one layer1 output (H0*W0*256 elements per clip) within ~env_chunk MB
    const size_t l1_bytes = (size_t)H0 * W0 * 256 * es;
    int chunk = (int)std::max<size_t>(1, ((size_t)env_chunk << 20) / l1_bytes);
    const size_t n_front = 7;                                   // layer1 (3 blocks) + layer2 (4 blocks)
    if (env_chunk <= 0 || chunk >= N) chunk = N;
    const int nchunks = (int)ceil_div(N, chunk);

    int H2, W2, H4, W4; size_t front_last, back_last;
    // full-batch layer2 output, then layer3/4 scratch; per-chunk front scratch
    BlockBufs fb = alloc_block_bufs(0, n_front, chunk, H0, W0, (size_t)chunk * H0 * W0 * 64, ws, &H2, &W2, &front_last);
    const size_t l2_per_clip = front_last / chunk;
    void* l2_out = ws.alloc((size_t)N * l2_per_clip * es);
    BlockBufs bk = alloc_block_bufs(n_front, blocks_.size(), N, H2, W2, 0, ws, &H4, &W4, &back_last);

    StemWeights sw{(const float*)(packed_ + off_weff), (const float*)(packed_ + off_sat),
                   (const float*)(packed_ + off_stem_scale), (const float*)(packed_ + off_stem_bias)};
    for (int c = 0; c < nchunks; ++c) {
        const int n0 = c * chunk, nc = std::min(chunk, N - n0);
        if (!dry) {
            if (cfg_.precision == 1 && cfg_.use_tensor_cores)
                SEDT_TRY(launch_stem_tc(x + (size_t)n0 * T * F, packed_ + off_stem_wtc, sw.bias, sw.scale, sw.sat, fb.pong, nc, T, F, s));
            else
                SEDT_TRY(launch_stem(x + (size_t)n0 * T * F, sw, fb.pong, dt, nc, T, F, s));
        }
        int h = H0, w = W0;
        void* o = nullptr;
        // stem output sits in `pong`; the first block writes to `ping`
        SEDT_TRY(run_blocks(0, n_front, fb.pong, nc, &h, &w, fb, (char*)l2_out + (size_t)n0 * l2_per_clip * es, &o, s, dry));
    }
    int h = H2, w = W2;
    SEDT_TRY(run_blocks(n_front, blocks_.size(), l2_out, N, &h, &w, bk, nullptr, feat, s, dry));
    *Hout = h; *Wout = w;
    return SEDT_OK;
}

// Cross attention with K / V already projected (all decoder layers at once, see forward()).
int Model::cross_mha(const Mha& A, const void* q_in, const void* Kp, const void* Vp, int ldkv, int64_t B, int Lq, int Lk,
                     const uint8_t* kpm, float* resid32, float* out32, Arena& ws, cudaStream_t s, bool dry)
{
    const int d = cfg_.hidden_dim, dt = act_dt();
    const size_t es = dtype_size(dt);
    const size_t mark = ws.off;
    const float scale = (float)std::sqrt(1.0 / (double)(d / cfg_.nheads));
    void* qb = ws.alloc((size_t)B * Lq * d * es);
    void* ao = ws.alloc((size_t)B * Lq * d * es);
    SEDT_TRY(linear(A.in_proj, 0, d, q_in, dt, d, B * Lq, nullptr, qb, dt, d, 0, s, dry));
    if (!dry) SEDT_TRY(launch_attention(qb, d, Kp, ldkv, Vp, ldkv, ao, d, dt, kpm, nullptr, (int)B, cfg_.nheads, Lq, Lk, scale, s));
    SEDT_TRY(linear(A.out_proj, 0, d, ao, dt, d, B * Lq, resid32, out32, DT_F32, d, 0, s, dry));
    ws.off = mark;
    return SEDT_OK;
}

int Model::mha(const Mha& A, bool self_attn, const void* q_in, const void* k_in, const void* v_in, int64_t B, int Lq, int Lk,
               const uint8_t* kpm, const float* amask, float* resid32, float* out32, Arena& ws, cudaStream_t s, bool dry,
               const Norm* post_norm, void* post_norm_out, bool* post_norm_done)
{
    if (post_norm_done != nullptr) *post_norm_done = false;
    const int d = cfg_.hidden_dim, dt = act_dt();
    const size_t es = dtype_size(dt);
    const size_t mark = ws.off;
    const float scale = (float)std::sqrt(1.0 / (double)(d / cfg_.nheads));
    // self-attention of a <= 128-token clip with q = k: the whole block (three projections, attention core, out_proj + residual)
    // as one launch with every intermediate on chip (enc_attn_fused.cu); the launches below remain for the other shapes,
    // the SP-SEDT decoder mask and the fp32 tier
    // (one 128-row tile per clip: pays from ~64 tokens per clip on; the decoder's 11 / 21 queries pack 6 clips into one GEMM tile
    // in the separate launches below -- measured 47 vs 51 us for the decoder self-attention at B = 256)
    if (!dry && self_attn && q_in == k_in && amask == nullptr && Lq == Lk && Lq >= 64 && resid32 == out32 && cfg_.use_tensor_cores &&
        enc_attn_fused_enabled() && !A.in_proj.f32_only &&
        enc_attn_fused_supported(d, cfg_.nheads, Lq, v_in, q_in, packed_ + A.in_proj.off_w, packed_ + A.out_proj.off_w, out32, dt))
    {
        // the layer's next LayerNorm (norm2 of the pre-norm encoder layer) rides in the same launch when the caller asks for it
        // (SEDT_ENC_ATTN_LN=1; off by default: measured on B200 at B = 256 the fold makes the forward 40 us SLOWER -- 4.445 vs
        // 4.405 ms -- although it removes six 14 us launches: the two extra passes over the staging tile and the row-strided
        // bf16 stores sit on the per-clip critical path between the output stores and the next clip's loads)
        static const bool ln_on = [] { const char* e = getenv("SEDT_ENC_ATTN_LN"); return e != nullptr && atoi(e) != 0; }();
        const bool ln = ln_on && post_norm != nullptr && post_norm_out != nullptr && post_norm_done != nullptr;
        if (ln) *post_norm_done = true;
        return launch_enc_attn_fused(v_in, q_in, packed_ + A.in_proj.off_w, (const float*)(packed_ + A.in_proj.off_b),
                                     packed_ + A.out_proj.off_w, (const float*)(packed_ + A.out_proj.off_b), kpm, out32, (int)B, Lq,
                                     scale, s, ln ? (const float*)(packed_ + post_norm->off_g) : nullptr,
                                     ln ? (const float*)(packed_ + post_norm->off_b) : nullptr, ln ? post_norm_out : nullptr);
    }
    const void *Qp, *Kp, *Vp; int ldq, ldk;
    void* vbuf = ws.alloc((size_t)B * Lk * d * es);
    SEDT_TRY(linear(A.in_proj, 2 * d, d, v_in, dt, d, B * Lk, nullptr, vbuf, dt, d, 0, s, dry));
    Vp = vbuf;
    if (self_attn) {           // q and k share their input: one GEMM for Q and K (N = 512)
        void* qk = ws.alloc((size_t)B * Lq * 2 * d * es);
        SEDT_TRY(linear(A.in_proj, 0, 2 * d, q_in, dt, d, B * Lq, nullptr, qk, dt, 2 * d, 0, s, dry));
        Qp = qk; Kp = (const char*)qk + (size_t)d * es; ldq = ldk = 2 * d;
    } else {
        void* qb = ws.alloc((size_t)B * Lq * d * es);
        void* kb = ws.alloc((size_t)B * Lk * d * es);
        SEDT_TRY(linear(A.in_proj, 0, d, q_in, dt, d, B * Lq, nullptr, qb, dt, d, 0, s, dry));
        SEDT_TRY(linear(A.in_proj, d, d, k_in, dt, d, B * Lk, nullptr, kb, dt, d, 0, s, dry));
        Qp = qb; Kp = kb; ldq = ldk = d;
    }
    void* ao = ws.alloc((size_t)B * Lq * d * es);
    if (!dry) SEDT_TRY(launch_attention(Qp, ldq, Kp, ldk, Vp, d, ao, d, dt, kpm, amask, (int)B, cfg_.nheads, Lq, Lk, scale, s));
    SEDT_TRY(linear(A.out_proj, 0, d, ao, dt, d, B * Lq, resid32, out32, DT_F32, d, 0, s, dry));
    ws.off = mark;
    return SEDT_OK;
}

// ---- forward ---------------------------------------------------------------------------
int Model::forward(const float* x, const uint8_t* mask, int B, int T, int F, const float* patches, int P, int PT,
                   Arena& ws, const ForwardOut& out, cudaStream_t s, bool dry)
{
    SEDT_REQUIRE(cfg_.hidden_dim == 256 && cfg_.nheads == 8, "forward: kernels are built for hidden_dim 256 / 8 heads");
    SEDT_REQUIRE(B >= 1 && T >= 1, "forward: B=%d T=%d", B, T);
    SEDT_REQUIRE(dry || packed_ != nullptr, "forward: sedt_model_pack has not been called");
    SEDT_REQUIRE(!cfg_.self_sup || (patches != nullptr || dry), "forward: SP-SEDT needs patches");
    SEDT_REQUIRE(!(cfg_.self_sup && cfg_.dec_at), "forward: SP-SEDT has no audio query (sedt/spsedt.py:59,72)");
    const int d = cfg_.hidden_dim, ff = cfg_.dim_feedforward, dt = act_dt();
    const size_t es = dtype_size(dt);
    auto P_ = [&](size_t off) { return (const float*)(packed_ + off); };
    auto LN = [&](const Norm& n, const float* xin, const float* pos, int64_t pos_rows, void* y, void* ypos, float* y32,
                  int64_t rows) -> int {
        if (dry) return SEDT_OK;
        return launch_layernorm(xin, P_(n.off_g), P_(n.off_b), pos, pos_rows, y, ypos, y32, dt, rows, s);
    };

    // ---- backbone on the clips
    void* feat = nullptr; int H = 0, W = 0;
    SEDT_TRY(backbone(x, B, T, F, ws, &feat, &H, &W, s, dry));
    const int S = H * W;
    const int64_t rows = (int64_t)B * S;
    if (out.feat != nullptr && !dry) {
        SEDT_REQUIRE(dt == DT_F32, "forward: feature-map copy-out is only available in the fp32 tier");
        SEDT_CHECK_CUDA(cudaMemcpyAsync(out.feat, feat, (size_t)rows * 2048 * 4, cudaMemcpyDeviceToDevice, s));
    }

    // ---- SP-SEDT: backbone on the patches, avgpool, patch2query (spsedt.py:46-57)
    float* qpos_batched = nullptr;
    int Qall = qall_;
    if (cfg_.self_sup) {
        SEDT_REQUIRE(P >= 1 && cfg_.num_queries % cfg_.num_patches == 0, "forward: bad patch configuration");
        const int qpp = cfg_.num_queries / cfg_.num_patches;
        Qall = P * qpp;                                     // eval branch, spsedt.py:72
        SEDT_REQUIRE(Qall <= cfg_.num_queries, "forward: %d patches exceed num_patches", P);
        float* gt = out.gt_feature != nullptr ? out.gt_feature : (float*)ws.alloc((size_t)B * P * 2048 * 4);
        float* pq = (float*)ws.alloc((size_t)B * P * d * 4);
        qpos_batched = (float*)ws.alloc((size_t)B * Qall * d * 4);
        const size_t mark = ws.off;
        void* pfeat = nullptr; int ph = 0, pw = 0;
        SEDT_TRY(backbone(patches, B * P, PT, F, ws, &pfeat, &ph, &pw, s, dry));
        if (!dry) {
            SEDT_TRY(launch_avgpool(pfeat, dt, gt, B * P, ph * pw, 2048, s));
            SEDT_TRY(linear(patch2query_, 0, d, gt, DT_F32, 2048, (int64_t)B * P, nullptr, pq, DT_F32, d, 0, s, dry));
            SEDT_TRY(launch_patch_query(pq, P_(off_query_embed), qpos_batched, B, P, qpp, cfg_.dec_at ? 1 : 0, s));
        }
        ws.off = mark;
    }

    // ---- padding mask and position table (backbone.py:81, position_encoding.py:28-47)
    uint8_t* mask_ds = nullptr; float* pos; int64_t pos_rows;
    if (mask != nullptr) {
        mask_ds = (uint8_t*)ws.alloc((size_t)rows);
        pos = (float*)ws.alloc((size_t)rows * d * 4);
        pos_rows = rows;
        if (!dry) {
            SEDT_TRY(launch_mask_downsample(mask, mask_ds, B, T, F, H, W, s));
            SEDT_TRY(launch_pos_table(mask_ds, pos, B, H, W, s));
        }
    } else {
        pos = (float*)ws.alloc((size_t)S * d * 4);
        pos_rows = S;
        if (!dry) SEDT_TRY(launch_pos_table(nullptr, pos, 1, H, W, s));
    }

    // ---- input_proj (sedt.py:88): NHWC feature map == token-major [B, S, 2048]
    float* x32 = (float*)ws.alloc((size_t)rows * d * 4);
    SEDT_TRY(linear(input_proj_, 0, d, feat, dt, 2048, rows, nullptr, x32, DT_F32, d, 0, s, dry));

    // FFN: x32 += linear2(relu(linear1(in))).  Where it pays (ffn_fused_preferred: whole rounds of 128-row tiles over the SMs) the
    // fused kernel keeps the hidden activation on chip (90 vs 97 us per encoder layer at B = 256); SEDT_FFN_FUSED=0 / 1 forces it
    // off / on for every size.
    static const int ffn_mode = [] { const char* e = getenv("SEDT_FFN_FUSED"); return e == nullptr ? -1 : atoi(e); }();
    auto ffn = [&](const Linear& l1, const Linear& l2, const void* in, int64_t nrows, void* hidden, float* x) -> int {
        const bool ffn_fused = ffn_mode == 1 || (ffn_mode == -1 && ffn_fused_preferred(nrows));
        if (ffn_fused && !dry && dt == DT_BF16 && cfg_.use_tensor_cores && !l1.f32_only && !l2.f32_only &&
            ffn_fused_supported(d, ff, nrows, in, packed_ + l1.off_w, packed_ + l2.off_w, x, x, d, d))
            return launch_ffn_fused(in, packed_ + l1.off_w, (const float*)(packed_ + l1.off_b), packed_ + l2.off_w,
                                    (const float*)(packed_ + l2.off_b), x, d, x, d, nrows, ff, s);
        SEDT_TRY(linear(l1, 0, ff, in, dt, d, nrows, nullptr, hidden, dt, ff, 1, s, dry));
        return linear(l2, 0, d, hidden, dt, ff, nrows, x, x, DT_F32, d, 0, s, dry);
    };

    // ---- encoder (transformer.py:98-111, :177-204)
    void* na = ws.alloc((size_t)rows * d * es);
    void* nap = ws.alloc((size_t)rows * d * es);
    void* na2 = ws.alloc((size_t)rows * d * es);
    void* ffh = ws.alloc((size_t)rows * ff * es);
    for (auto& e : enc_) {
        if (cfg_.pre_norm) {
            SEDT_TRY(LN(e.n1, x32, pos, pos_rows, na, nap, nullptr, rows));
            bool ln2_done = false;               // the fused attention block also writes LN2(x) (into na2: na is still being read)
            SEDT_TRY(mha(e.attn, true, nap, nap, na, B, S, S, mask_ds, nullptr, x32, x32, ws, s, dry, &e.n2, na2, &ln2_done));
            if (!ln2_done) SEDT_TRY(LN(e.n2, x32, nullptr, 1, na2, nullptr, nullptr, rows));
            SEDT_TRY(ffn(e.lin1, e.lin2, na2, rows, ffh, x32));
        } else {
            if (!dry) SEDT_TRY(launch_cast_addpos(x32, pos, pos_rows, na, nap, dt, rows, s));
            SEDT_TRY(mha(e.attn, true, nap, nap, na, B, S, S, mask_ds, nullptr, x32, x32, ws, s, dry));
            SEDT_TRY(LN(e.n1, x32, nullptr, 1, na, nullptr, x32, rows));
            SEDT_TRY(ffn(e.lin1, e.lin2, na, rows, ffh, x32));
            SEDT_TRY(LN(e.n2, x32, nullptr, 1, nullptr, nullptr, x32, rows));
        }
    }
    // memory (T-typed) and memory + pos for the decoder's cross attention
    void* mem = ws.alloc((size_t)rows * d * es);
    void* mempos = ws.alloc((size_t)rows * d * es);
    if (cfg_.pre_norm) SEDT_TRY(LN(enc_norm_, x32, pos, pos_rows, mem, mempos, out.memory, rows));
    else if (!dry) {
        SEDT_TRY(launch_cast_addpos(x32, pos, pos_rows, mem, mempos, dt, rows, s));
        if (out.memory != nullptr) SEDT_CHECK_CUDA(cudaMemcpyAsync(out.memory, x32, (size_t)rows * d * 4, cudaMemcpyDeviceToDevice, s));
    }

    // cross-attention keys / values of every decoder layer: K_l = (memory + pos) Wk_l^T, V_l = memory Wv_l^T
    const int Dn_ = (int)dec_.size();
    void* ck_all = ws.alloc((size_t)rows * Dn_ * d * es);
    void* cv_all = ws.alloc((size_t)rows * Dn_ * d * es);
    {
        ConvGemm g;
        g.in_dt = g.out_dt = dt; g.B = (int)rows; g.H = g.W = g.Ho = g.Wo = 1; g.Cin = d; g.lda = d;
        g.Cout = Dn_ * d; g.ldc = Dn_ * d; g.ld_res = Dn_ * d;
        g.in = mempos; g.w = packed_ + off_ck_w; g.bias = (const float*)(packed_ + off_ck_b); g.out = ck_all;
        SEDT_TRY(gemm(g, s, dry));
        g.in = mem; g.w = packed_ + off_cv_w; g.bias = (const float*)(packed_ + off_cv_b); g.out = cv_all;
        SEDT_TRY(gemm(g, s, dry));
    }

    // ---- decoder (transformer.py:123-152, :240-284)
    const int64_t qrows = (int64_t)B * Qall;
    const float* qpos = cfg_.self_sup ? qpos_batched : P_(off_query_embed);
    const int64_t qpos_rows = cfg_.self_sup ? qrows : Qall;
    float* amask = nullptr;
    if (cfg_.self_sup) {                                    // block-diagonal 0/-inf mask, spsedt.py:27-32
        amask = (float*)ws.alloc((size_t)Qall * Qall * 4);
        if (!dry) SEDT_TRY(launch_blockdiag_mask(amask, Qall, cfg_.num_queries / cfg_.num_patches, s));
    }
    float* t32 = (float*)ws.alloc((size_t)qrows * d * 4);
    void* da = ws.alloc((size_t)qrows * d * es);
    void* dap = ws.alloc((size_t)qrows * d * es);
    void* dffh = ws.alloc((size_t)qrows * ff * es);
    if (!dry) SEDT_TRY(launch_fill_zero(t32, (size_t)qrows * d * 4, s));
    // decoder states in the tier dtype (GEMM operand of the box / feature MLPs) next to the fp32 copy the caller gets
    void* hs_t = dt == DT_F32 ? (void*)out.hs : ws.alloc((size_t)dec_.size() * qrows * d * es);
    for (size_t l = 0; l < dec_.size(); ++l) {
        auto& e = dec_[l];
        float* hs_l = out.hs + l * (size_t)qrows * d;
        void* hs_tl = dt == DT_F32 ? nullptr : (void*)((char*)hs_t + l * (size_t)qrows * d * es);
        if (cfg_.pre_norm) {
            SEDT_TRY(LN(e.n1, t32, qpos, qpos_rows, da, dap, nullptr, qrows));
            SEDT_TRY(mha(e.self_attn, true, dap, dap, da, B, Qall, Qall, nullptr, amask, t32, t32, ws, s, dry));
            SEDT_TRY(LN(e.n2, t32, qpos, qpos_rows, nullptr, dap, nullptr, qrows));
            SEDT_TRY(cross_mha(e.cross_attn, dap, (const char*)ck_all + l * d * es, (const char*)cv_all + l * d * es, Dn_ * d, B, Qall, S,
                               mask_ds, t32, t32, ws, s, dry));
            SEDT_TRY(LN(e.n3, t32, nullptr, 1, da, nullptr, nullptr, qrows));
            SEDT_TRY(ffn(e.lin1, e.lin2, da, qrows, dffh, t32));
        } else {
            if (!dry) SEDT_TRY(launch_cast_addpos(t32, qpos, qpos_rows, da, dap, dt, qrows, s));
            SEDT_TRY(mha(e.self_attn, true, dap, dap, da, B, Qall, Qall, nullptr, amask, t32, t32, ws, s, dry));
            SEDT_TRY(LN(e.n1, t32, qpos, qpos_rows, nullptr, dap, t32, qrows));
            SEDT_TRY(cross_mha(e.cross_attn, dap, (const char*)ck_all + l * d * es, (const char*)cv_all + l * d * es, Dn_ * d, B, Qall, S,
                               mask_ds, t32, t32, ws, s, dry));
            SEDT_TRY(LN(e.n2, t32, nullptr, 1, da, nullptr, t32, qrows));
            SEDT_TRY(ffn(e.lin1, e.lin2, da, qrows, dffh, t32));
            SEDT_TRY(LN(e.n3, t32, nullptr, 1, nullptr, nullptr, t32, qrows));
        }
        SEDT_TRY(LN(dec_norm_, t32, nullptr, 1, hs_tl, nullptr, hs_l, qrows));     // transformer.py:140-147
    }

    // ---- heads (sedt.py:89-95, spsedt.py:77-85): fp32 CUDA-core GEMMs on the tiny [D*B*Q, 256] matrix
    const int Dn = (int)dec_.size();
    const int64_t hrows = (int64_t)Dn * qrows;
    const int ncls = cfg_.self_sup ? 1 : cfg_.num_classes, C1 = ncls + 1;
    const int start = cfg_.dec_at ? 1 : 0;
    float* cls_raw = (float*)ws.alloc((size_t)hrows * C1 * 4);
    void* h1 = ws.alloc((size_t)hrows * d * es);
    float* h2 = (float*)ws.alloc((size_t)hrows * d * 4);
    float* box_raw = (float*)ws.alloc((size_t)hrows * 2 * 4);
    float* weak_raw = cfg_.dec_at ? (float*)ws.alloc((size_t)B * ncls * 4) : nullptr;
    static const bool heads_simt = [] { const char* e = getenv("SEDT_HEADS_SIMT"); return e != nullptr && e[0] == '1'; }();
    SEDT_TRY(linear(bbox0_, 0, d, hs_t, dt, d, hrows, nullptr, h1, dt, d, 1, s, dry));
    SEDT_TRY(linear(bbox1_, 0, d, h1, dt, d, hrows, nullptr, h2, DT_F32, d, 1, s, dry));
    if (!heads_simt) {
        auto Pf = [&](size_t off) { return (const float*)(packed_ + off); };
        if (!dry)
            SEDT_TRY(launch_heads_out(out.hs, h2, Pf(class_embed_.off_w), Pf(class_embed_.off_b), Pf(bbox2_.off_w), Pf(bbox2_.off_b),
                                      cfg_.dec_at ? Pf(weak_.off_w) : nullptr, cfg_.dec_at ? Pf(weak_.off_b) : nullptr, out.logits,
                                      out.boxes, cfg_.dec_at ? out.at : nullptr, Dn, B, Qall, start, C1, ncls, s));
    } else {
        SEDT_TRY(linear(class_embed_, 0, C1, out.hs, DT_F32, d, hrows, nullptr, cls_raw, DT_F32, C1, 0, s, dry));
        SEDT_TRY(linear(bbox2_, 0, 2, h2, DT_F32, d, hrows, nullptr, box_raw, DT_F32, 2, 0, s, dry));
        if (cfg_.dec_at)     // slot 0 of the last layer: rows at stride Qall*d (sedt.py:92)
            SEDT_TRY(linear(weak_, 0, ncls, out.hs + (size_t)(Dn - 1) * qrows * d, DT_F32, Qall * d, B, nullptr, weak_raw,
                            DT_F32, ncls, 0, s, dry));
        if (!dry)
            SEDT_TRY(launch_heads_finalize(cls_raw, box_raw, weak_raw, out.logits, out.boxes, cfg_.dec_at ? out.at : nullptr,
                                           Dn, B, Qall, start, C1, ncls, s));
    }
    if (cfg_.self_sup && cfg_.feature_recon && out.pred_feature != nullptr) {
        SEDT_TRY(linear(falign0_, 0, d, hs_t, dt, d, hrows, nullptr, h1, dt, d, 1, s, dry));
        SEDT_TRY(linear(falign1_, 0, 2048, h1, dt, d, hrows, nullptr, out.pred_feature, DT_F32, 2048, 0, s, dry));
    }
    if (ws.overflow && !dry) {
        set_error("forward: workspace too small (%zu bytes needed, %zu given)", ws.peak, ws.cap);
        return SEDT_ERR_WORKSPACE;
    }
    return SEDT_OK;
}

}  // namespace sedt
