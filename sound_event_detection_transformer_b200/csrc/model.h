// Host-side runtime of the SEDT forward hot path: owns the layer table (which
// reference state_dict entry feeds which kernel), the packed-weight layout and
// the launch sequence.  All device memory belongs to the caller (PyTorch):
// weights, the packed buffer, the workspace and the outputs are raw pointers.
#pragma once
#include "kernels.h"
#include <string>
#include <unordered_map>
#include <vector>

namespace sedt {

struct Config {                 // mirrors include/sedt_b200.h : sedt_config
    int enc_layers, dec_layers, num_queries, num_classes, hidden_dim, nheads, dim_feedforward;
    int dec_at, pre_norm, dilation, self_sup, feature_recon, num_patches, aux_loss;
    int precision;              // 0 = fp32 CUDA-core tier, 1 = bf16 tcgen05 tier
    int use_tensor_cores;       // bf16 tier only: 0 forces the CUDA-core kernels (debug / A-B tests)
};

struct WeightSlot {
    std::string name;           // reference state_dict key
    int64_t numel;
};

struct ConvLayer {
    int cin, cout, k, stride, dil, pad, relu;
    int w_slot, bn_slot;        // bn_slot: weight, +1 bias, +2 running_mean, +3 running_var
    size_t off_w, off_scale, off_bias;
};

struct Block { ConvLayer c1, c2, c3, ds; bool has_ds; };

struct Linear {                 // y = x W^T + b ; W [out, in]
    int in, out, w_slot, b_slot;
    size_t off_w, off_b;        // off_w in tier dtype (or fp32 when f32_only)
    bool f32_only;
};

struct Norm { int w_slot, b_slot; size_t off_g, off_b; };

struct Mha { Linear in_proj, out_proj; };        // in_proj rows: [Wq; Wk; Wv]

struct EncLayer { Mha attn; Linear lin1, lin2; Norm n1, n2; };
struct DecLayer { Mha self_attn, cross_attn; Linear lin1, lin2; Norm n1, n2, n3; };

struct Arena {
    char* base; size_t off, cap, peak; bool overflow;
    Arena(void* b, size_t c) : base((char*)b), off(0), cap(c), peak(0), overflow(false) {}
    void* alloc(size_t bytes) {
        size_t o = (off + 255) & ~(size_t)255;
        off = o + bytes;
        if (off > peak) peak = off;
        if (base == nullptr || off > cap) { overflow = true; return base == nullptr ? nullptr : base; }
        return base + o;
    }
};

struct ForwardOut {             // all fp32, caller-allocated
    float* hs;                  // [D, B, Qall, 256] decoder states after decoder.norm
    float* logits;              // [D, B, Q, C+1]
    float* boxes;               // [D, B, Q, 2]   sigmoid(center, width)
    float* at;                  // [B, C] or null (dec_at only)
    float* memory;              // optional [B, S, 256] encoder output (fp32 copy) or null
    float* pred_feature;        // SP-SEDT: [D, B, Q, 2048] or null
    float* gt_feature;          // SP-SEDT: [B*P, 2048] or null
    float* feat;                // optional backbone feature map copy [B, H, W, 2048] fp32 (tests) or null
};

class Model {
public:
    explicit Model(const Config& c);
    const Config& cfg() const { return cfg_; }
    const std::vector<WeightSlot>& slots() const { return slots_; }
    size_t packed_bytes() const { return packed_bytes_; }
    int act_dt() const { return cfg_.precision ? DT_BF16 : DT_F32; }
    bool fold_scale() const { return cfg_.precision != 0; }

    int pack(const void* const* weights, void* packed, size_t bytes, cudaStream_t stream);
    // dry = true only sizes the workspace (no launches)
    int forward(const float* x, const uint8_t* mask, int B, int T, int F, const float* patches, int P, int PT,
                Arena& ws, const ForwardOut& out, cudaStream_t stream, bool dry);
    static void feature_shape(int T, int F, bool dilation, int* H, int* W);
    // data-parallel training: backward() records this event on its stream once every gradient outside the backbone (transformer,
    // heads, input_proj, query_embed) is final, so that the caller can all-reduce that bucket while the backbone backward runs
    void set_bucket_event(cudaEvent_t ev) { bucket_event_ = ev; }

    // ---- training step (train.cu): bf16 tier, pre-norm; SEDT (backbone trainable from layer2 / conv0) and SP-SEDT pretraining
    // (frozen backbone, train_spsedt.py:50; patches + query-drop mask through SpTrain)
    struct SpTrain {                       // SP-SEDT extras of one training step (all null / 0 for SEDT)
        const float* patches = nullptr;    // [B, P, 1, PT, F]
        int P = 0, PT = 0;
        const uint8_t* query_keep = nullptr;     // [B, Q] 1 = the patch feature is added to this query (spsedt.py:65-67)
        const float* d_pred_feature = nullptr;   // backward: gradient of pred_feature [D, B, Q, 2048] or null
    };
    int64_t tape_bytes(int B, int T, int F, bool has_mask, int P = 0, int PT = 0);
    int64_t backward_workspace_bytes(int B, int T, int F, int P = 0, int PT = 0);
    int64_t grad_offset(int slot) const;       // element offset of a state_dict entry in the flat fp32 gradient buffer
    int64_t grad_numel() const;
    int forward_train(const float* x, const uint8_t* mask, int B, int T, int F, void* tape, size_t tape_bytes,
                      const ForwardOut& out, float dropout, unsigned long long seed, cudaStream_t stream, const SpTrain* sp = nullptr);
    int backward(const void* const* weights, const float* x, const uint8_t* mask, int B, int T, int F, void* tape,
                 size_t tape_bytes, void* workspace, size_t ws_bytes, const float* d_logits, const float* d_boxes,
                 const float* d_at, float* grads, int train_backbone, float dropout, cudaStream_t stream, const SpTrain* sp = nullptr);

private:
    struct BlockTape; struct EncTape; struct DecTape; struct Tape; struct BwdBufs;
    void tape_layout(int B, int T, int F, bool has_mask, int P, int PT, Arena& a, Tape& tp);
    void bwd_layout(int B, const Tape& tp, Arena& a, BwdBufs& bb) const;
    int check_train_config() const;
    int add_slot(const std::string& name, int64_t numel);
    size_t reserve(size_t bytes);
    ConvLayer make_conv(const std::string& conv, const std::string& bn, int cin, int cout, int k, int stride, int dil,
                        int pad, int relu);
    Linear make_linear(const std::string& name, int in, int out, bool f32_only);
    Linear make_linear_slots(int w_slot, int b_slot, int in, int out, bool f32_only);
    Norm make_norm(const std::string& name);
    Mha make_mha(const std::string& name);

    int gemm(const ConvGemm& g, cudaStream_t s, bool dry);
    ConvGemm conv_gemm(const ConvLayer& L, const void* in, int N, int H, int W, const void* residual, void* out) const;
    int conv(const ConvLayer& L, const void* in, int N, int H, int W, const void* residual, void* out, int* Ho, int* Wo,
             cudaStream_t s, bool dry);
    // rows x L.in -> rows x L.out.  row0/nrows select a slice of W's rows (packed in_proj).
    int linear(const Linear& L, int row0, int nrows, const void* in, int in_dt, int lda, int64_t rows, const void* residual,
               void* out, int out_dt, int ldc, int relu, cudaStream_t s, bool dry);
    struct BlockBufs { void *ping, *pong, *t1, *t2, *ds; };
    BlockBufs alloc_block_bufs(size_t b0, size_t b1, int N, int h, int w, size_t extra_out_elems, Arena& ws, int* Hout, int* Wout,
                               size_t* last_elems);
    int run_blocks(size_t b0, size_t b1, const void* in, int N, int* H, int* W, const BlockBufs& bb, void* final_out, void** out,
                   cudaStream_t s, bool dry);
    int backbone(const float* x, int N, int T, int F, Arena& ws, void** feat, int* H, int* W, cudaStream_t s, bool dry);
    int cross_mha(const Mha& A, const void* q_in, const void* Kp, const void* Vp, int ldkv, int64_t B, int Lq, int Lk,
                  const uint8_t* kpm, float* resid32, float* out32, Arena& ws, cudaStream_t s, bool dry);
    int mha(const Mha& A, bool self_attn, const void* q_in, const void* k_in, const void* v_in, int64_t B, int Lq, int Lk,
            const uint8_t* kpm, const float* amask, float* resid32, float* out32, Arena& ws, cudaStream_t s, bool dry,
            const Norm* post_norm = nullptr, void* post_norm_out = nullptr, bool* post_norm_done = nullptr);

    Config cfg_;
    std::vector<WeightSlot> slots_;
    size_t packed_bytes_ = 0;
    char* packed_ = nullptr;

    // layer table
    int s_conv0_w, s_conv0_b, s_conv1_w, s_bn1;
    size_t off_weff, off_sat, off_stem_scale, off_stem_bias, off_stem_wtc;
    std::vector<Block> blocks_;
    Linear input_proj_;
    std::vector<EncLayer> enc_;
    Norm enc_norm_;
    std::vector<DecLayer> dec_;
    Norm dec_norm_;
    Linear class_embed_, bbox0_, bbox1_, bbox2_, weak_, patch2query_, falign0_, falign1_;
    int s_query_embed; size_t off_query_embed;
    // cross-attention K / V projection weights of all decoder layers, concatenated ([D*256, 256]) so that the
    // encoder memory is projected by two GEMMs instead of 2*D
    size_t off_ck_w, off_ck_b, off_cv_w, off_cv_b;
    int qall_;
    // data-gradient weight re-layouts of one backward pass, recorded by the first call and replayed as one batched launch
    // by the following ones while the weight / workspace addresses stay put (train.cu: Model::backward)
    struct DgradPlan { std::vector<DgradJob> jobs; unsigned long long key = 0; bool ready = false; };
    DgradPlan dgrad_plan_;
    cudaEvent_t bucket_event_ = nullptr;
    std::unordered_map<const void*, unsigned long long> rng_tapes_;   // tape -> seed its dropout RNG state {seed, step} was initialised with
};

}  // namespace sedt
