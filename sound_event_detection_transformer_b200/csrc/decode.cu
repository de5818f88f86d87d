// The step right after the eval forward, on the device: PostProcess.forward (sedt/sedt.py:359-396: softmax, the at_m 1/2/3
// audio-tag fusion, per-query score / label, (center, width) -> (onset, offset) in seconds) followed by
// BoxEncoder.decode_strong with del_overlap (utilities/BoxEncoder.py:179-226: score >= threshold, duration >= 0.2 s,
// per class sort by onset and drop the lower-scored event of every overlapping neighbour pair).  The reference does this
// with per-clip Python / numpy loops (engine.py:277-291); here one warp handles one clip and one launch handles the batch.
//
// Output per clip: the PostProcess triplet (scores, labels, boxes in seconds) and the decoded event list in the
// reference's order (classes in order of first appearance among the kept queries, events of a class by onset).
#include "common.cuh"
#include "kernels.h"
#include <math_constants.h>

namespace sedt {
namespace {

constexpr int kDecWarps = 4;

__global__ void __launch_bounds__(kDecWarps * 32)
decode_events_kernel(const float* __restrict__ logits, const float* __restrict__ boxes, const float* __restrict__ sizes,
                     const float* __restrict__ tags, int B, int Q, int C1, int at_m, float fuse_threshold, int is_semi,
                     float score_threshold, float min_duration,
                     float* __restrict__ out_scores, int64_t* __restrict__ out_labels, float* __restrict__ out_boxes,
                     int32_t* __restrict__ ev_class, float* __restrict__ ev_onset, float* __restrict__ ev_offset,
                     float* __restrict__ ev_score, int32_t* __restrict__ ev_count)
{
    extern __shared__ float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x * kDecWarps + warp;
    if (b >= B) return;
    const int C = C1 - 1;
    const int per_warp = Q * C1 + 5 * Q;
    float* prob = smem + warp * per_warp;          // [Q][C1]
    float* sc = prob + Q * C1;                     // [Q] score
    float* on = sc + Q;                            // [Q] onset (seconds)
    float* off = on + Q;                           // [Q] offset
    int* lab = reinterpret_cast<int*>(off + Q);    // [Q] label
    int* ord = lab + Q;                            // [Q] scratch: members of one class
    const float* lg = logits + (size_t)b * Q * C1;

    // F.softmax(out_logits, -1)
    for (int q = lane; q < Q; q += 32) {
        float m = -CUDART_INF_F;
        for (int c = 0; c < C1; ++c) m = fmaxf(m, lg[q * C1 + c]);
        float s = 0.f;
        for (int c = 0; c < C1; ++c) { const float e = expf(lg[q * C1 + c] - m); prob[q * C1 + c] = e; s += e; }
        for (int c = 0; c < C1; ++c) prob[q * C1 + c] = prob[q * C1 + c] / s;
    }
    __syncwarp();
    // audio-tag fusion (sedt.py:371-387): per class the strongest query (first maximum over queries)
    if (tags != nullptr) {
        const float* tg = tags + (size_t)b * C;
        for (int c = lane; c < C; c += 32) {
            int best = 0; float bv = prob[c];
            for (int q = 1; q < Q; ++q) { const float v = prob[q * C1 + c]; if (v > bv) { bv = v; best = q; } }
            const float t = tg[c];
            if (at_m == 2 || (at_m == 3 && t != 0.f)) {
                if (bv < fuse_threshold) prob[best * C1 + c] = fuse_threshold;
            }
            if (at_m == 1 || at_m == 2)
                for (int q = 0; q < Q; ++q) prob[q * C1 + c] = prob[q * C1 + c] * t;
        }
        __syncwarp();
    }
    // scores, labels = prob[..., :-1].max(-1); boxes -> (onset, offset) * clip length
    const float len = is_semi ? 1.f : sizes[b];
    for (int q = lane; q < Q; q += 32) {
        int best = 0; float bv = prob[q * C1];
        for (int c = 1; c < C; ++c) { const float v = prob[q * C1 + c]; if (v > bv) { bv = v; best = c; } }
        const float cx = boxes[((size_t)b * Q + q) * 2], w = boxes[((size_t)b * Q + q) * 2 + 1];
        float s0, e0;
        if (is_semi) { s0 = cx; e0 = w; }                              // sedt.py:393-394: boxes passed through
        else { s0 = __fmul_rn(cx - w / 2.f, len); e0 = __fmul_rn(cx + w / 2.f, len); }
        sc[q] = bv; lab[q] = best; on[q] = s0; off[q] = e0;
        out_scores[(size_t)b * Q + q] = bv;
        out_labels[(size_t)b * Q + q] = best;
        out_boxes[((size_t)b * Q + q) * 2] = s0;
        out_boxes[((size_t)b * Q + q) * 2 + 1] = e0;
    }
    __syncwarp();
    if (ev_count == nullptr || lane != 0) return;

    // decode_strong, del_overlap = True (BoxEncoder.py:201-225); sequential per clip, Q is small
    int n_out = 0;
    int32_t* oc = ev_class + (size_t)b * Q;
    float* oo = ev_onset + (size_t)b * Q; float* of = ev_offset + (size_t)b * Q; float* os = ev_score + (size_t)b * Q;
    // a kept query is marked by a non-negative label; consumed ones are flipped to -1 - label
    for (int q = 0; q < Q; ++q) {
        const bool keep = sc[q] >= score_threshold && (off[q] - on[q]) >= min_duration;
        if (!keep) lab[q] = -1 - C1;                                   // never selected
    }
    for (int q0 = 0; q0 < Q; ++q0) {
        if (lab[q0] < 0) continue;
        const int cls = lab[q0];
        int n = 0;
        for (int q = q0; q < Q; ++q)
            if (lab[q] == cls) { ord[n++] = q; lab[q] = -1 - cls; }
        // sort the class's events by onset (insertion sort: stable, like numpy's small-array path)
        for (int i = 1; i < n; ++i) {
            const int v = ord[i];
            int j = i - 1;
            while (j >= 0 && on[ord[j]] > on[v]) { ord[j + 1] = ord[j]; --j; }
            ord[j + 1] = v;
        }
        // drop the weaker of every overlapping neighbour pair
        int i = 1;
        while (i < n) {
            if (on[ord[i]] < off[ord[i - 1]]) {
                const int del = sc[ord[i]] > sc[ord[i - 1]] ? i - 1 : i;
                for (int k = del; k + 1 < n; ++k) ord[k] = ord[k + 1];
                --n;
                continue;
            }
            ++i;
        }
        for (int k = 0; k < n; ++k) {
            oc[n_out] = cls; oo[n_out] = on[ord[k]]; of[n_out] = off[ord[k]]; os[n_out] = sc[ord[k]];
            ++n_out;
        }
    }
    ev_count[b] = n_out;
}

// get_pseudo_labels (engine.py:300-348, the teacher's targets for the unlabelled clips in train_ss_sedt.py): PostProcess with
// at_m = 1 and is_semi (boxes stay (center, width)), class-wise score thresholds, minimum width, then a greedy same-class
// overlap suppression in descending score order.  One warp per clip; the greedy pass is sequential on lane 0.
__global__ void __launch_bounds__(kDecWarps * 32)
pseudo_labels_kernel(const float* __restrict__ logits, const float* __restrict__ boxes, const float* __restrict__ tags,
                     const float* __restrict__ class_thr, int B, int Q, int C1, float min_width, int del_overlap,
                     int64_t* __restrict__ out_labels, float* __restrict__ out_boxes, float* __restrict__ out_scores,
                     int32_t* __restrict__ out_count)
{
    extern __shared__ float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x * kDecWarps + warp;
    if (b >= B) return;
    const int C = C1 - 1;
    const int per_warp = Q * C1 + 5 * Q;
    float* prob = smem + warp * per_warp;
    float* sc = prob + Q * C1;
    float* xs = sc + Q;                            // onset  = c - l / 2
    float* ys = xs + Q;                            // offset = c + l / 2
    int* lab = reinterpret_cast<int*>(ys + Q);
    int* ord = lab + Q;
    const float* lg = logits + (size_t)b * Q * C1;
    for (int q = lane; q < Q; q += 32) {
        float m = -CUDART_INF_F;
        for (int c = 0; c < C1; ++c) m = fmaxf(m, lg[q * C1 + c]);
        float s = 0.f;
        for (int c = 0; c < C1; ++c) { const float e = expf(lg[q * C1 + c] - m); prob[q * C1 + c] = e; s += e; }
        for (int c = 0; c < C1; ++c) prob[q * C1 + c] = prob[q * C1 + c] / s;
    }
    __syncwarp();
    for (int q = lane; q < Q; q += 32) {
        int best = 0; float bv = -CUDART_INF_F;
        for (int c = 0; c < C; ++c) {
            const float v = tags != nullptr ? prob[q * C1 + c] * tags[(size_t)b * C + c] : prob[q * C1 + c];   // at_m = 1
            if (v > bv) { bv = v; best = c; }
        }
        const float cx = boxes[((size_t)b * Q + q) * 2], w = boxes[((size_t)b * Q + q) * 2 + 1];
        const bool keep = bv >= class_thr[best] && w > min_width;
        sc[q] = bv; lab[q] = keep ? best : -1;
        xs[q] = cx - w / 2.f; ys[q] = cx + w / 2.f;
    }
    __syncwarp();
    if (lane != 0) return;
    int n = 0;
    for (int q = 0; q < Q; ++q) if (lab[q] >= 0) ord[n++] = q;
    int n_out = 0;
    auto emit = [&](int q) {
        out_labels[(size_t)b * Q + n_out] = lab[q];
        out_boxes[((size_t)b * Q + n_out) * 2] = boxes[((size_t)b * Q + q) * 2];
        out_boxes[((size_t)b * Q + n_out) * 2 + 1] = boxes[((size_t)b * Q + q) * 2 + 1];
        out_scores[(size_t)b * Q + n_out] = sc[q];
        ++n_out;
    };
    if (!del_overlap) {
        for (int i = 0; i < n; ++i) emit(ord[i]);
    } else {
        for (int i = 1; i < n; ++i) {              // descending score (stable)
            const int v = ord[i];
            int j = i - 1;
            while (j >= 0 && sc[ord[j]] < sc[v]) { ord[j + 1] = ord[j]; --j; }
            ord[j + 1] = v;
        }
        while (n > 0) {
            const int k = ord[0];
            emit(k);
            int m = 0;
            for (int i = 1; i < n; ++i) {
                const int j = ord[i];
                const float overlap = fmaxf(fminf(ys[j], ys[k]) - fmaxf(xs[j], xs[k]), 0.f);
                if (overlap == 0.f || lab[j] != lab[k]) ord[m++] = j;
            }
            n = m;
        }
    }
    out_count[b] = n_out;
}

}  // namespace

int launch_pseudo_labels(const float* logits, const float* boxes, const float* tags, const float* class_thr, int B, int Q, int C1,
                         float min_width, int del_overlap, int64_t* out_labels, float* out_boxes, float* out_scores,
                         int32_t* out_count, cudaStream_t stream)
{
    if (B == 0) return SEDT_OK;
    SEDT_REQUIRE(logits && boxes && class_thr && out_labels && out_boxes && out_scores && out_count, "pseudo_labels: null argument");
    SEDT_REQUIRE(Q >= 1 && C1 >= 2, "pseudo_labels: bad sizes Q=%d C1=%d", Q, C1);
    const size_t smem = (size_t)kDecWarps * (Q * C1 + 5 * Q) * sizeof(float);
    SEDT_REQUIRE(smem <= 200 * 1024, "pseudo_labels: Q=%d x C1=%d does not fit shared memory", Q, C1);
    if (smem > 48 * 1024)
        SEDT_CHECK_CUDA(cudaFuncSetAttribute(pseudo_labels_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ProfScope _prof(PROF_OTHER, stream);
    pseudo_labels_kernel<<<(unsigned)ceil_div(B, kDecWarps), kDecWarps * 32, smem, stream>>>(
        logits, boxes, tags, class_thr, B, Q, C1, min_width, del_overlap, out_labels, out_boxes, out_scores, out_count);
    SEDT_COUNT_LAUNCH();
    SEDT_CHECK_CUDA(cudaGetLastError());
    return SEDT_OK;
}

int launch_decode_events(const float* logits, const float* boxes, const float* sizes, const float* tags, int B, int Q, int C1,
                         int at_m, float fuse_threshold, int is_semi, float score_threshold, float min_duration,
                         float* out_scores, int64_t* out_labels, float* out_boxes, int32_t* ev_class, float* ev_onset,
                         float* ev_offset, float* ev_score, int32_t* ev_count, cudaStream_t stream)
{
    if (B == 0) return SEDT_OK;
    SEDT_REQUIRE(logits && boxes && out_scores && out_labels && out_boxes, "decode: null argument");
    SEDT_REQUIRE(is_semi || sizes != nullptr, "decode: target sizes are required unless is_semi");
    SEDT_REQUIRE(Q >= 1 && C1 >= 2, "decode: bad sizes Q=%d C1=%d", Q, C1);
    SEDT_REQUIRE(tags == nullptr || (at_m >= 1 && at_m <= 3), "decode: at_m=%d (1, 2 or 3)", at_m);
    SEDT_REQUIRE(ev_count == nullptr || (ev_class && ev_onset && ev_offset && ev_score), "decode: event buffers go together");
    const size_t smem = (size_t)kDecWarps * (Q * C1 + 5 * Q) * sizeof(float);
    SEDT_REQUIRE(smem <= 200 * 1024, "decode: Q=%d x C1=%d does not fit shared memory", Q, C1);
    if (smem > 48 * 1024)
        SEDT_CHECK_CUDA(cudaFuncSetAttribute(decode_events_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ProfScope _prof(PROF_OTHER, stream);
    decode_events_kernel<<<(unsigned)ceil_div(B, kDecWarps), kDecWarps * 32, smem, stream>>>(
        logits, boxes, sizes, tags, B, Q, C1, at_m, fuse_threshold, is_semi, score_threshold, min_duration, out_scores, out_labels,
        out_boxes, ev_class, ev_onset, ev_offset, ev_score, ev_count);
    SEDT_COUNT_LAUNCH();
    SEDT_CHECK_CUDA(cudaGetLastError());
    return SEDT_OK;
}

}  // namespace sedt
