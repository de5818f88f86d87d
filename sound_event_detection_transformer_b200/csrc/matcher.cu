// HungarianMatcher hot path on the GPU: one warp per clip.
//
// Replaces sedt/matcher.py:65-97 of the reference (cost matrix over the whole
// batch, .cpu(), per-clip scipy.optimize.linear_sum_assignment).  Here only the
// block diagonal is ever formed: a warp builds its clip's [Q,K] fp32 cost block
// in shared memory (same fp32 operation order as utilities/box_ops.py:9-56 and
// matcher.py:91) and solves it with the shortest-augmenting-path algorithm of
// Crouse (the algorithm behind scipy's solver) in fp64, columns spread over the
// lanes.  Tie-breaking replays scipy's `remaining[]` iteration order, so the
// output indices are bit-identical to scipy's on the same cost block, ties
// included.
#include "common.cuh"
#include <math_constants.h>

namespace sedt {

namespace {

constexpr int kWarpsPerCta = 4;

struct Cand {           // candidate column for the arg-min with scipy's tie rule
    double val;
    int pos;            // position in scipy's remaining[] array (iteration order)
    int col;
    int unassigned;
};

__device__ __forceinline__ bool better(const Cand& a, const Cand& b) {
    // true if a wins over b.  Sequential rule (rectangular_lsap.cpp): strictly
    // lower wins; among equal values the LAST unassigned one in iteration order
    // wins, and if none is unassigned the FIRST one wins.
    if (a.col < 0) return false;
    if (b.col < 0) return true;
    if (a.val < b.val) return true;
    if (b.val < a.val) return false;
    if (a.unassigned != b.unassigned) return a.unassigned != 0;
    return a.unassigned ? (a.pos > b.pos) : (a.pos < b.pos);
}

__device__ __forceinline__ Cand shfl_cand(const Cand& c, int src_xor) {
    Cand o;
    o.val = __shfl_xor_sync(0xffffffffu, c.val, src_xor);
    o.pos = __shfl_xor_sync(0xffffffffu, c.pos, src_xor);
    o.col = __shfl_xor_sync(0xffffffffu, c.col, src_xor);
    o.unassigned = __shfl_xor_sync(0xffffffffu, c.unassigned, src_xor);
    return o;
}

// Select element `idx` (warp-uniform) of a lane-distributed array: element e
// lives in lane e%32, slot e/32.
template <int CPL, typename T>
__device__ __forceinline__ T gather(const T (&arr)[CPL], int idx) {
    T out = T();
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
        T t = __shfl_sync(0xffffffffu, arr[c], idx & 31);
        if (c == (idx >> 5)) out = t;
    }
    return out;
}

// cost block element as scipy sees it: cost[i][j] of the (possibly transposed) problem
__device__ __forceinline__ double cost_at(const float* sc, int i, int j, int ldk, bool transposed) {
    return transposed ? (double)sc[j * ldk + i] : (double)sc[i * ldk + j];
}

template <int CPL>
__device__ int lsap_warp(const float* sc, int Q, int K, int lane, int64_t* rows_out, int64_t* cols_out) {
    // sc: [Q][K] fp32 in shared memory (ld = K).  Returns pair count or <0.
    const bool transposed = K < Q;
    const int nr = transposed ? K : Q;
    const int nc = transposed ? Q : K;
    if (nr == 0) return 0;

    double v[CPL], spc[CPL], u[CPL];
    int path[CPL], row4col[CPL], pos[CPL], col4row[CPL];
    unsigned sr_bits = 0;     // bit c: row (lane + 32c) in SR
    unsigned sc_bits = 0;     // bit c: column (lane + 32c) in SC
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
        v[c] = 0.0; u[c] = 0.0; spc[c] = CUDART_INF; path[c] = -1; row4col[c] = -1; col4row[c] = -1; pos[c] = -1;
    }

    for (int cur = 0; cur < nr; ++cur) {
        double minVal = 0.0;
        int i = cur;
        int num_remaining = nc;
        sr_bits = 0; sc_bits = 0;
#pragma unroll
        for (int c = 0; c < CPL; ++c) {
            int j = lane + 32 * c;
            pos[c] = (j < nc) ? (nc - 1 - j) : -1;       // remaining[it] = nc - it - 1
            spc[c] = CUDART_INF;
        }
        int sink = -1;
        while (sink == -1) {
            if ((i & 31) == lane) sr_bits |= 1u << (i >> 5);
            const double ui = gather<CPL>(u, i);
            Cand best; best.val = CUDART_INF; best.pos = 0; best.col = -1; best.unassigned = 0;
#pragma unroll
            for (int c = 0; c < CPL; ++c) {
                int j = lane + 32 * c;
                if (j < nc && pos[c] >= 0) {
                    double r = minVal + cost_at(sc, i, j, K, transposed) - ui - v[c];
                    if (r < spc[c]) { path[c] = i; spc[c] = r; }
                    Cand me; me.val = spc[c]; me.pos = pos[c]; me.col = j; me.unassigned = (row4col[c] == -1);
                    if (better(me, best)) best = me;
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                Cand other = shfl_cand(best, o);
                if (better(other, best)) best = other;
            }
            minVal = best.val;
            if (best.col < 0 || minVal == CUDART_INF) return SEDT_ERR_INFEASIBLE;
            const int j = best.col;
            if (best.unassigned) sink = j; else i = gather<CPL>(row4col, j);
            // SC[j] = true; remaining[index] = remaining[--num_remaining]
            const int idx = best.pos;
            --num_remaining;
#pragma unroll
            for (int c = 0; c < CPL; ++c) {
                if (pos[c] == num_remaining) pos[c] = idx;
            }
            if ((j & 31) == lane) {
#pragma unroll
                for (int c = 0; c < CPL; ++c)
                    if (c == (j >> 5)) { pos[c] = -1; sc_bits |= 1u << c; }
            }
        }
        // dual updates
#pragma unroll
        for (int c = 0; c < CPL; ++c) {
            int r = lane + 32 * c;
            // every lane takes part in the gather shuffles; only SR rows use the value
            int cr = (r < nr) ? col4row[c] : -1;
            double s = 0.0;
#pragma unroll
            for (int c2 = 0; c2 < CPL; ++c2) {
                double t = __shfl_sync(0xffffffffu, spc[c2], (cr < 0 ? 0 : cr) & 31);
                if (c2 == ((cr < 0 ? 0 : cr) >> 5)) s = t;
            }
            if (r < nr) {
                if (r == cur) u[c] += minVal;
                else if ((sr_bits >> c) & 1u) u[c] += minVal - s;
            }
        }
#pragma unroll
        for (int c = 0; c < CPL; ++c)
            if ((sc_bits >> c) & 1u) v[c] -= minVal - spc[c];
        // augment
        int j = sink;
        for (;;) {
            const int r = gather<CPL>(path, j);
            if ((j & 31) == lane) {
#pragma unroll
                for (int c = 0; c < CPL; ++c) if (c == (j >> 5)) row4col[c] = r;
            }
            const int old = gather<CPL>(col4row, r);
            if ((r & 31) == lane) {
#pragma unroll
                for (int c = 0; c < CPL; ++c) if (c == (r >> 5)) col4row[c] = j;
            }
            j = old;
            if (r == cur) break;
        }
    }

    // emit pairs sorted by query index
    if (!transposed) {
#pragma unroll
        for (int c = 0; c < CPL; ++c) {
            int r = lane + 32 * c;
            if (r < nr) { rows_out[r] = r; cols_out[r] = col4row[c]; }
        }
    } else {
        int base = 0;
#pragma unroll
        for (int c = 0; c < CPL; ++c) {
            int q = lane + 32 * c;
            bool has = (q < nc) && row4col[c] != -1;
            unsigned m = __ballot_sync(0xffffffffu, has);
            if (has) {
                int o = base + __popc(m & ((1u << lane) - 1u));
                rows_out[o] = q; cols_out[o] = row4col[c];
            }
            base += __popc(m);
        }
    }
    return nr;
}

// One clip's [Q,K] cost block (matcher.py:65-91) into sc (ld = K) and its class probabilities into sp [Q][C1];
// logits / boxes / targets already offset to the clip.  Returns true if an entry is NaN or -inf.
__device__ bool build_cost_block(const float* __restrict__ logits, const float* __restrict__ boxes,
                                 const int64_t* __restrict__ tgt_labels, const float* __restrict__ tgt_boxes,
                                 int Q, int C1, int K, float w_class, float w_bbox, float w_giou,
                                 float* sc, float* sp, int lane, int fl = 0, float alpha_fl = 0.f, float gamma_fl = 0.f)
{
    bool bad = false;
    // softmax over classes, one query row per lane (matcher.py:65); sigmoid under fl (focal class cost)
    for (int q = lane; q < Q; q += 32) {
        const float* lg = logits + (size_t)q * C1;
        if (fl) {
            for (int c = 0; c < C1; ++c) sp[q * C1 + c] = 1.f / (1.f + expf(-lg[c]));
            continue;
        }
        float m = -CUDART_INF_F;
        for (int c = 0; c < C1; ++c) m = fmaxf(m, lg[c]);
        float s = 0.f;
        for (int c = 0; c < C1; ++c) { float e = expf(lg[c] - m); sp[q * C1 + c] = e; s += e; }
        for (int c = 0; c < C1; ++c) sp[q * C1 + c] = sp[q * C1 + c] / s;
    }
    __syncwarp();
    for (int e = lane; e < Q * K; e += 32) {
        const int q = e / K, t = e % K;
        const float cp = boxes[(size_t)q * 2 + 0], lp = boxes[(size_t)q * 2 + 1];
        const float ct = tgt_boxes[(size_t)t * 2 + 0], lt = tgt_boxes[(size_t)t * 2 + 1];
        const int64_t lab = tgt_labels[t];
        // box_ops.py:9-14  (c,l) -> (c - l/2, 0, c + l/2, 1)
        const float s_p = cp - lp / 2.f, e_p = cp + lp / 2.f;
        const float s_t = ct - lt / 2.f, e_t = ct + lt / 2.f;
        float cost_class = -sp[q * C1 + (int)lab];                             // matcher.py:76
        if (fl) {                                                              // matcher.py:78-82
            const float p = sp[q * C1 + (int)lab];
            const float pg = gamma_fl == 1.f ? p : (gamma_fl == 2.f ? p * p : powf(p, gamma_fl));
            const float qg = gamma_fl == 1.f ? 1.f - p : (gamma_fl == 2.f ? (1.f - p) * (1.f - p) : powf(1.f - p, gamma_fl));
            const float neg = __fmul_rn(__fmul_rn(1.f - alpha_fl, pg), -logf(__fadd_rn(1.f - p, 1e-8f)));
            const float pos = __fmul_rn(__fmul_rn(alpha_fl, qg), -logf(__fadd_rn(p, 1e-8f)));
            cost_class = __fsub_rn(pos, neg);
        }
        const float cost_bbox = __fadd_rn(fabsf(s_p - s_t), fabsf(e_p - e_t));   // matcher.py:85 (cdist p=1)
        // box_ops.py:29-42 with y-extent [0,1]
        const float area_p = e_p - s_p, area_t = e_t - s_t;
        const float inter = fmaxf(fminf(e_p, e_t) - fmaxf(s_p, s_t), 0.f);
        const float uni = __fsub_rn(__fadd_rn(area_p, area_t), inter);
        const float iou = __fdiv_rn(inter, uni);
        // box_ops.py:45-56
        const float enc = fmaxf(fmaxf(e_p, e_t) - fminf(s_p, s_t), 0.f);
        const float giou = __fsub_rn(iou, __fdiv_rn(__fsub_rn(enc, uni), enc));
        // matcher.py:91  cost_bbox*w + cost_class*w + cost_giou*w, left to right
        float cv = __fadd_rn(__fadd_rn(__fmul_rn(w_bbox, cost_bbox), __fmul_rn(w_class, cost_class)),
                             __fmul_rn(w_giou, -giou));
        sc[e] = cv;
        bad |= (cv != cv) || (cv == -CUDART_INF_F);
    }
    return bad;
}

template <int CPL>
__global__ void __launch_bounds__(kWarpsPerCta * 32)
matcher_kernel(const float* __restrict__ logits, const float* __restrict__ boxes,
               const int64_t* __restrict__ tgt_labels, const float* __restrict__ tgt_boxes,
               const int32_t* __restrict__ offsets, int B, int Q, int C1, int Kmax,
               float w_class, float w_bbox, float w_giou,
               const float* __restrict__ cost_in, int ld_in,
               float* __restrict__ cost_out, int ld_out,
               int64_t* __restrict__ rows, int64_t* __restrict__ cols, int32_t* __restrict__ counts,
               int32_t* __restrict__ status, int solve, int fl, float alpha_fl, float gamma_fl,
               float* __restrict__ lmin, int64_t* __restrict__ largmin)
{
    extern __shared__ float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x * kWarpsPerCta + warp;
    if (b >= B) return;
    const int per_warp = Q * Kmax + Q * C1;
    float* sc = smem + warp * per_warp;          // [Q][K]
    float* sp = sc + Q * Kmax;                   // [Q][C1] probabilities
    const int k0 = offsets[b];
    const int K = offsets[b + 1] - k0;
    bool bad = false;

    if (cost_in != nullptr) {
        for (int e = lane; e < Q * K; e += 32) {
            float cv = cost_in[((size_t)b * Q + e / K) * ld_in + (e % K)];
            sc[e] = cv;
            bad |= (cv != cv) || (cv == -CUDART_INF_F);
        }
    } else {
        bad = build_cost_block(logits + (size_t)b * Q * C1, boxes + (size_t)b * Q * 2, tgt_labels + k0, tgt_boxes + (size_t)k0 * 2,
                               Q, C1, K, w_class, w_bbox, w_giou, sc, sp, lane, fl, alpha_fl, gamma_fl);
        if (lmin != nullptr) {
            // fine_tune (matcher.py:99-106): per query the smallest location cost C_l = w_bbox * L1 + w_giou * (-GIoU) and its
            // target (first minimum, as torch.min(-1))
            for (int q = lane; q < Q; q += 32) {
                const float cp = boxes[((size_t)b * Q + q) * 2], lp = boxes[((size_t)b * Q + q) * 2 + 1];
                const float s_p = cp - lp / 2.f, e_p = cp + lp / 2.f;
                float best = CUDART_INF_F; int bt = -1;
                for (int t = 0; t < K; ++t) {
                    const float ct = tgt_boxes[(size_t)(k0 + t) * 2], lt = tgt_boxes[(size_t)(k0 + t) * 2 + 1];
                    const float s_t = ct - lt / 2.f, e_t = ct + lt / 2.f;
                    const float cost_bbox = __fadd_rn(fabsf(s_p - s_t), fabsf(e_p - e_t));
                    const float inter = fmaxf(fminf(e_p, e_t) - fmaxf(s_p, s_t), 0.f);
                    const float uni = __fsub_rn(__fadd_rn(e_p - s_p, e_t - s_t), inter);
                    const float enc = fmaxf(fmaxf(e_p, e_t) - fminf(s_p, s_t), 0.f);
                    const float giou = __fsub_rn(__fdiv_rn(inter, uni), __fdiv_rn(__fsub_rn(enc, uni), enc));
                    const float cl = __fadd_rn(__fmul_rn(w_bbox, cost_bbox), __fmul_rn(w_giou, -giou));
                    if (cl < best || bt < 0) { best = cl; bt = t; }
                }
                lmin[(size_t)b * Q + q] = best;
                largmin[(size_t)b * Q + q] = bt;
            }
        }
    }
    __syncwarp();
    if (cost_out != nullptr) {
        for (int e = lane; e < Q * K; e += 32) cost_out[((size_t)b * Q + e / K) * ld_out + (e % K)] = sc[e];
    }
    if (!solve) return;
    bad = __any_sync(0xffffffffu, bad);
    int64_t* ro = rows + (size_t)b * Q;
    int64_t* co = cols + (size_t)b * Q;
    for (int e = lane; e < Q; e += 32) { ro[e] = -1; co[e] = -1; }
    __syncwarp();
    int n;
    if (bad) n = SEDT_ERR_NUMERIC;
    else n = lsap_warp<CPL>(sc, Q, K, lane, ro, co);
    if (lane == 0) {
        counts[b] = n < 0 ? 0 : n;
        if (n < 0) atomicMin(status, n);
    }
}


// ---- SetCriterion (sedt/sedt.py:309-352, default path: fine_tune = normalize = fl = False, no mixup ratio) -----------------
// One warp per (decoder layer l, clip b): the matcher above, then loss_labels (:188-221), loss_boxes (:238-261) and
// loss_cardinality (:223-236) of that clip AND their gradients w.r.t. pred_logits / pred_boxes, so that the training step
// needs no autograd graph (and no ~150 small torch kernels) between the model's backward and the matched pairs.
// partials[l][b][8] = { sum w*ce, sum L1, sum (1-GIoU), #correct, #non-empty predictions, #pairs, 0, 0 }; a second kernel
// reduces them in a fixed order.  Clips b >= Bs (not in strong_mask) only contribute to the cardinality metric.
constexpr int kPartials = 8;

template <int CPL>
__global__ void __launch_bounds__(kWarpsPerCta * 32)
set_criterion_kernel(const float* __restrict__ logits, const float* __restrict__ boxes,
                     const int64_t* __restrict__ tgt_labels, const float* __restrict__ tgt_boxes,
                     const int32_t* __restrict__ offsets, int L, int B, int Bs, int Q, int C1, int Kmax,
                     float w_class, float w_bbox, float w_giou, float eos_coef, float inv_num_boxes,
                     int64_t* __restrict__ rows, int64_t* __restrict__ cols, int32_t* __restrict__ status,
                     float* __restrict__ partials, float* __restrict__ g_logits, float* __restrict__ g_l1,
                     float* __restrict__ g_giou)
{
    extern __shared__ float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wid = blockIdx.x * kWarpsPerCta + warp;
    if (wid >= L * B) return;
    const int l = wid / B, b = wid % B;
    const int per_warp = Q * Kmax + Q * C1 + 2 * Q;
    float* sc = smem + warp * per_warp;          // [Q][K] cost block
    float* sp = sc + Q * Kmax;                   // [Q][C1] probabilities
    float* slse = sp + Q * C1;                   // [Q] max + log(sum exp)
    int* tcls = reinterpret_cast<int*>(slse + Q);   // [Q] target class per query
    const float* lg = logits + ((size_t)l * B + b) * Q * C1;
    const float* bx = boxes + ((size_t)l * B + b) * Q * 2;
    float* gl = g_logits + ((size_t)l * B + b) * Q * C1;
    float* g1 = g_l1 + ((size_t)l * B + b) * Q * 2;
    float* g2 = g_giou + ((size_t)l * B + b) * Q * 2;
    float* part = partials + ((size_t)l * B + b) * kPartials;
    const int noobj = C1 - 1;

    // cardinality (sedt.py:232): queries whose arg-max (first maximum) is not the no-object class
    int npred = 0;
    for (int q = lane; q < Q; q += 32) {
        float m = -CUDART_INF_F;
        for (int c = 0; c < noobj; ++c) m = fmaxf(m, lg[q * C1 + c]);
        npred += (noobj > 0 && m >= lg[q * C1 + noobj]) ? 1 : 0;
    }
    npred = __reduce_add_sync(0xffffffffu, npred);
    for (int e = lane; e < Q * 2; e += 32) { g1[e] = 0.f; g2[e] = 0.f; }
    if (b >= Bs) {
        for (int e = lane; e < Q * C1; e += 32) gl[e] = 0.f;
        if (lane < kPartials) part[lane] = lane == 4 ? (float)npred : 0.f;
        return;
    }

    const int k0 = offsets[b];
    const int K = offsets[b + 1] - k0;
    bool bad = build_cost_block(lg, bx, tgt_labels + k0, tgt_boxes + (size_t)k0 * 2, Q, C1, K, w_class, w_bbox, w_giou, sc, sp, lane);
    __syncwarp();
    bad = __any_sync(0xffffffffu, bad);
    int64_t* ro = rows + ((size_t)l * Bs + b) * Q;
    int64_t* co = cols + ((size_t)l * Bs + b) * Q;
    for (int e = lane; e < Q; e += 32) { ro[e] = -1; co[e] = -1; tcls[e] = noobj; }
    __syncwarp();
    int n = bad ? (int)SEDT_ERR_NUMERIC : lsap_warp<CPL>(sc, Q, K, lane, ro, co);
    if (n < 0) { if (lane == 0) atomicMin(status, n); n = 0; }
    __syncwarp();

    // matched pairs: target classes (sedt.py:202-207), box losses and their gradients (sedt.py:246-260)
    float l1_sum = 0.f, giou_sum = 0.f;
    int correct = 0;
    for (int i = lane; i < n; i += 32) {
        const int q = (int)ro[i], t = (int)co[i];
        const int lab = (int)tgt_labels[k0 + t];
        tcls[q] = lab;
        // class_error (utilities/utils.py:564-579): top-1 of the matched queries
        int am = 0; float mv = lg[q * C1];
        for (int c = 1; c < C1; ++c) { const float v = lg[q * C1 + c]; if (v > mv) { mv = v; am = c; } }
        correct += (am == lab) ? 1 : 0;
        const float cp = bx[q * 2], lp = bx[q * 2 + 1];
        const float ct = tgt_boxes[(size_t)(k0 + t) * 2], lt = tgt_boxes[(size_t)(k0 + t) * 2 + 1];
        const float s1 = cp - lp / 2.f, e1 = cp + lp / 2.f, s2 = ct - lt / 2.f, e2 = ct + lt / 2.f;
        // L1 on (s, 0, e, 1)
        const float ds = s1 - s2, de = e1 - e2;
        l1_sum += fabsf(ds) + fabsf(de);
        const float sgs = (ds > 0.f) - (ds < 0.f), sge = (de > 0.f) - (de < 0.f);
        g1[q * 2] = (sgs + sge) * inv_num_boxes;
        g1[q * 2 + 1] = (sge - sgs) * 0.5f * inv_num_boxes;
        // GIoU (utilities/box_ops.py:29-56 on y-extent [0,1]): giou = I/U - (E - U)/E
        const float iraw = fminf(e1, e2) - fmaxf(s1, s2);
        const float I = fmaxf(iraw, 0.f);
        const float U = (e1 - s1) + (e2 - s2) - I;
        const float eraw = fmaxf(e1, e2) - fminf(s1, s2);
        const float E = fmaxf(eraw, 0.f);
        const float giou = I / U - (E - U) / E;
        giou_sum += 1.f - giou;
        // d giou: through I (direct and via U), the predicted area (via U) and E
        const float dU = -I / (U * U) + 1.f / E;
        const float dI = 1.f / U - dU;
        const float dE = -U / (E * E);
        const float ion = iraw >= 0.f ? 1.f : 0.f, eon = eraw >= 0.f ? 1.f : 0.f;
        // min / max hand the gradient to the selected operand (ties: split evenly, as torch does)
        const float e1_is_min = e1 < e2 ? 1.f : (e1 == e2 ? 0.5f : 0.f);
        const float s1_is_max = s1 > s2 ? 1.f : (s1 == s2 ? 0.5f : 0.f);
        const float e1_is_max = e1 > e2 ? 1.f : (e1 == e2 ? 0.5f : 0.f);
        const float s1_is_min = s1 < s2 ? 1.f : (s1 == s2 ? 0.5f : 0.f);
        const float de1 = dI * ion * e1_is_min + dU + dE * eon * e1_is_max;
        const float ds1 = -dI * ion * s1_is_max - dU - dE * eon * s1_is_min;
        // loss_giou = 1 - giou
        g2[q * 2] = -(ds1 + de1) * inv_num_boxes;
        g2[q * 2 + 1] = -(de1 - ds1) * 0.5f * inv_num_boxes;
    }
    __syncwarp();

    // weighted cross entropy over all queries (sedt.py:219-220) and d/d logits = w * (softmax - onehot) / num_boxes
    float ce_sum = 0.f;
    for (int q = lane; q < Q; q += 32) {
        float m = -CUDART_INF_F;
        for (int c = 0; c < C1; ++c) m = fmaxf(m, lg[q * C1 + c]);
        float ssum = 0.f;
        for (int c = 0; c < C1; ++c) ssum += expf(lg[q * C1 + c] - m);
        const float lse = logf(ssum);
        const int t = tcls[q];
        const float w = t == noobj ? eos_coef : 1.f;
        ce_sum += -w * ((lg[q * C1 + t] - m) - lse);
        for (int c = 0; c < C1; ++c)
            gl[q * C1 + c] = w * (sp[q * C1 + c] - (c == t ? 1.f : 0.f)) * inv_num_boxes;
    }
    ce_sum = warp_sum(ce_sum); l1_sum = warp_sum(l1_sum); giou_sum = warp_sum(giou_sum);
    correct = __reduce_add_sync(0xffffffffu, correct);
    if (lane == 0) {
        part[0] = ce_sum; part[1] = l1_sum; part[2] = giou_sum; part[3] = (float)correct; part[4] = (float)npred;
        part[5] = (float)n; part[6] = 0.f; part[7] = 0.f;
    }
}

// losses[l][8] = { loss_ce, loss_bbox, loss_giou, class_error, cardinality_error, loss_weak (last layer only), 0, 0 }.
// CTA l < L reduces layer l's partials; CTA L computes the weak (audio tagging) loss, sedt.py:161-186:
// BCELoss(at[:Bw], clamp(multi-hot(labels), 0, 1)) and its gradient (log clamped at -100 as torch.nn.BCELoss does).
__global__ void __launch_bounds__(256)
set_criterion_finish_kernel(const float* __restrict__ partials, const float* __restrict__ n_tgt, int L, int B, float inv_num_boxes,
                            const float* __restrict__ at, const int64_t* __restrict__ wl_labels,
                            const int32_t* __restrict__ wl_offsets, int Bw, int C,
                            float* __restrict__ losses, float* __restrict__ g_at)
{
    __shared__ double red[256];
    const int tid = threadIdx.x;
    const int l = blockIdx.x;
    auto block_sum = [&](double v) -> double {
        __syncthreads();
        red[tid] = v;
        __syncthreads();
        for (int o = 128; o > 0; o >>= 1) {
            if (tid < o) red[tid] += red[tid + o];
            __syncthreads();
        }
        return red[0];
    };
    if (l < L) {
        double acc[6] = {0, 0, 0, 0, 0, 0};
        for (int b = tid; b < B; b += 256) {
            const float* p = partials + ((size_t)l * B + b) * kPartials;
            acc[0] += p[0]; acc[1] += p[1]; acc[2] += p[2]; acc[3] += p[3]; acc[5] += p[5];
            acc[4] += fabs((double)p[4] - (double)n_tgt[b]);
        }
        double tot[6];
        for (int j = 0; j < 6; ++j) tot[j] = block_sum(acc[j]);
        if (tid == 0) {
            float* o = losses + (size_t)l * kPartials;
            o[0] = (float)tot[0] * inv_num_boxes;
            o[1] = (float)tot[1] * inv_num_boxes;
            o[2] = (float)tot[2] * inv_num_boxes;
            o[3] = tot[5] > 0 ? 100.f - (float)(tot[3] * 100.0 / tot[5]) : 100.f;
            o[4] = (float)(tot[4] / (double)B);
            if (l != L - 1 || at == nullptr) o[5] = 0.f;
            o[6] = 0.f; o[7] = 0.f;
        }
        return;
    }
    if (at == nullptr) return;
    const int N = Bw * C;
    double acc = 0.0;
    for (int e = tid; e < N; e += 256) {
        const int b = e / C, c = e % C;
        float g = 0.f;
        for (int j = wl_offsets[b]; j < wl_offsets[b + 1]; ++j) g = (wl_labels[j] == c) ? 1.f : g;
        const float p = at[e];
        const float lp = fmaxf(logf(p), -100.f), l1p = fmaxf(logf(1.f - p), -100.f);
        acc += (double)(-(g * lp + (1.f - g) * l1p));
        g_at[e] = (p - g) / fmaxf((1.f - p) * p, 1e-12f) / (float)N;
    }
    const double tot = block_sum(acc);
    if (tid == 0) losses[(size_t)(L - 1) * kPartials + 5] = (float)(tot / (double)N);
}

}  // namespace

// Host launcher (also used by the C-ABI in api.cu).
int launch_matcher(const float* logits, const float* boxes, const int64_t* tgt_labels, const float* tgt_boxes,
                   const int32_t* offsets, int B, int Q, int C1, int Kmax, float w_class, float w_bbox, float w_giou,
                   const float* cost_in, int ld_in, float* cost_out, int ld_out,
                   int64_t* rows, int64_t* cols, int32_t* counts, int32_t* status, int solve, cudaStream_t stream,
                   int fl, float alpha_fl, float gamma_fl, float* lmin, int64_t* largmin)
{
    if (B == 0) return SEDT_OK;
    SEDT_REQUIRE((lmin == nullptr) == (largmin == nullptr), "matcher: lmin and largmin go together");
    SEDT_REQUIRE(lmin == nullptr || cost_in == nullptr, "matcher: the location-cost minimum needs boxes, not a cost matrix");
    SEDT_REQUIRE(Q >= 1 && Kmax >= 0 && C1 >= 1, "matcher: bad sizes Q=%d Kmax=%d C1=%d", Q, Kmax, C1);
    const int nmax = Q > Kmax ? Q : Kmax;
    SEDT_REQUIRE(nmax <= 128, "matcher: max(Q, K)=%d exceeds the 128 supported by the warp solve", nmax);
    const int kpad = Kmax > 0 ? Kmax : 1;
    const size_t smem = (size_t)kWarpsPerCta * (Q * kpad + Q * C1) * sizeof(float);
    SEDT_REQUIRE(smem <= 200 * 1024, "matcher: cost block too large for shared memory (%zu bytes)", smem);
    dim3 grid((unsigned)ceil_div(B, kWarpsPerCta)), block(kWarpsPerCta * 32);
    ProfScope _prof(PROF_MATCHER, stream);
#define SEDT_MATCHER_LAUNCH(CPL)                                                                          \
    do {                                                                                                  \
        if (smem > 48 * 1024)                                                                             \
            SEDT_CHECK_CUDA(cudaFuncSetAttribute(matcher_kernel<CPL>,                                     \
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        matcher_kernel<CPL><<<grid, block, smem, stream>>>(logits, boxes, tgt_labels, tgt_boxes, offsets, B, Q, C1, \
            kpad, w_class, w_bbox, w_giou, cost_in, ld_in, cost_out, ld_out, rows, cols, counts, status, solve,   \
            fl, alpha_fl, gamma_fl, lmin, largmin);                                                               \
    } while (0)
    if (nmax <= 32) SEDT_MATCHER_LAUNCH(1);
    else if (nmax <= 64) SEDT_MATCHER_LAUNCH(2);
    else SEDT_MATCHER_LAUNCH(4);
#undef SEDT_MATCHER_LAUNCH
    SEDT_COUNT_LAUNCH();
    SEDT_CHECK_CUDA(cudaGetLastError());
    return SEDT_OK;
}

int launch_set_criterion(const float* logits, const float* boxes, const float* at, const int64_t* tgt_labels, const float* tgt_boxes,
                         const int32_t* offsets, const float* n_tgt, const int64_t* wl_labels, const int32_t* wl_offsets,
                         int L, int B, int Bs, int Bw, int Q, int C1, int Kmax, float w_class, float w_bbox, float w_giou,
                         float eos_coef, float num_boxes, int64_t* rows, int64_t* cols, int32_t* status, float* partials,
                         float* losses, float* g_logits, float* g_l1, float* g_giou, float* g_at, cudaStream_t stream)
{
    SEDT_REQUIRE(L >= 1 && B >= 1 && Bs >= 0 && Bs <= B && Bw >= 0 && Bw <= B, "set_criterion: bad sizes L=%d B=%d Bs=%d Bw=%d", L, B, Bs, Bw);
    SEDT_REQUIRE(Q >= 1 && Kmax >= 0 && C1 >= 2, "set_criterion: bad sizes Q=%d Kmax=%d C1=%d", Q, Kmax, C1);
    SEDT_REQUIRE(logits && boxes && offsets && n_tgt && rows && cols && status && partials && losses && g_logits && g_l1 && g_giou,
                 "set_criterion: null argument");
    SEDT_REQUIRE(at == nullptr || (g_at && wl_offsets && (wl_labels || Bw == 0)), "set_criterion: weak loss needs g_at and its label table");
    const int nmax = Q > Kmax ? Q : Kmax;
    SEDT_REQUIRE(nmax <= 128, "set_criterion: max(Q, K)=%d exceeds the 128 supported by the warp solve", nmax);
    const int kpad = Kmax > 0 ? Kmax : 1;
    const size_t smem = (size_t)kWarpsPerCta * (Q * kpad + Q * C1 + 2 * Q) * sizeof(float);
    SEDT_REQUIRE(smem <= 200 * 1024, "set_criterion: cost block too large for shared memory (%zu bytes)", smem);
    const float inv_nb = 1.f / num_boxes;          // num_boxes = 0 gives inf / nan losses, as in the reference
    dim3 grid((unsigned)ceil_div((int64_t)L * B, kWarpsPerCta)), block(kWarpsPerCta * 32);
    ProfScope _prof(PROF_MATCHER, stream);
#define SEDT_SETCRIT_LAUNCH(CPL)                                                                               \
    do {                                                                                                       \
        if (smem > 48 * 1024)                                                                                  \
            SEDT_CHECK_CUDA(cudaFuncSetAttribute(set_criterion_kernel<CPL>,                                    \
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));     \
        set_criterion_kernel<CPL><<<grid, block, smem, stream>>>(logits, boxes, tgt_labels, tgt_boxes, offsets, L, B, Bs, Q, C1, \
            kpad, w_class, w_bbox, w_giou, eos_coef, inv_nb, rows, cols, status, partials, g_logits, g_l1, g_giou); \
    } while (0)
    if (nmax <= 32) SEDT_SETCRIT_LAUNCH(1);
    else if (nmax <= 64) SEDT_SETCRIT_LAUNCH(2);
    else SEDT_SETCRIT_LAUNCH(4);
#undef SEDT_SETCRIT_LAUNCH
    SEDT_COUNT_LAUNCH();
    set_criterion_finish_kernel<<<L + (at != nullptr ? 1 : 0), 256, 0, stream>>>(partials, n_tgt, L, B, inv_nb, at, wl_labels,
                                                                              wl_offsets, Bw, C1 - 1, losses, g_at);
    SEDT_COUNT_LAUNCH();
    SEDT_CHECK_CUDA(cudaGetLastError());
    return SEDT_OK;
}

}  // namespace sedt
