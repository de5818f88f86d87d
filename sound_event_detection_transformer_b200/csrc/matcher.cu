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

template <int CPL>
__global__ void __launch_bounds__(kWarpsPerCta * 32)
matcher_kernel(const float* __restrict__ logits, const float* __restrict__ boxes,
               const int64_t* __restrict__ tgt_labels, const float* __restrict__ tgt_boxes,
               const int32_t* __restrict__ offsets, int B, int Q, int C1, int Kmax,
               float w_class, float w_bbox, float w_giou,
               const float* __restrict__ cost_in, int ld_in,
               float* __restrict__ cost_out, int ld_out,
               int64_t* __restrict__ rows, int64_t* __restrict__ cols, int32_t* __restrict__ counts,
               int32_t* __restrict__ status, int solve)
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
        // softmax over classes, one query row per lane (matcher.py:65)
        for (int q = lane; q < Q; q += 32) {
            const float* lg = logits + ((size_t)b * Q + q) * C1;
            float m = -CUDART_INF_F;
            for (int c = 0; c < C1; ++c) m = fmaxf(m, lg[c]);
            float s = 0.f;
            for (int c = 0; c < C1; ++c) { float e = expf(lg[c] - m); sp[q * C1 + c] = e; s += e; }
            for (int c = 0; c < C1; ++c) sp[q * C1 + c] = sp[q * C1 + c] / s;
        }
        __syncwarp();
        for (int e = lane; e < Q * K; e += 32) {
            const int q = e / K, t = e % K;
            const float cp = boxes[((size_t)b * Q + q) * 2 + 0], lp = boxes[((size_t)b * Q + q) * 2 + 1];
            const float ct = tgt_boxes[(size_t)(k0 + t) * 2 + 0], lt = tgt_boxes[(size_t)(k0 + t) * 2 + 1];
            const int64_t lab = tgt_labels[k0 + t];
            // box_ops.py:9-14  (c,l) -> (c - l/2, 0, c + l/2, 1)
            const float s_p = cp - lp / 2.f, e_p = cp + lp / 2.f;
            const float s_t = ct - lt / 2.f, e_t = ct + lt / 2.f;
            const float cost_class = -sp[q * C1 + (int)lab];                       // matcher.py:76
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

}  // namespace

// Host launcher (also used by the C-ABI in api.cu).
int launch_matcher(const float* logits, const float* boxes, const int64_t* tgt_labels, const float* tgt_boxes,
                   const int32_t* offsets, int B, int Q, int C1, int Kmax, float w_class, float w_bbox, float w_giou,
                   const float* cost_in, int ld_in, float* cost_out, int ld_out,
                   int64_t* rows, int64_t* cols, int32_t* counts, int32_t* status, int solve, cudaStream_t stream)
{
    if (B == 0) return SEDT_OK;
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
            kpad, w_class, w_bbox, w_giou, cost_in, ld_in, cost_out, ld_out, rows, cols, counts, status, solve);  \
    } while (0)
    if (nmax <= 32) SEDT_MATCHER_LAUNCH(1);
    else if (nmax <= 64) SEDT_MATCHER_LAUNCH(2);
    else SEDT_MATCHER_LAUNCH(4);
#undef SEDT_MATCHER_LAUNCH
    SEDT_COUNT_LAUNCH();
    SEDT_CHECK_CUDA(cudaGetLastError());
    return SEDT_OK;
}

}  // namespace sedt
