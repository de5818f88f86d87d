// Training-time input transforms of the reference on the device:
//
//   augment_clips_kernel   TimeMask -> FreqMask(fill "mean" / "constant") -> FreqShift on the padded log-mel clip, in place
//                          (utilities/BoxTransforms.py:363-452); the random draws stay on the host in the reference's order
//                          (augment.py: draw_augment_params), the kernel gets the resulting integer bands per clip
//   mix_rows_kernel        mixup's data path: out[k] = a_k x[i1_k] + b_k x[i2_k] (utilities/mixup.py:35: lam * data_1 +
//                          (1 - lam) * data_2, two rounded fp32 products and one rounded sum like the eager expression)
//   query_patches_kernel   SP-SEDT's patch crop + resize (utilities/BoxTransforms.py:315-360, Query): crop the box's frames,
//                          min / max normalise, quantise to 8 bits (ToPILImage), Pillow's antialiased bilinear resample to 128
//                          frames in 22-bit fixed point (src/libImaging/Resample.c, vertical pass: the width stays 64),
//                          back to float (ToTensor) and de-normalise.  Integer arithmetic: bit-exact with the reference.
//
// All three are HBM-bound elementwise passes (4 B in + 4 B out per element); one CTA per clip / row / patch.
#include "common.cuh"
#include "kernels.h"
#include <math_constants.h>

namespace sedt {
namespace {

// np.mean over a [T, n] float32 slice (FreqMask's fill value must match the reference bit for bit).  What numpy does (verified
// against numpy 2.3 on 300 random shapes, tests/test_oracle_golden.py): the slice is read in C order in chunks of
// (8192 / n) * n elements (whole rows that fit its 8192-element buffer); every chunk goes through FLOAT_pairwise_sum
// (numpy/core/src/umath/loops_utils.h.src: blocks of <= 128 elements with 8 interleaved accumulators, larger runs split
// recursively at n/2 rounded down to a multiple of 8); the chunk sums are accumulated in order in float32; the mean is
// float32(double(sum) / double(count)).
struct SliceView { const float* base; int ld, n, f0; };            // element i of the flattened slice = base[(i / n) * ld + f0 + i % n]
__device__ __forceinline__ float slice_at(const SliceView& v, int i) { return v.base[(size_t)(i / v.n) * v.ld + v.f0 + i % v.n]; }

__device__ float np_leaf_sum(const SliceView& v, int start, int n)   // n <= 128
{
    if (n < 8) {
        float res = 0.f;
        for (int i = 0; i < n; ++i) res = __fadd_rn(res, slice_at(v, start + i));
        return res;
    }
    float r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = slice_at(v, start + j);
    int i;
    for (i = 8; i < n - (n % 8); i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = __fadd_rn(r[j], slice_at(v, start + i + j));
    }
    float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])), __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
    for (; i < n; ++i) res = __fadd_rn(res, slice_at(v, start + i));
    return res;
}

constexpr int kMaxLeaves = 256;                 // leaves of one 8192-element chunk: >= 64 elements each
// leaves of pairwise_sum(start, n) in left-to-right order (explicit stack instead of recursion)
__device__ int np_list_leaves(int start, int n, int* leaf_start, int* leaf_len)
{
    int st_s[16], st_n[16], sp = 0, cnt = 0;
    st_s[0] = start; st_n[0] = n; sp = 1;
    while (sp > 0) {
        --sp;
        const int s0 = st_s[sp], m = st_n[sp];
        if (m <= 128) { leaf_start[cnt] = s0; leaf_len[cnt] = m; ++cnt; continue; }
        int n2 = m / 2; n2 -= n2 % 8;
        st_s[sp] = s0 + n2; st_n[sp] = m - n2; ++sp;          // right half below the left half: the left one pops first
        st_s[sp] = s0; st_n[sp] = n2; ++sp;
    }
    return cnt;
}
// the same tree folded over the leaf sums (post-order: res(node) = res(left) + res(right))
__device__ float np_combine(int n, const float* leaf_sum, int& next)
{
    if (n <= 128) return leaf_sum[next++];
    int n2 = n / 2; n2 -= n2 % 8;
    const float l = np_combine(n2, leaf_sum, next);
    const float r = np_combine(n - n2, leaf_sum, next);
    return __fadd_rn(l, r);
}

__global__ void __launch_bounds__(256)
augment_clips_kernel(float* __restrict__ x, const AugmentParams* __restrict__ params, int T, int F)
{
    const int b = blockIdx.x;
    const AugmentParams p = params[b];
    float* clip = x + (size_t)b * T * F;
    __shared__ float s_fill;
    // TimeMask: data[t0:t0+t, :] *= 0 (fade = False, the reference's default)
    if (p.tm_t > 0) {
        const int r0 = p.tm_t0, r1 = min(T, p.tm_t0 + p.tm_t);
        for (int i = threadIdx.x + r0 * F; i < r1 * F; i += 256) clip[i] = __fmul_rn(clip[i], 0.f);     // keeps -0.0 / NaN like numpy
        __syncthreads();
    }
    // FreqMask: data[:, f0:f0+f] = mean(data[:, f0:f0+f]) (np.mean in float32, numpy's exact summation order: see above)
    if (p.fm_mode != 0 && p.fm_f > 0) {
        const int f0 = p.fm_f0, f1 = min(F, p.fm_f0 + p.fm_f), n = f1 - f0;
        if (p.fm_mode == 2) {
            __shared__ int s_ls[kMaxLeaves], s_ll[kMaxLeaves], s_nleaf;
            __shared__ float s_lsum[kMaxLeaves], s_acc;
            const SliceView v{clip, F, n, f0};
            const int total = T * n, chunk = (8192 / n) * n;
            if (threadIdx.x == 0) s_acc = 0.f;
            for (int c0 = 0; c0 < total; c0 += chunk) {
                const int len = min(chunk, total - c0);
                if (threadIdx.x == 0) s_nleaf = np_list_leaves(c0, len, s_ls, s_ll);
                __syncthreads();
                for (int i = threadIdx.x; i < s_nleaf; i += 256) s_lsum[i] = np_leaf_sum(v, s_ls[i], s_ll[i]);
                __syncthreads();
                if (threadIdx.x == 0) { int next = 0; s_acc = __fadd_rn(s_acc, np_combine(len, s_lsum, next)); }
                __syncthreads();
            }
            if (threadIdx.x == 0) s_fill = (float)((double)s_acc / (double)total);
        } else if (threadIdx.x == 0) {
            s_fill = p.fm_const;
        }
        __syncthreads();
        const float fill = s_fill;
        for (int i = threadIdx.x; i < T * n; i += 256) clip[(size_t)(i / n) * F + f0 + i % n] = fill;
        __syncthreads();
    }
    // FreqShift: np.roll along the mel axis, the wrapped bins zeroed
    if (p.fs_shift != 0) {
        const int s = p.fs_shift;
        for (int r = threadIdx.x >> 5; r < T; r += 8) {                  // one warp per row: read the row, then write it shifted
            float* row = clip + (size_t)r * F;
            float v[8];                                                   // F <= 256
            const int lane = threadIdx.x & 31;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int f = lane + 32 * k, src = f - s;
                v[k] = (f < F && src >= 0 && src < F) ? row[src] : 0.f;
            }
            __syncwarp();
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int f = lane + 32 * k;
                if (f < F) row[f] = v[k];
            }
        }
    }
}

__global__ void __launch_bounds__(256)
mix_rows_kernel(const float* __restrict__ x, float* __restrict__ out, const MixRow* __restrict__ rows, int64_t row_elems)
{
    const MixRow r = rows[blockIdx.y];
    const float4* a = reinterpret_cast<const float4*>(x + (size_t)r.i1 * row_elems);
    const float4* b = reinterpret_cast<const float4*>(x + (size_t)r.i2 * row_elems);
    float4* o = reinterpret_cast<float4*>(out + (size_t)blockIdx.y * row_elems);
    const int64_t n4 = row_elems >> 2;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (int64_t)gridDim.x * 256) {
        float4 u = a[i];
        if (r.b != 0.f) {
            const float4 w = b[i];
            u.x = __fadd_rn(__fmul_rn(r.a, u.x), __fmul_rn(r.b, w.x)); u.y = __fadd_rn(__fmul_rn(r.a, u.y), __fmul_rn(r.b, w.y));
            u.z = __fadd_rn(__fmul_rn(r.a, u.z), __fmul_rn(r.b, w.z)); u.w = __fadd_rn(__fmul_rn(r.a, u.w), __fmul_rn(r.b, w.w));
        }
        o[i] = u;
    }
}

constexpr int kQueryRows = 128;        // transforms.Resize((128, 64))
constexpr int kMaxTaps = 24;           // ceil(support) * 2 + 1 with support = max(1, L / 128): clips of up to ~1400 frames

// One CTA per (clip, patch).  bounds: [B*P][2] = (s_idx, e_idx) from the host (float32 arithmetic of Query.transform_label)
__global__ void __launch_bounds__(256)
query_patches_kernel(const float* __restrict__ x, const int32_t* __restrict__ bounds, float* __restrict__ out, int P, int T, int F,
                     int fixed)
{
    __shared__ float red_min[8], red_max[8];
    __shared__ float s_min, s_max;
    __shared__ int s_first[kQueryRows], s_n[kQueryRows];
    __shared__ int s_w[kQueryRows][kMaxTaps];
    const int bp = blockIdx.x, b = bp / P;
    const int s_idx = bounds[2 * bp], e_idx = bounds[2 * bp + 1];
    const int L = e_idx - s_idx;
    const float* src = x + ((size_t)b * T + s_idx) * F;
    float* dst = out + (size_t)bp * kQueryRows * F;
    if (fixed) {                                            // data[:, e_idx-128:e_idx, :]
        for (int i = threadIdx.x; i < kQueryRows * F; i += 256) dst[i] = src[i];
        return;
    }
    float mn = CUDART_INF_F, mx = -CUDART_INF_F;
    for (int i = threadIdx.x; i < L * F; i += 256) { const float v = src[i]; mn = fminf(mn, v); mx = fmaxf(mx, v); }
    mn = -warp_max(-mn); mx = warp_max(mx);
    if ((threadIdx.x & 31) == 0) { red_min[threadIdx.x >> 5] = mn; red_max[threadIdx.x >> 5] = mx; }
    // Pillow precompute_coeffs (bilinear, support scaled by the down-sampling factor) + normalize_coeffs_8bpc, in double like
    // the C code; explicit _rn intrinsics keep the compiler from contracting to FMAs
    if (threadIdx.x < kQueryRows) {
        const int xx = threadIdx.x;
        const double scale = __ddiv_rn((double)L, (double)kQueryRows);
        const double filterscale = scale < 1.0 ? 1.0 : scale;
        const double support = filterscale;                  // bilinear support 1.0 * filterscale
        const double ss = __ddiv_rn(1.0, filterscale);
        const double center = __dmul_rn((double)xx + 0.5, scale);
        int xmin = (int)__dadd_rn(__dsub_rn(center, support), 0.5);
        if (xmin < 0) xmin = 0;
        int xmax = (int)__dadd_rn(__dadd_rn(center, support), 0.5);
        if (xmax > L) xmax = L;
        xmax -= xmin;
        if (xmax > kMaxTaps) xmax = kMaxTaps;               // not reached for L <= 1400 (checked by the launcher)
        double k[kMaxTaps], ww = 0.0;
        for (int j = 0; j < xmax; ++j) {
            double a = __dmul_rn(__dadd_rn(__dsub_rn((double)(j + xmin), center), 0.5), ss);
            a = fabs(a);
            k[j] = a < 1.0 ? __dsub_rn(1.0, a) : 0.0;
            ww = __dadd_rn(ww, k[j]);
        }
        for (int j = 0; j < xmax; ++j) {
            const double w = ww != 0.0 ? __ddiv_rn(k[j], ww) : k[j];
            s_w[xx][j] = w < 0 ? (int)__dsub_rn(__dmul_rn(w, 4194304.0), 0.5) : (int)__dadd_rn(__dmul_rn(w, 4194304.0), 0.5);
        }
        s_first[xx] = xmin; s_n[xx] = xmax;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = red_min[0], c = red_max[0];
        for (int w = 1; w < 8; ++w) { a = fminf(a, red_min[w]); c = fmaxf(c, red_max[w]); }
        s_min = a; s_max = c;
    }
    __syncthreads();
    mn = s_min; mx = s_max;
    const float range = __fsub_rn(mx, mn);
    for (int i = threadIdx.x; i < kQueryRows * F; i += 256) {
        const int yy = i / F, f = i - yy * F;
        int acc = 1 << 21;
        const int first = s_first[yy], n = s_n[yy];
        for (int j = 0; j < n; ++j) {
            const float norm = __fdiv_rn(__fsub_rn(src[(size_t)(first + j) * F + f], mn), range);
            const int u8 = (int)(unsigned char)(int)__fmul_rn(norm, 255.f);          // ToPILImage: pic.mul(255).byte()
            acc += u8 * s_w[yy][j];
        }
        int q = acc >> 22;
        q = q < 0 ? 0 : (q > 255 ? 255 : q);
        dst[i] = __fadd_rn(__fmul_rn(__fdiv_rn((float)q, 255.f), range), mn);
    }
}

}  // namespace

int launch_augment_clips(float* x, const AugmentParams* params, int B, int T, int F, float* row_sums, cudaStream_t stream)
{
    if (B == 0) return SEDT_OK;
    (void)row_sums;
    SEDT_REQUIRE(x && params && T >= 1 && F >= 1 && F <= 256, "augment_clips: bad arguments (F <= 256)");
    ProfScope _prof(PROF_OTHER, stream);
    augment_clips_kernel<<<(unsigned)B, 256, 0, stream>>>(x, params, T, F);
    SEDT_COUNT_LAUNCH();
    SEDT_CHECK_CUDA(cudaGetLastError());
    return SEDT_OK;
}

int launch_mix_rows(const float* x, float* out, const MixRow* rows, int n_out, int64_t row_elems, cudaStream_t stream)
{
    if (n_out == 0) return SEDT_OK;
    SEDT_REQUIRE(x && out && rows && row_elems > 0 && row_elems % 4 == 0, "mix_rows: row_elems must be a multiple of 4");
    SEDT_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)out & 15) == 0, "mix_rows: 16-byte aligned tensors");
    const unsigned gx = (unsigned)std::min<int64_t>(ceil_div(row_elems / 4, 256), 64);
    ProfScope _prof(PROF_OTHER, stream);
    mix_rows_kernel<<<dim3(gx, (unsigned)n_out), 256, 0, stream>>>(x, out, rows, row_elems);
    SEDT_COUNT_LAUNCH();
    SEDT_CHECK_CUDA(cudaGetLastError());
    return SEDT_OK;
}

int launch_query_patches(const float* x, const int32_t* bounds, float* out, int B, int P, int T, int F, int fixed, cudaStream_t stream)
{
    if (B * P == 0) return SEDT_OK;
    SEDT_REQUIRE(x && bounds && out && T >= 1 && F >= 1, "query_patches: bad arguments");
    SEDT_REQUIRE(T <= 1400, "query_patches: clips of up to 1400 frames (tap table)");
    ProfScope _prof(PROF_OTHER, stream);
    query_patches_kernel<<<(unsigned)(B * P), 256, 0, stream>>>(x, bounds, out, P, T, F, fixed);
    SEDT_COUNT_LAUNCH();
    SEDT_CHECK_CUDA(cudaGetLastError());
    return SEDT_OK;
}

}  // namespace sedt
