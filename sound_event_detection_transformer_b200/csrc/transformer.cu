// Transformer-side kernels that are bandwidth / latency bound (no contraction
// large enough for a tensor-core tile): LayerNorm(+pos), the sine position
// table, padding-mask resize, the multi-head attention core on <=128-token
// sequences, head finalisation, and the two SP-SEDT glue ops.
// Numerical recipe: SURVEY.md Appendix A.3, A.5-A.9 (sedt/transformer.py,
// sedt/position_encoding.py:28-47, torch functional.py:6630-6659).
#include "kernels.h"
#include <math_constants.h>
#include <cstdlib>

namespace sedt {
namespace {

constexpr int D = 256;   // hidden_dim of every documented recipe (train_sedt.py:86)

template <typename T> __device__ __forceinline__ void store8(T* p, const float (&v)[8]);
template <> __device__ __forceinline__ void store8<float>(float* p, const float (&v)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
template <> __device__ __forceinline__ void store8<__nv_bfloat16>(__nv_bfloat16* p, const float (&v)[8]) {
    uint4 u;
    __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
    __nv_bfloat162 c = __floats2bfloat162_rn(v[4], v[5]), d = __floats2bfloat162_rn(v[6], v[7]);
    u.x = *reinterpret_cast<uint32_t*>(&a); u.y = *reinterpret_cast<uint32_t*>(&b);
    u.z = *reinterpret_cast<uint32_t*>(&c); u.w = *reinterpret_cast<uint32_t*>(&d);
    *reinterpret_cast<uint4*>(p) = u;
}
__device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

// ---- LayerNorm: one warp per row, 8 contiguous elements per lane -------------
template <typename T, bool kNorm>
__global__ void __launch_bounds__(256)
layernorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                 const float* __restrict__ pos, int64_t pos_rows, T* __restrict__ y, T* __restrict__ ypos,
                 float* __restrict__ y32, int64_t rows)
{
    pdl_trigger();
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    float v[8];
    load8(x + row * D + lane * 8, v);
    if (kNorm) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) s += v[i];
        const float mean = warp_sum(s) * (1.f / D);
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) { const float d = v[i] - mean; q = fmaf(d, d, q); }
        const float rstd = rsqrtf(warp_sum(q) * (1.f / D) + 1e-5f);
        float g[8], bt[8];
        load8(gamma + lane * 8, g);
        load8(beta + lane * 8, bt);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = (v[i] - mean) * rstd * g[i] + bt[i];
    }
    if (y != nullptr) store8<T>(y + row * D + lane * 8, v);
    if (y32 != nullptr) store8<float>(y32 + row * D + lane * 8, v);
    if (ypos != nullptr) {
        float p[8];
        load8(pos + (row % pos_rows) * D + lane * 8, p);
#pragma unroll
        for (int i = 0; i < 8; ++i) p[i] += v[i];
        store8<T>(ypos + row * D + lane * 8, p);
    }
}

// ---- attention core ------------------------------------------------------------
// CTA = (query block, head group, clip); thread = (head in group, query row).
// K/V head slices of one 128-key tile are staged in shared memory as fp32 and
// read as warp-wide broadcasts; softmax is the online (running max / sum) form
// updated every 8 keys, all in fp32.
constexpr int KT = 128;   // keys per shared-memory tile
constexpr int HD = 32;    // head_dim

template <typename T, int HG, int QB>
__global__ void __launch_bounds__(HG * QB)
attention_kernel(const T* __restrict__ Q, int ldq, const T* __restrict__ K, int ldk, const T* __restrict__ V, int ldv,
                 T* __restrict__ O, int ldo, const uint8_t* __restrict__ kpm, const float* __restrict__ amask,
                 int Lq, int Lk, float scale)
{
    extern __shared__ __align__(16) float smem[];
    float* Ks = smem;                       // [HG][KT][HD]
    float* Vs = smem + HG * KT * HD;
    const int tid = threadIdx.x;
    const int hl = tid / QB, ql = tid % QB;
    const int b = blockIdx.z, h0 = blockIdx.y * HG;
    const int qrow = blockIdx.x * QB + ql;
    const bool active = qrow < Lq;

    float q[HD], o[HD];
#pragma unroll
    for (int d = 0; d < HD; ++d) { q[d] = 0.f; o[d] = 0.f; }
    if (active) {
        const T* qp = Q + ((size_t)b * Lq + qrow) * ldq + (h0 + hl) * HD;
#pragma unroll
        for (int d = 0; d < HD; ++d) q[d] = to_f32<T>(qp[d]) * scale;     // q * sqrt(1/hd) first (functional.py:6636)
    }
    float m = -CUDART_INF_F, l = 0.f;

    for (int k0 = 0; k0 < Lk; k0 += KT) {
        const int kt = min(KT, Lk - k0);
        __syncthreads();
        for (int i = tid; i < HG * kt * (HD / 4); i += HG * QB) {
            const int d4 = i % (HD / 4);
            const int j = (i / (HD / 4)) % kt;
            const int hh = i / ((HD / 4) * kt);
            const T* kp = K + ((size_t)b * Lk + k0 + j) * ldk + (h0 + hh) * HD + d4 * 4;
            const T* vp = V + ((size_t)b * Lk + k0 + j) * ldv + (h0 + hh) * HD + d4 * 4;
            float4 kv = make_float4(to_f32<T>(kp[0]), to_f32<T>(kp[1]), to_f32<T>(kp[2]), to_f32<T>(kp[3]));
            float4 vv = make_float4(to_f32<T>(vp[0]), to_f32<T>(vp[1]), to_f32<T>(vp[2]), to_f32<T>(vp[3]));
            *reinterpret_cast<float4*>(Ks + ((size_t)hh * KT + j) * HD + d4 * 4) = kv;
            *reinterpret_cast<float4*>(Vs + ((size_t)hh * KT + j) * HD + d4 * 4) = vv;
        }
        __syncthreads();
        if (!active) continue;
        const float* kh = Ks + (size_t)hl * KT * HD;
        const float* vh = Vs + (size_t)hl * KT * HD;
        for (int j0 = 0; j0 < kt; j0 += 8) {
            float s[8];
            float mb = -CUDART_INF_F;
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                const int j = j0 + t;
                float acc = -CUDART_INF_F;
                if (j < kt) {
                    acc = 0.f;
                    const float4* kr = reinterpret_cast<const float4*>(kh + j * HD);
#pragma unroll
                    for (int d4 = 0; d4 < HD / 4; ++d4) {
                        const float4 kk = kr[d4];
                        acc = fmaf(q[4 * d4 + 0], kk.x, acc); acc = fmaf(q[4 * d4 + 1], kk.y, acc);
                        acc = fmaf(q[4 * d4 + 2], kk.z, acc); acc = fmaf(q[4 * d4 + 3], kk.w, acc);
                    }
                    if (amask != nullptr) acc += amask[(size_t)qrow * Lk + k0 + j];
                    if (kpm != nullptr && kpm[(size_t)b * Lk + k0 + j]) acc = -CUDART_INF_F;
                }
                s[t] = acc;
                mb = fmaxf(mb, acc);
            }
            const float mn = fmaxf(m, mb);
            if (mn == -CUDART_INF_F) continue;            // everything masked so far
            const float corr = expf(m - mn);              // m = -inf -> 0
            l *= corr;
#pragma unroll
            for (int d = 0; d < HD; ++d) o[d] *= corr;
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                const int j = j0 + t;
                if (j >= kt) break;
                const float p = expf(s[t] - mn);
                l += p;
                const float4* vr = reinterpret_cast<const float4*>(vh + j * HD);
#pragma unroll
                for (int d4 = 0; d4 < HD / 4; ++d4) {
                    const float4 vv = vr[d4];
                    o[4 * d4 + 0] = fmaf(p, vv.x, o[4 * d4 + 0]); o[4 * d4 + 1] = fmaf(p, vv.y, o[4 * d4 + 1]);
                    o[4 * d4 + 2] = fmaf(p, vv.z, o[4 * d4 + 2]); o[4 * d4 + 3] = fmaf(p, vv.w, o[4 * d4 + 3]);
                }
            }
            m = mn;
        }
    }
    if (active) {
        const float inv = 1.f / l;
        T* op = O + ((size_t)b * Lq + qrow) * ldo + (h0 + hl) * HD;
#pragma unroll
        for (int d = 0; d < HD; ++d) op[d] = from_f32<T>(o[d] * inv);
    }
}

// ---- padding mask nearest-resize (sedt/backbone.py:81) ----------------------------
__global__ void mask_downsample_kernel(const uint8_t* __restrict__ mask, uint8_t* __restrict__ out, int B, int T, int F,
                                       int H, int W)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * H * W) return;
    const int w = i % W, h = (i / W) % H, b = i / (W * H);
    // torch nearest: src = min(floor(dst * (float)in / out), in - 1), in fp32
    const float sh = (float)T / (float)H, sw = (float)F / (float)W;
    const int th = min((int)floorf((float)h * sh), T - 1);
    const int tw = min((int)floorf((float)w * sw), F - 1);
    out[i] = mask[((size_t)b * T + th) * F + tw];
}

// ---- sine position table (sedt/position_encoding.py:28-47) -------------------------
__global__ void pos_table_kernel(const uint8_t* __restrict__ mask_ds, float* __restrict__ pos, int H, int W)
{
    // grid (H*W, nb); block 256 = feature index
    const int s = blockIdx.x, b = blockIdx.y, i = threadIdx.x;
    const int h = s / W, w = s % W;
    float y = (float)(h + 1), last = (float)H;
    if (mask_ds != nullptr) {
        int cy = 0, cl = 0;
        for (int hh = 0; hh < H; ++hh) {
            const int nm = mask_ds[((size_t)b * H + hh) * W + w] ? 0 : 1;
            cl += nm;
            if (hh <= h) cy += nm;
        }
        y = (float)cy; last = (float)cl;
    }
    const float ye = __fmul_rn(__fdiv_rn(y, __fadd_rn(last, 1e-6f)), 6.283185307179586f);
    const float expo = __fdiv_rn(2.f * (float)(i / 2), 256.f);
    const float dim_t = powf(10000.f, expo);
    const float v = __fdiv_rn(ye, dim_t);
    pos[((size_t)b * H * W + s) * D + i] = (i & 1) ? cosf(v) : sinf(v);
}

// ---- heads: slice decoder slots, sigmoid (sedt/sedt.py:89-95) ----------------------
__device__ __forceinline__ float sigmoidf(float x) { return 1.f / (1.f + expf(-x)); }

__global__ void heads_finalize_kernel(const float* __restrict__ cls_raw, const float* __restrict__ box_raw,
                                      const float* __restrict__ weak_raw, float* __restrict__ logits,
                                      float* __restrict__ boxes, float* __restrict__ at, int D_, int B, int Qall, int start,
                                      int C1, int C)
{
    const int Q = Qall - start;
    const int64_t nl = (int64_t)D_ * B * Q * C1, nb = (int64_t)D_ * B * Q * 2, na = (at != nullptr) ? (int64_t)B * C : 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nl + nb + na; i += (int64_t)gridDim.x * blockDim.x) {
        if (i < nl) {
            const int c = (int)(i % C1);
            const int64_t r = i / C1;                 // (d*B + b)*Q + q
            const int q = (int)(r % Q);
            const int64_t db = r / Q;
            logits[i] = cls_raw[(db * Qall + q + start) * C1 + c];
        } else if (i < nl + nb) {
            const int64_t k = i - nl;
            const int c = (int)(k % 2);
            const int64_t r = k / 2;
            const int q = (int)(r % Q);
            const int64_t db = r / Q;
            boxes[k] = sigmoidf(box_raw[(db * Qall + q + start) * 2 + c]);
        } else {
            const int64_t k = i - nl - nb;
            at[k] = sigmoidf(weak_raw[k]);
        }
    }
}

// The three output projections in one pass (sedt/sedt.py:89-95): class_embed on the event slots, the last bbox_embed layer
// + sigmoid, weak_class_embed + sigmoid on slot 0 of the last decoder layer.  One warp per decoder-state row; the 13..23
// weight rows live in shared memory; replaces three 64x64-tile CUDA-core GEMMs (29 + 28 + 22 us at B = 256) and the
// slice / sigmoid pass that followed them.
// four output features at a time: independent shuffle chains instead of one serial reduction per output
__device__ __forceinline__ void dot4(const float (&x)[8], const float* w, int n, int lane, float (&out)[4])
{
    float a[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (j < n) {
            const float4 w0 = *reinterpret_cast<const float4*>(w + j * D + lane * 8);
            const float4 w1 = *reinterpret_cast<const float4*>(w + j * D + lane * 8 + 4);
            a[j] = fmaf(x[0], w0.x, fmaf(x[1], w0.y, fmaf(x[2], w0.z, fmaf(x[3], w0.w,
                   fmaf(x[4], w1.x, fmaf(x[5], w1.y, fmaf(x[6], w1.z, x[7] * w1.w)))))));
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int j = 0; j < 4; ++j) a[j] += __shfl_xor_sync(0xffffffffu, a[j], o);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) out[j] = a[j];
}

__device__ __forceinline__ float pick4(const float (&a)[4], int i) { return i == 0 ? a[0] : (i == 1 ? a[1] : (i == 2 ? a[2] : a[3])); }

__global__ void __launch_bounds__(256)
heads_out_kernel(const float* __restrict__ hs, const float* __restrict__ h2, const float* __restrict__ wc, const float* __restrict__ bc,
                 const float* __restrict__ wb, const float* __restrict__ bb, const float* __restrict__ ww, const float* __restrict__ bw,
                 float* __restrict__ logits, float* __restrict__ boxes, float* __restrict__ at, int D_, int B, int Qall, int start,
                 int C1, int C)
{
    extern __shared__ float hsm[];
    float* s_wc = hsm;                       // [C1][256]
    float* s_wb = s_wc + C1 * D;             // [2][256]
    float* s_ww = s_wb + 2 * D;              // [C][256] (dec_at only)
    pdl_trigger();
    for (int i = threadIdx.x; i < C1 * D; i += blockDim.x) s_wc[i] = wc[i];
    for (int i = threadIdx.x; i < 2 * D; i += blockDim.x) s_wb[i] = wb[i];
    if (at != nullptr) for (int i = threadIdx.x; i < C * D; i += blockDim.x) s_ww[i] = ww[i];
    pdl_wait();
    __syncthreads();
    const int lane = threadIdx.x & 31, warps = blockDim.x >> 5;
    const int Q = Qall - start;
    const int64_t rows = (int64_t)D_ * B * Qall;
    for (int64_t r = (int64_t)blockIdx.x * warps + (threadIdx.x >> 5); r < rows; r += (int64_t)gridDim.x * warps) {
        const int q = (int)(r % Qall);
        const int64_t lb = r / Qall;                         // l * B + b
        const bool is_at = at != nullptr && q == 0 && lb >= (int64_t)(D_ - 1) * B;
        if (q < start && !is_at) continue;
        float x[8], y[8];
        load8(hs + r * D + lane * 8, x);
        if (q >= start) load8(h2 + r * D + lane * 8, y);          // both rows in flight before the first reduction
        if (q >= start) {
            const int64_t o = lb * Q + (q - start);
            for (int c0 = 0; c0 < C1; c0 += 4) {
                float a[4];
                dot4(x, s_wc + c0 * D, C1 - c0, lane, a);
                if (lane < 4 && c0 + lane < C1) logits[o * C1 + c0 + lane] = pick4(a, lane) + bc[c0 + lane];
            }
            float a[4];
            dot4(y, s_wb, 2, lane, a);
            if (lane < 2) boxes[o * 2 + lane] = sigmoidf(pick4(a, lane) + bb[lane]);
        }
        if (is_at) {
            const int64_t b = lb - (int64_t)(D_ - 1) * B;
            for (int c0 = 0; c0 < C; c0 += 4) {
                float a[4];
                dot4(x, s_ww + c0 * D, C - c0, lane, a);
                if (lane < 4 && c0 + lane < C) at[b * C + c0 + lane] = sigmoidf(pick4(a, lane) + bw[c0 + lane]);
            }
        }
    }
}

template <typename T>
__global__ void avgpool_kernel(const T* __restrict__ x, float* __restrict__ out, int HW, int C)
{
    // grid (C/256, N); thread = channel
    const int c = blockIdx.x * blockDim.x + threadIdx.x, n = blockIdx.y;
    if (c >= C) return;
    float s = 0.f;
    for (int p = 0; p < HW; ++p) s += to_f32<T>(x[((size_t)n * HW + p) * C + c]);
    out[(size_t)n * C + c] = s / (float)HW;
}

__global__ void patch_query_kernel(const float* __restrict__ pq, const float* __restrict__ qe, float* __restrict__ out,
                                   int B, int P, int qpp, int start, const uint8_t* __restrict__ keep, float qe_scale)
{
    const int Q = P * qpp;
    const int64_t total = (int64_t)B * Q * D;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int d = (int)(i % D);
        const int q = (int)((i / D) % Q);
        const int b = (int)(i / ((int64_t)D * Q));
        const float pf = pq[((size_t)b * P + q / qpp) * D + d];
        const float k = keep == nullptr ? 1.f : (keep[(size_t)b * Q + q] ? 1.f : 0.f);
        out[i] = pf * k + qe_scale * qe[(size_t)(start + q) * D + d];
    }
}

__global__ void accum_bf16_kernel(const __nv_bfloat16* __restrict__ g, float* __restrict__ acc, int64_t n)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        acc[i] += __bfloat162float(g[i]);
}

// one thread per (q or (b, p), d): fixed summation order, no atomics
__global__ void patch_query_bwd_kernel(const float* __restrict__ dq, const uint8_t* __restrict__ keep, float* __restrict__ d_qe,
                                       float* __restrict__ d_pq, int B, int P, int qpp, float qe_scale)
{
    const int Q = P * qpp;
    const int64_t n_qe = (int64_t)Q * D, n_pq = (int64_t)B * P * D;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_qe + n_pq; i += (int64_t)gridDim.x * blockDim.x) {
        if (i < n_qe) {
            const int d = (int)(i % D), q = (int)(i / D);
            float s = 0.f;
            for (int b = 0; b < B; ++b) s += dq[((size_t)b * Q + q) * D + d];
            d_qe[i] += qe_scale * s;
        } else {
            const int64_t j = i - n_qe;
            const int d = (int)(j % D), p = (int)((j / D) % P), b = (int)(j / ((int64_t)D * P));
            float s = 0.f;
            for (int k = 0; k < qpp; ++k) {
                const int q = p * qpp + k;
                if (keep == nullptr || keep[(size_t)b * Q + q]) s += dq[((size_t)b * Q + q) * D + d];
            }
            d_pq[j] = s;
        }
    }
}

__global__ void blockdiag_mask_kernel(float* __restrict__ m, int Q, int qpp)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Q * Q) return;
    m[i] = ((i / Q) / qpp == (i % Q) / qpp) ? 0.f : -CUDART_INF_F;      // sedt/spsedt.py:27-32
}

}  // namespace

int launch_blockdiag_mask(float* m, int Q, int qpp, cudaStream_t stream)
{
    ProfScope _prof(PROF_OTHER, stream);
    blockdiag_mask_kernel<<<(unsigned)ceil_div((int64_t)Q * Q, 256), 256, 0, stream>>>(m, Q, qpp);
    SEDT_COUNT_LAUNCH();
    SEDT_CHECK_CUDA(cudaGetLastError());
    return SEDT_OK;
}

int launch_layernorm(const float* x, const float* gamma, const float* beta, const float* pos, int64_t pos_rows,
                     void* y, void* ypos, float* y32, int dt, int64_t rows, cudaStream_t stream)
{
    if (rows == 0) return SEDT_OK;
    dim3 grid((unsigned)ceil_div(rows, 8)), block(256);
    ProfScope _prof(PROF_NORM, stream);
    if (dt == DT_F32) SEDT_CHECK_CUDA(launch_pdl(layernorm_kernel<float, true>, grid, block, 0, stream, 1, x, gamma, beta, pos, pos_rows, (float*)y, (float*)ypos, y32, rows));
    else SEDT_CHECK_CUDA(launch_pdl(layernorm_kernel<__nv_bfloat16, true>, grid, block, 0, stream, 1, x, gamma, beta, pos, pos_rows, (__nv_bfloat16*)y, (__nv_bfloat16*)ypos, y32, rows));
    SEDT_COUNT_LAUNCH();
    SEDT_CHECK_CUDA(cudaGetLastError());
    return SEDT_OK;
}

int launch_cast_addpos(const float* x, const float* pos, int64_t pos_rows, void* y, void* ypos, int dt, int64_t rows,
                       cudaStream_t stream)
{
    if (rows == 0) return SEDT_OK;
    dim3 grid((unsigned)ceil_div(rows, 8)), block(256);
    ProfScope _prof(PROF_NORM, stream);
    if (dt == DT_F32) SEDT_CHECK_CUDA(launch_pdl(layernorm_kernel<float, false>, grid, block, 0, stream, 1, x, (const float*)nullptr, (const float*)nullptr, pos, pos_rows, (float*)y, (float*)ypos, (float*)nullptr, rows));
    else SEDT_CHECK_CUDA(launch_pdl(layernorm_kernel<__nv_bfloat16, false>, grid, block, 0, stream, 1, x, (const float*)nullptr, (const float*)nullptr, pos, pos_rows, (__nv_bfloat16*)y, (__nv_bfloat16*)ypos, (float*)nullptr, rows));
    SEDT_COUNT_LAUNCH();
    SEDT_CHECK_CUDA(cudaGetLastError());
    return SEDT_OK;
}

template <typename T, int HG, int QB>
static int attention_launch(const void* Q, int ldq, const void* K, int ldk, const void* V, int ldv, void* O, int ldo,
                            const uint8_t* kpm, const float* amask, int B, int nheads, int Lq, int Lk, float scale,
                            cudaStream_t stream)
{
    const size_t smem = (size_t)2 * HG * KT * HD * sizeof(float);
    SEDT_CHECK_CUDA(cudaFuncSetAttribute(attention_kernel<T, HG, QB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)ceil_div(Lq, QB), (unsigned)(nheads / HG), (unsigned)B), block(HG * QB);
    ProfScope _prof(PROF_ATTENTION, stream);
    attention_kernel<T, HG, QB><<<grid, block, smem, stream>>>((const T*)Q, ldq, (const T*)K, ldk, (const T*)V, ldv,
                                                               (T*)O, ldo, kpm, amask, Lq, Lk, scale);
    SEDT_COUNT_LAUNCH();
    SEDT_CHECK_CUDA(cudaGetLastError());
    return SEDT_OK;
}

int launch_attention(const void* Q, int ldq, const void* K, int ldk, const void* V, int ldv, void* O, int ldo, int dt,
                     const uint8_t* kpm, const float* amask, int B, int nheads, int Lq, int Lk, float scale,
                     cudaStream_t stream)
{
    if (B == 0 || Lq == 0) return SEDT_OK;
    SEDT_REQUIRE(Lk >= 1, "attention: Lk=%d", Lk);
    SEDT_REQUIRE(nheads % 4 == 0, "attention: nheads=%d must be a multiple of 4", nheads);
    static const bool force_simt = [] { const char* e = getenv("SEDT_ATT_SIMT"); return e != nullptr && e[0] == '1'; }();
    if (!force_simt && attention_tc_supported(Q, ldq, K, ldk, V, ldv, O, ldo, dt, nheads, Lq, Lk))
        return launch_attention_tc(Q, ldq, K, ldk, V, ldv, O, ldo, kpm, amask, B, nheads, Lq, Lk, scale, stream);
    const bool small_q = Lq <= 32;      // decoder: 11 / 21 queries
#define SEDT_ATT(T)                                                                                                   \
    (small_q ? attention_launch<T, 4, 32>(Q, ldq, K, ldk, V, ldv, O, ldo, kpm, amask, B, nheads, Lq, Lk, scale, stream) \
             : attention_launch<T, 1, 128>(Q, ldq, K, ldk, V, ldv, O, ldo, kpm, amask, B, nheads, Lq, Lk, scale, stream))
    return dt == DT_F32 ? SEDT_ATT(float) : SEDT_ATT(__nv_bfloat16);
#undef SEDT_ATT
}

int launch_mask_downsample(const uint8_t* mask, uint8_t* out, int B, int T, int F, int H, int W, cudaStream_t stream)
{
    const int n = B * H * W;
    if (n == 0) return SEDT_OK;
    ProfScope _prof(PROF_OTHER, stream);
    mask_downsample_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, stream>>>(mask, out, B, T, F, H, W);
    SEDT_COUNT_LAUNCH();
    SEDT_CHECK_CUDA(cudaGetLastError());
    return SEDT_OK;
}

int launch_pos_table(const uint8_t* mask_ds, float* pos, int nb, int H, int W, cudaStream_t stream)
{
    SEDT_REQUIRE(mask_ds != nullptr || nb == 1, "pos_table: an unpadded table is batch-invariant (nb must be 1)");
    dim3 grid((unsigned)(H * W), (unsigned)nb), block(D);
    ProfScope _prof(PROF_OTHER, stream);
    pos_table_kernel<<<grid, block, 0, stream>>>(mask_ds, pos, H, W);
    SEDT_COUNT_LAUNCH();
    SEDT_CHECK_CUDA(cudaGetLastError());
    return SEDT_OK;
}

int launch_heads_out(const float* hs, const float* h2, const float* wc, const float* bc, const float* wb, const float* bb,
                     const float* ww, const float* bw, float* logits, float* boxes, float* at, int D_, int B, int Qall, int start,
                     int C1, int C, cudaStream_t stream)
{
    const int64_t rows = (int64_t)D_ * B * Qall;
    if (rows == 0) return SEDT_OK;
    const size_t smem = (size_t)(C1 + 2 + (at != nullptr ? C : 0)) * D * sizeof(float);
    SEDT_REQUIRE(smem <= 200 * 1024, "heads: %d classes do not fit shared memory", C1);
    if (smem > 48 * 1024)
        SEDT_CHECK_CUDA(cudaFuncSetAttribute(heads_out_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const unsigned grid = (unsigned)std::min<int64_t>(ceil_div(rows, 8), 148 * 8);
    ProfScope _prof(PROF_OTHER, stream);
    SEDT_CHECK_CUDA(launch_pdl(heads_out_kernel, dim3(grid), dim3(256), smem, stream, 1, hs, h2, wc, bc, wb, bb, ww, bw, logits, boxes,
                               at, D_, B, Qall, start, C1, C));
    SEDT_COUNT_LAUNCH();
    SEDT_CHECK_CUDA(cudaGetLastError());
    return SEDT_OK;
}

int launch_heads_finalize(const float* cls_raw, const float* box_raw, const float* weak_raw, float* logits, float* boxes,
                          float* at, int D_, int B, int Qall, int start, int C1, int C, cudaStream_t stream)
{
    const int64_t n = (int64_t)D_ * B * (Qall - start) * (C1 + 2) + (at ? (int64_t)B * C : 0);
    if (n == 0) return SEDT_OK;
    int64_t g = ceil_div(n, 256); if (g > 148 * 8) g = 148 * 8;
    ProfScope _prof(PROF_OTHER, stream);
    heads_finalize_kernel<<<(unsigned)g, 256, 0, stream>>>(cls_raw, box_raw, weak_raw, logits, boxes, at, D_, B, Qall, start, C1, C);
    SEDT_COUNT_LAUNCH();
    SEDT_CHECK_CUDA(cudaGetLastError());
    return SEDT_OK;
}

int launch_avgpool(const void* x, int dt, float* out, int N, int HW, int C, cudaStream_t stream)
{
    if (N == 0) return SEDT_OK;
    dim3 grid((unsigned)ceil_div(C, 256), (unsigned)N), block(256);
    ProfScope _prof(PROF_OTHER, stream);
    if (dt == DT_F32) avgpool_kernel<float><<<grid, block, 0, stream>>>((const float*)x, out, HW, C);
    else avgpool_kernel<__nv_bfloat16><<<grid, block, 0, stream>>>((const __nv_bfloat16*)x, out, HW, C);
    SEDT_COUNT_LAUNCH();
    SEDT_CHECK_CUDA(cudaGetLastError());
    return SEDT_OK;
}

int launch_accum_bf16(const void* g, float* acc, int64_t n, cudaStream_t stream)
{
    if (n == 0) return SEDT_OK;
    int64_t gr = ceil_div(n, 256); if (gr > 148 * 8) gr = 148 * 8;
    accum_bf16_kernel<<<(unsigned)gr, 256, 0, stream>>>((const __nv_bfloat16*)g, acc, n);
    SEDT_COUNT_LAUNCH();
    SEDT_CHECK_CUDA(cudaGetLastError());
    return SEDT_OK;
}

int launch_patch_query_bwd(const float* dq, const uint8_t* keep, float* d_qe, float* d_pq, int B, int P, int qpp, float qe_scale,
                           cudaStream_t stream)
{
    const int64_t n = (int64_t)(P * qpp + B * P) * D;
    if (n == 0) return SEDT_OK;
    int64_t gr = ceil_div(n, 256); if (gr > 148 * 8) gr = 148 * 8;
    patch_query_bwd_kernel<<<(unsigned)gr, 256, 0, stream>>>(dq, keep, d_qe, d_pq, B, P, qpp, qe_scale);
    SEDT_COUNT_LAUNCH();
    SEDT_CHECK_CUDA(cudaGetLastError());
    return SEDT_OK;
}

int launch_patch_query(const float* pq, const float* query_embed, float* out, int B, int P, int qpp, int start,
                       cudaStream_t stream, const uint8_t* keep, float qe_scale)
{
    const int64_t n = (int64_t)B * P * qpp * D;
    if (n == 0) return SEDT_OK;
    int64_t g = ceil_div(n, 256); if (g > 148 * 8) g = 148 * 8;
    ProfScope _prof(PROF_OTHER, stream);
    patch_query_kernel<<<(unsigned)g, 256, 0, stream>>>(pq, query_embed, out, B, P, qpp, start, keep, qe_scale);
    SEDT_COUNT_LAUNCH();
    SEDT_CHECK_CUDA(cudaGetLastError());
    return SEDT_OK;
}

}  // namespace sedt
