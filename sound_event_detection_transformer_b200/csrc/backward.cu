// Backward-pass kernels that are not GEMMs (bf16 tier): weight re-layout for the data-gradient GEMMs,
// zero insertion for stride-2 transposed convolutions, ReLU masking, column sums (bias gradients),
// LayerNorm backward and the attention-core backward.  The GEMM halves of the backward pass reuse the
// forward implicit-GEMM kernels (data gradients, with the re-laid-out weights) and gemm_wgrad.cu.
// Together they replace autograd's backward of sedt/backbone.py, sedt/transformer.py and the heads
// in sedt/sedt.py:89-95.
#include "kernels.h"
#include <math_constants.h>

namespace sedt {
namespace {

using bf16 = __nv_bfloat16;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- weights for the data gradient -----------------------------------------------------------
// dX = conv(dY, Wd) with Wd[ci][r'][s'][co] = scale[co] * W[co][ci][R-1-r'][S-1-s']   (W is OIHW fp32)
// The Cout axis (the reduction axis of the data-gradient GEMM) is zero-padded to Cout_pad.
template <typename T>
__global__ void repack_dgrad_kernel(const float* __restrict__ w, const float* __restrict__ scale, T* __restrict__ out, int Cout,
                                    int Cout_pad, int Cin, int R, int S)
{
    const int64_t total = (int64_t)Cout_pad * Cin * R * S;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int co = (int)(i % Cout_pad);
        int64_t t = i / Cout_pad;
        const int s = (int)(t % S); t /= S;
        const int r = (int)(t % R);
        const int ci = (int)(t / R);
        float v = 0.f;
        if (co < Cout) {
            v = w[(((int64_t)co * Cin + ci) * R + (R - 1 - r)) * S + (S - 1 - s)];
            if (scale != nullptr) v *= scale[co];
        }
        out[i] = from_f32<T>(v);
    }
}

// Same re-layout through a 32 x 32 shared-memory tile: reads run along Cin (stride R*S floats), writes along Cout.
// grid (ceil(Cin/32), ceil(Cout_pad/32), R*S), block (32, 8)
template <typename T>
__global__ void repack_dgrad_tiled_kernel(const float* __restrict__ w, const float* __restrict__ scale, T* __restrict__ out, int Cout,
                                          int Cout_pad, int Cin, int RS)
{
    __shared__ float tile[32][33];
    const int tap = blockIdx.z, src_tap = RS - 1 - tap;          // 180-degree rotation = reversed tap order
    const int ci0 = blockIdx.x * 32, co0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += 8) {
        const int co = co0 + j, ci = ci0 + threadIdx.x;
        float v = 0.f;
        if (co < Cout && ci < Cin) {
            v = w[((int64_t)co * Cin + ci) * RS + src_tap];
            if (scale != nullptr) v *= scale[co];
        }
        tile[j][threadIdx.x] = v;
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += 8) {
        const int ci = ci0 + j, co = co0 + threadIdx.x;
        if (ci < Cin && co < Cout_pad) out[((int64_t)ci * RS + tap) * Cout_pad + co] = from_f32<T>(tile[threadIdx.x][j]);
    }
}

// All data-gradient weight re-layouts of one backward pass in one launch (square filters: R == S); a CTA is one 32 x 32
// tile of one job, found through the tile prefix.  Jobs travel in the kernel parameters.
__global__ void __launch_bounds__(256)
repack_dgrad_batched_kernel(const __grid_constant__ DgradJobs jobs)
{
    __shared__ float tile[32][33];
    int lo = 0, hi = jobs.n - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (jobs.j[mid].tile0 <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
    }
    const DgradJob& J = jobs.j[lo];
    const int Cout = J.Cout, Cout_pad = J.Cout_pad, Cin = J.Cin, RS = J.RS;
    const int tx = (Cin + 31) / 32, ty = (Cout_pad + 31) / 32;
    int t = (int)blockIdx.x - J.tile0;
    const int bx = t % tx; t /= tx;
    const int by = t % ty;
    const int tap = t / ty, src_tap = RS - 1 - tap;
    const int ci0 = bx * 32, co0 = by * 32;
    const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
    for (int j = ly; j < 32; j += 8) {
        const int co = co0 + j, ci = ci0 + lx;
        float v = 0.f;
        if (co < Cout && ci < Cin) {
            v = J.w[((int64_t)co * Cin + ci) * RS + src_tap];
            if (J.scale != nullptr) v *= J.scale[co];
        }
        tile[j][lx] = v;
    }
    __syncthreads();
    bf16* out = (bf16*)J.out;
    for (int j = ly; j < 32; j += 8) {
        const int ci = ci0 + j, co = co0 + lx;
        if (ci < Cin && co < Cout_pad) out[((int64_t)ci * RS + tap) * Cout_pad + co] = __float2bfloat16_rn(tile[lx][j]);
    }
}

// grad[co][ci][r][s] = scale[co] * dw[co][(r,s)][ci]   (weight gradient back in the reference's OIHW layout)
__global__ void unpack_wgrad_kernel(const float* __restrict__ dw, const float* __restrict__ scale, float* __restrict__ grad, int Cout,
                                    int Cin, int RS)
{
    const int64_t total = (int64_t)Cout * Cin * RS;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int tap = (int)(i % RS);
        const int64_t t = i / RS;
        const int ci = (int)(t % Cin), co = (int)(t / Cin);
        float v = dw[((int64_t)co * RS + tap) * Cin + ci];
        if (scale != nullptr) v *= scale[co];
        grad[i] = v;
    }
}

// Gradients of the head outputs -> zero-padded bf16 GEMM operands (128 columns each):
//   dcls[(l,b,q), c]  = d_logits[l,b,q-start,c]                       (0 for the audio query slot)
//   dbox[(l,b,q), c]  = d_boxes[l,b,q-start,c] * s(1-s), s = boxes   (sigmoid backward)
//   dweak[b, c]       = d_at[b,c] * a(1-a)
__global__ void heads_bwd_prepare_kernel(const float* __restrict__ d_logits, const float* __restrict__ d_boxes,
                                         const float* __restrict__ d_at, const float* __restrict__ boxes,
                                         const float* __restrict__ at, bf16* __restrict__ dcls, bf16* __restrict__ dbox,
                                         bf16* __restrict__ dweak, int D_, int B, int Qall, int start, int C1, int C)
{
    const int Q = Qall - start;
    const int64_t hrows = (int64_t)D_ * B * Qall;
    const int64_t n1 = hrows * 128, n3 = dweak != nullptr ? (int64_t)B * 128 : 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < 2 * n1 + n3; i += (int64_t)gridDim.x * blockDim.x) {
        if (i < 2 * n1) {
            const bool is_box = i >= n1;
            const int64_t k = is_box ? i - n1 : i;
            const int c = (int)(k % 128);
            const int64_t r = k / 128;
            const int q = (int)(r % Qall) - start;
            const int64_t db = r / Qall;
            float v = 0.f;
            if (q >= 0) {
                if (!is_box) { if (c < C1 && d_logits != nullptr) v = d_logits[(db * Q + q) * C1 + c]; }
                else if (c < 2 && d_boxes != nullptr) {
                    const float sg = boxes[(db * Q + q) * 2 + c];
                    v = d_boxes[(db * Q + q) * 2 + c] * sg * (1.f - sg);
                }
            }
            (is_box ? dbox : dcls)[k] = __float2bfloat16_rn(v);
        } else {
            const int64_t k = i - 2 * n1;
            const int c = (int)(k % 128);
            const int64_t b = k / 128;
            float v = 0.f;
            if (c < C && d_at != nullptr) { const float a = at[b * C + c]; v = d_at[b * C + c] * a * (1.f - a); }
            dweak[k] = __float2bfloat16_rn(v);
        }
    }
}

// ---- U[b, 2*ho, 2*wo, :] = dY[b, ho, wo, :], zero elsewhere (U is [B,H,W,C]); 8 channels per thread
__global__ void upsample2_kernel(const uint4* __restrict__ dy, uint4* __restrict__ u, int B, int H, int W, int Ho, int Wo, int C8)
{
    const int64_t total = (int64_t)B * H * W * C8;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C8);
        int64_t t = i / C8;
        const int w = (int)(t % W); t /= W;
        const int h = (int)(t % H);
        const int b = (int)(t / H);
        uint4 v = make_uint4(0, 0, 0, 0);
        if (!(h & 1) && !(w & 1) && (h >> 1) < Ho && (w >> 1) < Wo)
            v = dy[(((int64_t)b * Ho + (h >> 1)) * Wo + (w >> 1)) * C8 + c];
        u[i] = v;
    }
}

// ---- out = act > 0 ? (g1 + g2) : 0 ; g2 may be null; out may alias g1 -------------------------------
__global__ void relu_mask_kernel(const uint4* __restrict__ act, const uint4* g1, const uint4* __restrict__ g2, uint4* out,
                                 int64_t n8)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
        const uint4 a = act[i], x = g1[i];
        uint4 y = make_uint4(0, 0, 0, 0);
        if (g2 != nullptr) y = g2[i];
        const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, xw[4] = {x.x, x.y, x.z, x.w}, yw[4] = {y.x, y.y, y.z, y.w};
        uint32_t ow[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float2 af = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&aw[q]));
            float2 xf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&xw[q]));
            const float2 yf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&yw[q]));
            xf.x = af.x > 0.f ? xf.x + yf.x : 0.f;
            xf.y = af.y > 0.f ? xf.y + yf.y : 0.f;
            const __nv_bfloat162 h = __floats2bfloat162_rn(xf.x, xf.y);
            ow[q] = *reinterpret_cast<const uint32_t*>(&h);
        }
        out[i] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
    }
}

// ---- out[n] += sum_m in[m, n]  (bias gradients; also sums over the batch with N = everything else) -----
template <typename T>
__global__ void colsum_kernel(const T* __restrict__ in, int64_t ld, float* __restrict__ out, int64_t M, int N, int rows_per_block)
{
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const int64_t m0 = (int64_t)blockIdx.y * rows_per_block;
    const int64_t m1 = m0 + rows_per_block < M ? m0 + rows_per_block : M;
    float s = 0.f;
    for (int64_t m = m0; m < m1; ++m) s += to_f32<T>(in[m * ld + n]);
    atomicAdd(out + n, s);
}

// ---- LayerNorm backward over D = 256 ------------------------------------------------------------
// g = g1 + g2 + g3 (gradients of the up-to-three forward outputs y, ypos, y32);  xh = (x - mean) * rstd
// dx = rstd * (g*gamma - mean(g*gamma) - xh * mean(g*gamma*xh)) (+ dres);  dgamma += sum g*xh;  dbeta += sum g
constexpr int D = 256;

__device__ __forceinline__ void load8f(const float* p, float (&v)[8]) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void load8b(const bf16* p, float (&v)[8]) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[q]));
        v[2 * q] = f.x; v[2 * q + 1] = f.y;
    }
}

__global__ void __launch_bounds__(256)
layernorm_bwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const bf16* __restrict__ g1,
                     const bf16* __restrict__ g2, const float* __restrict__ g3, const float* __restrict__ dres,
                     float* __restrict__ dx, float* __restrict__ dgamma, float* __restrict__ dbeta, int64_t rows)
{
    __shared__ float red[2][8][D];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float gm[8];
    load8f(gamma + lane * 8, gm);
    float ag[8], ab[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { ag[i] = 0.f; ab[i] = 0.f; }
    for (int64_t row = (int64_t)blockIdx.x * 8 + warp; row < rows; row += (int64_t)gridDim.x * 8) {
        float v[8], g[8], t[8];
        load8f(x + row * D + lane * 8, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) g[i] = 0.f;
        if (g1 != nullptr) { load8b(g1 + row * D + lane * 8, t);
#pragma unroll
            for (int i = 0; i < 8; ++i) g[i] += t[i]; }
        if (g2 != nullptr) { load8b(g2 + row * D + lane * 8, t);
#pragma unroll
            for (int i = 0; i < 8; ++i) g[i] += t[i]; }
        if (g3 != nullptr) { load8f(g3 + row * D + lane * 8, t);
#pragma unroll
            for (int i = 0; i < 8; ++i) g[i] += t[i]; }
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) s += v[i];
        const float mean = warp_sum(s) * (1.f / D);
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) { v[i] -= mean; q = fmaf(v[i], v[i], q); }
        const float rstd = rsqrtf(warp_sum(q) * (1.f / D) + 1e-5f);
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            v[i] *= rstd;                                   // xh
            ag[i] = fmaf(g[i], v[i], ag[i]);
            ab[i] += g[i];
            g[i] *= gm[i];                                  // g * gamma
            s1 += g[i];
            s2 = fmaf(g[i], v[i], s2);
        }
        s1 = warp_sum(s1) * (1.f / D);
        s2 = warp_sum(s2) * (1.f / D);
        float o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = rstd * (g[i] - s1 - v[i] * s2);
        if (dres != nullptr) {
            load8f(dres + row * D + lane * 8, t);
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] += t[i];
        }
        *reinterpret_cast<float4*>(dx + row * D + lane * 8) = make_float4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<float4*>(dx + row * D + lane * 8 + 4) = make_float4(o[4], o[5], o[6], o[7]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) { red[0][warp][lane * 8 + i] = ag[i]; red[1][warp][lane * 8 + i] = ab[i]; }
    __syncthreads();
    const int c = threadIdx.x;
    float sg = 0.f, sb = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) { sg += red[0][w][c]; sb += red[1][w][c]; }
    if (dgamma != nullptr) atomicAdd(dgamma + c, sg);
    if (dbeta != nullptr) atomicAdd(dbeta + c, sb);
}

// ---- attention core backward: one CTA of 128 threads per (clip, head), S <= 128, head_dim 32 -------------
// phase 1 (thread = query row i): p_ij = softmax_j(scale * q_i.k_j + amask_ij) over the valid keys, dP_ij = dO_i.v_j,
//   dS_ij = p_ij (dP_ij - sum_j p_ij dP_ij), dQ_i = scale * sum_j dS_ij k_j;  P and dS go to shared memory.
// phase 2 (thread = key j): dV_j = sum_i p_ij dO_i,  dK_j = scale * sum_i dS_ij q_i.
constexpr int HD = 32;
constexpr int AB_LD = HD + 1;          // padded row of the fp32 Q/K/V/dO tiles
constexpr int AB_LDP = 129;            // padded row of P / dS
constexpr int AB_SMEM = (4 * 128 * AB_LD + 2 * 128 * AB_LDP + 128) * 4;

__global__ void __launch_bounds__(128)
attention_bwd_kernel(const bf16* __restrict__ Q, int ldq, const bf16* __restrict__ K, int ldk, const bf16* __restrict__ V, int ldv,
                     const bf16* __restrict__ dO, int ldo, bf16* __restrict__ dQ, int lddq, bf16* __restrict__ dK, int lddk,
                     bf16* __restrict__ dV, int lddv, const uint8_t* __restrict__ kpm, const float* __restrict__ amask, int Lq,
                     int Lk, float scale)
{
    extern __shared__ __align__(16) float sm[];
    float* Qs = sm;
    float* Ks = Qs + 128 * AB_LD;
    float* Vs = Ks + 128 * AB_LD;
    float* Gs = Vs + 128 * AB_LD;              // dO
    float* Ps = Gs + 128 * AB_LD;
    float* Ss = Ps + 128 * AB_LDP;             // dS
    float* valid = Ss + 128 * AB_LDP;          // 1 = real key
    const int t = threadIdx.x, h = blockIdx.x, b = blockIdx.y;

    {
        float v[8];
        for (int c = 0; c < 4; ++c) {
            if (t < Lq) {
                load8b(Q + ((size_t)b * Lq + t) * ldq + h * HD + c * 8, v);
#pragma unroll
                for (int i = 0; i < 8; ++i) Qs[t * AB_LD + c * 8 + i] = v[i];
                load8b(dO + ((size_t)b * Lq + t) * ldo + h * HD + c * 8, v);
#pragma unroll
                for (int i = 0; i < 8; ++i) Gs[t * AB_LD + c * 8 + i] = v[i];
            }
            if (t < Lk) {
                load8b(K + ((size_t)b * Lk + t) * ldk + h * HD + c * 8, v);
#pragma unroll
                for (int i = 0; i < 8; ++i) Ks[t * AB_LD + c * 8 + i] = v[i];
                load8b(V + ((size_t)b * Lk + t) * ldv + h * HD + c * 8, v);
#pragma unroll
                for (int i = 0; i < 8; ++i) Vs[t * AB_LD + c * 8 + i] = v[i];
            }
        }
        valid[t] = (t < Lk && !(kpm != nullptr && kpm[(size_t)b * Lk + t])) ? 1.f : 0.f;
    }
    __syncthreads();

    if (t < Lq) {
        float q[HD], g[HD], dq[HD];
#pragma unroll
        for (int d = 0; d < HD; ++d) { q[d] = Qs[t * AB_LD + d]; g[d] = Gs[t * AB_LD + d]; dq[d] = 0.f; }
        const float* arow = amask != nullptr ? amask + (size_t)t * Lk : nullptr;
        float m = -CUDART_INF_F;
        for (int j = 0; j < Lk; ++j) {
            float s = 0.f;
#pragma unroll
            for (int d = 0; d < HD; ++d) s = fmaf(q[d], Ks[j * AB_LD + d], s);
            s *= scale;
            if (arow != nullptr) s += fmaxf(arow[j], -1e30f);
            Ps[t * AB_LDP + j] = s;
            if (valid[j] != 0.f) m = fmaxf(m, s);
        }
        float l = 0.f;
        for (int j = 0; j < Lk; ++j) {
            const float e = valid[j] != 0.f ? __expf(Ps[t * AB_LDP + j] - m) : 0.f;
            Ps[t * AB_LDP + j] = e;
            l += e;
        }
        const float inv = l > 0.f ? 1.f / l : 0.f;
        float delta = 0.f;
        for (int j = 0; j < Lk; ++j) {
            float dp = 0.f;
#pragma unroll
            for (int d = 0; d < HD; ++d) dp = fmaf(g[d], Vs[j * AB_LD + d], dp);
            const float p = Ps[t * AB_LDP + j] * inv;
            Ps[t * AB_LDP + j] = p;
            Ss[t * AB_LDP + j] = dp;
            delta = fmaf(p, dp, delta);
        }
        for (int j = 0; j < Lk; ++j) {
            const float ds = Ps[t * AB_LDP + j] * (Ss[t * AB_LDP + j] - delta);
            Ss[t * AB_LDP + j] = ds;
#pragma unroll
            for (int d = 0; d < HD; ++d) dq[d] = fmaf(ds, Ks[j * AB_LD + d], dq[d]);
        }
        bf16* o = dQ + ((size_t)b * Lq + t) * lddq + h * HD;
#pragma unroll
        for (int d = 0; d < HD; d += 2)
            *reinterpret_cast<__nv_bfloat162*>(o + d) = __floats2bfloat162_rn(dq[d] * scale, dq[d + 1] * scale);
    }
    __syncthreads();

    if (t < Lk) {
        float dk[HD], dv[HD];
#pragma unroll
        for (int d = 0; d < HD; ++d) { dk[d] = 0.f; dv[d] = 0.f; }
        for (int i = 0; i < Lq; ++i) {
            const float p = Ps[i * AB_LDP + t], ds = Ss[i * AB_LDP + t];
#pragma unroll
            for (int d = 0; d < HD; ++d) {
                dv[d] = fmaf(p, Gs[i * AB_LD + d], dv[d]);
                dk[d] = fmaf(ds, Qs[i * AB_LD + d], dk[d]);
            }
        }
        bf16* ok = dK + ((size_t)b * Lk + t) * lddk + h * HD;
        bf16* ov = dV + ((size_t)b * Lk + t) * lddv + h * HD;
#pragma unroll
        for (int d = 0; d < HD; d += 2) {
            *reinterpret_cast<__nv_bfloat162*>(ok + d) = __floats2bfloat162_rn(dk[d] * scale, dk[d + 1] * scale);
            *reinterpret_cast<__nv_bfloat162*>(ov + d) = __floats2bfloat162_rn(dv[d], dv[d + 1]);
        }
    }
}

// ---- stem backward: gradients of conv0 (1 -> 3 channels, 1x1, bias), the only trainable parameters below layer2 ---------
// Forward (sedt/backbone.py:102 + resnet conv1/bn1/relu/maxpool): z[o] = sum_taps Weff[o][tap] x[tap] + sum_{taps inside
// the image} Beff[o][tap], Weff = sum_c conv1[o][c][tap] w0[c], Beff = sum_c conv1[o][c][tap] b0[c]; a = z * bn_scale + bn_bias;
// out = maxpool3x3s2(relu(a)).  Given G = d/d(out) and the window arg-max the training forward recorded (stem_tc_kernel<true>:
// 0..8 = position, 9 = dead), the kernel routes G to that conv pixel and accumulates dWeff[o][tap] += dz * x[tap],
// dBeff[o][tap] += dz (dz = G * bn_scale) in registers; thread = output channel o, CTA = (clip, SB_PH pooled rows).
constexpr int SB_PH = 4;            // pooled rows per CTA
constexpr int SB_XC = 64 + 10;      // padded input row

__global__ void __launch_bounds__(64)
stem_bwd_kernel(const float* __restrict__ x, const float* __restrict__ bn_scale, const bf16* __restrict__ G,
                const uint8_t* __restrict__ amax, float* __restrict__ acc, int T, int Hc, int Hp)
{
    __shared__ float xs[11][SB_XC];
    __shared__ int rowin[11];
    const int o = threadIdx.x, b = blockIdx.y, hp0 = blockIdx.x * SB_PH;
    const float sc = bn_scale[o];
    float dW[49], dB[49];
#pragma unroll
    for (int i = 0; i < 49; ++i) { dW[i] = 0.f; dB[i] = 0.f; }
    const float* xb = x + (size_t)b * T * 64;
    for (int hp = hp0; hp < hp0 + SB_PH && hp < Hp; ++hp) {
        const int irow0 = 4 * hp - 5;
        __syncthreads();
        for (int i = o; i < 11 * SB_XC; i += 64) {
            const int lr = i / SB_XC, lc = i - lr * SB_XC;
            const int row = irow0 + lr, col = lc - 5;
            xs[lr][lc] = (row >= 0 && row < T && col >= 0 && col < 64) ? xb[(size_t)row * 64 + col] : 0.f;
        }
        if (o < 11) rowin[o] = (irow0 + o >= 0 && irow0 + o < T) ? 1 : 0;
        __syncthreads();
        for (int wp = 0; wp < 16; ++wp) {
            const size_t e = (((size_t)b * Hp + hp) * 16 + wp) * 64 + o;
            const float g = __bfloat162float(G[e]);
            const int a = amax[e];
            if (a < 9 && g != 0.f) {
                const int bdy = a / 3, bdx = a - 3 * bdy;
                const float dz = g * sc;
                const int wc = 2 * wp - 1 + bdx;
#pragma unroll
                for (int r = 0; r < 7; ++r) {
                    if (!rowin[2 * bdy + r]) continue;
                    const float* xr = &xs[2 * bdy + r][2 * wc - 3 + 5];
#pragma unroll
                    for (int q = 0; q < 7; ++q) {
                        const int ic = 2 * wc - 3 + q;
                        if (ic >= 0 && ic < 64) { dW[r * 7 + q] = fmaf(dz, xr[q], dW[r * 7 + q]); dB[r * 7 + q] += dz; }
                    }
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 49; ++i) {
        if (dW[i] != 0.f) atomicAdd(acc + i * 64 + o, dW[i]);
        if (dB[i] != 0.f) atomicAdd(acc + 49 * 64 + i * 64 + o, dB[i]);
    }
}

// d(conv0.weight)[c] = sum_{o,tap} dWeff[o][tap] conv1[o][c][tap];  d(conv0.bias)[c] likewise with dBeff
__global__ void __launch_bounds__(256)
stem_bwd_finish_kernel(const float* __restrict__ acc, const float* __restrict__ conv1_w, float* __restrict__ g_w, float* __restrict__ g_b)
{
    __shared__ float red[6][256];
    float s[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int i = threadIdx.x; i < 49 * 64; i += 256) {
        const int tap = i / 64, o = i % 64;
        const float dw = acc[i], db = acc[49 * 64 + i];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float k = conv1_w[(o * 3 + c) * 49 + tap];
            s[c] = fmaf(dw, k, s[c]); s[3 + c] = fmaf(db, k, s[3 + c]);
        }
    }
#pragma unroll
    for (int c = 0; c < 6; ++c) red[c][threadIdx.x] = s[c];
    __syncthreads();
    if (threadIdx.x < 6) {
        float t = 0.f;
        for (int i = 0; i < 256; ++i) t += red[threadIdx.x][i];
        if (threadIdx.x < 3) g_w[threadIdx.x] = t; else g_b[threadIdx.x - 3] = t;
    }
}

static inline unsigned grid_for(int64_t n, int per_block = 256) { return (unsigned)std::min<int64_t>(ceil_div(n, per_block), 148 * 16); }

}  // namespace

int launch_repack_dgrad(const float* w_oihw, const float* scale, void* out, int dt, int Cout, int Cout_pad, int Cin, int R, int S,
                        cudaStream_t stream)
{
    SEDT_REQUIRE(Cout_pad >= Cout, "repack_dgrad: Cout_pad=%d < Cout=%d", Cout_pad, Cout);
    const int64_t total = (int64_t)Cout_pad * Cin * R * S;
    if (total == 0) return SEDT_OK;
    if (R == S) {          // rotating both axes of a square filter reverses the flattened tap index
        dim3 grid((unsigned)ceil_div(Cin, 32), (unsigned)ceil_div(Cout_pad, 32), (unsigned)(R * S)), block(32, 8);
        if (dt == DT_F32) repack_dgrad_tiled_kernel<float><<<grid, block, 0, stream>>>(w_oihw, scale, (float*)out, Cout, Cout_pad, Cin, R * S);
        else repack_dgrad_tiled_kernel<bf16><<<grid, block, 0, stream>>>(w_oihw, scale, (bf16*)out, Cout, Cout_pad, Cin, R * S);
    } else if (dt == DT_F32) repack_dgrad_kernel<float><<<grid_for(total), 256, 0, stream>>>(w_oihw, scale, (float*)out, Cout, Cout_pad, Cin, R, S);
    else repack_dgrad_kernel<bf16><<<grid_for(total), 256, 0, stream>>>(w_oihw, scale, (bf16*)out, Cout, Cout_pad, Cin, R, S);
    SEDT_COUNT_LAUNCH();
    SEDT_CHECK_CUDA(cudaGetLastError());
    return SEDT_OK;
}

int launch_repack_dgrad_batched(const DgradJob* jobs, int n, cudaStream_t stream)
{
    for (int j0 = 0; j0 < n; j0 += kDgradMaxJobs) {
        DgradJobs P;
        P.n = std::min(kDgradMaxJobs, n - j0);
        int tiles = 0;
        for (int j = 0; j < P.n; ++j) {
            P.j[j] = jobs[j0 + j];
            P.j[j].tile0 = tiles;
            tiles += (int)(ceil_div(P.j[j].Cin, 32) * ceil_div(P.j[j].Cout_pad, 32) * P.j[j].RS);
        }
        if (tiles == 0) continue;
        repack_dgrad_batched_kernel<<<(unsigned)tiles, 256, 0, stream>>>(P);
        SEDT_COUNT_LAUNCH();
    }
    SEDT_CHECK_CUDA(cudaGetLastError());
    return SEDT_OK;
}

int launch_unpack_wgrad(const float* dw, const float* scale, float* grad, int Cout, int Cin, int RS, cudaStream_t stream)
{
    const int64_t total = (int64_t)Cout * Cin * RS;
    if (total == 0) return SEDT_OK;
    unpack_wgrad_kernel<<<grid_for(total), 256, 0, stream>>>(dw, scale, grad, Cout, Cin, RS);
    SEDT_COUNT_LAUNCH();
    SEDT_CHECK_CUDA(cudaGetLastError());
    return SEDT_OK;
}

int launch_heads_bwd_prepare(const float* d_logits, const float* d_boxes, const float* d_at, const float* boxes, const float* at,
                             void* dcls, void* dbox, void* dweak, int D_, int B, int Qall, int start, int C1, int C,
                             cudaStream_t stream)
{
    const int64_t n = (int64_t)D_ * B * Qall * 256 + (dweak != nullptr ? (int64_t)B * 128 : 0);
    if (n == 0) return SEDT_OK;
    heads_bwd_prepare_kernel<<<grid_for(n), 256, 0, stream>>>(d_logits, d_boxes, d_at, boxes, at, (bf16*)dcls, (bf16*)dbox,
                                                             (bf16*)dweak, D_, B, Qall, start, C1, C);
    SEDT_COUNT_LAUNCH();
    SEDT_CHECK_CUDA(cudaGetLastError());
    return SEDT_OK;
}

int launch_upsample2(const void* dy, void* u, int B, int H, int W, int Ho, int Wo, int C, cudaStream_t stream)
{
    SEDT_REQUIRE(C % 8 == 0, "upsample2: C=%d must be a multiple of 8", C);
    const int64_t total = (int64_t)B * H * W * (C / 8);
    if (total == 0) return SEDT_OK;
    upsample2_kernel<<<grid_for(total), 256, 0, stream>>>((const uint4*)dy, (uint4*)u, B, H, W, Ho, Wo, C / 8);
    SEDT_COUNT_LAUNCH();
    SEDT_CHECK_CUDA(cudaGetLastError());
    return SEDT_OK;
}

int launch_relu_mask(const void* act, const void* g1, const void* g2, void* out, int64_t n, cudaStream_t stream)
{
    SEDT_REQUIRE(n % 8 == 0, "relu_mask: n must be a multiple of 8");
    if (n == 0) return SEDT_OK;
    relu_mask_kernel<<<grid_for(n / 8), 256, 0, stream>>>((const uint4*)act, (const uint4*)g1, (const uint4*)g2, (uint4*)out, n / 8);
    SEDT_COUNT_LAUNCH();
    SEDT_CHECK_CUDA(cudaGetLastError());
    return SEDT_OK;
}

int launch_colsum(const void* in, int dt, int64_t ld, float* out, int64_t M, int N, cudaStream_t stream)
{
    if (M == 0 || N == 0) return SEDT_OK;
    const int rows_per_block = (int)std::max<int64_t>(16, ceil_div(M, 512));
    dim3 grid((unsigned)ceil_div(N, 128), (unsigned)ceil_div(M, rows_per_block)), block(128);
    if (dt == DT_F32) colsum_kernel<float><<<grid, block, 0, stream>>>((const float*)in, ld, out, M, N, rows_per_block);
    else colsum_kernel<bf16><<<grid, block, 0, stream>>>((const bf16*)in, ld, out, M, N, rows_per_block);
    SEDT_COUNT_LAUNCH();
    SEDT_CHECK_CUDA(cudaGetLastError());
    return SEDT_OK;
}

int launch_layernorm_bwd(const float* x, const float* gamma, const void* g1, const void* g2, const float* g3, const float* dres,
                         float* dx, float* dgamma, float* dbeta, int64_t rows, cudaStream_t stream)
{
    if (rows == 0) return SEDT_OK;
    const unsigned grid = (unsigned)std::min<int64_t>(ceil_div(rows, 8), 148 * 4);
    ProfScope _prof(PROF_NORM, stream);
    layernorm_bwd_kernel<<<grid, 256, 0, stream>>>(x, gamma, (const bf16*)g1, (const bf16*)g2, g3, dres, dx, dgamma, dbeta, rows);
    SEDT_COUNT_LAUNCH();
    SEDT_CHECK_CUDA(cudaGetLastError());
    return SEDT_OK;
}

int launch_stem_bwd(const float* x, const float* conv1_w, const float* bn_scale, const void* G, const uint8_t* amax,
                    float* scratch, float* g_conv0_w, float* g_conv0_b, int B, int T, int F, cudaStream_t stream)
{
    SEDT_REQUIRE(F == 64, "stem_bwd: F=%d must be 64", F);
    SEDT_REQUIRE(amax != nullptr, "stem_bwd: needs the pooling arg-max recorded by the training forward");
    if (B == 0) return SEDT_OK;
    const int Hc = (T + 2 * 3 - 7) / 2 + 1, Hp = (Hc + 2 - 3) / 2 + 1;
    SEDT_TRY(launch_fill_zero(scratch, (size_t)2 * 49 * 64 * 4, stream));
    stem_bwd_kernel<<<dim3((unsigned)ceil_div(Hp, SB_PH), (unsigned)B), 64, 0, stream>>>(x, bn_scale, (const bf16*)G, amax, scratch, T, Hc,
                                                                                      Hp);
    SEDT_COUNT_LAUNCH();
    stem_bwd_finish_kernel<<<1, 256, 0, stream>>>(scratch, conv1_w, g_conv0_w, g_conv0_b);
    SEDT_COUNT_LAUNCH();
    SEDT_CHECK_CUDA(cudaGetLastError());
    return SEDT_OK;
}

int launch_attention_bwd(const void* Q, int ldq, const void* K, int ldk, const void* V, int ldv, const void* dO, int ldo,
                         void* dQ, int lddq, void* dK, int lddk, void* dV, int lddv, const uint8_t* kpm, const float* amask,
                         int B, int nheads, int Lq, int Lk, float scale, cudaStream_t stream)
{
    if (B == 0) return SEDT_OK;
    SEDT_REQUIRE(Lq >= 1 && Lk >= 1 && Lq <= 128 && Lk <= 128, "attention_bwd: Lq=%d Lk=%d (at most 128)", Lq, Lk);
    static bool attr_set = false;
    if (!attr_set) {
        SEDT_CHECK_CUDA(cudaFuncSetAttribute(attention_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AB_SMEM));
        attr_set = true;
    }
    ProfScope _prof(PROF_ATTENTION, stream);
    attention_bwd_kernel<<<dim3((unsigned)nheads, (unsigned)B), 128, AB_SMEM, stream>>>(
        (const bf16*)Q, ldq, (const bf16*)K, ldk, (const bf16*)V, ldv, (const bf16*)dO, ldo, (bf16*)dQ, lddq, (bf16*)dK, lddk,
        (bf16*)dV, lddv, kpm, amask, Lq, Lk, scale);
    SEDT_COUNT_LAUNCH();
    SEDT_CHECK_CUDA(cudaGetLastError());
    return SEDT_OK;
}

}  // namespace sedt
