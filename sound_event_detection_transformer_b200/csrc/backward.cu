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
template <typename T>
__global__ void repack_dgrad_kernel(const float* __restrict__ w, const float* __restrict__ scale, T* __restrict__ out, int Cout,
                                    int Cin, int R, int S)
{
    const int64_t total = (int64_t)Cout * Cin * R * S;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int co = (int)(i % Cout);
        int64_t t = i / Cout;
        const int s = (int)(t % S); t /= S;
        const int r = (int)(t % R);
        const int ci = (int)(t / R);
        float v = w[(((int64_t)co * Cin + ci) * R + (R - 1 - r)) * S + (S - 1 - s)];
        if (scale != nullptr) v *= scale[co];
        out[i] = from_f32<T>(v);
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

static inline unsigned grid_for(int64_t n, int per_block = 256) { return (unsigned)std::min<int64_t>(ceil_div(n, per_block), 148 * 16); }

}  // namespace

int launch_repack_dgrad(const float* w_oihw, const float* scale, void* out, int dt, int Cout, int Cin, int R, int S,
                        cudaStream_t stream)
{
    const int64_t total = (int64_t)Cout * Cin * R * S;
    if (total == 0) return SEDT_OK;
    if (dt == DT_F32) repack_dgrad_kernel<float><<<grid_for(total), 256, 0, stream>>>(w_oihw, scale, (float*)out, Cout, Cin, R, S);
    else repack_dgrad_kernel<bf16><<<grid_for(total), 256, 0, stream>>>(w_oihw, scale, (bf16*)out, Cout, Cin, R, S);
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
