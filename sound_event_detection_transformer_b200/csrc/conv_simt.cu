// CUDA-core implicit-GEMM convolution / linear with fp32 accumulation.
//
// This is the "precise" tier of the hot path (fp32 operands, fp32 FMA: the tier
// held to the 1e-4 parity bar against the reference's fp32 PyTorch path) and
// the fallback for shapes too small for a 128-row tcgen05 tile (the class / box
// / audio-tag heads, sedt/sedt.py:89-95).  Epilogue fuses FrozenBatchNorm
// (sedt/backbone.py:43-53 folded to scale/bias), bias, residual add and ReLU.
#include "kernels.h"

namespace sedt {
namespace {

constexpr int BM = 64, BN = 64, BK = 16, NT = 256;

template <typename T> struct Vec4;
template <> struct Vec4<float> {
    static __device__ __forceinline__ void load(const float* p, float (&v)[4]) {
        float4 t = *reinterpret_cast<const float4*>(p);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
};
template <> struct Vec4<__nv_bfloat16> {
    static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[4]) {
        uint2 t = *reinterpret_cast<const uint2*>(p);
        __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&t.x);
        __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&t.y);
        v[0] = __low2float(a); v[1] = __high2float(a); v[2] = __low2float(b); v[3] = __high2float(b);
    }
};

struct Geo {
    int B, H, W, Cin, lda, Ho, Wo, Cout, ldc, ld_res, R, S, stride, dil, pad, relu;
};

template <typename TA, typename TO>
__global__ void __launch_bounds__(NT)
conv_simt_kernel(const TA* __restrict__ in, const TA* __restrict__ w, const float* __restrict__ scale,
                 const float* __restrict__ bias, const TO* __restrict__ residual, TO* __restrict__ out, Geo g)
{
    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Bs[BK][BN + 4];
    pdl_trigger();
    pdl_wait();
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int64_t M = (int64_t)g.B * g.Ho * g.Wo;
    const int64_t m0 = (int64_t)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const int K = g.R * g.S * g.Cin;

    // loader mapping: one 4-wide k slice of one row per thread
    const int lr = tid >> 2, kq = (tid & 3) * 4;
    const int64_t m = m0 + lr;
    const bool mvalid = m < M;
    int b = 0, ho = 0, wo = 0;
    if (mvalid) {
        wo = (int)(m % g.Wo);
        int64_t t = m / g.Wo;
        ho = (int)(t % g.Ho);
        b = (int)(t / g.Ho);
    }
    const int n_ld = n0 + lr;
    const bool nvalid = n_ld < g.Cout;
    const TA* wrow = w + (size_t)(nvalid ? n_ld : 0) * K;

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int tap = 0; tap < g.R * g.S; ++tap) {
        const int r = tap / g.S, s = tap % g.S;
        const int hi = ho * g.stride + r * g.dil - g.pad;
        const int wi = wo * g.stride + s * g.dil - g.pad;
        const bool avalid = mvalid && hi >= 0 && hi < g.H && wi >= 0 && wi < g.W;
        const TA* arow = in + ((size_t)((size_t)b * g.H + (avalid ? hi : 0)) * g.W + (avalid ? wi : 0)) * g.lda;
        for (int c0 = 0; c0 < g.Cin; c0 += BK) {
            float av[4] = {0.f, 0.f, 0.f, 0.f}, bv[4] = {0.f, 0.f, 0.f, 0.f};
            if (avalid) Vec4<TA>::load(arow + c0 + kq, av);
            if (nvalid) Vec4<TA>::load(wrow + (size_t)tap * g.Cin + c0 + kq, bv);
#pragma unroll
            for (int i = 0; i < 4; ++i) { As[kq + i][lr] = av[i]; Bs[kq + i][lr] = bv[i]; }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < BK; ++k) {
                const float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
                const float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
                const float a[4] = {a4.x, a4.y, a4.z, a4.w};
                const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
            }
            __syncthreads();
        }
    }

#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t mo = m0 + ty * 4 + i;
        if (mo >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= g.Cout) continue;
            float v = acc[i][j];
            if (scale != nullptr) v *= scale[n];
            if (bias != nullptr) v += bias[n];
            if (residual != nullptr) {
                const float rv = to_f32<TO>(residual[(size_t)mo * g.ld_res + n]);
                if (g.relu == 2) { if (!(rv > 0.f)) v = 0.f; }        // residual is a ReLU mask (data-gradient GEMMs)
                else v += rv;
            }
            if (g.relu == 1) v = fmaxf(v, 0.f);
            out[(size_t)mo * g.ldc + n] = from_f32<TO>(v);
        }
    }
}

}  // namespace

int launch_conv_simt(const ConvGemm& c, cudaStream_t stream)
{
    SEDT_REQUIRE(c.Cin % BK == 0, "conv_simt: Cin=%d must be a multiple of %d", c.Cin, BK);
    SEDT_REQUIRE(c.lda % 4 == 0, "conv_simt: lda=%d must be a multiple of 4", c.lda);
    const int64_t M = (int64_t)c.B * c.Ho * c.Wo;
    if (M == 0 || c.Cout == 0) return SEDT_OK;
    Geo g{c.B, c.H, c.W, c.Cin, c.lda, c.Ho, c.Wo, c.Cout, c.ldc, c.ld_res, c.R, c.S, c.stride, c.dil, c.pad, c.relu};
    dim3 grid((unsigned)ceil_div(M, BM), (unsigned)ceil_div(c.Cout, BN)), block(NT);
    ProfScope _prof(PROF_GEMM_SIMT, stream);
    if (c.in_dt == DT_F32 && c.out_dt == DT_F32) {
        SEDT_CHECK_CUDA(launch_pdl(conv_simt_kernel<float, float>, grid, block, 0, stream, 1,
                                   (const float*)c.in, (const float*)c.w, c.scale, c.bias, (const float*)c.residual, (float*)c.out, g));
    } else if (c.in_dt == DT_BF16 && c.out_dt == DT_BF16) {
        SEDT_CHECK_CUDA(launch_pdl(conv_simt_kernel<__nv_bfloat16, __nv_bfloat16>, grid, block, 0, stream, 1,
                                   (const __nv_bfloat16*)c.in, (const __nv_bfloat16*)c.w, c.scale, c.bias,
                                   (const __nv_bfloat16*)c.residual, (__nv_bfloat16*)c.out, g));
    } else if (c.in_dt == DT_BF16 && c.out_dt == DT_F32) {
        SEDT_CHECK_CUDA(launch_pdl(conv_simt_kernel<__nv_bfloat16, float>, grid, block, 0, stream, 1,
                                   (const __nv_bfloat16*)c.in, (const __nv_bfloat16*)c.w, c.scale, c.bias,
                                   (const float*)c.residual, (float*)c.out, g));
    } else {
        SEDT_REQUIRE(false, "conv_simt: unsupported dtype combination in=%d out=%d", c.in_dt, c.out_dt);
    }
    SEDT_COUNT_LAUNCH();
    SEDT_CHECK_CUDA(cudaGetLastError());
    return SEDT_OK;
}

}  // namespace sedt
