// Elementwise dropout kernels of the training step (sedt/transformer.py:160-175, 220-240: dropout, dropout1-3):
// applied after the out-proj / linear2 GEMMs (together with the residual add), on the FFN hidden layer, and on the
// gradient that re-enters those GEMMs in the backward pass.  Masks come from dropout.cuh, so forward and
// backward agree by construction.
#include "kernels.h"
#include "dropout.cuh"

namespace sedt {
namespace {

using bf16 = __nv_bfloat16;

__global__ void rng_init_kernel(unsigned long long* state, unsigned long long seed) { state[0] = seed; state[1] = 0ull; }
__global__ void rng_step_kernel(unsigned long long* state) { state[1] += 1ull; }

// out = resid + y * keep / (1 - p)      (fp32, 4 elements per thread)
__global__ void dropout_add_kernel(const float4* __restrict__ y, const float4* __restrict__ resid, float4* __restrict__ out, int64_t n4,
                                   DropSite d)
{
    const unsigned long long seed = d.state[0], step = d.state[1];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const uint4 r = drop_draw4(d, seed, step, (unsigned long long)i);
        const float4 a = y[i], b = resid[i];
        float4 o;
        o.x = b.x + (r.x < d.thresh ? a.x * d.inv_keep : 0.f);
        o.y = b.y + (r.y < d.thresh ? a.y * d.inv_keep : 0.f);
        o.z = b.z + (r.z < d.thresh ? a.z * d.inv_keep : 0.f);
        o.w = b.w + (r.w < d.thresh ? a.w * d.inv_keep : 0.f);
        out[i] = o;
    }
}

// h = h * keep / (1 - p) in place (bf16, 4 elements per thread)
__global__ void dropout_bf16_kernel(uint2* __restrict__ h, int64_t n4, DropSite d)
{
    const unsigned long long seed = d.state[0], step = d.state[1];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const uint4 r = drop_draw4(d, seed, step, (unsigned long long)i);
        uint2 u = h[i];
        float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
        float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
        a.x = r.x < d.thresh ? a.x * d.inv_keep : 0.f; a.y = r.y < d.thresh ? a.y * d.inv_keep : 0.f;
        b.x = r.z < d.thresh ? b.x * d.inv_keep : 0.f; b.y = r.w < d.thresh ? b.y * d.inv_keep : 0.f;
        const __nv_bfloat162 ha = __floats2bfloat162_rn(a.x, a.y), hb = __floats2bfloat162_rn(b.x, b.y);
        u.x = *reinterpret_cast<const uint32_t*>(&ha); u.y = *reinterpret_cast<const uint32_t*>(&hb);
        h[i] = u;
    }
}

// g16 = bf16(g32 * keep / (1 - p)): the gradient entering a GEMM whose forward output was dropped
__global__ void cast_dropout_kernel(const float4* __restrict__ g, uint2* __restrict__ out, int64_t n4, DropSite d)
{
    const unsigned long long seed = d.state[0], step = d.state[1];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const uint4 r = drop_draw4(d, seed, step, (unsigned long long)i);
        const float4 a = g[i];
        const __nv_bfloat162 h0 = __floats2bfloat162_rn(r.x < d.thresh ? a.x * d.inv_keep : 0.f, r.y < d.thresh ? a.y * d.inv_keep : 0.f);
        const __nv_bfloat162 h1 = __floats2bfloat162_rn(r.z < d.thresh ? a.z * d.inv_keep : 0.f, r.w < d.thresh ? a.w * d.inv_keep : 0.f);
        uint2 u;
        u.x = *reinterpret_cast<const uint32_t*>(&h0); u.y = *reinterpret_cast<const uint32_t*>(&h1);
        out[i] = u;
    }
}

__global__ void fill_value_kernel(float* p, float v, int n) { const int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) p[i] = v; }

// keep flags of elements [0, n) of one site at an explicit (seed, step): test hook for the parity tests
__global__ void dropout_mask_kernel(unsigned char* out, int64_t n, DropSite d, unsigned long long seed, unsigned long long step)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i * 4 < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint4 r = drop_draw4(d, seed, step, (unsigned long long)i);
        const uint32_t v[4] = {r.x, r.y, r.z, r.w};
        for (int q = 0; q < 4; ++q)
            if (i * 4 + q < n) out[i * 4 + q] = v[q] < d.thresh ? 1 : 0;
    }
}

inline unsigned grid_for(int64_t n) { return (unsigned)std::min<int64_t>(ceil_div(n, 256), 148 * 16); }

}  // namespace

int launch_fill_value(float* p, float v, int n, cudaStream_t stream)
{
    if (n == 0) return SEDT_OK;
    fill_value_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, stream>>>(p, v, n);
    SEDT_COUNT_LAUNCH();
    SEDT_CHECK_CUDA(cudaGetLastError());
    return SEDT_OK;
}

int launch_dropout_mask(unsigned char* out, int64_t n, unsigned long long seed, unsigned long long step, uint32_t site, float p,
                        cudaStream_t stream)
{
    if (n == 0) return SEDT_OK;
    DropSite d = make_drop_site(nullptr, site, p);
    dropout_mask_kernel<<<grid_for((n + 3) / 4), 256, 0, stream>>>(out, n, d, seed, step);
    SEDT_COUNT_LAUNCH();
    SEDT_CHECK_CUDA(cudaGetLastError());
    return SEDT_OK;
}

int launch_rng_init(unsigned long long* state, unsigned long long seed, cudaStream_t stream)
{
    rng_init_kernel<<<1, 1, 0, stream>>>(state, seed);
    SEDT_COUNT_LAUNCH();
    SEDT_CHECK_CUDA(cudaGetLastError());
    return SEDT_OK;
}

int launch_rng_step(unsigned long long* state, cudaStream_t stream)
{
    rng_step_kernel<<<1, 1, 0, stream>>>(state);
    SEDT_COUNT_LAUNCH();
    SEDT_CHECK_CUDA(cudaGetLastError());
    return SEDT_OK;
}

int launch_dropout_add(const float* y, const float* resid, float* out, int64_t n, const DropSite& d, cudaStream_t stream)
{
    SEDT_REQUIRE(n % 4 == 0 && d.state != nullptr, "dropout_add: n must be a multiple of 4 and the RNG state set");
    if (n == 0) return SEDT_OK;
    dropout_add_kernel<<<grid_for(n / 4), 256, 0, stream>>>((const float4*)y, (const float4*)resid, (float4*)out, n / 4, d);
    SEDT_COUNT_LAUNCH();
    SEDT_CHECK_CUDA(cudaGetLastError());
    return SEDT_OK;
}

int launch_dropout_bf16(void* h, int64_t n, const DropSite& d, cudaStream_t stream)
{
    SEDT_REQUIRE(n % 4 == 0 && d.state != nullptr, "dropout_bf16: n must be a multiple of 4 and the RNG state set");
    if (n == 0) return SEDT_OK;
    dropout_bf16_kernel<<<grid_for(n / 4), 256, 0, stream>>>((uint2*)h, n / 4, d);
    SEDT_COUNT_LAUNCH();
    SEDT_CHECK_CUDA(cudaGetLastError());
    return SEDT_OK;
}

int launch_cast_dropout(const float* g, void* out16, int64_t n, const DropSite& d, cudaStream_t stream)
{
    SEDT_REQUIRE(n % 4 == 0 && d.state != nullptr, "cast_dropout: n must be a multiple of 4 and the RNG state set");
    if (n == 0) return SEDT_OK;
    cast_dropout_kernel<<<grid_for(n / 4), 256, 0, stream>>>((const float4*)g, (uint2*)out16, n / 4, d);
    SEDT_COUNT_LAUNCH();
    SEDT_CHECK_CUDA(cudaGetLastError());
    return SEDT_OK;
}

}  // namespace sedt
