// One-time weight packing: FrozenBatchNorm fold, OIHW -> O(HW)I repack, casts.
// FrozenBatchNorm2d.forward (sedt/backbone.py:43-53):
//   scale = weight * rsqrt(running_var + 1e-5); bias = bias - running_mean * scale
#include "kernels.h"

namespace sedt {
namespace {

__global__ void bn_fold_kernel(const float* __restrict__ w, const float* __restrict__ b, const float* __restrict__ mean,
                               const float* __restrict__ var, float* __restrict__ scale, float* __restrict__ bias, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float s = w[i] * rsqrtf(var[i] + 1e-5f);
    scale[i] = s;
    bias[i] = b[i] - mean[i] * s;
}

template <typename T>
__global__ void repack_conv_kernel(const float* __restrict__ w, const float* __restrict__ scale, T* __restrict__ out, int Cout,
                                   int Cin, int RS)
{
    // out[o][tap][c] = w[o][c][tap]; consecutive threads -> consecutive c (coalesced writes)
    const int64_t total = (int64_t)Cout * Cin * RS;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % Cin);
        const int tap = (int)((i / Cin) % RS);
        const int o = (int)(i / ((int64_t)Cin * RS));
        const float v = w[((int64_t)o * Cin + c) * RS + tap];
        out[i] = from_f32<T>(scale != nullptr ? v * scale[o] : v);
    }
}

template <typename T>
__global__ void cast_kernel(const float* __restrict__ in, T* __restrict__ out, int64_t n)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = from_f32<T>(in[i]);
}

__global__ void fill_zero_kernel(uint4* __restrict__ p, size_t n16, unsigned char* tail, size_t ntail)
{
    const uint4 z = make_uint4(0, 0, 0, 0);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) p[i] = z;
    if (blockIdx.x == 0 && threadIdx.x < ntail) tail[threadIdx.x] = 0;
}

static inline unsigned grid_for(int64_t n, int block = 256)
{
    int64_t g = ceil_div(n, block);
    return (unsigned)(g < 1 ? 1 : (g > 148 * 16 ? 148 * 16 : g));
}

}  // namespace

int launch_bn_fold(const float* w, const float* b, const float* mean, const float* var, float* scale, float* bias,
                   int n, cudaStream_t stream)
{
    bn_fold_kernel<<<(unsigned)ceil_div(n, 128), 128, 0, stream>>>(w, b, mean, var, scale, bias, n);
    SEDT_COUNT_LAUNCH();
    SEDT_CHECK_CUDA(cudaGetLastError());
    return SEDT_OK;
}

int launch_repack_conv(const float* w_oihw, const float* scale, void* out, int dt, int Cout, int Cin, int R, int S,
                       cudaStream_t stream)
{
    const int64_t total = (int64_t)Cout * Cin * R * S;
    if (dt == DT_F32) repack_conv_kernel<float><<<grid_for(total), 256, 0, stream>>>(w_oihw, scale, (float*)out, Cout, Cin, R * S);
    else repack_conv_kernel<__nv_bfloat16><<<grid_for(total), 256, 0, stream>>>(w_oihw, scale, (__nv_bfloat16*)out, Cout, Cin, R * S);
    SEDT_COUNT_LAUNCH();
    SEDT_CHECK_CUDA(cudaGetLastError());
    return SEDT_OK;
}

int launch_cast(const float* in, void* out, int dt, int64_t n, cudaStream_t stream)
{
    if (n == 0) return SEDT_OK;
    if (dt == DT_F32) cast_kernel<float><<<grid_for(n), 256, 0, stream>>>(in, (float*)out, n);
    else cast_kernel<__nv_bfloat16><<<grid_for(n), 256, 0, stream>>>(in, (__nv_bfloat16*)out, n);
    SEDT_COUNT_LAUNCH();
    SEDT_CHECK_CUDA(cudaGetLastError());
    return SEDT_OK;
}

int launch_fill_zero(void* p, size_t bytes, cudaStream_t stream)
{
    if (bytes == 0) return SEDT_OK;
    SEDT_REQUIRE(((uintptr_t)p & 15) == 0, "fill_zero: pointer must be 16-byte aligned");
    const size_t n16 = bytes / 16, ntail = bytes % 16;
    ProfScope _prof(PROF_OTHER, stream);
    fill_zero_kernel<<<grid_for((int64_t)(n16 ? n16 : 1)), 256, 0, stream>>>((uint4*)p, n16, (unsigned char*)p + n16 * 16, ntail);
    SEDT_COUNT_LAUNCH();
    SEDT_CHECK_CUDA(cudaGetLastError());
    return SEDT_OK;
}

}  // namespace sedt
