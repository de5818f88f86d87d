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

// ---- batched packing: every cast / repack / BN fold of one sedt_model_pack call in one launch ---------------------
// The training loop re-packs after every optimizer step; as ~260 separate launches that costs ~0.7 ms of launch
// latency per step for ~0.05 ms of memory traffic.  Jobs travel in the kernel parameters (no table in memory, so the
// launch is safe to capture in a CUDA graph); CTA -> job through the chunk prefix.
constexpr int kPackChunk = 4096;           // elements per CTA

__device__ __forceinline__ void pack_store(void* dst, int dt, int64_t i, float v)
{
    if (dt == DT_F32) ((float*)dst)[i] = v;
    else ((__nv_bfloat16*)dst)[i] = __float2bfloat16_rn(v);
}

__global__ void __launch_bounds__(256)
pack_jobs_kernel(const __grid_constant__ PackJobs jobs)
{
    // job of this CTA: last j with chunk0[j] <= blockIdx.x
    int lo = 0, hi = jobs.n - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (jobs.j[mid].chunk0 <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
    }
    const PackJob& J = jobs.j[lo];
    const int64_t beg = (int64_t)((int)blockIdx.x - J.chunk0) * kPackChunk;
    if (J.kind == PACK_CAST) {
        const int64_t n = J.a;
        for (int64_t i = beg + threadIdx.x; i < beg + kPackChunk && i < n; i += 256) pack_store(J.dst, J.dt, i, J.src[i]);
    } else if (J.kind == PACK_REPACK) {
        // out[o][tap][c] = w[o][c][tap] (* FrozenBN scale of channel o, folded before rounding)
        const int Cin = J.b, RS = J.c;
        const int64_t n = (int64_t)J.a * Cin * RS;
        for (int64_t i = beg + threadIdx.x; i < beg + kPackChunk && i < n; i += 256) {
            const int c = (int)(i % Cin);
            const int tap = (int)((i / Cin) % RS);
            const int o = (int)(i / ((int64_t)Cin * RS));
            float v = J.src[((int64_t)o * Cin + c) * RS + tap];
            if (J.bn_w != nullptr) v *= J.bn_w[o] * rsqrtf(J.bn_var[o] + 1e-5f);
            pack_store(J.dst, J.dt, i, v);
        }
    } else {      // PACK_BNFOLD: scale -> dst, bias -> dst2 (sedt/backbone.py:43-53)
        const int64_t n = J.a;
        for (int64_t i = beg + threadIdx.x; i < beg + kPackChunk && i < n; i += 256) {
            const float sc = J.bn_w[i] * rsqrtf(J.bn_var[i] + 1e-5f);
            ((float*)J.dst)[i] = sc;
            J.dst2[i] = J.src[i] - J.bn_mean[i] * sc;
        }
    }
}

static inline unsigned grid_for(int64_t n, int block = 256)
{
    int64_t g = ceil_div(n, block);
    return (unsigned)(g < 1 ? 1 : (g > 148 * 16 ? 148 * 16 : g));
}

}  // namespace

int PackBatch::add(const PackJob& job, int64_t elems)
{
    if (elems <= 0) return SEDT_OK;
    if (cur_.n == kPackMaxJobs) SEDT_TRY(flush());
    PackJob j = job;
    j.chunk0 = chunks_;
    cur_.j[cur_.n++] = j;
    chunks_ += (int)ceil_div(elems, kPackChunk);
    return SEDT_OK;
}

int PackBatch::flush()
{
    if (cur_.n == 0) return SEDT_OK;
    pack_jobs_kernel<<<(unsigned)chunks_, 256, 0, stream_>>>(cur_);
    SEDT_COUNT_LAUNCH();
    SEDT_CHECK_CUDA(cudaGetLastError());
    cur_.n = 0; chunks_ = 0;
    return SEDT_OK;
}

int PackBatch::cast(const float* in, void* out, int dt, int64_t n)
{
    PackJob j{}; j.kind = PACK_CAST; j.src = in; j.dst = out; j.dt = dt; j.a = (int)n;
    return add(j, n);
}

int PackBatch::repack_conv(const float* w_oihw, const float* bn_w, const float* bn_var, void* out, int dt, int Cout, int Cin, int RS)
{
    PackJob j{}; j.kind = PACK_REPACK; j.src = w_oihw; j.bn_w = bn_w; j.bn_var = bn_var; j.dst = out; j.dt = dt;
    j.a = Cout; j.b = Cin; j.c = RS;
    return add(j, (int64_t)Cout * Cin * RS);
}

int PackBatch::bn_fold(const float* w, const float* b, const float* mean, const float* var, float* scale, float* bias, int n)
{
    PackJob j{}; j.kind = PACK_BNFOLD; j.bn_w = w; j.src = b; j.bn_mean = mean; j.bn_var = var; j.dst = scale; j.dst2 = bias; j.a = n;
    return add(j, n);
}

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
