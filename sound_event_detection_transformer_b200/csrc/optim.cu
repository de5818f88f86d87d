// Gradient clipping + AdamW over a table of tensors: the step right after loss.backward() in the
// reference's training loop (engine.py:76-80: clip_grad_norm_(model.parameters(), 0.1); optimizer.step(),
// optimizer = AdamW with two lr groups and weight decay 1e-4, train_sedt.py:234-240,269-270).
//
// Stock PyTorch runs ~10 foreach passes over ~200 tensors (norm, clip, mul, lerp, mul, addcmul, sqrt, div, add,
// addcdiv).  Here: one pass for the global L2 norm (fixed-order reduction, no atomics) and one pass that reads
// p, g, m, v once and writes p, m, v once -- 28 bytes per parameter element, HBM-bound.  The clip coefficient is
// read from device memory, so nothing synchronises with the host.  Work is split in chunks of kChunk elements of
// one tensor; the (tensor, chunk) table is built once by the caller.
#include "common.cuh"
#include "kernels.h"

namespace sedt {

namespace {

constexpr int kThreads = 256;
constexpr int kChunk = 4096;       // elements of one tensor per CTA

struct OptimTensor {               // == sedt_optim_tensor (include/sedt_b200.h)
    float* param; float* grad; float* exp_avg; float* exp_avg_sq;
    int64_t numel; int32_t group; int32_t reserved;
};

// per parameter group, every scalar computed in double on the host and rounded once, as torch does when it hands a
// Python float to a tensor op: decay = 1 - lr*weight_decay, w1 = 1 - beta1, w2 = 1 - beta2,
// neg_step = -lr / (1 - beta1^step), bc2_sqrt = sqrt(1 - beta2^step)        (== sedt_adamw_group)
struct AdamWGroup { float decay, w1, beta2, w2, bc2_sqrt, eps, neg_step, reserved; };
struct AdamWGroups { AdamWGroup g[8]; };

__device__ __forceinline__ float block_sum(float v, float* red) {
    v = warp_sum(v);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float t = 0.f;
    if (warp == 0) {
        t = lane < kThreads / 32 ? red[lane] : 0.f;
        t = warp_sum(t);
    }
    return t;                      // valid in warp 0
}

__global__ void __launch_bounds__(kThreads)
grad_sumsq_kernel(const OptimTensor* __restrict__ tensors, const int32_t* __restrict__ chunks, float* __restrict__ partials)
{
    __shared__ float red[kThreads / 32];
    const int ti = chunks[2 * blockIdx.x], ci = chunks[2 * blockIdx.x + 1];
    const OptimTensor t = tensors[ti];
    const int64_t beg = (int64_t)ci * kChunk;
    const int n = (int)min((int64_t)kChunk, t.numel - beg);
    const float* g = t.grad + beg;
    float acc = 0.f;
    if ((reinterpret_cast<uintptr_t>(g) & 15) == 0) {
        const float4* g4 = reinterpret_cast<const float4*>(g);
        for (int i = threadIdx.x; i < n / 4; i += kThreads) {
            const float4 v = g4[i];
            acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
        }
        for (int i = (n & ~3) + threadIdx.x; i < n; i += kThreads) acc += g[i] * g[i];
    } else {
        for (int i = threadIdx.x; i < n; i += kThreads) acc += g[i] * g[i];
    }
    const float s = block_sum(acc, red);
    if (threadIdx.x == 0) partials[blockIdx.x] = s;
}

// one CTA: sqrt of the sum of the chunk partials, accumulated in fp64 in a fixed order
__global__ void __launch_bounds__(kThreads)
grad_norm_finish_kernel(const float* __restrict__ partials, int n, float* __restrict__ norm_out)
{
    __shared__ double red[kThreads];
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += kThreads) acc += (double)partials[i];
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int o = kThreads / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) norm_out[0] = (float)sqrt(red[0]);
}

// torch.nn.utils.clip_grad_norm_: coef = min(1, max_norm / (total_norm + 1e-6))
__device__ __forceinline__ float clip_coef(const float* norm, float max_norm) {
    if (norm == nullptr || !(max_norm > 0.f)) return 1.f;
    const float c = max_norm / (norm[0] + 1e-6f);
    return c < 1.f ? c : (c != c ? c : 1.f);            // torch clamps with max = 1: a NaN norm propagates to every gradient
}

__global__ void __launch_bounds__(kThreads)
grad_scale_kernel(const OptimTensor* __restrict__ tensors, const int32_t* __restrict__ chunks,
                  const float* __restrict__ norm, float max_norm)
{
    const int ti = chunks[2 * blockIdx.x], ci = chunks[2 * blockIdx.x + 1];
    const OptimTensor t = tensors[ti];
    const int64_t beg = (int64_t)ci * kChunk;
    const int n = (int)min((int64_t)kChunk, t.numel - beg);
    const float c = clip_coef(norm, max_norm);
    float* g = t.grad + beg;
    for (int i = threadIdx.x; i < n; i += kThreads) g[i] *= c;
}

// torch.optim.AdamW (_single_tensor_adamw, amsgrad = False, maximize = False), same operation order in fp32
__device__ __forceinline__ void adamw_one(float& p, float g, float& m, float& v, const AdamWGroup& h, float coef) {
    g *= coef;
    p *= h.decay;                                       // param.mul_(1 - lr * weight_decay)
    m = m + h.w1 * (g - m);                             // exp_avg.lerp_(grad, 1 - beta1)   (weight < 0.5 branch)
    v = v * h.beta2;                                    // exp_avg_sq.mul_(beta2)
    v = v + (h.w2 * g) * g;                             //           .addcmul_(grad, grad, value = 1 - beta2)
    const float denom = sqrtf(v) / h.bc2_sqrt + h.eps;
    p = p + h.neg_step * (m / denom);                   // param.addcdiv_(exp_avg, denom, value = -lr / bias_correction1)
}

__global__ void __launch_bounds__(kThreads)
adamw_kernel(const OptimTensor* __restrict__ tensors, const int32_t* __restrict__ chunks, const AdamWGroups groups,
             const float* __restrict__ norm, float max_norm)
{
    const int ti = chunks[2 * blockIdx.x], ci = chunks[2 * blockIdx.x + 1];
    const OptimTensor t = tensors[ti];
    const AdamWGroup h = groups.g[t.group & 7];
    const int64_t beg = (int64_t)ci * kChunk;
    const int n = (int)min((int64_t)kChunk, t.numel - beg);
    const float coef = clip_coef(norm, max_norm);
    float* p = t.param + beg; const float* g = t.grad + beg; float* m = t.exp_avg + beg; float* v = t.exp_avg_sq + beg;
    const uintptr_t al = reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                         reinterpret_cast<uintptr_t>(v);
    int done = 0;
    if ((al & 15) == 0) {
        float4* p4 = reinterpret_cast<float4*>(p); const float4* g4 = reinterpret_cast<const float4*>(g);
        float4* m4 = reinterpret_cast<float4*>(m); float4* v4 = reinterpret_cast<float4*>(v);
        for (int i = threadIdx.x; i < n / 4; i += kThreads) {
            float4 pp = p4[i], gg = g4[i], mm = m4[i], vv = v4[i];
            adamw_one(pp.x, gg.x, mm.x, vv.x, h, coef);
            adamw_one(pp.y, gg.y, mm.y, vv.y, h, coef);
            adamw_one(pp.z, gg.z, mm.z, vv.z, h, coef);
            adamw_one(pp.w, gg.w, mm.w, vv.w, h, coef);
            p4[i] = pp; m4[i] = mm; v4[i] = vv;
        }
        done = n & ~3;
    }
    for (int i = done + threadIdx.x; i < n; i += kThreads) {
        float pp = p[i], mm = m[i], vv = v[i];
        adamw_one(pp, g[i], mm, vv, h, coef);
        p[i] = pp; m[i] = mm; v[i] = vv;
    }
}

}  // namespace

int optim_chunk_elems() { return kChunk; }

int launch_grad_norm(const void* tensors, const int32_t* chunks, int nchunks, float* partials, float* norm_out,
                     cudaStream_t stream)
{
    SEDT_REQUIRE(nchunks >= 0 && norm_out != nullptr, "grad_norm: bad arguments");
    ProfScope _prof(PROF_OTHER, stream);
    if (nchunks > 0) {
        SEDT_REQUIRE(tensors && chunks && partials, "grad_norm: null table");
        grad_sumsq_kernel<<<nchunks, kThreads, 0, stream>>>((const OptimTensor*)tensors, chunks, partials);
        SEDT_COUNT_LAUNCH();
    }
    grad_norm_finish_kernel<<<1, kThreads, 0, stream>>>(partials, nchunks, norm_out);
    SEDT_COUNT_LAUNCH();
    SEDT_CHECK_CUDA(cudaGetLastError());
    return SEDT_OK;
}

int launch_grad_scale(const void* tensors, const int32_t* chunks, int nchunks, const float* norm, float max_norm,
                      cudaStream_t stream)
{
    if (nchunks <= 0) return SEDT_OK;
    SEDT_REQUIRE(tensors && chunks && norm, "clip_grads: null argument");
    ProfScope _prof(PROF_OTHER, stream);
    grad_scale_kernel<<<nchunks, kThreads, 0, stream>>>((const OptimTensor*)tensors, chunks, norm, max_norm);
    SEDT_COUNT_LAUNCH();
    SEDT_CHECK_CUDA(cudaGetLastError());
    return SEDT_OK;
}

int launch_adamw(const void* tensors, const int32_t* chunks, int nchunks, const float* groups_host, int ngroups,
                 const float* norm, float max_norm, cudaStream_t stream)
{
    if (nchunks <= 0) return SEDT_OK;
    SEDT_REQUIRE(tensors && chunks && groups_host, "adamw: null argument");
    SEDT_REQUIRE(ngroups >= 1 && ngroups <= 8, "adamw: %d parameter groups (1..8 supported)", ngroups);
    AdamWGroups gs;
    memset(&gs, 0, sizeof(gs));
    memcpy(&gs, groups_host, sizeof(AdamWGroup) * ngroups);
    for (int i = 0; i < ngroups; ++i) {
        const AdamWGroup& h = gs.g[i];
        SEDT_REQUIRE(h.w1 > 0.f && h.w1 < 0.5f, "adamw: 1 - beta1 = %g (the lerp form used here needs 0.5 < beta1 < 1)", h.w1);
        SEDT_REQUIRE(h.bc2_sqrt > 0.f, "adamw: bias correction must be positive");
    }
    ProfScope _prof(PROF_OTHER, stream);
    adamw_kernel<<<nchunks, kThreads, 0, stream>>>((const OptimTensor*)tensors, chunks, gs, norm, max_norm);
    SEDT_COUNT_LAUNCH();
    SEDT_CHECK_CUDA(cudaGetLastError());
    return SEDT_OK;
}

}  // namespace sedt
