// Shared helpers for the sedt_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <utility>

namespace sedt {

// ---- error plumbing: C-ABI functions return int and never throw -------------
enum : int {
    SEDT_OK = 0,
    SEDT_ERR_INVALID = -1,      // bad argument / unsupported shape
    SEDT_ERR_CUDA = -2,         // a CUDA runtime/driver call failed
    SEDT_ERR_WORKSPACE = -3,    // workspace or packed buffer too small
    SEDT_ERR_NUMERIC = -4,      // NaN/-inf cost entries (matcher) -> ValueError upstream
    SEDT_ERR_INFEASIBLE = -5,   // LSAP infeasible
    SEDT_ERR_UNSUPPORTED = -6,
};

void set_error(const char* fmt, ...);
const char* get_error();

#define SEDT_CHECK_CUDA(expr)                                                                  \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            ::sedt::set_error("%s:%d CUDA error %s: %s", __FILE__, __LINE__, #expr,            \
                              cudaGetErrorString(_e));                                         \
            return ::sedt::SEDT_ERR_CUDA;                                                      \
        }                                                                                      \
    } while (0)

#define SEDT_REQUIRE(cond, ...)                                                                \
    do {                                                                                       \
        if (!(cond)) {                                                                         \
            ::sedt::set_error(__VA_ARGS__);                                                    \
            return ::sedt::SEDT_ERR_INVALID;                                                   \
        }                                                                                      \
    } while (0)

#define SEDT_TRY(expr)                                                                         \
    do {                                                                                       \
        int _rc = (expr);                                                                      \
        if (_rc != 0) return _rc;                                                              \
    } while (0)

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int64_t align_up(int64_t a, int64_t b) { return ceil_div(a, b) * b; }

// launch counter: bench.py reports how many of OUR kernels ran in the timed region
extern unsigned long long g_launch_count;
#define SEDT_COUNT_LAUNCH() (++::sedt::g_launch_count)
// per-kernel launch counters for the kernels whose selection depends on the problem size (launch_conv_tc dispatch, fused
// blocks): the parity tests assert that the launch sequence they checked is the one the benchmark times
enum KernelKind : int { KK_CONV_TC2 = 0, KK_CONV_TC3_2SM, KK_CONV_TC4_WS, KK_CONV_TC_V1, KK_FFN_FUSED, KK_ATTENTION_TC, KK_STEM_TC,
                        KK_WGRAD_TC, KK_ATTENTION_BWD_TC, KK_ENC_ATTN_FUSED, KK_BOTTLENECK_FUSED, KK_DEC_LAYER_FUSED, KK_NKINDS };
extern unsigned long long g_kind_count[KK_NKINDS];
const char* kernel_kind_name(int kind);
#define SEDT_COUNT_KIND(kind) (++::sedt::g_launch_count, ++::sedt::g_kind_count[kind])

// ---- optional per-kernel-class device timing (bench.py roofline evidence) ---------------
// When enabled, every launcher brackets its kernel with two CUDA events on the launching
// stream; sedt_profile_read() synchronises and sums elapsed time per class.  Off by default
// (no events are recorded and nothing is synchronised).
enum ProfClass : int { PROF_GEMM_TC = 0, PROF_GEMM_SIMT, PROF_STEM, PROF_ATTENTION, PROF_NORM, PROF_MATCHER, PROF_OTHER, PROF_NCLASS };
extern bool g_prof_on;
void prof_begin(int cls, cudaStream_t s);
void prof_end(cudaStream_t s);
struct ProfScope {
    cudaStream_t s; bool on;
    ProfScope(int cls, cudaStream_t st) : s(st), on(g_prof_on) { if (on) prof_begin(cls, s); }
    ~ProfScope() { if (on) prof_end(s); }
};

// ---- programmatic dependent launch (PDL) -------------------------------------------------
// A kernel launched through launch_pdl() may start while its predecessor on the stream is still draining: its
// prologue (barrier init, TMEM allocation, tensor-map prefetch, launch latency itself) overlaps the predecessor's
// tail.  Contract: such a kernel executes pdl_wait() before it touches any global memory another kernel produces or
// reads (the wait returns once the predecessor grid has completed and its writes are visible; every kernel in the
// chain waits the same way, so completion is transitive), and pdl_trigger() at its top so that ITS successor may be
// scheduled early.  Both are no-ops when the kernel was launched without the attribute.  Opt-in with SEDT_PDL=1 (measured slower, see
// pdl_enabled()).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
bool pdl_enabled();

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, int cluster_x,
                              Args&&... args)
{
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    int n = 0;
    if (cluster_x > 1) {
        attr[n].id = cudaLaunchAttributeClusterDimension;
        attr[n].val.clusterDim.x = (unsigned)cluster_x; attr[n].val.clusterDim.y = 1; attr[n].val.clusterDim.z = 1;
        ++n;
    }
    if (pdl_enabled()) {
        attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[n].val.programmaticStreamSerializationAllowed = 1;
        ++n;
    }
    cfg.attrs = attr; cfg.numAttrs = (unsigned)n;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(std::forward<Args>(args))...);
}

// ---- device-side conversions -------------------------------------------------
template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

}  // namespace sedt
