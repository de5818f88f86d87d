// Deterministic part of the reference's input pipeline on the device (utilities/BoxTransforms.py:454-490 get_transforms with
// frames + scaler, i.e. the evaluation recipe: ApplyLog -> PadOrTrunc -> ToTensor -> Normalize):
//   ApplyLog      librosa.amplitude_to_db(S) = max(10 log10(max(1e-10, S^2)), max over the clip - 80)   (:55-67; third party)
//   PadOrTrunc    zero-pad (in the dB domain) or truncate to `frames` rows                                  (:70-118)
//   Normalize     (x - mean_[f]) / std_[f] in float64, rounded to float32                                   (utilities/Scaler.py:102-108)
// One CTA per clip: a block reduction for the clip maximum, then one pass that writes the [frames, F] fp32 clip the model
// reads.  HBM-bound: 4 B in + 4 B out per element.  The random augmentations (time / frequency masks, frequency shift,
// mixup) and the SP-SEDT patch crop + resize are not built.
#include "common.cuh"
#include "kernels.h"
#include <math_constants.h>

namespace sedt {
namespace {

__device__ __forceinline__ float amp_to_db(float s) { return 10.0f * log10f(fmaxf(1e-10f, s * s)); }

__global__ void __launch_bounds__(256)
prepare_clips_kernel(const float* __restrict__ raw, const int64_t* __restrict__ offsets, const double* __restrict__ mean,
                     const double* __restrict__ stdv, float* __restrict__ out, int frames, int F, int apply_log)
{
    __shared__ float red[8];
    __shared__ float s_floor;
    const int b = blockIdx.x;
    const int64_t r0 = offsets[b];
    const int T = (int)(offsets[b + 1] - r0);
    const float* src = raw + r0 * F;
    if (apply_log) {
        float m = -CUDART_INF_F;
        for (int64_t i = threadIdx.x; i < (int64_t)T * F; i += 256) m = fmaxf(m, amp_to_db(src[i]));
        m = warp_max(m);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
        __syncthreads();
        if (threadIdx.x == 0) {
            float t = red[0];
            for (int w = 1; w < 8; ++w) t = fmaxf(t, red[w]);
            s_floor = t - 80.0f;                       // top_db
        }
        __syncthreads();
    }
    const float floor_db = apply_log ? s_floor : 0.f;
    float* dst = out + (size_t)b * frames * F;
    for (int64_t i = threadIdx.x; i < (int64_t)frames * F; i += 256) {
        const int row = (int)(i / F), f = (int)(i - (int64_t)row * F);
        float v = 0.f;                                 // np.pad(..., mode="constant") after the log
        if (row < T) {
            v = src[i];
            if (apply_log) v = fmaxf(amp_to_db(v), floor_db);
        }
        dst[i] = mean != nullptr ? (float)(((double)v - mean[f]) / stdv[f]) : v;
    }
}

}  // namespace

int launch_prepare_clips(const float* raw, const int64_t* offsets, const double* mean, const double* stdv, float* out, int B,
                         int frames, int F, int apply_log, cudaStream_t stream)
{
    if (B == 0) return SEDT_OK;
    SEDT_REQUIRE(raw && offsets && out && frames >= 1 && F >= 1, "prepare_clips: bad arguments");
    SEDT_REQUIRE((mean == nullptr) == (stdv == nullptr), "prepare_clips: mean and std go together");
    ProfScope _prof(PROF_OTHER, stream);
    prepare_clips_kernel<<<(unsigned)B, 256, 0, stream>>>(raw, offsets, mean, stdv, out, frames, F, apply_log);
    SEDT_COUNT_LAUNCH();
    SEDT_CHECK_CUDA(cudaGetLastError());
    return SEDT_OK;
}

}  // namespace sedt
