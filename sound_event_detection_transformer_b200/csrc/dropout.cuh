// Counter-based dropout masks shared by the forward and backward kernels of the training step.
// keep(e) for element e of dropout site `site` at training step `step` is bit-identical wherever it is
// evaluated: Philox4x32-10 with key = seed, counter = (e / 4, site, step), lane e % 4; kept with probability
// 1 - p (uniform 32-bit draw < thresh).  Replaces torch's dropout RNG (same distribution, different stream).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace sedt {

struct DropSite {
    const unsigned long long* state;   // device: [0] = seed, [1] = step counter; nullptr = no dropout
    uint32_t site;
    uint32_t thresh;                   // keep iff draw < thresh
    float inv_keep;                    // 1 / (1 - p)
};

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k)
{
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
        const uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += W0; k.y += W1;
    }
    return c;
}

// four draws for elements 4*e4 .. 4*e4+3
__device__ __forceinline__ uint4 drop_draw4(const DropSite& d, unsigned long long seed, unsigned long long step, unsigned long long e4)
{
    return philox4x32_10(make_uint4((uint32_t)e4, (uint32_t)(e4 >> 32), d.site, (uint32_t)step),
                         make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
}

inline DropSite make_drop_site(const unsigned long long* state, uint32_t site, float p)
{
    DropSite d;
    d.state = p > 0.f ? state : nullptr;
    d.site = site;
    const double keep = 1.0 - (double)p;
    d.thresh = keep >= 1.0 ? 0xFFFFFFFFu : (uint32_t)(keep * 4294967296.0);
    d.inv_keep = p > 0.f ? (float)(1.0 / keep) : 1.f;
    return d;
}

}  // namespace sedt
