// Fused backbone stem: conv0 (1x1, 1->3, bias) -> conv1 (7x7, s2, p3, 3->64)
// -> FrozenBatchNorm -> ReLU -> maxpool (3x3, s2, p1), written NHWC.
// Reference: sedt/backbone.py:97-111 + torchvision resnet.py:266-272.
//
// conv0 is folded into conv1: Weff[o][tap] = sum_c W1[o][c][tap] * w0[c] acts on
// the single input channel, and conv0's bias contributes
// sum_c W1[o][c][tap] * b0[c] for every tap that lands INSIDE the clip (conv1's
// zero padding is applied after conv0, so padded taps contribute nothing).  The
// per-pixel bias sum over the valid tap rectangle is read from an 8x8
// summed-area table, which makes the fold exact at the borders.
//
// One CTA produces PH=4 pooled rows x 16 pooled columns x 64 channels: 9 warps
// compute the 9 conv rows feeding them (8 channels x 4 pixels per lane, inputs
// and weights staged in shared memory), the tile is pooled from shared memory.
#include "kernels.h"
#include <math_constants.h>

namespace sedt {
namespace {

constexpr int PH = 4;                 // pooled rows per CTA
constexpr int CR = 2 * PH + 1;        // conv rows per CTA (9)
constexpr int XR = 2 * CR + 5;        // input rows per CTA (23)
constexpr int XC = 72;                // input cols -3..66 padded to 72
constexpr int WC = 32, WP = 16;       // conv / pooled width for F = 64
constexpr int NTHREADS = CR * 32;

constexpr size_t kStemSmem = sizeof(float) * (XR * XC + 49 * 64 + 64 * 64 + CR * WC * 64);

template <typename TO>
__global__ void __launch_bounds__(NTHREADS)
stem_kernel(const float* __restrict__ x, const float* __restrict__ weff, const float* __restrict__ sat,
            const float* __restrict__ scale, const float* __restrict__ bias, TO* __restrict__ out,
            int T, int Hc, int Hp)
{
    extern __shared__ __align__(16) float smem[];
    float* xs = smem;                       // [XR][XC]
    float* ws = xs + XR * XC;               // [49][64]
    float* ss = ws + 49 * 64;               // [8][8][64]
    float* cs = ss + 64 * 64;               // [CR][WC][64]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.y;
    const int hp0 = blockIdx.x * PH;
    const int row0 = 4 * hp0 - 5;           // first input row held in xs
    const float* xb = x + (size_t)b * T * 64;

    for (int i = tid; i < XR * XC; i += NTHREADS) {
        const int ri = i / XC, ci = i % XC;
        const int row = row0 + ri, col = ci - 3;
        xs[i] = (row >= 0 && row < T && col >= 0 && col < 64) ? xb[(size_t)row * 64 + col] : 0.f;
    }
    for (int i = tid; i < 49 * 64; i += NTHREADS) ws[i] = weff[i];
    for (int i = tid; i < 64 * 64; i += NTHREADS) ss[i] = sat[i];
    __syncthreads();

    // ---- conv phase: warp <-> conv row
    const int hc = 2 * hp0 - 1 + warp;
    if (hc >= 0 && hc < Hc) {
        const int cg = lane & 7, quad = lane >> 3;
        // valid tap rows for this conv row (conv1 padding 3, stride 2)
        const int rlo = max(0, 3 - 2 * hc), rhi = min(6, T + 2 - 2 * hc);
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
            const int wc0 = half * 16 + quad * 4;
            float acc[4][8];
#pragma unroll
            for (int p = 0; p < 4; ++p)
#pragma unroll
                for (int c = 0; c < 8; ++c) acc[p][c] = 0.f;
#pragma unroll 1
            for (int r = 0; r < 7; ++r) {
                const float* xr = xs + (2 * warp + r) * XC + 2 * wc0;
                float xv[13];
#pragma unroll
                for (int i = 0; i < 13; ++i) xv[i] = xr[i];
#pragma unroll
                for (int s = 0; s < 7; ++s) {
                    const float4 w0 = *reinterpret_cast<const float4*>(ws + (r * 7 + s) * 64 + cg * 8);
                    const float4 w1 = *reinterpret_cast<const float4*>(ws + (r * 7 + s) * 64 + cg * 8 + 4);
                    const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                    for (int p = 0; p < 4; ++p)
#pragma unroll
                        for (int c = 0; c < 8; ++c) acc[p][c] = fmaf(xv[2 * p + s], wv[c], acc[p][c]);
                }
            }
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                const int wc = wc0 + p;
                const int slo = max(0, 3 - 2 * wc), shi = min(6, 66 - 2 * wc);
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const int ch = cg * 8 + c;
                    // inclusive SAT with a zero border: S[r+1][s+1]
                    const float bsum = ss[((rhi + 1) * 8 + (shi + 1)) * 64 + ch] - ss[(rlo * 8 + (shi + 1)) * 64 + ch]
                                     - ss[((rhi + 1) * 8 + slo) * 64 + ch] + ss[(rlo * 8 + slo) * 64 + ch];
                    float v = (acc[p][c] + bsum) * scale[ch] + bias[ch];
                    cs[(warp * WC + wc) * 64 + ch] = fmaxf(v, 0.f);
                }
            }
        }
    }
    __syncthreads();

    // ---- pool phase
    for (int i = tid; i < PH * WP * 64; i += NTHREADS) {
        const int ch = i & 63, wp = (i >> 6) & 15, hl = i >> 10;
        const int hp = hp0 + hl;
        if (hp >= Hp) continue;
        float m = -CUDART_INF_F;
#pragma unroll
        for (int dr = 0; dr < 3; ++dr) {
            const int hcc = 2 * hp - 1 + dr;
            if (hcc < 0 || hcc >= Hc) continue;
#pragma unroll
            for (int dc = 0; dc < 3; ++dc) {
                const int wcc = 2 * wp - 1 + dc;
                if (wcc < 0 || wcc >= WC) continue;
                m = fmaxf(m, cs[((2 * hl + dr) * WC + wcc) * 64 + ch]);
            }
        }
        out[(((size_t)b * Hp + hp) * WP + wp) * 64 + ch] = from_f32<TO>(m);
    }
}

// one thread per output channel: fold conv0 into conv1 and build the SAT
__global__ void stem_pack_kernel(const float* __restrict__ w0, const float* __restrict__ b0,
                                 const float* __restrict__ w1, float* __restrict__ weff, float* __restrict__ sat)
{
    const int o = threadIdx.x;
    if (o >= 64) return;
    float wb[7][7];
    for (int r = 0; r < 7; ++r)
        for (int s = 0; s < 7; ++s) {
            float we = 0.f, wbv = 0.f;
            for (int c = 0; c < 3; ++c) {
                const float v = w1[((o * 3 + c) * 7 + r) * 7 + s];
                we = fmaf(v, w0[c], we);
                wbv = fmaf(v, b0[c], wbv);
            }
            weff[(r * 7 + s) * 64 + o] = we;
            wb[r][s] = wbv;
        }
    for (int r = 0; r <= 7; ++r)
        for (int s = 0; s <= 7; ++s) {
            double acc = 0.0;        // prefix sums in fp64, stored fp32
            for (int rr = 0; rr < r; ++rr)
                for (int s2 = 0; s2 < s; ++s2) acc += (double)wb[rr][s2];
            sat[(r * 8 + s) * 64 + o] = (float)acc;
        }
}

}  // namespace

int launch_stem_pack(const float* conv0_w, const float* conv0_b, const float* conv1_w, float* weff, float* sat,
                     cudaStream_t stream)
{
    stem_pack_kernel<<<1, 64, 0, stream>>>(conv0_w, conv0_b, conv1_w, weff, sat);
    SEDT_COUNT_LAUNCH();
    SEDT_CHECK_CUDA(cudaGetLastError());
    return SEDT_OK;
}

int launch_stem(const float* x, const StemWeights& w, void* out, int out_dt, int B, int T, int F, cudaStream_t stream)
{
    SEDT_REQUIRE(F == 64, "stem: the fused stem kernel needs 64 mel bins (config.py n_mels), got F=%d", F);
    SEDT_REQUIRE(T >= 1, "stem: T=%d", T);
    if (B == 0) return SEDT_OK;
    const int Hc = (T - 1) / 2 + 1;
    const int Hp = (Hc - 1) / 2 + 1;
    dim3 grid((unsigned)ceil_div(Hp, PH), (unsigned)B), block(NTHREADS);
    ProfScope _prof(PROF_STEM, stream);
    if (out_dt == DT_F32) {
        SEDT_CHECK_CUDA(cudaFuncSetAttribute(stem_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)kStemSmem));
        stem_kernel<float><<<grid, block, kStemSmem, stream>>>(x, w.weff, w.sat, w.scale, w.bias, (float*)out, T, Hc, Hp);
    } else {
        SEDT_CHECK_CUDA(cudaFuncSetAttribute(stem_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)kStemSmem));
        stem_kernel<__nv_bfloat16><<<grid, block, kStemSmem, stream>>>(x, w.weff, w.sat, w.scale, w.bias,
                                                                       (__nv_bfloat16*)out, T, Hc, Hp);
    }
    SEDT_COUNT_LAUNCH();
    SEDT_CHECK_CUDA(cudaGetLastError());
    return SEDT_OK;
}

}  // namespace sedt
