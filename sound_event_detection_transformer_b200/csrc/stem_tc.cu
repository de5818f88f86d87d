// Backbone stem on tcgen05 (bf16 tier): conv0 (1x1, bias) + conv1 (7x7 s2 p3) + FrozenBN + ReLU +
// maxpool (3x3 s2 p1) in one kernel.  Reference: sedt/backbone.py:97-111, torchvision resnet.py:266-272.
//
// conv0 folded into conv1 turns the stem into a 2-input-channel 7x7 convolution: channel 0 is the
// log-mel value, channel 1 is an indicator that is 1 inside the clip and 0 in conv1's zero padding
// (conv0's bias only reaches taps that land inside the clip).  As a GEMM that is K = 2 x 49, padded
// to 2 x 64: the A tile [128 conv pixels x 128] is built in shared memory by the CTA's threads
// (im2col of a staged fp32 patch, written directly in the 128B-swizzled K-major layout), B
// [64 channels x 128] is a pre-swizzled 16 KiB image produced at pack time (BN scale folded in) and
// fetched with one bulk copy.  D lives in TMEM; the epilogue adds the BN bias, applies ReLU and
// parks the conv tile in shared memory as bf16, from where the 3x3/s2 max-pool is taken.
//
// One CTA = 4 pooled rows of one clip = 9 conv rows = 3 M-tiles of 4 conv rows.
#include "tc_common.cuh"

namespace sedt {
namespace {

using namespace tc;

constexpr int ST_PH = 4;                    // pooled rows per CTA
constexpr int ST_TILES = 3;                 // 12 conv rows are computed, 9 are used
constexpr int ST_XR = 2 * (4 * ST_TILES) + 5;   // 29 input rows
constexpr int ST_XC = 72;
constexpr int A_OFF = 0;                    // 2 chunks x [128 px][64 k] bf16 = 32 KiB
constexpr int B_OFF = 32768;                // 2 chunks x [64 ch][64 k] bf16 = 16 KiB
constexpr int C_OFF = 49152;                // [12 conv rows][32 px][64 ch] bf16 = 48 KiB
constexpr int X_OFF = 98304;                // [29][72] fp32
constexpr int BIAS_OFF = X_OFF + ST_XR * ST_XC * 4;
constexpr int STBAR_OFF = ((BIAS_OFF + 256 + 15) / 16) * 16;
constexpr int ST_SMEM = STBAR_OFF + 64 + 1024;

__device__ __forceinline__ uint32_t swz128(int row, int piece) { return (uint32_t)(row * 128 + ((piece ^ (row & 7)) << 4)); }

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}

// kArgmax (training forward): also records, per pooled element, which of the 9 window positions (dr * 3 + dc, first
// maximum) supplied the value, or 9 when the window's maximum is <= 0 (dead ReLU: no gradient) -- the backward of
// max-pool + ReLU (stem_bwd_kernel) then needs no recomputation of the convolution.
template <bool kArgmax>
__global__ void __launch_bounds__(128)
stem_tc_kernel(const float* __restrict__ x, const uint8_t* __restrict__ wtc, const float* __restrict__ bias,
               __nv_bfloat16* __restrict__ out, uint8_t* __restrict__ amax, int T, int Hc, int Hp)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    float* xs = (float*)(smem + X_OFF);
    float* sbias = (float*)(smem + BIAS_OFF);
    uint64_t* bar_w = (uint64_t*)(smem + STBAR_OFF);
    uint64_t* bar_mma = bar_w + 1;
    uint32_t* tmem_slot = (uint32_t*)(bar_mma + 1);

    const int t = threadIdx.x, warp = t >> 5;
    const int b = blockIdx.y, hp0 = blockIdx.x * ST_PH;
    const int cr0 = 2 * hp0 - 1;                 // first conv row of this CTA
    const int row0 = 2 * cr0 - 3;                // first input row held in xs

    pdl_wait();            // the output buffer may still be read by the previous kernel (workspace reuse across steps)
    if (t == 0) {
        mbar_init(bar_w, 1); mbar_init(bar_mma, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(bar_w, 16384);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(smem + B_OFF)), "l"(wtc), "r"(16384), "r"(smem_u32(bar_w)) : "memory");
    }
    if (warp == 0) tmem_alloc<64>(tmem_slot);

    const float* xb = x + (size_t)b * T * 64;
    for (int i = t; i < ST_XR * ST_XC; i += 128) {
        const int ri = i / ST_XC, ci = i - ri * ST_XC;
        const int row = row0 + ri, col = ci - 3;
        xs[i] = (row >= 0 && row < T && col >= 0 && col < 64) ? __ldg(xb + (size_t)row * 64 + col) : 0.f;
    }
    if (t < 64) sbias[t] = bias[t];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    pdl_trigger();         // TMEM is owned: a successor scheduled next to this CTA cannot starve it
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);

    const int wc = t & 31;
#pragma unroll 1
    for (int tile = 0; tile < ST_TILES; ++tile) {
        const int lr = tile * 4 + (t >> 5);      // local conv row of this thread's pixel
        const int hc = cr0 + lr;
        // ---- im2col row of pixel (hc, wc): chunk 0 = log-mel taps, chunk 1 = inside indicator ----
        {
            uint32_t av[32], iv[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) { av[i] = 0u; iv[i] = 0u; }
            float prev = 0.f; float pin = 0.f;
#pragma unroll
            for (int r = 0; r < 7; ++r) {
                const float* xr = xs + (2 * lr + r) * ST_XC + 2 * wc;
                const int irow = 2 * hc - 3 + r;
                const bool rin = irow >= 0 && irow < T;
#pragma unroll
                for (int s = 0; s < 7; ++s) {
                    const int k = r * 7 + s;
                    const int icol = 2 * wc - 3 + s;
                    const float v = xr[s];
                    const float ind = (rin && icol >= 0 && icol < 64) ? 1.f : 0.f;
                    if (k & 1) { av[k >> 1] = pack_bf16(prev, v); iv[k >> 1] = pack_bf16(pin, ind); }
                    else { prev = v; pin = ind; }
                }
            }
            av[24] = pack_bf16(prev, 0.f); iv[24] = pack_bf16(pin, 0.f);      // k = 48 is the last tap
#pragma unroll
            for (int pc = 0; pc < 8; ++pc) {
                *reinterpret_cast<uint4*>(smem + A_OFF + swz128(t, pc)) = make_uint4(av[4 * pc], av[4 * pc + 1], av[4 * pc + 2], av[4 * pc + 3]);
                *reinterpret_cast<uint4*>(smem + A_OFF + 16384 + swz128(t, pc)) = make_uint4(iv[4 * pc], iv[4 * pc + 1], iv[4 * pc + 2], iv[4 * pc + 3]);
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        tc_fence_before();
        __syncthreads();
        if (t == 0) {
            if (tile == 0) mbar_wait(bar_w, 0);
            tc_fence_after();
            constexpr uint32_t idesc = make_idesc(128, 64);
            const uint32_t sa = smem_u32(smem + A_OFF), sb = smem_u32(smem + B_OFF);
#pragma unroll
            for (int c = 0; c < 2; ++c)
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_bf16(tmem_base, make_smem_desc(sa + c * 16384 + k * 32), make_smem_desc(sb + c * 8192 + k * 32), idesc,
                              (c > 0 || k > 0) ? 1u : 0u);
            umma_commit(bar_mma);
        }
        mbar_wait(bar_mma, tile & 1);
        tc_fence_after();
        // ---- epilogue: + BN bias, ReLU, bf16, park the conv pixel (64 channels = 128 bytes) ----
        {
            const int px = lr * 32 + wc;
            uint8_t* crow = smem + C_OFF + px * 128;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                uint32_t acc[32];
                tmem_ld32(lane_addr + half * 32, acc);
#pragma unroll
                for (int j8 = 0; j8 < 4; ++j8) {
                    uint32_t w[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int ch = half * 32 + j8 * 8 + 2 * q;
                        const float a = fmaxf(__uint_as_float(acc[j8 * 8 + 2 * q]) + sbias[ch], 0.f);
                        const float c = fmaxf(__uint_as_float(acc[j8 * 8 + 2 * q + 1]) + sbias[ch + 1], 0.f);
                        w[q] = pack_bf16(a, c);
                    }
                    *reinterpret_cast<uint4*>(crow + (((half * 4 + j8) ^ (px & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
                }
            }
        }
        tc_fence_before();
        __syncthreads();          // A tile and the TMEM accumulator may be overwritten by the next tile
        tc_fence_after();
    }

    // ---- 3x3 / stride 2 / pad 1 max-pool over the parked conv rows (values are >= 0 after ReLU) ----
    for (int i = t; i < ST_PH * 16 * 8; i += 128) {
        const int cg = i & 7, wp = (i >> 3) & 15, hl = i >> 7;
        const int hp = hp0 + hl;
        if (hp >= Hp) continue;
        if constexpr (kArgmax) {
            float best[8]; uint32_t arg[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) { best[q] = 0.f; arg[q] = 9u; }
#pragma unroll
            for (int dr = 0; dr < 3; ++dr) {
                const int hcc = 2 * hp - 1 + dr;
                if (hcc < 0 || hcc >= Hc) continue;
#pragma unroll
                for (int dc = 0; dc < 3; ++dc) {
                    const int wcc = 2 * wp - 1 + dc;
                    if (wcc < 0 || wcc >= 32) continue;
                    const int px = (2 * hl + dr) * 32 + wcc;
                    const uint4 u = *reinterpret_cast<const uint4*>(smem + C_OFF + px * 128 + ((cg ^ (px & 7)) << 4));
                    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[q]));
                        if (f.x > best[2 * q]) { best[2 * q] = f.x; arg[2 * q] = (uint32_t)(dr * 3 + dc); }
                        if (f.y > best[2 * q + 1]) { best[2 * q + 1] = f.y; arg[2 * q + 1] = (uint32_t)(dr * 3 + dc); }
                    }
                }
            }
            const size_t e = (((size_t)b * Hp + hp) * 16 + wp) * 64 + cg * 8;
            uint4 o;
            o.x = pack_bf16(best[0], best[1]); o.y = pack_bf16(best[2], best[3]);
            o.z = pack_bf16(best[4], best[5]); o.w = pack_bf16(best[6], best[7]);
            *reinterpret_cast<uint4*>(out + e) = o;
            uint2 a;
            a.x = arg[0] | (arg[1] << 8) | (arg[2] << 16) | (arg[3] << 24);
            a.y = arg[4] | (arg[5] << 8) | (arg[6] << 16) | (arg[7] << 24);
            *reinterpret_cast<uint2*>(amax + e) = a;
            continue;
        }
        __nv_bfloat162 m[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) m[q] = __floats2bfloat162_rn(0.f, 0.f);
#pragma unroll
        for (int dr = 0; dr < 3; ++dr) {
            const int hcc = 2 * hp - 1 + dr;
            if (hcc < 0 || hcc >= Hc) continue;
#pragma unroll
            for (int dc = 0; dc < 3; ++dc) {
                const int wcc = 2 * wp - 1 + dc;
                if (wcc < 0 || wcc >= 32) continue;
                const int px = (2 * hl + dr) * 32 + wcc;
                const uint4 u = *reinterpret_cast<const uint4*>(smem + C_OFF + px * 128 + ((cg ^ (px & 7)) << 4));
                const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                for (int q = 0; q < 4; ++q) m[q] = __hmax2(m[q], *reinterpret_cast<const __nv_bfloat162*>(&w[q]));
            }
        }
        uint4 o;
        o.x = *reinterpret_cast<uint32_t*>(&m[0]); o.y = *reinterpret_cast<uint32_t*>(&m[1]);
        o.z = *reinterpret_cast<uint32_t*>(&m[2]); o.w = *reinterpret_cast<uint32_t*>(&m[3]);
        *reinterpret_cast<uint4*>(out + (((size_t)b * Hp + hp) * 16 + wp) * 64 + cg * 8) = o;
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc<64>(tmem_base);
    }
}

// Pre-swizzled B image: chunk 0 = conv1 folded with conv0.weight, chunk 1 = conv1 folded with conv0.bias,
// both scaled by the FrozenBN scale.  One thread per output channel.
__global__ void stem_tc_pack_kernel(const float* __restrict__ w0, const float* __restrict__ b0, const float* __restrict__ w1,
                                    const float* __restrict__ scale, uint8_t* __restrict__ wtc)
{
    const int o = threadIdx.x;
    if (o >= 64) return;
    const float sc = scale[o];
    for (int k = 0; k < 64; ++k) {
        float we = 0.f, wb = 0.f;
        if (k < 49) {
            for (int c = 0; c < 3; ++c) {
                const float v = w1[(o * 3 + c) * 49 + k];
                we = fmaf(v, w0[c], we);
                wb = fmaf(v, b0[c], wb);
            }
        }
        const uint32_t off = (uint32_t)(o * 128 + (((k >> 3) ^ (o & 7)) << 4) + (k & 7) * 2);
        *reinterpret_cast<__nv_bfloat16*>(wtc + off) = __float2bfloat16_rn(we * sc);
        *reinterpret_cast<__nv_bfloat16*>(wtc + 8192 + off) = __float2bfloat16_rn(wb * sc);
    }
}

}  // namespace

int launch_stem_tc_pack(const float* conv0_w, const float* conv0_b, const float* conv1_w, const float* bn_scale, void* wtc,
                        cudaStream_t stream)
{
    stem_tc_pack_kernel<<<1, 64, 0, stream>>>(conv0_w, conv0_b, conv1_w, bn_scale, (uint8_t*)wtc);
    SEDT_COUNT_LAUNCH();
    SEDT_CHECK_CUDA(cudaGetLastError());
    return SEDT_OK;
}

int launch_stem_tc(const float* x, const void* wtc, const float* bn_bias, void* out, int B, int T, int F, cudaStream_t stream,
                   uint8_t* amax)
{
    SEDT_REQUIRE(F == 64, "stem: the fused stem kernel needs 64 mel bins (config.py n_mels), got F=%d", F);
    SEDT_REQUIRE(((uintptr_t)wtc & 15) == 0, "stem_tc: weight image must be 16-byte aligned");
    if (B == 0) return SEDT_OK;
    const int Hc = (T - 1) / 2 + 1, Hp = (Hc - 1) / 2 + 1;
    static bool attr_set = false;
    if (!attr_set) {
        SEDT_CHECK_CUDA(cudaFuncSetAttribute(stem_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ST_SMEM));
        SEDT_CHECK_CUDA(cudaFuncSetAttribute(stem_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ST_SMEM));
        attr_set = true;
    }
    dim3 grid((unsigned)ceil_div(Hp, ST_PH), (unsigned)B), block(128);
    ProfScope _prof(PROF_STEM, stream);
    if (amax != nullptr) {
        SEDT_REQUIRE(((uintptr_t)amax & 7) == 0, "stem_tc: arg-max buffer must be 8-byte aligned");
        SEDT_CHECK_CUDA(launch_pdl(stem_tc_kernel<true>, grid, block, ST_SMEM, stream, 1, x, (const uint8_t*)wtc, bn_bias,
                                   (__nv_bfloat16*)out, amax, T, Hc, Hp));
    } else {
        SEDT_CHECK_CUDA(launch_pdl(stem_tc_kernel<false>, grid, block, ST_SMEM, stream, 1, x, (const uint8_t*)wtc, bn_bias,
                                   (__nv_bfloat16*)out, (uint8_t*)nullptr, T, Hc, Hp));
    }
    SEDT_COUNT_LAUNCH();
    SEDT_CHECK_CUDA(cudaGetLastError());
    return SEDT_OK;
}

}  // namespace sedt
