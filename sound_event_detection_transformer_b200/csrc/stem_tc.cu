// Backbone stem on tcgen05 (bf16 tier): conv0 (1x1, bias) + conv1 (7x7 s2 p3) + FrozenBN + ReLU +
// maxpool (3x3 s2 p1) in one kernel.  Reference: sedt/backbone.py:97-111, torchvision resnet.py:266-272.
//
// conv0 folded into conv1 is a 1-input-channel 7x7 convolution (weights Weff = sum_c conv1[:, c] * w0[c]) plus a bias term
// that only counts the taps landing INSIDE the clip (conv1's zero padding comes after conv0's bias).  As a GEMM the
// convolution is K = 49, padded to 64: the A tile [128 conv pixels x 64] is built in shared memory by the CTA's threads
// (im2col of a staged fp32 patch, written directly in the 128B-swizzled K-major layout), B [64 channels x 64] is a
// pre-swizzled 8 KiB image produced at pack time (BN scale folded in) and fetched with one bulk copy.  The inside-the-clip
// bias depends only on the pixel's border class (row class x {col 0, col 1, col 31, interior}); it is evaluated in fp32
// from the 8x8 summed-area table of the bias taps (stem_pack_kernel) into a small shared-memory table per CTA and added
// in the epilogue together with the BN bias.  (The first version carried it as a second, inside-indicator input channel:
// K = 2 x 64, twice the im2col work and MMA time.)  D lives in TMEM; the epilogue applies ReLU and parks the conv tile in
// shared memory as bf16, from where the 3x3/s2 max-pool is taken.
//
// One CTA = 4 pooled rows of one clip = 9 conv rows = 2 M-tiles of 4 conv rows + 1 M-tile of which only the first row
// (32 pixels, one warp) is built and kept.  ~70 KiB of shared memory: three CTAs per SM.
#include "tc_common.cuh"

namespace sedt {
namespace {

using namespace tc;

constexpr int ST_PH = 4;                    // pooled rows per CTA
constexpr int ST_CR = 2 * ST_PH + 1;        // 9 conv rows
constexpr int ST_TILES = 3;
constexpr int ST_XR = 2 * (4 * ST_TILES) + 5;   // 29 input rows
constexpr int ST_XC = 72;
constexpr int ST_NCLS = 5;                  // bias tables: interior rows + up to 4 border rows per CTA
constexpr int A_OFF = 0;                    // [128 px][64 k] bf16 = 16 KiB
constexpr int B_OFF = 16384;                // [64 ch][64 k] bf16 = 8 KiB
constexpr int C_OFF = 24576;                // [9 conv rows][32 px][64 ch] bf16 = 36 KiB
constexpr int X_OFF = C_OFF + ST_CR * 32 * 128;                 // [29][72] fp32
constexpr int TAB_OFF = X_OFF + ST_XR * ST_XC * 4;              // [ST_NCLS][4 col classes][64] fp32 bias tables
constexpr int ROWCLS_OFF = TAB_OFF + ST_NCLS * 4 * 64 * 4;      // [12] int: bias table of each local conv row
constexpr int STBAR_OFF = ((ROWCLS_OFF + 64 + 15) / 16) * 16;
constexpr int ST_SMEM = STBAR_OFF + 64 + 1024;
static_assert(ST_SMEM * 3 <= 232448, "three stem CTAs per SM");

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
               const float* __restrict__ scale, const float* __restrict__ sat,
               __nv_bfloat16* __restrict__ out, uint8_t* __restrict__ amax, int T, int Hc, int Hp)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    float* xs = (float*)(smem + X_OFF);
    float* tab = (float*)(smem + TAB_OFF);
    int* rowcls = (int*)(smem + ROWCLS_OFF);
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
        mbar_expect_tx(bar_w, 8192);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(smem + B_OFF)), "l"(wtc), "r"(8192), "r"(smem_u32(bar_w)) : "memory");
    }
    if (warp == 0) tmem_alloc<64>(tmem_slot);

    // ---- stage the 29 x 64 input patch (columns -3..68 in xs, zero outside the clip): all loads of a thread are issued
    // before the first store (ncu: 35 % of the stall samples sat on a load-then-store loop here) ----
    const float* xb = x + (size_t)b * T * 64;
    {
        float4 v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int i = t + k * 128, ri = i >> 4, c4 = i & 15;
            const int row = row0 + ri;
            v[k] = (i < ST_XR * 16 && row >= 0 && row < T) ? __ldg(reinterpret_cast<const float4*>(xb + (size_t)row * 64) + c4)
                                                          : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        for (int i = t; i < ST_XR * 8; i += 128) {            // the 3 + 5 padding columns of every row
            const int ri = i >> 3, k = i & 7;
            xs[ri * ST_XC + (k < 3 ? k : 64 + k)] = 0.f;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int i = t + k * 128, ri = i >> 4, c4 = i & 15;
            if (i < ST_XR * 16) {
                float* d = xs + ri * ST_XC + 3 + c4 * 4;
                d[0] = v[k].x; d[1] = v[k].y; d[2] = v[k].z; d[3] = v[k].w;
            }
        }
    }
    // ---- bias tables: table 0 = rows whose 7 tap rows all lie inside the clip; border rows get their own ----
    // valid tap rows of conv row hc: [max(0, 3 - 2 hc), min(6, T + 2 - 2 hc)]; tap columns of conv column wc likewise
    // with 64 mel bins: wc = 0 -> [3, 6], wc = 1 -> [1, 6], wc = 31 -> [0, 4], else [0, 6]
    int ncls = 1;
    int cls_rlo[ST_NCLS], cls_rhi[ST_NCLS];
    cls_rlo[0] = 0; cls_rhi[0] = 6;
#pragma unroll 1
    for (int lr = 0; lr < ST_CR; ++lr) {
        const int hc = cr0 + lr;
        int cls = 0;
        if (hc >= 0 && hc < Hc) {
            const int rlo = max(0, 3 - 2 * hc), rhi = min(6, T + 2 - 2 * hc);
            if (rlo != 0 || rhi != 6) {
                cls = -1;
                for (int k = 1; k < ncls; ++k) if (cls_rlo[k] == rlo && cls_rhi[k] == rhi) cls = k;
                if (cls < 0 && ncls < ST_NCLS) { cls = ncls; cls_rlo[ncls] = rlo; cls_rhi[ncls] = rhi; ++ncls; }
                if (cls < 0) cls = 0;            // cannot happen: at most 2 top + 2 bottom border rows exist
            }
        }
        if (t == 0) rowcls[lr] = cls;
    }
    for (int i = t; i < ncls * 4 * 64; i += 128) {
        const int ch = i & 63, cc = (i >> 6) & 3, k = i >> 8;
        const int rlo = cls_rlo[k], rhi = cls_rhi[k];
        const int slo = cc == 1 ? 3 : (cc == 2 ? 1 : 0), shi = cc == 3 ? 4 : 6;
        // inclusive summed-area table with a zero border: S[r + 1][s + 1]
        const float bsum = __ldg(sat + ((rhi + 1) * 8 + (shi + 1)) * 64 + ch) - __ldg(sat + (rlo * 8 + (shi + 1)) * 64 + ch)
                         - __ldg(sat + ((rhi + 1) * 8 + slo) * 64 + ch) + __ldg(sat + (rlo * 8 + slo) * 64 + ch);
        tab[i] = fmaf(bsum, __ldg(scale + ch), __ldg(bias + ch));
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    pdl_trigger();         // TMEM is owned: a successor scheduled next to this CTA cannot starve it
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);

    const int wc = t & 31;
    const int cc = wc == 0 ? 1 : (wc == 1 ? 2 : (wc == 31 ? 3 : 0));
#pragma unroll 1
    for (int tile = 0; tile < ST_TILES; ++tile) {
        const int lr = tile * 4 + (t >> 5);      // local conv row of this thread's pixel
        const bool active = lr < ST_CR;          // the last tile only carries conv row 8 (warp 0)
        // ---- im2col row of pixel (cr0 + lr, wc): 49 log-mel taps, zero padded to 64 ----
        if (active) {
            uint32_t av[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) av[i] = 0u;
            float prev = 0.f;
#pragma unroll
            for (int r = 0; r < 7; ++r) {
                const float* xr = xs + (2 * lr + r) * ST_XC + 2 * wc;
#pragma unroll
                for (int s_ = 0; s_ < 7; ++s_) {
                    const int k = r * 7 + s_;
                    const float v = xr[s_];
                    if (k & 1) av[k >> 1] = pack_bf16(prev, v);
                    else prev = v;
                }
            }
            av[24] = pack_bf16(prev, 0.f);       // k = 48 is the last tap
#pragma unroll
            for (int pc = 0; pc < 8; ++pc)
                *reinterpret_cast<uint4*>(smem + A_OFF + swz128(t, pc)) = make_uint4(av[4 * pc], av[4 * pc + 1], av[4 * pc + 2], av[4 * pc + 3]);
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
            for (int k = 0; k < 4; ++k)
                umma_bf16(tmem_base, make_smem_desc(sa + k * 32), make_smem_desc(sb + k * 32), idesc, k > 0 ? 1u : 0u);
            umma_commit(bar_mma);
        }
        mbar_wait(bar_mma, tile & 1);
        tc_fence_after();
        // ---- epilogue: + (BN bias + inside-the-clip conv0 bias of this pixel's border class), ReLU, bf16, park the
        // conv pixel (64 channels = 128 bytes) ----
        if (active) {
            const int px = lr * 32 + wc;
            uint8_t* crow = smem + C_OFF + px * 128;
            const float* pb = tab + (rowcls[lr] * 4 + cc) * 64;
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
                        const float a = fmaxf(__uint_as_float(acc[j8 * 8 + 2 * q]) + pb[ch], 0.f);
                        const float c = fmaxf(__uint_as_float(acc[j8 * 8 + 2 * q + 1]) + pb[ch + 1], 0.f);
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

int launch_stem_tc(const float* x, const void* wtc, const float* bn_bias, const float* bn_scale, const float* sat, void* out, int B,
                   int T, int F, cudaStream_t stream, uint8_t* amax)
{
    SEDT_REQUIRE(bn_scale != nullptr && sat != nullptr, "stem_tc: needs the BN scale and the bias summed-area table");
    SEDT_REQUIRE(T >= 1, "stem_tc: T=%d", T);
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
        SEDT_CHECK_CUDA(launch_pdl(stem_tc_kernel<true>, grid, block, ST_SMEM, stream, 1, x, (const uint8_t*)wtc, bn_bias, bn_scale, sat,
                                   (__nv_bfloat16*)out, amax, T, Hc, Hp));
    } else {
        SEDT_CHECK_CUDA(launch_pdl(stem_tc_kernel<false>, grid, block, ST_SMEM, stream, 1, x, (const uint8_t*)wtc, bn_bias, bn_scale, sat,
                                   (__nv_bfloat16*)out, (uint8_t*)nullptr, T, Hc, Hp));
    }
    SEDT_COUNT_KIND(KK_STEM_TC);
    SEDT_CHECK_CUDA(cudaGetLastError());
    return SEDT_OK;
}

}  // namespace sedt
