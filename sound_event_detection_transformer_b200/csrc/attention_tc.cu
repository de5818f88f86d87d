// Fused multi-head attention core on tcgen05 / TMEM for sequences of up to 128 tokens
// (encoder: S = 124 / 128 tokens; decoder: 11 / 21 queries x 21 / 124 keys), head_dim 32.
//
// One CTA of 128 threads per (clip, head); thread t owns query row t = TMEM lane t.
//   1. Q_h, K_h rows are copied into shared memory in the 128B-swizzled K-major operand layout
//      (head_dim 32 fills the first 64 bytes of each 128-byte row), V_h is transposed on the fly
//      into [head_dim][keys] so that it is a K-major B operand as well.
//   2. S = Q K^T       one 128x128x32 tcgen05.mma chain, fp32 accumulator in TMEM.
//   3. softmax         two passes over the thread's own TMEM lane (max, then exp2 / sum), masks
//      (additive attn_mask, key_padding_mask, keys >= Lk) applied in fp32; P is written as bf16
//      into shared memory in A-operand layout.
//   4. O = P V         128x32x128 tcgen05.mma chain into the same TMEM columns, scaled by 1/sum
//      and stored as bf16.
// Replaces the need_weights=True eager path of nn.MultiheadAttention
// (torch/nn/functional.py:6630-6659: q*sqrt(1/hd), baddbmm, softmax, bmm).
#include "tc_common.cuh"
#include "dropout.cuh"
#include <math_constants.h>

namespace sedt {
namespace {

using namespace tc;

constexpr int ATT_THREADS = 128;
constexpr int HD = 32;
constexpr int SQ_OFF = 0;                   // [128 rows][128 B]
constexpr int SK_OFF = 16384;
constexpr int SP_OFF = 0;                   // 2 chunks of [128 rows][64 keys]; reuses Q/K once S = QK^T has retired
constexpr int SV_OFF = 32768;               // 2 chunks of [32 rows][64 keys]
constexpr int SM_OFF = 32768 + 8192;        // key mask as float [128]
constexpr int SN_OFF = SM_OFF + 512;         // 0 / -1e30 per key (keeps masked keys out of the row maximum)
constexpr int BAR_OFF = SN_OFF + 512;
constexpr int ATT_SMEM = BAR_OFF + 64 + 1024;

__device__ __forceinline__ uint32_t swz(int row, int byte_in_row) {
    // 128-byte swizzle: 16-byte piece index XOR (row mod 8)
    return (uint32_t)(row * 128 + ((((byte_in_row >> 4) ^ (row & 7)) << 4) | (byte_in_row & 15)));
}

__device__ __forceinline__ uint32_t pack2(float a, float b) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}

// DROP (training forward): the attention weights are dropped (nn.MultiheadAttention dropout, functional.py:6650):
// O = sum_j (p_j * keep_j / (1 - p)) v_j with p_j normalised over ALL valid keys.
template <bool HAS_AMASK, bool DROP>
__global__ void __launch_bounds__(ATT_THREADS)
attention_tc_kernel(const __nv_bfloat16* __restrict__ Q, int ldq, const __nv_bfloat16* __restrict__ K, int ldk,
                    const __nv_bfloat16* __restrict__ V, int ldv, __nv_bfloat16* __restrict__ O, int ldo,
                    const uint8_t* __restrict__ kpm, const float* __restrict__ amask, int Lq, int Lk, float scale, DropSite drop)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // stays a __shared__ pointer (LDS/STS)
    float* s_mask = (float*)(smem + SM_OFF);
    float* s_neg = (float*)(smem + SN_OFF);
    uint64_t* bar_s = (uint64_t*)(smem + BAR_OFF);
    uint64_t* bar_o = bar_s + 1;
    uint32_t* tmem_slot = (uint32_t*)(bar_o + 1);

    const int t = threadIdx.x, warp = t >> 5;
    const int h = blockIdx.x, b = blockIdx.y;

    if (t == 0) {
        mbar_init(bar_s, 1); mbar_init(bar_o, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc<128>(tmem_slot);
    pdl_wait();            // Q, K, V come from the preceding projection kernels

    // ---- stage Q, K (row t) and V^T (column t) ----------------------------------------------
    {
        const uint4 z = make_uint4(0, 0, 0, 0);
        uint4 q4[4] = {z, z, z, z}, k4[4] = {z, z, z, z}, v4[4] = {z, z, z, z};
        if (t < Lq) {
            const uint4* src = reinterpret_cast<const uint4*>(Q + ((size_t)b * Lq + t) * ldq + h * HD);
#pragma unroll
            for (int j = 0; j < 4; ++j) q4[j] = src[j];
        }
        if (t < Lk) {
            const uint4* ks = reinterpret_cast<const uint4*>(K + ((size_t)b * Lk + t) * ldk + h * HD);
            const uint4* vs = reinterpret_cast<const uint4*>(V + ((size_t)b * Lk + t) * ldv + h * HD);
#pragma unroll
            for (int j = 0; j < 4; ++j) { k4[j] = ks[j]; v4[j] = vs[j]; }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            *reinterpret_cast<uint4*>(smem + SQ_OFF + swz(t, j * 16)) = j < 4 ? q4[j] : z;
            *reinterpret_cast<uint4*>(smem + SK_OFF + swz(t, j * 16)) = j < 4 ? k4[j] : z;
        }
        // V^T: element (d, key t) -> chunk t/64, row d, column t%64
        const __nv_bfloat16* vh = reinterpret_cast<const __nv_bfloat16*>(v4);
        uint8_t* vbase = smem + SV_OFF + (t >> 6) * 4096;
        const int col_b = (t & 63) * 2;
#pragma unroll
        for (int d = 0; d < HD; ++d) *reinterpret_cast<__nv_bfloat16*>(vbase + swz(d, col_b)) = vh[d];
        float mk = 1.f;                                    // validity of key t
        if (t >= Lk) mk = 0.f;
        else if (kpm != nullptr && kpm[(size_t)b * Lk + t]) mk = 0.f;
        s_mask[t] = mk;
        s_neg[t] = (mk - 1.f) * 1e30f;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    pdl_trigger();         // TMEM is owned: a successor scheduled next to this CTA cannot starve it
    const uint32_t tmem_base = *tmem_slot;

    // ---- S = Q K^T -----------------------------------------------------------------------------
    if (t == 0) {
        constexpr uint32_t idesc = make_idesc(128, 128);
        const uint32_t sq = smem_u32(smem + SQ_OFF), sk = smem_u32(smem + SK_OFF);
#pragma unroll
        for (int k = 0; k < HD / UMMA_K; ++k)
            umma_bf16(tmem_base, make_smem_desc(sq + k * UMMA_K * 2), make_smem_desc(sk + k * UMMA_K * 2), idesc, k > 0 ? 1u : 0u);
        umma_commit(bar_s);
    }
    mbar_wait(bar_s, 0);
    tc_fence_after();

    // ---- softmax over this thread's row -------------------------------------------------------
    // p_j = 2^(c*(s_j + a_j) - c*m) * valid_j with c = scale*log2(e).  The row maximum m is taken over
    // the scores of the valid keys;
    // masked keys are kept out of the maximum by a 0 / -1e30 vector and zeroed by a 0 / 1 validity vector, both
    // in shared memory.  Only key chunks that hold real keys, and only warps that
    // own real query rows, do any work (decoder: 11 / 21 queries, 21 keys in self-attention).
    const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
    const int nkc = (Lk + 31) >> 5;
    const bool row_warp = warp * 32 < Lq;
    const float cs = scale * 1.4426950408889634f;
    const float* arow = nullptr;
    if (HAS_AMASK) arow = amask + (size_t)min(t, Lq - 1) * Lk;
    float m = -CUDART_INF_F;
#pragma unroll 1
    for (int c = 0; c < (row_warp ? nkc : 0); ++c) {
        uint32_t acc[32];
        tmem_ld32(lane_addr + c * 32, acc);
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
            const float4 ng = *reinterpret_cast<const float4*>(s_neg + c * 32 + j4 * 4);
            const float ngv[4] = {ng.x, ng.y, ng.z, ng.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int j = j4 * 4 + q;
                float sv = __uint_as_float(acc[j]) + ngv[q];
                if (HAS_AMASK) { const int key = c * 32 + j; sv += key < Lk ? fmaxf(arow[key], -1e30f) / scale : 0.f; }
                m = fmaxf(m, sv);
            }
        }
    }
    const float mc = m * cs;
    float l = 0.f;
    unsigned long long d_seed = 0ull, d_step = 0ull;
    if (DROP) { d_seed = drop.state[0]; d_step = drop.state[1]; }
    const unsigned long long d_row = ((unsigned long long)(blockIdx.y * gridDim.x + blockIdx.x) * 128ull + (unsigned long long)t) * 32ull;
#pragma unroll 1
    for (int c = 0; c < (row_warp ? nkc : 0); ++c) {
        uint32_t acc[32];
        tmem_ld32(lane_addr + c * 32, acc);
        uint32_t packed[16];
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
            const float4 vm = *reinterpret_cast<const float4*>(s_mask + c * 32 + j4 * 4);      // 1 = real key, 0 = masked
            const float vmv[4] = {vm.x, vm.y, vm.z, vm.w};
            float pv[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int j = j4 * 4 + q;
                float sv = __uint_as_float(acc[j]);
                if (HAS_AMASK) { const int key = c * 32 + j; sv += key < Lk ? fmaxf(arow[key], -1e30f) / scale : 0.f; }
                float e;
                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fmaf(sv, cs, -mc)));
                pv[q] = e * vmv[q];
                l += pv[q];
            }
            if (DROP) {
                const uint4 r = drop_draw4(drop, d_seed, d_step, d_row + (unsigned long long)(c * 8 + j4));
                pv[0] = r.x < drop.thresh ? pv[0] * drop.inv_keep : 0.f;
                pv[1] = r.y < drop.thresh ? pv[1] * drop.inv_keep : 0.f;
                pv[2] = r.z < drop.thresh ? pv[2] * drop.inv_keep : 0.f;
                pv[3] = r.w < drop.thresh ? pv[3] * drop.inv_keep : 0.f;
            }
            packed[j4 * 2] = pack2(pv[0], pv[1]);
            packed[j4 * 2 + 1] = pack2(pv[2], pv[3]);
        }
        // 32 keys = 64 bytes = pieces (c&1)*4 .. +3 of row t in P chunk c/2
        uint8_t* prow = smem + SP_OFF + (c >> 1) * 16384;
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4)
            *reinterpret_cast<uint4*>(prow + swz(t, ((c & 1) * 4 + j4) * 16)) =
                make_uint4(packed[4 * j4], packed[4 * j4 + 1], packed[4 * j4 + 2], packed[4 * j4 + 3]);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    // ---- O = P V (accumulates into TMEM columns 0..31, S is dead) ------------------------------
    if (t == 0) {
        constexpr uint32_t idesc = make_idesc(128, HD);
        const int ksteps = (Lk + UMMA_K - 1) / UMMA_K;          // P is exactly 0 beyond Lk inside a processed chunk
        for (int ks = 0; ks < ksteps; ++ks) {
            const int c = ks >> 2, k = ks & 3;
            const uint32_t sp = smem_u32(smem + SP_OFF + c * 16384), sv = smem_u32(smem + SV_OFF + c * 4096);
            umma_bf16(tmem_base, make_smem_desc(sp + k * UMMA_K * 2), make_smem_desc(sv + k * UMMA_K * 2), idesc, ks > 0 ? 1u : 0u);
        }
        umma_commit(bar_o);
    }
    mbar_wait(bar_o, 0);
    tc_fence_after();
    {
        uint32_t acc[32];
        tmem_ld32(lane_addr, acc);
        if (t < Lq) {
            const float inv = l > 0.f ? 1.f / l : 0.f;
            uint32_t w[16];
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
                const __nv_bfloat162 ho = __floats2bfloat162_rn(__uint_as_float(acc[j]) * inv, __uint_as_float(acc[j + 1]) * inv);
                w[j >> 1] = *reinterpret_cast<const uint32_t*>(&ho);
            }
            uint4* dst = reinterpret_cast<uint4*>(O + ((size_t)b * Lq + t) * ldo + h * HD);
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) dst[j4] = make_uint4(w[4 * j4], w[4 * j4 + 1], w[4 * j4 + 2], w[4 * j4 + 3]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc<128>(tmem_base);
    }
}

}  // namespace

bool attention_tc_supported(const void* Q, int ldq, const void* K, int ldk, const void* V, int ldv, const void* O, int ldo,
                            int dt, int nheads, int Lq, int Lk)
{
    if (dt != DT_BF16 || Lq < 1 || Lk < 1 || Lq > 128 || Lk > 128 || nheads < 1) return false;
    if ((ldq % 8) || (ldk % 8) || (ldv % 8) || (ldo % 8)) return false;
    if (((uintptr_t)Q & 15) || ((uintptr_t)K & 15) || ((uintptr_t)V & 15) || ((uintptr_t)O & 15)) return false;
    return true;
}

template <bool AM, bool DR>
static int launch_att_variant(const void* Q, int ldq, const void* K, int ldk, const void* V, int ldv, void* O, int ldo,
                              const uint8_t* kpm, const float* amask, int B, int nheads, int Lq, int Lk, float scale,
                              const DropSite& drop, cudaStream_t stream)
{
    static bool attr_set = false;
    if (!attr_set) {
        SEDT_CHECK_CUDA(cudaFuncSetAttribute(attention_tc_kernel<AM, DR>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
        attr_set = true;
    }
    dim3 grid((unsigned)nheads, (unsigned)B), block(ATT_THREADS);
    ProfScope _prof(PROF_ATTENTION, stream);
    SEDT_CHECK_CUDA(launch_pdl(attention_tc_kernel<AM, DR>, grid, block, ATT_SMEM, stream, 1, (const __nv_bfloat16*)Q, ldq,
                               (const __nv_bfloat16*)K, ldk, (const __nv_bfloat16*)V, ldv, (__nv_bfloat16*)O, ldo, kpm, amask, Lq, Lk,
                               scale, drop));
    SEDT_COUNT_KIND(KK_ATTENTION_TC);
    SEDT_CHECK_CUDA(cudaGetLastError());
    return SEDT_OK;
}

int launch_attention_tc(const void* Q, int ldq, const void* K, int ldk, const void* V, int ldv, void* O, int ldo,
                        const uint8_t* kpm, const float* amask, int B, int nheads, int Lq, int Lk, float scale,
                        cudaStream_t stream)
{
    const DropSite none = make_drop_site(nullptr, 0, 0.f);
    if (amask != nullptr) return launch_att_variant<true, false>(Q, ldq, K, ldk, V, ldv, O, ldo, kpm, amask, B, nheads, Lq, Lk, scale, none, stream);
    return launch_att_variant<false, false>(Q, ldq, K, ldk, V, ldv, O, ldo, kpm, amask, B, nheads, Lq, Lk, scale, none, stream);
}

// training forward with dropout on the attention weights
int launch_attention_tc_drop(const void* Q, int ldq, const void* K, int ldk, const void* V, int ldv, void* O, int ldo,
                             const uint8_t* kpm, const float* amask, int B, int nheads, int Lq, int Lk, float scale,
                             const DropSite& drop, cudaStream_t stream)
{
    SEDT_REQUIRE(drop.state != nullptr, "attention_tc_drop: RNG state missing");
    if (amask != nullptr) return launch_att_variant<true, true>(Q, ldq, K, ldk, V, ldv, O, ldo, kpm, amask, B, nheads, Lq, Lk, scale, drop, stream);
    return launch_att_variant<false, true>(Q, ldq, K, ldk, V, ldv, O, ldo, kpm, amask, B, nheads, Lq, Lk, scale, drop, stream);
}

}  // namespace sedt
