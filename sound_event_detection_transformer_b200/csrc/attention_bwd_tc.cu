// Backward of the attention core on tcgen05 / TMEM (bf16, head_dim 32, Lq, Lk <= 128): one CTA of 128
// threads per (clip, head), thread t = query row t = key row t = TMEM lane t.
//
//   S  = Q K^T, dP = dO V^T                 two 128x128x32 MMA chains (K-major operands)
//   P  = softmax(scale*S + masks)           recomputed in registers from TMEM (no P is saved by the forward)
//   dS = P o (dP - rowsum(P o dP))          P and dS written to shared memory as bf16 operand tiles
//   dV = P^T dO, dK = scale * dS^T Q        the SAME tiles read as MN-major operands (rows = reduction index)
//   dQ = scale * dS K                       dS K-major, K MN-major
//
// Q, K, V, dO rows are staged once into 128-byte-swizzled [128][64] tiles (head_dim 32 fills the first
// 64 bytes of a row, the rest is zero), which serve both as K-major operands (reduction over head_dim)
// and as MN-major operands with N = 64 (reduction over the rows).
// Replaces autograd's backward of torch/nn/functional.py:6630-6659 (baddbmm, softmax, bmm).
#include "tc_common.cuh"
#include "dropout.cuh"
#include <math_constants.h>

namespace sedt {
namespace {

using namespace tc;
using bf16 = __nv_bfloat16;

constexpr int AB_THREADS = 128;
constexpr int HD = 32;
constexpr int TILE = 16384;                  // [128 rows][128 B]
constexpr int Q_OFF = 0, K_OFF = TILE, V_OFF = 2 * TILE, G_OFF = 3 * TILE;
constexpr int P_OFF = 4 * TILE;              // 2 chunks of [128 queries][64 keys]
constexpr int S_OFF = 6 * TILE;              // dS, same shape
constexpr int M_OFF = 8 * TILE;              // key validity 0 / 1 as float [128]
constexpr int N_OFF = M_OFF + 512;           // 0 / -1e30 per key
constexpr int BAR_OFF2 = N_OFF + 512;
constexpr int AB_SMEM_TC = BAR_OFF2 + 64 + 1024;

// TMEM columns
constexpr int C_S = 0, C_DP = 128, C_DV = 256, C_DK = 320, C_DQ = 384;

__device__ __forceinline__ uint32_t swz(int row, int byte_in_row) {
    return (uint32_t)(row * 128 + ((((byte_in_row >> 4) ^ (row & 7)) << 4) | (byte_in_row & 15)));
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}
// MN-major SWIZZLE_128B descriptor: 64 contiguous MN elements per 128-byte row, rows = reduction index,
// 8-row groups 1024 B apart (SBO), next 64-element MN chunk lbo bytes away
__device__ __forceinline__ uint64_t desc_mn(uint32_t saddr, uint32_t lbo_bytes) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

// DROP: the forward dropped the attention weights (P_drop = P o keep / (1-p) fed the P V product), so
// dP_eff = dP o keep / (1-p), dS = P o (dP_eff - rowsum(P o dP_eff)), dV = P_drop^T dO.
template <bool HAS_AMASK, bool DROP>
__global__ void __launch_bounds__(AB_THREADS)
attention_bwd_tc_kernel(const bf16* __restrict__ Q, int ldq, const bf16* __restrict__ K, int ldk, const bf16* __restrict__ V, int ldv,
                        const bf16* __restrict__ dO, int ldo, bf16* __restrict__ dQ, int lddq, bf16* __restrict__ dK, int lddk,
                        bf16* __restrict__ dV, int lddv, const uint8_t* __restrict__ kpm, const float* __restrict__ amask, int Lq,
                        int Lk, float scale, DropSite drop)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    float* s_mask = (float*)(smem + M_OFF);
    float* s_neg = (float*)(smem + N_OFF);
    uint64_t* bar1 = (uint64_t*)(smem + BAR_OFF2);
    uint64_t* bar2 = bar1 + 1;
    uint32_t* tmem_slot = (uint32_t*)(bar2 + 1);

    const int t = threadIdx.x, warp = t >> 5;
    const int h = blockIdx.x, b = blockIdx.y;

    if (t == 0) {
        mbar_init(bar1, 1); mbar_init(bar2, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc<512>(tmem_slot);

    // ---- stage row t of Q, dO (queries) and K, V (keys); the upper 64 bytes of every row are zero ----
    {
        const uint4 z = make_uint4(0, 0, 0, 0);
        uint4 q4[4] = {z, z, z, z}, g4[4] = {z, z, z, z}, k4[4] = {z, z, z, z}, v4[4] = {z, z, z, z};
        if (t < Lq) {
            const uint4* qs = reinterpret_cast<const uint4*>(Q + ((size_t)b * Lq + t) * ldq + h * HD);
            const uint4* gs = reinterpret_cast<const uint4*>(dO + ((size_t)b * Lq + t) * ldo + h * HD);
#pragma unroll
            for (int j = 0; j < 4; ++j) { q4[j] = qs[j]; g4[j] = gs[j]; }
        }
        if (t < Lk) {
            const uint4* ks = reinterpret_cast<const uint4*>(K + ((size_t)b * Lk + t) * ldk + h * HD);
            const uint4* vs = reinterpret_cast<const uint4*>(V + ((size_t)b * Lk + t) * ldv + h * HD);
#pragma unroll
            for (int j = 0; j < 4; ++j) { k4[j] = ks[j]; v4[j] = vs[j]; }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            *reinterpret_cast<uint4*>(smem + Q_OFF + swz(t, j * 16)) = j < 4 ? q4[j] : z;
            *reinterpret_cast<uint4*>(smem + G_OFF + swz(t, j * 16)) = j < 4 ? g4[j] : z;
            *reinterpret_cast<uint4*>(smem + K_OFF + swz(t, j * 16)) = j < 4 ? k4[j] : z;
            *reinterpret_cast<uint4*>(smem + V_OFF + swz(t, j * 16)) = j < 4 ? v4[j] : z;
        }
        float mk = 1.f;
        if (t >= Lk) mk = 0.f;
        else if (kpm != nullptr && kpm[(size_t)b * Lk + t]) mk = 0.f;
        s_mask[t] = mk;
        s_neg[t] = (mk - 1.f) * 1e30f;
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // ---- S = Q K^T and dP = dO V^T --------------------------------------------------------------------
    if (t == 0) {
        constexpr uint32_t idesc = make_idesc(128, 128);
        const uint32_t sq = smem_u32(smem + Q_OFF), sk = smem_u32(smem + K_OFF), sv = smem_u32(smem + V_OFF), sg = smem_u32(smem + G_OFF);
#pragma unroll
        for (int k = 0; k < HD / UMMA_K; ++k)
            umma_bf16(tmem_base + C_S, make_smem_desc(sq + k * UMMA_K * 2), make_smem_desc(sk + k * UMMA_K * 2), idesc, k > 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < HD / UMMA_K; ++k)
            umma_bf16(tmem_base + C_DP, make_smem_desc(sg + k * UMMA_K * 2), make_smem_desc(sv + k * UMMA_K * 2), idesc, k > 0 ? 1u : 0u);
        umma_commit(bar1);
    }
    mbar_wait(bar1, 0);
    tc_fence_after();

    // ---- P and dS for query row t ------------------------------------------------------------------------
    const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
    const int nkc = (Lk + 31) >> 5;
    const bool row_ok = t < Lq;
    const float cs = scale * 1.4426950408889634f;
    const float* arow = nullptr;
    if (HAS_AMASK) arow = amask + (size_t)min(t, Lq - 1) * Lk;
    auto score = [&](float raw, int key) -> float {          // raw score + additive mask, in units of the raw score
        if (HAS_AMASK) return raw + (key < Lk ? fmaxf(arow[key], -1e30f) / scale : 0.f);
        return raw;
    };
    float m = -CUDART_INF_F;
#pragma unroll 1
    for (int c = 0; c < nkc; ++c) {
        uint32_t acc[32];
        tmem_ld32(lane_addr + C_S + c * 32, acc);
#pragma unroll
        for (int j = 0; j < 32; ++j) m = fmaxf(m, score(__uint_as_float(acc[j]), c * 32 + j) + s_neg[c * 32 + j]);
    }
    const float mc = m * cs;
    float l = 0.f, pd = 0.f;
    unsigned long long d_seed = 0ull, d_step = 0ull;
    if (DROP) { d_seed = drop.state[0]; d_step = drop.state[1]; }
    const unsigned long long d_row = ((unsigned long long)(blockIdx.y * gridDim.x + blockIdx.x) * 128ull + (unsigned long long)t) * 32ull;
    // keep / (1-p) factors of keys c*32 + 4*g .. +3 of this row (same counters as the forward kernel)
    auto keep4 = [&](int c, int g, float (&k)[4]) {
        const uint4 r = drop_draw4(drop, d_seed, d_step, d_row + (unsigned long long)(c * 8 + g));
        k[0] = r.x < drop.thresh ? drop.inv_keep : 0.f; k[1] = r.y < drop.thresh ? drop.inv_keep : 0.f;
        k[2] = r.z < drop.thresh ? drop.inv_keep : 0.f; k[3] = r.w < drop.thresh ? drop.inv_keep : 0.f;
    };
#pragma unroll 1
    for (int c = 0; c < nkc; ++c) {
        uint32_t acc[32], dp[32];
        tmem_ld32_nowait(lane_addr + C_S + c * 32, acc);
        tmem_ld32_nowait(lane_addr + C_DP + c * 32, dp);
        tmem_ld_wait();
        float kq[4] = {1.f, 1.f, 1.f, 1.f};
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            float e;
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fmaf(score(__uint_as_float(acc[j]), c * 32 + j), cs, -mc)));
            e *= s_mask[c * 32 + j];
            l += e;
            float dpj = __uint_as_float(dp[j]);
            if (DROP) {
                float k4[4];
                if ((j & 3) == 0) { keep4(c, j >> 2, k4); kq[0] = k4[0]; kq[1] = k4[1]; kq[2] = k4[2]; kq[3] = k4[3]; }
                dpj *= kq[j & 3];
            }
            pd = fmaf(e, dpj, pd);
        }
    }
    const float inv = (row_ok && l > 0.f) ? 1.f / l : 0.f;       // rows beyond Lq contribute nothing
    const float delta = pd * inv;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
        uint32_t pk[16], dk[16];
        if (c < nkc) {
            uint32_t acc[32], dp[32];
            tmem_ld32_nowait(lane_addr + C_S + c * 32, acc);
            tmem_ld32_nowait(lane_addr + C_DP + c * 32, dp);
            tmem_ld_wait();
            float kq3[4] = {1.f, 1.f, 1.f, 1.f};
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
                float e0, e1;
                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(fmaf(score(__uint_as_float(acc[j]), c * 32 + j), cs, -mc)));
                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(fmaf(score(__uint_as_float(acc[j + 1]), c * 32 + j + 1), cs, -mc)));
                const float p0 = e0 * s_mask[c * 32 + j] * inv, p1 = e1 * s_mask[c * 32 + j + 1] * inv;
                float k0 = 1.f, k1 = 1.f;
                if (DROP) {
                    if ((j & 3) == 0) keep4(c, j >> 2, kq3);
                    k0 = kq3[j & 3]; k1 = kq3[(j & 3) + 1];
                }
                pk[j >> 1] = pack2(p0 * k0, p1 * k1);
                dk[j >> 1] = pack2(p0 * (__uint_as_float(dp[j]) * k0 - delta), p1 * (__uint_as_float(dp[j + 1]) * k1 - delta));
            }
        } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) { pk[j] = 0u; dk[j] = 0u; }
        }
        // 32 keys = 64 bytes = pieces (c&1)*4 .. +3 of row t in chunk c/2
        uint8_t* prow = smem + P_OFF + (c >> 1) * TILE;
        uint8_t* srow = smem + S_OFF + (c >> 1) * TILE;
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
            const uint32_t o = swz(t, ((c & 1) * 4 + j4) * 16);
            *reinterpret_cast<uint4*>(prow + o) = make_uint4(pk[4 * j4], pk[4 * j4 + 1], pk[4 * j4 + 2], pk[4 * j4 + 3]);
            *reinterpret_cast<uint4*>(srow + o) = make_uint4(dk[4 * j4], dk[4 * j4 + 1], dk[4 * j4 + 2], dk[4 * j4 + 3]);
        }
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    // ---- dV = P^T dO, dK = dS^T Q (A MN-major over the keys, reduction over the queries), dQ = dS K ------
    if (t == 0) {
        const uint32_t sp = smem_u32(smem + P_OFF), ss = smem_u32(smem + S_OFF), sq = smem_u32(smem + Q_OFF),
                       sk = smem_u32(smem + K_OFF), sg = smem_u32(smem + G_OFF);
        constexpr uint32_t id_mn_mn = make_idesc(128, 64) | (1u << 15) | (1u << 16);
        constexpr uint32_t id_k_mn = make_idesc(128, 64) | (1u << 16);
        const int qsteps = (Lq + UMMA_K - 1) / UMMA_K;           // reduction over the queries
        for (int ks = 0; ks < qsteps; ++ks) {
            const uint32_t ro = (uint32_t)ks * UMMA_K * 128;      // 16 rows further down
            umma_bf16(tmem_base + C_DV, desc_mn(sp + ro, TILE), desc_mn(sg + ro, TILE), id_mn_mn, ks > 0 ? 1u : 0u);
            umma_bf16(tmem_base + C_DK, desc_mn(ss + ro, TILE), desc_mn(sq + ro, TILE), id_mn_mn, ks > 0 ? 1u : 0u);
        }
        const int ksteps = (Lk + UMMA_K - 1) / UMMA_K;           // reduction over the keys
        for (int ks = 0; ks < ksteps; ++ks) {
            const uint32_t a = ss + (uint32_t)(ks >> 2) * TILE + (uint32_t)(ks & 3) * UMMA_K * 2;
            umma_bf16(tmem_base + C_DQ, make_smem_desc(a), desc_mn(sk + (uint32_t)ks * UMMA_K * 128, TILE), id_k_mn, ks > 0 ? 1u : 0u);
        }
        umma_commit(bar2);
    }
    mbar_wait(bar2, 0);
    tc_fence_after();

    // the tcgen05.ld are warp-collective: every thread loads, only valid rows store
    {
        uint32_t acc[32];
        tmem_ld32(lane_addr + C_DQ, acc);
        if (t < Lq) {
            uint32_t w[16];
#pragma unroll
            for (int j = 0; j < 32; j += 2) w[j >> 1] = pack2(__uint_as_float(acc[j]) * scale, __uint_as_float(acc[j + 1]) * scale);
            uint4* d4 = reinterpret_cast<uint4*>(dQ + ((size_t)b * Lq + t) * lddq + h * HD);
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) d4[j4] = make_uint4(w[4 * j4], w[4 * j4 + 1], w[4 * j4 + 2], w[4 * j4 + 3]);
        }
        tmem_ld32(lane_addr + C_DK, acc);
        if (t < Lk) {
            uint32_t w[16];
#pragma unroll
            for (int j = 0; j < 32; j += 2) w[j >> 1] = pack2(__uint_as_float(acc[j]) * scale, __uint_as_float(acc[j + 1]) * scale);
            uint4* d4 = reinterpret_cast<uint4*>(dK + ((size_t)b * Lk + t) * lddk + h * HD);
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) d4[j4] = make_uint4(w[4 * j4], w[4 * j4 + 1], w[4 * j4 + 2], w[4 * j4 + 3]);
        }
        tmem_ld32(lane_addr + C_DV, acc);
        if (t < Lk) {
            uint32_t w[16];
#pragma unroll
            for (int j = 0; j < 32; j += 2) w[j >> 1] = pack2(__uint_as_float(acc[j]), __uint_as_float(acc[j + 1]));
            uint4* d4 = reinterpret_cast<uint4*>(dV + ((size_t)b * Lk + t) * lddv + h * HD);
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) d4[j4] = make_uint4(w[4 * j4], w[4 * j4 + 1], w[4 * j4 + 2], w[4 * j4 + 3]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

}  // namespace

bool attention_bwd_tc_supported(const void* Q, int ldq, const void* K, int ldk, const void* V, int ldv, const void* dO, int ldo,
                                const void* dQ, int lddq, const void* dK, int lddk, const void* dV, int lddv, int Lq, int Lk)
{
    if (Lq < 1 || Lk < 1 || Lq > 128 || Lk > 128) return false;
    const int lds[7] = {ldq, ldk, ldv, ldo, lddq, lddk, lddv};
    for (int v : lds) if (v % 8) return false;
    const void* ps[7] = {Q, K, V, dO, dQ, dK, dV};
    for (const void* p : ps) if ((uintptr_t)p & 15) return false;
    return true;
}

template <bool AM, bool DR>
static int launch_ab_variant(const void* Q, int ldq, const void* K, int ldk, const void* V, int ldv, const void* dO, int ldo,
                             void* dQ, int lddq, void* dK, int lddk, void* dV, int lddv, const uint8_t* kpm, const float* amask,
                             int B, int nheads, int Lq, int Lk, float scale, const DropSite& drop, cudaStream_t stream)
{
    static bool attr_set = false;
    if (!attr_set) {
        SEDT_CHECK_CUDA(cudaFuncSetAttribute(attention_bwd_tc_kernel<AM, DR>, cudaFuncAttributeMaxDynamicSharedMemorySize, AB_SMEM_TC));
        attr_set = true;
    }
    dim3 grid((unsigned)nheads, (unsigned)B), block(AB_THREADS);
    ProfScope _prof(PROF_ATTENTION, stream);
    attention_bwd_tc_kernel<AM, DR><<<grid, block, AB_SMEM_TC, stream>>>(
        (const bf16*)Q, ldq, (const bf16*)K, ldk, (const bf16*)V, ldv, (const bf16*)dO, ldo, (bf16*)dQ, lddq, (bf16*)dK, lddk,
        (bf16*)dV, lddv, kpm, amask, Lq, Lk, scale, drop);
    SEDT_COUNT_KIND(KK_ATTENTION_BWD_TC);
    SEDT_CHECK_CUDA(cudaGetLastError());
    return SEDT_OK;
}

int launch_attention_bwd_tc(const void* Q, int ldq, const void* K, int ldk, const void* V, int ldv, const void* dO, int ldo,
                            void* dQ, int lddq, void* dK, int lddk, void* dV, int lddv, const uint8_t* kpm, const float* amask,
                            int B, int nheads, int Lq, int Lk, float scale, const DropSite* drop, cudaStream_t stream)
{
    if (B == 0) return SEDT_OK;
#define SEDT_AB_CALL(AM, DR, D) launch_ab_variant<AM, DR>(Q, ldq, K, ldk, V, ldv, dO, ldo, dQ, lddq, dK, lddk, dV, lddv, kpm, amask, B, \
                                                         nheads, Lq, Lk, scale, D, stream)
    if (drop != nullptr && drop->state != nullptr)
        return amask != nullptr ? SEDT_AB_CALL(true, true, *drop) : SEDT_AB_CALL(false, true, *drop);
    const DropSite none = make_drop_site(nullptr, 0, 0.f);
    return amask != nullptr ? SEDT_AB_CALL(true, false, none) : SEDT_AB_CALL(false, false, none);
#undef SEDT_AB_CALL
}

}  // namespace sedt
