// Fused encoder self-attention block (eval):   x += out_proj( MHA( q = k = (LN(x)+pos) Wq/Wk, v = LN(x) Wv ) )
// (sedt/transformer.py:192-198 forward_pre, torch/nn/functional.py:5833-5867 in-proj split, :6630-6659 attention core.)
//
// One persistent CTA per SM walks over the clips; a clip is ONE 128-row tile (S = H*W <= 128 tokens), so every intermediate of
// the block stays on chip:
//
//   TMA      LN(x)+pos  -> bufA, LN(x) -> bufB          (bf16 [128 x 256] K-major A operands, 64 KiB each, from layernorm_kernel)
//   tcgen05  Q = bufA Wq^T -> TMEM [0,256)              K = bufA Wk^T -> TMEM [256,512)        (weights streamed through a TMA ring)
//   rows     Q + bq -> bf16 packed IN PLACE in TMEM [0,128): the A operand of S = Q K^T is read from tensor memory
//            K + bk -> bf16 -> bufA (LN(x)+pos is dead)   [128 keys x 256] K-major = B operand of S
//   tcgen05  V = bufB Wv^T -> TMEM [256,512)
//   rows     V + bv -> bf16 -> bufB as stored ([128 keys x 256] = an MN-major B operand of P V: no transpose; LN(x) is dead)
//   per head h (two teams of four warps take alternate heads, each with its own S / O buffers):
//     tcgen05  S_h = Q_h K_h^T                          -> TMEM S[team] (128 columns)
//     rows     softmax over the thread's own TMEM lane (masks as shared-memory vectors), un-normalised P packed to bf16 IN PLACE
//     tcgen05  O_h = P V_h  (A from TMEM)                -> TMEM Otmp[team] (32 columns)
//     rows     O_h / sum -> bf16 -> TMEM [16h, 16h+16)  (the columns Q_h occupied: dead once S_h has been issued)
//   tcgen05  Y = O Wo^T  (A = the packed [128 x 256] O in TMEM, Wo streamed)   -> TMEM [128,384)
//   rows     Y + bo + residual x (TMA-prefetched into bufA/bufB as soon as K / V are dead) -> fp32 -> TMA store (S rows only)
//            optionally LayerNorm2 of the new row (two-pass mean / variance like layernorm_kernel; the two teams own half a row each
//            and exchange partial sums through shared memory) -> bf16 -> global: the FFN's input, without a separate launch
//
// Replaces four launches per encoder layer (V projection, Q/K projection, attention_tc_kernel, out_proj + residual: 102 us at
// B = 256, 11 % of the tensor peak, six HBM round trips of [rows, 256..512] tensors) by one that reads LN(x), LN(x)+pos and x once
// and writes x once.  Rounding points are those of the unfused path (Q, K, V, P, O in bf16; everything else fp32), so the two
// agree up to fp32 summation order.
//
// 320 threads: warp 0 TMA producer, warp 1 MMA issuer, warps 2..9 "row" warps (warp & 3 = TMEM lane quadrant; team = (warp-2)/4).
#include "tc_common.cuh"
#include <algorithm>
#include <cstdlib>
#include <math_constants.h>

namespace sedt {
namespace {

using namespace tc;

constexpr int EA_THREADS = 320;
constexpr int EA_SLOT = 16384;                   // one [128 rows][64 k] bf16 box / one [128 rows][32 fp32] box
constexpr int EA_NSLOTS = 5;
constexpr int EA_BUFA = 0;
constexpr int EA_BUFB = 65536;
constexpr int EA_RING = 131072;
constexpr int EA_BIAS = EA_RING + EA_NSLOTS * EA_SLOT;          // in_proj bias [768] + out_proj bias [256] fp32
constexpr int EA_MASK = EA_BIAS + 1024 * 4;                     // key validity 0/1 [128], 0/-1e30 [128]
constexpr int EA_LN = EA_MASK + 1024;                           // LayerNorm2 gamma [256], beta [256]
constexpr int EA_STAT = EA_LN + 2048;                           // per-row partial sums of the two teams: [2 passes][2 teams][128]
constexpr int EA_BAR = EA_STAT + 2048;
constexpr int EA_NBARS = 2 * EA_NSLOTS + 23;
constexpr int EA_SMEM = EA_BAR + EA_NBARS * 8 + 16 + 1024;
static_assert(EA_SMEM <= 232448, "shared memory budget exceeded");

// TMEM columns
constexpr uint32_t TM_Q = 0;          // Q accumulator [0,256) -> packed Q / O bf16 [0,128)
constexpr uint32_t TM_KV = 256;       // K, then V accumulator [256,512)
constexpr uint32_t TM_S = 128;        // S / P buffers of the two teams: [128,256), [256,384)
constexpr uint32_t TM_OT = 384;       // O_h accumulators of the two teams: [384,416), [416,448)
constexpr uint32_t TM_Y = 128;        // out_proj accumulator [128,384)

struct EaParams {
    const float* b_in;        // [768]
    const float* b_out;       // [256]
    const uint8_t* kpm;       // [B, S] (1 = padded key) or null
    int B, S;
    float scale;
    const float* ln_g;        // optional LayerNorm applied to the updated rows (the encoder layer's norm2): gamma / beta [256]
    const float* ln_b;
    __nv_bfloat16* ln_out;    // [B*S, 256] bf16 or null
    long long* dbg;           // optional phase timestamps of CTA 0 (SEDT_EA_DEBUG=1, see launch_enc_attn_fused); null in production
};

// debug timeline: slot = role * 64 + event, clock64() of the first two clips of CTA 0
#define EA_T(role, ev) do { if (p.dbg != nullptr && blockIdx.x == 0 && it < 2) p.dbg[(it * 3 + (role)) * 64 + (ev)] = clock64(); } while (0)

__device__ __forceinline__ uint32_t ea_swz(int row, int byte_in_row) {
    return (uint32_t)(row * 128 + ((((byte_in_row >> 4) ^ (row & 7)) << 4) | (byte_in_row & 15)));
}
__device__ __forceinline__ uint32_t ea_pack2(float a, float b) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}

__global__ void __launch_bounds__(EA_THREADS, 1)
enc_attn_fused_kernel(const __grid_constant__ CUtensorMap map_nap, const __grid_constant__ CUtensorMap map_na,
                      const __grid_constant__ CUtensorMap map_win, const __grid_constant__ CUtensorMap map_wo,
                      const __grid_constant__ CUtensorMap map_res, const __grid_constant__ CUtensorMap map_out,
                      const __grid_constant__ EaParams p)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* slot_full = (uint64_t*)(smem + EA_BAR);
    uint64_t* slot_empty = slot_full + EA_NSLOTS;
    uint64_t* nap_full = slot_empty + EA_NSLOTS;      // LN(x)+pos has landed in bufA
    uint64_t* na_full = nap_full + 1;                 // LN(x) has landed in bufB
    uint64_t* q_full = na_full + 1;
    uint64_t* k_full = q_full + 1;
    uint64_t* v_full = k_full + 1;
    uint64_t* q_drained = v_full + 1;
    uint64_t* k_drained = q_drained + 1;
    uint64_t* v_drained = k_drained + 1;
    uint64_t* s_full = v_drained + 1;        // [2]
    uint64_t* p_ready = s_full + 2;          // [2]
    uint64_t* o_full = p_ready + 2;          // [2]
    uint64_t* o_done = o_full + 2;           // [2]
    uint64_t* kv_dead = o_done + 2;
    uint64_t* y_full = kv_dead + 1;
    uint64_t* res_full = y_full + 1;         // [2]
    uint64_t* y_drained = res_full + 2;
    uint64_t* stage_free = y_drained + 1;    // [2] the output stores of bufA / bufB have been read out
    uint32_t* tmem_slot = (uint32_t*)(stage_free + 2);
    float* s_bias = (float*)(smem + EA_BIAS);
    float* s_mask = (float*)(smem + EA_MASK);
    float* s_neg = s_mask + 128;
    float* s_ln = (float*)(smem + EA_LN);
    float* s_stat = (float*)(smem + EA_STAT);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int S = p.S;
    const int iters = (int)blockIdx.x < p.B ? (p.B - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&map_nap); prefetch_tmap(&map_na); prefetch_tmap(&map_win); prefetch_tmap(&map_wo);
        prefetch_tmap(&map_res); prefetch_tmap(&map_out);
        for (int s = 0; s < EA_NSLOTS; ++s) { mbar_init(&slot_full[s], 1); mbar_init(&slot_empty[s], 1); }
        mbar_init(nap_full, 1); mbar_init(na_full, 1); mbar_init(q_full, 1); mbar_init(k_full, 1); mbar_init(v_full, 1);
        mbar_init(q_drained, 8); mbar_init(k_drained, 8); mbar_init(v_drained, 8);
        for (int g = 0; g < 2; ++g) {
            mbar_init(&s_full[g], 1); mbar_init(&p_ready[g], 4); mbar_init(&o_full[g], 1); mbar_init(&o_done[g], 4);
            mbar_init(&res_full[g], 1);
        }
        mbar_init(kv_dead, 1); mbar_init(y_full, 1); mbar_init(y_drained, 8);
        mbar_init(&stage_free[0], 1); mbar_init(&stage_free[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc<512>(tmem_slot);
    for (int i = threadIdx.x; i < 1024; i += EA_THREADS) s_bias[i] = i < 768 ? p.b_in[i] : p.b_out[i - 768];   // weights: not produced by the predecessor
    if (p.ln_out != nullptr)
        for (int i = threadIdx.x; i < 512; i += EA_THREADS) s_ln[i] = i < 256 ? p.ln_g[i] : p.ln_b[i - 256];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    pdl_trigger();
    pdl_wait();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int slot = 0, pre = 0; uint32_t sphase = 0;
            auto next_slot = [&](const CUtensorMap* m, int c0, int c1) {
                mbar_wait(&slot_empty[slot], sphase ^ 1);
                mbar_expect_tx(&slot_full[slot], EA_SLOT);
                tma_load_2d(m, smem + EA_RING + slot * EA_SLOT, &slot_full[slot], c0, c1);
                if (++slot == EA_NSLOTS) { slot = 0; sphase ^= 1; }
            };
            for (int it = 0; it < iters; ++it) {
                const int b = (int)blockIdx.x + it * (int)gridDim.x;
                const int row0 = b * S;
                // bufA / bufB are free once the previous clip's output stores have been read out of them; Q / K only need bufA
                mbar_wait(&stage_free[0], (it & 1) ^ 1);
                EA_T(0, 0);
                mbar_expect_tx(nap_full, 4 * EA_SLOT);
                for (int kb = 0; kb < 4; ++kb) tma_load_2d(&map_nap, smem + EA_BUFA + kb * EA_SLOT, nap_full, kb * 64, row0);
                mbar_wait(&stage_free[1], (it & 1) ^ 1);
                mbar_expect_tx(na_full, 4 * EA_SLOT);
                for (int kb = 0; kb < 4; ++kb) tma_load_2d(&map_na, smem + EA_BUFB + kb * EA_SLOT, na_full, kb * 64, row0);
                // weight boxes of one clip in consumption order: 0..23 = Wq, Wk, Wv as [N half][k block], 24..31 = Wo as [k block][N half];
                // the first `pre` boxes of this clip were issued during the previous clip's tail (they do not need bufA / bufB)
                auto issue_w = [&](int i) {
                    if (i < 24) next_slot(&map_win, (i & 3) * 64, (i >> 3) * 256 + ((i >> 2) & 1) * 128);
                    else next_slot(&map_wo, ((i - 24) >> 1) * 64, ((i - 24) & 1) * 128);
                };
                for (int i = pre; i < 24; ++i) issue_w(i);
                EA_T(0, 1);
                for (int i = 24; i < 32; ++i) {
                    if (i == 24 + EA_NSLOTS) {
                        // K and V^T are dead once the last P V has retired: prefetch the residual tile over them
                        EA_T(0, 2);
                        mbar_wait(kv_dead, it & 1);
                        EA_T(0, 3);
                        for (int g = 0; g < 2; ++g) {
                            mbar_expect_tx(&res_full[g], 4 * EA_SLOT);
                            for (int c = 0; c < 4; ++c)
                                tma_load_2d(&map_res, smem + (g ? EA_BUFB : EA_BUFA) + c * EA_SLOT, &res_full[g], (4 * g + c) * 32, row0);
                        }
                    }
                    issue_w(i);
                }
                pre = 0;
                if (it + 1 < iters)                                 // the ring drains while the epilogue runs: the next clip's first weights go now
                    for (; pre < EA_NSLOTS; ++pre) issue_w(pre);
                EA_T(0, 4);
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc128 = make_idesc(128, 128);
            constexpr uint32_t idesc32 = make_idesc(128, 32) | (1u << 16);      // B (= V, [keys][dims]) is MN-major
            int slot = 0; uint32_t sphase = 0;
            const uint32_t sA = smem_u32(smem + EA_BUFA), sB = smem_u32(smem + EA_BUFB);
            const int ksteps_pv = (S + UMMA_K - 1) / UMMA_K;
            // one projection: D[128 x 256] = A[128 x 256] W^T, W rows streamed as [N half][k block] boxes
            auto project = [&](uint32_t a_base, uint32_t d_col) {
                for (int nh = 0; nh < 2; ++nh)
                    for (int kb = 0; kb < 4; ++kb) {
                        mbar_wait(&slot_full[slot], sphase);
                        tc_fence_after();
                        const uint32_t sa = a_base + kb * EA_SLOT;
                        const uint32_t sb = smem_u32(smem + EA_RING + slot * EA_SLOT);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_bf16(tmem_base + d_col + (uint32_t)(nh * 128), make_smem_desc(sa + k * 32), make_smem_desc(sb + k * 32),
                                      idesc128, (kb > 0 || k > 0) ? 1u : 0u);
                        umma_commit(&slot_empty[slot]);
                        if (++slot == EA_NSLOTS) { slot = 0; sphase ^= 1; }
                    }
            };
            for (int it = 0; it < iters; ++it) {
                const uint32_t par = (uint32_t)(it & 1);
                EA_T(1, 0);
                mbar_wait(nap_full, par);
                EA_T(1, 1);
                mbar_wait(y_drained, par ^ 1);                      // the previous clip's Y has left TMEM
                tc_fence_after();
                EA_T(1, 2);
                project(sA, TM_Q);  umma_commit(q_full);
                EA_T(1, 3);
                project(sA, TM_KV); umma_commit(k_full);
                EA_T(1, 4);
                mbar_wait(k_drained, par);                          // K is in bufA, its accumulator may be overwritten
                mbar_wait(na_full, par);
                tc_fence_after();
                EA_T(1, 5);
                project(sB, TM_KV); umma_commit(v_full);
                EA_T(1, 6);
                mbar_wait(q_drained, par);
                mbar_wait(v_drained, par);
                tc_fence_after();
                EA_T(1, 7);
                auto issue_s = [&](int h) {
                    const int g = h & 1;
                    const uint32_t d = tmem_base + TM_S + (uint32_t)(g * 128);
                    const uint32_t kaddr = sA + (h >> 1) * EA_SLOT + (h & 1) * 64;
#pragma unroll
                    for (int k = 0; k < 2; ++k)
                        umma_bf16_ts(d, tmem_base + TM_Q + (uint32_t)(16 * h + 8 * k), make_smem_desc(kaddr + k * 32), idesc128, k > 0 ? 1u : 0u);
                    umma_commit(&s_full[g]);
                };
                issue_s(0);
                issue_s(1);
                for (int h = 0; h < 8; ++h) {
                    const int g = h & 1;
                    const uint32_t u = (uint32_t)(it * 4 + (h >> 1));
                    mbar_wait(&p_ready[g], u & 1);
                    if (h >= 2) mbar_wait(&o_done[g], (u - 1) & 1);  // this team's O accumulator has been read
                    tc_fence_after();
                    const uint32_t d = tmem_base + TM_OT + (uint32_t)(g * 32);
                    const uint32_t pa = tmem_base + TM_S + (uint32_t)(g * 128);
                    const uint32_t vaddr = sB + (h >> 1) * EA_SLOT + (h & 1) * 64;       // 32 dims = 64 bytes of every 128-byte key row
                    for (int ks = 0; ks < ksteps_pv; ++ks)                               // 16 keys = 16 rows of 128 bytes per step
                        umma_bf16_ts(d, pa + (uint32_t)(8 * ks), make_smem_desc_mn(vaddr + ks * UMMA_K * 128, EA_SLOT), idesc32, ks > 0 ? 1u : 0u);
                    umma_commit(&o_full[g]);
                    if (h + 2 < 8) issue_s(h + 2);                  // in-order execution: P_h has been consumed by then
                }
                umma_commit(kv_dead);
                EA_T(1, 8);
                mbar_wait(&o_done[0], (uint32_t)(it * 4 + 3) & 1);
                mbar_wait(&o_done[1], (uint32_t)(it * 4 + 3) & 1);
                tc_fence_after();
                EA_T(1, 9);
                for (int i = 0; i < 8; ++i) {                       // Y = O Wo^T, A = packed O in TMEM
                    const int kb = i >> 1, nh = i & 1;
                    mbar_wait(&slot_full[slot], sphase);
                    tc_fence_after();
                    const uint32_t sb = smem_u32(smem + EA_RING + slot * EA_SLOT);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16_ts(tmem_base + TM_Y + (uint32_t)(nh * 128), tmem_base + TM_Q + (uint32_t)(kb * 32 + k * 8),
                                     make_smem_desc(sb + k * 32), idesc128, (kb > 0 || k > 0) ? 1u : 0u);
                    umma_commit(&slot_empty[slot]);
                    if (++slot == EA_NSLOTS) { slot = 0; sphase ^= 1; }
                }
                umma_commit(y_full);
                EA_T(1, 10);
            }
        }
    } else {
        // ===================== row warps =====================
        const int e = warp - 2, team = e >> 2, quad = warp & 3;
        const int r = quad * 32 + lane;                              // row of the tile = TMEM lane
        const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
        const int nkc = (S + 31) >> 5;
        const bool row_warp = quad * 32 < S;
        const float cs = p.scale * 1.4426950408889634f;
        uint8_t* bufA = smem + EA_BUFA;
        uint8_t* bufB = smem + EA_BUFB;
        for (int it = 0; it < iters; ++it) {
            const int b = (int)blockIdx.x + it * (int)gridDim.x;
            const uint32_t par = (uint32_t)(it & 1);
            if (team == 0) {
                float mk = 1.f;                                      // validity of key r
                if (r >= S) mk = 0.f;
                else if (p.kpm != nullptr && p.kpm[(size_t)b * S + r]) mk = 0.f;
                s_mask[r] = mk;
                s_neg[r] = (mk - 1.f) * 1e30f;
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
#define EA_TT(ev) do { if (quad == 0 && lane == 0) EA_T(2, team * 32 + (ev)); } while (0)
            EA_TT(0);

            // ---- Q: fp32 accumulator + bias -> bf16, packed in place (team t: columns [128t, 128t+128) -> [64t, 64t+64))
            mbar_wait(q_full, par);
            tc_fence_after();
            EA_TT(1);
            {
                uint32_t w[64];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    uint32_t acc[32];
                    tmem_ld32(lane_base + TM_Q + (uint32_t)(team * 128 + c * 32), acc);
                    const float* bq = s_bias + team * 128 + c * 32;
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        w[16 * c + j] = ea_pack2(__uint_as_float(acc[2 * j]) + bq[2 * j], __uint_as_float(acc[2 * j + 1]) + bq[2 * j + 1]);
                }
                // the partner warp of the other team (same TMEM lanes) has read its fp32 columns too
                asm volatile("bar.sync %0, 64;" ::"r"(2 + quad) : "memory");
                tmem_st32(lane_base + TM_Q + (uint32_t)(team * 64), *reinterpret_cast<const uint32_t(*)[32]>(&w[0]));
                tmem_st32(lane_base + TM_Q + (uint32_t)(team * 64 + 32), *reinterpret_cast<const uint32_t(*)[32]>(&w[32]));
                tmem_st_wait();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(q_drained);
            EA_TT(2);

            // ---- K: + bias -> bf16 -> bufA in the K-major operand layout (LN(x)+pos is dead: both projections have retired)
            mbar_wait(k_full, par);
            tc_fence_after();
            EA_TT(3);
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                const int col = team * 128 + c * 32;
                uint32_t acc[32];
                tmem_ld32(lane_base + TM_KV + (uint32_t)col, acc);
                const float* bk = s_bias + 256 + col;
                uint8_t* rowp = bufA + (col >> 6) * EA_SLOT + r * 128;
#pragma unroll
                for (int j8 = 0; j8 < 4; ++j8) {
                    uint32_t q4[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        q4[q] = ea_pack2(__uint_as_float(acc[8 * j8 + 2 * q]) + bk[8 * j8 + 2 * q],
                                         __uint_as_float(acc[8 * j8 + 2 * q + 1]) + bk[8 * j8 + 2 * q + 1]);
                    *reinterpret_cast<uint4*>(rowp + ((((c & 1) * 4 + j8) ^ (r & 7)) << 4)) = make_uint4(q4[0], q4[1], q4[2], q4[3]);
                }
            }
            fence_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(k_drained);
            EA_TT(4);

            // ---- V: + bias -> bf16 -> bufB in the same [keys][256] image (read by P V as an MN-major operand; LN(x) is dead)
            mbar_wait(v_full, par);
            tc_fence_after();
            EA_TT(5);
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                const int col = team * 128 + c * 32;
                uint32_t acc[32];
                tmem_ld32(lane_base + TM_KV + (uint32_t)col, acc);
                const float* bv = s_bias + 512 + col;
                uint8_t* rowp = bufB + (col >> 6) * EA_SLOT + r * 128;
#pragma unroll
                for (int j8 = 0; j8 < 4; ++j8) {
                    uint32_t q4[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        q4[q] = ea_pack2(__uint_as_float(acc[8 * j8 + 2 * q]) + bv[8 * j8 + 2 * q],
                                         __uint_as_float(acc[8 * j8 + 2 * q + 1]) + bv[8 * j8 + 2 * q + 1]);
                    *reinterpret_cast<uint4*>(rowp + ((((c & 1) * 4 + j8) ^ (r & 7)) << 4)) = make_uint4(q4[0], q4[1], q4[2], q4[3]);
                }
            }
            fence_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(v_drained);
            EA_TT(6);

            // ---- heads h = 2 hh + team
#pragma unroll 1
            for (int hh = 0; hh < 4; ++hh) {
                const int h = 2 * hh + team;
                const uint32_t u = (uint32_t)(it * 4 + hh);
                const uint32_t sbase = lane_base + TM_S + (uint32_t)(team * 128);
                mbar_wait(&s_full[team], u & 1);
                tc_fence_after();
                EA_TT(8 + 4 * hh);
                // p_j = 2^(c (s_j - m)) * valid_j, c = scale * log2 e, m = max over the valid keys (attention_tc.cu)
                // chunks whose 32 keys are all real (no padding mask, below S) skip the mask vectors; the TMEM load of chunk c + 1 is
                // in flight while chunk c is processed
                const int nch = row_warp ? nkc : 0;
                const bool nomask = p.kpm == nullptr;
                // four independent partial maxima / sums: a single accumulator would serialise 128 dependent FMNMX / FADD per head
                float m0 = -CUDART_INF_F, m1 = -CUDART_INF_F, m2 = -CUDART_INF_F, m3 = -CUDART_INF_F;
                auto max_chunk = [&](const uint32_t (&a)[32], int cc) {
                    if (nomask && cc * 32 + 32 <= S) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            m0 = fmaxf(m0, __uint_as_float(a[j])); m1 = fmaxf(m1, __uint_as_float(a[j + 1]));
                            m2 = fmaxf(m2, __uint_as_float(a[j + 2])); m3 = fmaxf(m3, __uint_as_float(a[j + 3]));
                        }
                    } else {
#pragma unroll
                        for (int j4 = 0; j4 < 8; ++j4) {
                            const float4 ng = *reinterpret_cast<const float4*>(s_neg + cc * 32 + j4 * 4);
                            m0 = fmaxf(m0, __uint_as_float(a[4 * j4]) + ng.x); m1 = fmaxf(m1, __uint_as_float(a[4 * j4 + 1]) + ng.y);
                            m2 = fmaxf(m2, __uint_as_float(a[4 * j4 + 2]) + ng.z); m3 = fmaxf(m3, __uint_as_float(a[4 * j4 + 3]) + ng.w);
                        }
                    }
                };
                uint32_t acc0[32], acc1[32];
                if (nch > 0) tmem_ld32_nowait(sbase, acc0);
#pragma unroll 1
                for (int c = 0; c < nch; c += 2) {
                    tmem_ld_wait();
                    if (c + 1 < nch) tmem_ld32_nowait(sbase + (uint32_t)((c + 1) * 32), acc1);
                    max_chunk(acc0, c);
                    if (c + 1 < nch) {
                        tmem_ld_wait();
                        if (c + 2 < nch) tmem_ld32_nowait(sbase + (uint32_t)((c + 2) * 32), acc0);
                        max_chunk(acc1, c + 1);
                    }
                }
                const float mc = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)) * cs;
                float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
                // P chunk cc goes to columns [16 cc, 16 cc + 16): inside S chunks <= cc / 2, all consumed, and never the columns of
                // the load in flight (chunk cc + 1)
                auto exp_chunk = [&](const uint32_t (&a)[32], int cc) {
                    uint32_t pk[16];
                    if (nomask && cc * 32 + 32 <= S) {
#pragma unroll
                        for (int j = 0; j < 16; j += 2) {
                            float e0, e1, e2, e3;
                            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(fmaf(__uint_as_float(a[2 * j]), cs, -mc)));
                            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(fmaf(__uint_as_float(a[2 * j + 1]), cs, -mc)));
                            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e2) : "f"(fmaf(__uint_as_float(a[2 * j + 2]), cs, -mc)));
                            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e3) : "f"(fmaf(__uint_as_float(a[2 * j + 3]), cs, -mc)));
                            l0 += e0; l1 += e1; l2 += e2; l3 += e3;
                            pk[j] = ea_pack2(e0, e1);
                            pk[j + 1] = ea_pack2(e2, e3);
                        }
                    } else {
#pragma unroll
                        for (int j4 = 0; j4 < 8; ++j4) {
                            const float4 vm = *reinterpret_cast<const float4*>(s_mask + cc * 32 + j4 * 4);
                            const float vmv[4] = {vm.x, vm.y, vm.z, vm.w};
                            float pv[4];
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                float ex;
                                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(fmaf(__uint_as_float(a[4 * j4 + q]), cs, -mc)));
                                pv[q] = ex * vmv[q];
                            }
                            l0 += pv[0]; l1 += pv[1]; l2 += pv[2]; l3 += pv[3];
                            pk[2 * j4] = ea_pack2(pv[0], pv[1]);
                            pk[2 * j4 + 1] = ea_pack2(pv[2], pv[3]);
                        }
                    }
                    tmem_st16(sbase + (uint32_t)(cc * 16), pk);
                };
                if (nch > 0) tmem_ld32_nowait(sbase, acc0);
#pragma unroll 1
                for (int c = 0; c < nch; c += 2) {
                    tmem_ld_wait();
                    if (c + 1 < nch) tmem_ld32_nowait(sbase + (uint32_t)((c + 1) * 32), acc1);
                    exp_chunk(acc0, c);
                    if (c + 1 < nch) {
                        tmem_ld_wait();
                        if (c + 2 < nch) tmem_ld32_nowait(sbase + (uint32_t)((c + 2) * 32), acc0);
                        exp_chunk(acc1, c + 1);
                    }
                }
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&p_ready[team]);
                EA_TT(9 + 4 * hh);

                mbar_wait(&o_full[team], u & 1);
                tc_fence_after();
                EA_TT(10 + 4 * hh);
                {
                    uint32_t acc[32], pk[16];
                    tmem_ld32(lane_base + TM_OT + (uint32_t)(team * 32), acc);
                    const float l = (l0 + l1) + (l2 + l3);
                    const float inv = l > 0.f ? 1.f / l : 0.f;
#pragma unroll
                    for (int j = 0; j < 16; ++j) pk[j] = ea_pack2(__uint_as_float(acc[2 * j]) * inv, __uint_as_float(acc[2 * j + 1]) * inv);
                    tmem_st16(lane_base + TM_Q + (uint32_t)(16 * h), pk);
                    tmem_st_wait();
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&o_done[team]);
                EA_TT(11 + 4 * hh);
            }

            // ---- output: Y + bo + x -> fp32, staged through the residual tile (team t: columns [128t, 128t+128) in bufA / bufB)
            mbar_wait(y_full, par);
            tc_fence_after();
            EA_TT(24);
            uint8_t* stage = team ? bufB : bufA;
            mbar_wait(&res_full[team], par);
            EA_TT(25);
            float rsum = 0.f;
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                uint32_t acc[32];
                tmem_ld32(lane_base + TM_Y + (uint32_t)(team * 128 + c * 32), acc);
                if (c == 3) {                                                          // the whole accumulator is in registers / staged
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(y_drained);
                }
                const float* bo = s_bias + 768 + team * 128 + c * 32;
                uint8_t* rowp = stage + c * EA_SLOT + r * 128;                        // 32 fp32 columns = one 128-byte row of chunk c
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4) {
                    float4* slot = reinterpret_cast<float4*>(rowp + ((j4 ^ (r & 7)) << 4));
                    const float4 r4 = *slot;
                    const float4 b4 = *reinterpret_cast<const float4*>(bo + 4 * j4);
                    const float4 o4 = make_float4(__uint_as_float(acc[4 * j4]) + b4.x + r4.x, __uint_as_float(acc[4 * j4 + 1]) + b4.y + r4.y,
                                                  __uint_as_float(acc[4 * j4 + 2]) + b4.z + r4.z, __uint_as_float(acc[4 * j4 + 3]) + b4.w + r4.w);
                    *slot = o4;
                    rsum += (o4.x + o4.y) + (o4.z + o4.w);
                }
                // chunk c leaves through TMA while chunk c + 1 is being computed
                fence_async_smem();
                asm volatile("bar.sync %0, 128;" ::"r"(6 + team) : "memory");
                if ((e & 3) == 0 && lane == 0) {
                    tma_store_3d(&map_out, stage + c * EA_SLOT, (4 * team + c) * 32, 0, b);
                    tma_store_commit();
                }
            }
            if (p.ln_out != nullptr) {
                // LayerNorm2 of the updated row (sedt/transformer.py:199): this thread holds columns [128 team, 128 team + 128) of
                // row r in the staging tile; mean and centred variance in two passes, the halves meeting in shared memory
                s_stat[team * 128 + r] = rsum;
                asm volatile("bar.sync 1, 256;" ::: "memory");
                const float mean = (s_stat[r] + s_stat[128 + r]) * (1.f / 256.f);
                float q = 0.f;
#pragma unroll 1
                for (int c = 0; c < 4; ++c) {
                    const uint8_t* rowp = stage + c * EA_SLOT + r * 128;
#pragma unroll
                    for (int j4 = 0; j4 < 8; ++j4) {
                        const float4 v = *reinterpret_cast<const float4*>(rowp + ((j4 ^ (r & 7)) << 4));
                        const float d0 = v.x - mean, d1 = v.y - mean, d2 = v.z - mean, d3 = v.w - mean;
                        q = fmaf(d0, d0, q); q = fmaf(d1, d1, q); q = fmaf(d2, d2, q); q = fmaf(d3, d3, q);
                    }
                }
                s_stat[256 + team * 128 + r] = q;
                asm volatile("bar.sync 1, 256;" ::: "memory");
                const float rstd = rsqrtf((s_stat[256 + r] + s_stat[384 + r]) * (1.f / 256.f) + 1e-5f);
                if (r < S) {
                    __nv_bfloat16* orow = p.ln_out + ((size_t)b * S + r) * 256 + team * 128;
#pragma unroll 1
                    for (int c = 0; c < 4; ++c) {
                        const uint8_t* rowp = stage + c * EA_SLOT + r * 128;
                        const float* g = s_ln + team * 128 + c * 32;
                        const float* bt = s_ln + 256 + team * 128 + c * 32;
#pragma unroll
                        for (int j8 = 0; j8 < 4; ++j8) {
                            const float4 v0 = *reinterpret_cast<const float4*>(rowp + (((2 * j8) ^ (r & 7)) << 4));
                            const float4 v1 = *reinterpret_cast<const float4*>(rowp + (((2 * j8 + 1) ^ (r & 7)) << 4));
                            const float4 g0 = *reinterpret_cast<const float4*>(g + 8 * j8), g1 = *reinterpret_cast<const float4*>(g + 8 * j8 + 4);
                            const float4 b0 = *reinterpret_cast<const float4*>(bt + 8 * j8), b1 = *reinterpret_cast<const float4*>(bt + 8 * j8 + 4);
                            uint4 o;
                            o.x = ea_pack2((v0.x - mean) * rstd * g0.x + b0.x, (v0.y - mean) * rstd * g0.y + b0.y);
                            o.y = ea_pack2((v0.z - mean) * rstd * g0.z + b0.z, (v0.w - mean) * rstd * g0.w + b0.w);
                            o.z = ea_pack2((v1.x - mean) * rstd * g1.x + b1.x, (v1.y - mean) * rstd * g1.y + b1.y);
                            o.w = ea_pack2((v1.z - mean) * rstd * g1.z + b1.z, (v1.w - mean) * rstd * g1.w + b1.w);
                            *reinterpret_cast<uint4*>(orow + c * 32 + j8 * 8) = o;
                        }
                    }
                }
                // the staging tile has been re-read by every thread of the team: only now may it be handed back to the producer
                asm volatile("bar.sync %0, 128;" ::"r"(6 + team) : "memory");
            }
            EA_TT(26);
            if ((e & 3) == 0 && lane == 0) {
                tma_store_wait_read0();
                mbar_arrive(&stage_free[team]);
                if (p.dbg != nullptr && blockIdx.x == 0 && it < 2) p.dbg[(it * 3 + 2) * 64 + team * 32 + 27] = clock64();
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

}  // namespace

bool enc_attn_fused_supported(int d, int nheads, int S, const void* na, const void* nap, const void* w_in, const void* w_out,
                              const float* x, int dt)
{
    if (dt != DT_BF16 || d != 256 || nheads != 8 || S < 1 || S > 128) return false;
    if (((uintptr_t)na & 15) || ((uintptr_t)nap & 15) || ((uintptr_t)w_in & 15) || ((uintptr_t)w_out & 15) || ((uintptr_t)x & 15)) return false;
    return true;
}

// SEDT_ENC_ATTN_FUSED: unset / 1 = use the fused block wherever it is supported, 0 = the four separate launches
bool enc_attn_fused_enabled()
{
    static const bool on = [] { const char* e = getenv("SEDT_ENC_ATTN_FUSED"); return e == nullptr || atoi(e) != 0; }();
    return on;
}

// x[B*S, 256] (fp32, in place) += out_proj(attention(q = k = nap Wq/Wk^T, v = na Wv^T)); na / nap [B*S, 256] bf16,
// w_in [768, 256] / w_out [256, 256] bf16 (row-major [out, in] as packed), b_in [768] / b_out [256] fp32, kpm [B, S] or null
int launch_enc_attn_fused(const void* na, const void* nap, const void* w_in, const float* b_in, const void* w_out, const float* b_out,
                          const uint8_t* kpm, float* x, int B, int S, float scale, cudaStream_t stream, const float* ln_g,
                          const float* ln_b, void* ln_out)
{
    SEDT_REQUIRE(ln_out == nullptr || (ln_g != nullptr && ln_b != nullptr && ((uintptr_t)ln_out & 15) == 0), "enc_attn_fused: LayerNorm needs gamma, beta and a 16-byte aligned output");
    SEDT_REQUIRE(enc_attn_fused_supported(256, 8, S, na, nap, w_in, w_out, x, DT_BF16), "enc_attn_fused: unsupported shape / alignment");
    SEDT_REQUIRE(B >= 1, "enc_attn_fused: B=%d", B);
    SEDT_TRY(tc_init());
    const uint64_t rows = (uint64_t)B * (uint64_t)S;
    CUtensorMap mnap, mna, mwin, mwo, mres, mout;
    const uint32_t box[2] = {64u, 128u};
    {
        const uint64_t dims[2] = {256, rows}; const uint64_t strides[1] = {256 * 2};
        SEDT_TRY(encode_map(&mnap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, nap, 2, dims, strides, box));
        SEDT_TRY(encode_map(&mna, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, na, 2, dims, strides, box));
    }
    {
        const uint64_t dims[2] = {256, 768}; const uint64_t strides[1] = {256 * 2};
        SEDT_TRY(encode_map(&mwin, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, w_in, 2, dims, strides, box));
        const uint64_t dims2[2] = {256, 256};
        SEDT_TRY(encode_map(&mwo, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, w_out, 2, dims2, strides, box));
    }
    {
        const uint32_t rbox[2] = {32u, 128u};                     // 32 fp32 columns = one 128-byte swizzle row
        const uint64_t dims[2] = {256, rows}; const uint64_t strides[1] = {256 * 4};
        SEDT_TRY(encode_map(&mres, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, x, 2, dims, strides, rbox));
        // the store sees the clip as [B][S][256] so that a tile never writes into the next clip's rows
        const uint32_t obox[3] = {32u, (uint32_t)S, 1u};
        const uint64_t odims[3] = {256, (uint64_t)S, (uint64_t)B}; const uint64_t ostr[2] = {256 * 4, (uint64_t)S * 256 * 4};
        SEDT_TRY(encode_map(&mout, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, x, 3, odims, ostr, obox));
    }
    EaParams p;
    p.b_in = b_in; p.b_out = b_out; p.kpm = kpm; p.B = B; p.S = S; p.scale = scale; p.dbg = nullptr;
    p.ln_g = ln_g; p.ln_b = ln_b; p.ln_out = (__nv_bfloat16*)ln_out;
    // SEDT_EA_DEBUG=1 (development only; synchronises): phase timeline of CTA 0's first two clips on stderr, in cycles
    static const bool debug = [] { const char* e = getenv("SEDT_EA_DEBUG"); return e != nullptr && atoi(e) != 0; }();
    long long* dbg_dev = nullptr;
    if (debug) {
        SEDT_CHECK_CUDA(cudaMalloc(&dbg_dev, 6 * 64 * sizeof(long long)));
        SEDT_CHECK_CUDA(cudaMemsetAsync(dbg_dev, 0, 6 * 64 * sizeof(long long), stream));
        p.dbg = dbg_dev;
    }
    static bool attr_set = false;
    if (!attr_set) {
        SEDT_CHECK_CUDA(cudaFuncSetAttribute(enc_attn_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, EA_SMEM));
        attr_set = true;
    }
    const int grid = std::min(B, num_sms());
    ProfScope _prof(PROF_ATTENTION, stream);
    SEDT_CHECK_CUDA(launch_pdl(enc_attn_fused_kernel, dim3((unsigned)grid), dim3(EA_THREADS), EA_SMEM, stream, 1, mnap, mna, mwin, mwo, mres,
                               mout, p));
    SEDT_COUNT_KIND(KK_ENC_ATTN_FUSED);
    SEDT_CHECK_CUDA(cudaGetLastError());
    if (debug) {
        long long h[6 * 64];
        SEDT_CHECK_CUDA(cudaStreamSynchronize(stream));
        SEDT_CHECK_CUDA(cudaMemcpy(h, dbg_dev, sizeof(h), cudaMemcpyDeviceToHost));
        cudaFree(dbg_dev);
        long long t0 = 0;
        for (int i = 0; i < 6 * 64; ++i) if (h[i] != 0 && (t0 == 0 || h[i] < t0)) t0 = h[i];
        static const char* roles[3] = {"tma", "mma", "row"};
        fprintf(stderr, "[enc_attn_fused B=%d S=%d] cycles since the first event of CTA 0\n", B, S);
        for (int it = 0; it < 2; ++it)
            for (int role = 0; role < 3; ++role) {
                fprintf(stderr, "  clip %d %s:", it, roles[role]);
                for (int ev = 0; ev < 64; ++ev) if (h[(it * 3 + role) * 64 + ev] != 0) fprintf(stderr, " %d=%lld", ev, h[(it * 3 + role) * 64 + ev] - t0);
                fprintf(stderr, "\n");
            }
    }
    return SEDT_OK;
}

}  // namespace sedt
