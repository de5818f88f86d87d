// Fused transformer FFN (eval): out = residual + linear2(relu(linear1(x)))   (sedt/transformer.py:202-203, forward_pre)
//
//   x [M, 256] bf16 (LayerNorm output), W1 [ff, 256], W2 [256, ff] bf16 (K-major as packed), fp32 biases, fp32 residual / out.
//
// The unfused path writes the [M, ff] hidden activation to HBM (130 MB per encoder layer at B = 256: linear1 is bound by that
// write, linear2 by reading it back).  Here a CTA keeps a 128-row tile of x resident in shared memory and walks over the hidden
// dimension in chunks of 128: GEMM1 (x W1_j^T, K = 256) accumulates a 128 x 128 tile in TMEM; the epilogue warps add the bias,
// apply ReLU and pack the tile to bf16 IN PLACE in the accumulator's TMEM columns (tcgen05.st); GEMM2 (h_j W2_j^T, K = 128)
// reads that tile as its A operand straight from tensor memory (tcgen05.mma with [a_tmem]) and accumulates the 128 x 256
// output tile in TMEM over all chunks.  Software-pipelined by one chunk (GEMM1_{j+1} is issued before GEMM2_j) so the tensor
// core has work while chunk j goes through the epilogue; two epilogue groups of eight warps take alternate chunks.
//
// TMEM: Y 256 columns + 2 x 128 (hidden accumulator / packed tile, double buffered) = 512.  Shared memory: x 64 KiB, weight
// ring 9 x 16 KiB ([128 rows x 64 k] boxes of W1 / W2), b1.  Warps: 0 TMA producer, 1 MMA issuer, 2..17 epilogue.  The output
// tile goes through the x region (free once every GEMM1 of the tile has retired): TMA prefetches the residual half-tile, the
// threads add accumulator + b2 in place, TMA stores it.
//
// History (all variants bit-identical, B200, M = 31744, ff = 2048; the two GEMMs this replaces: 97-99 us):
//   * hidden tile parked in shared memory as the A operand (6-slot ring): 113 us; + 2-CTA weight multicast: 113 us;
//   * hidden tile in TMEM, 9-slot ring, b1 in smem, 16 epilogue warps: 106-108 us.  Timed with parts disabled: skeleton without
//     MMAs and output 40 us (520 MB of weights through TMA = 13 TB/s, the L2 -> SM limit); + MMAs 75 us (shared-memory bandwidth:
//     per chunk 128 KB of GEMM1 operand reads + 64 KB of GEMM2 B reads + 128 KB of TMA writes = 2560 cycles at 128 B/clk against
//     2048 cycles of MMA); + a register -> global fp32 output pass 111 us (128 B per thread at a 1 KB stride = 32 L1 wavefronts
//     per warp instruction, exposed at every tile boundary);
//   * output through the x region with TMA: 90-92 us (this file); + 2-CTA weight multicast on top: 92.8 us (dropped).
// Per 128-row tile the kernel takes 48 us (49.5 us at one tile per SM, 96.6 us at two, no tile-boundary cost); M = 31744 is 1.68
// tiles per SM, i.e. two rounds with a third of the SMs idle in the second.  Dropping the wait for GEMM2_{j-2} before GEMM1_j
// (in-order MMA execution would make it redundant) changed nothing (91.1 vs 91.3 us) and was not kept.
// Next: x as a TMEM-resident A operand for GEMM1 (halves its shared-memory reads), N = 256 MMAs for GEMM2, a balanced split
// of the last round.
#include "tc_common.cuh"
#include <algorithm>
#include <cstdlib>

namespace sedt {
namespace {

using namespace tc;

constexpr int FF_SLOT_BYTES = 16384;            // one [128 rows][64 k] bf16 box
constexpr int FX_OFF = 0;                       // x tile: 4 k-blocks; reused as the output staging area
constexpr int TS_THREADS = 576;                 // warp 0 TMA, warp 1 MMA, warps 2..17 epilogue (two groups of eight)
constexpr int TS_SLOTS = 9;
constexpr int TS_MAX_FF = 3072;                 // b1 staged in shared memory (ncu: the per-element __ldg of the bias stalled the epilogue)
constexpr int TSW_OFF = 65536;
constexpr int TSB1_OFF = TSW_OFF + TS_SLOTS * FF_SLOT_BYTES;
constexpr int TSBAR_OFF = TSB1_OFF + TS_MAX_FF * 4;
constexpr int TS_NBARS = 2 * TS_SLOTS + 2 + 6 + 2 + 3;
constexpr int TS_SMEM = TSBAR_OFF + TS_NBARS * 8 + 16 + 1024;
static_assert(TS_SMEM <= 232448, "shared memory budget exceeded");

struct FfnParams {
    const float* b1; const float* b2; const float* residual; float* out;
    int ld_res, ldo, M, nch, tiles_m;
};

__global__ void __launch_bounds__(TS_THREADS, 1)
ffn_fused_ts_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w1,
                    const __grid_constant__ CUtensorMap map_w2, const __grid_constant__ CUtensorMap map_res,
                    const __grid_constant__ CUtensorMap map_out, const __grid_constant__ FfnParams p)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* slot_full = (uint64_t*)(smem + TSBAR_OFF);
    uint64_t* slot_empty = slot_full + TS_SLOTS;
    uint64_t* x_full = slot_empty + TS_SLOTS;
    uint64_t* x_empty = x_full + 1;
    uint64_t* hacc_full = x_empty + 1;          // [2] MMA -> epilogue group b: fp32 hidden accumulator ready
    uint64_t* hts_full = hacc_full + 2;         // [2] epilogue group b -> MMA: bf16 hidden tile packed in TMEM
    uint64_t* hfree = hts_full + 2;             // [2] GEMM2 has read the packed tile: the accumulator may be overwritten
    uint64_t* y_full = hfree + 2;
    uint64_t* y_empty = y_full + 1;
    uint64_t* res_full = y_empty + 1;           // [2] residual half-tile has landed in the staging area (the x region)
    uint64_t* stage_free = res_full + 2;        // group 0's output stores have left the staging area
    uint32_t* tmem_slot = (uint32_t*)(stage_free + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nch = p.nch, half_uses = nch >> 1;
    const int iters = (int)blockIdx.x < p.tiles_m ? (p.tiles_m - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&map_x); prefetch_tmap(&map_w1); prefetch_tmap(&map_w2);
        for (int s = 0; s < TS_SLOTS; ++s) { mbar_init(&slot_full[s], 1); mbar_init(&slot_empty[s], 1); }
        mbar_init(x_full, 1); mbar_init(x_empty, 1);
        for (int b = 0; b < 2; ++b) { mbar_init(&hacc_full[b], 1); mbar_init(&hts_full[b], 8); mbar_init(&hfree[b], 1); }
        mbar_init(y_full, 1); mbar_init(y_empty, 16);
        mbar_init(&res_full[0], 1); mbar_init(&res_full[1], 1); mbar_init(stage_free, 1);
        prefetch_tmap(&map_res); prefetch_tmap(&map_out);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc<512>(tmem_slot);
    float* sb1 = (float*)(smem + TSB1_OFF);
    for (int i = threadIdx.x; i < nch * 128; i += TS_THREADS) sb1[i] = p.b1[i];        // weights: not produced by the predecessor
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    pdl_trigger();
    pdl_wait();
    const uint32_t tmem_base = *tmem_slot;
    constexpr uint32_t TM_Y = 0, TM_H = 256;

    if (warp == 0) {
        if (lane == 0) {
            int slot = 0; uint32_t sphase = 0;
            auto next_slot = [&](const CUtensorMap* m, int c0, int c1) {
                mbar_wait(&slot_empty[slot], sphase ^ 1);
                mbar_expect_tx(&slot_full[slot], FF_SLOT_BYTES);
                tma_load_2d(m, smem + TSW_OFF + slot * FF_SLOT_BYTES, &slot_full[slot], c0, c1);
                if (++slot == TS_SLOTS) { slot = 0; sphase ^= 1; }
            };
            for (int ti = 0; ti < iters; ++ti) {
                const int t = (int)blockIdx.x + ti * (int)gridDim.x;
                mbar_wait(x_empty, (ti & 1) ^ 1);
                mbar_expect_tx(x_full, 4 * FF_SLOT_BYTES);
                for (int kb = 0; kb < 4; ++kb) tma_load_2d(&map_x, smem + FX_OFF + kb * FF_SLOT_BYTES, x_full, kb * 64, t * 128);
                for (int s = 0; s <= nch; ++s) {
                    if (s < nch)
                        for (int kb = 0; kb < 4; ++kb) next_slot(&map_w1, kb * 64, s * 128);
                    if (s >= 1) {
                        const int j = s - 1;
                        for (int kb = 0; kb < 2; ++kb)
                            for (int nh = 0; nh < 2; ++nh) next_slot(&map_w2, j * 128 + kb * 64, nh * 128);
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(128, 128);
            int slot = 0; uint32_t sphase = 0;
            for (int ti = 0; ti < iters; ++ti) {
                mbar_wait(x_full, ti & 1);
                tc_fence_after();
                for (int s = 0; s <= nch; ++s) {
                    if (s < nch) {
                        const int b = s & 1;
                        const uint32_t u = (uint32_t)(ti * half_uses + (s >> 1));
                        mbar_wait(&hfree[b], (u & 1) ^ 1);               // GEMM2 of this buffer's previous chunk has retired
                        tc_fence_after();
                        const uint32_t d = tmem_base + TM_H + (uint32_t)(b * 128);
                        for (int kb = 0; kb < 4; ++kb) {
                            mbar_wait(&slot_full[slot], sphase);
                            tc_fence_after();
                            const uint32_t sa = smem_u32(smem + FX_OFF + kb * FF_SLOT_BYTES);
                            const uint32_t sb = smem_u32(smem + TSW_OFF + slot * FF_SLOT_BYTES);
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                umma_bf16(d, make_smem_desc(sa + k * 32), make_smem_desc(sb + k * 32), idesc, (kb > 0 || k > 0) ? 1u : 0u);
                            umma_commit(&slot_empty[slot]);
                            if (++slot == TS_SLOTS) { slot = 0; sphase ^= 1; }
                        }
                        umma_commit(&hacc_full[b]);
                    }
                    if (s >= 1) {
                        const int j = s - 1, b = j & 1;
                        const uint32_t u = (uint32_t)(ti * half_uses + (j >> 1));
                        if (j == 0) { mbar_wait(y_empty, (ti & 1) ^ 1); }
                        mbar_wait(&hts_full[b], u & 1);
                        tc_fence_after();
                        const uint32_t ta = tmem_base + TM_H + (uint32_t)(b * 128);      // bf16 pairs: 8 columns per K = 16
                        for (int kb = 0; kb < 2; ++kb) {
                            for (int nh = 0; nh < 2; ++nh) {
                                mbar_wait(&slot_full[slot], sphase);
                                tc_fence_after();
                                const uint32_t sb = smem_u32(smem + TSW_OFF + slot * FF_SLOT_BYTES);
                                const uint32_t d = tmem_base + TM_Y + (uint32_t)(nh * 128);
#pragma unroll
                                for (int k = 0; k < 4; ++k)
                                    umma_bf16_ts(d, ta + (uint32_t)((kb * 4 + k) * 8), make_smem_desc(sb + k * 32), idesc,
                                                 (j > 0 || kb > 0 || k > 0) ? 1u : 0u);
                                umma_commit(&slot_empty[slot]);
                                if (++slot == TS_SLOTS) { slot = 0; sphase ^= 1; }
                            }
                        }
                        umma_commit(&hfree[b]);
                        if (j == nch - 1) umma_commit(y_full);
                    }
                }
            }
        }
    } else {
        // ===== epilogue warps 2..17: group g = (warp - 2) >> 3 takes the chunks with j & 1 == g (accumulator buffer g); inside a
        // group, warp & 3 = TMEM lane quadrant and ((warp - 2) & 7) >> 2 = which 64 of the 128 hidden columns.  The two warps
        // of a quadrant meet at a named barrier between their loads and their in-place stores. =====
        const int e = warp - 2;
        const int quad = warp & 3, grp = e >> 3, half = (e & 7) >> 2;
        const int r = quad * 32 + lane;
        const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
        const int bar_id = 1 + grp * 4 + quad;
        for (int ti = 0; ti < iters; ++ti) {
            const int t = (int)blockIdx.x + ti * (int)gridDim.x;
            for (int j = grp; j < nch; j += 2) {
                const uint32_t u = (uint32_t)(ti * half_uses + (j >> 1));
                mbar_wait(&hacc_full[grp], u & 1);
                tc_fence_after();
                const uint32_t ta = lane_base + TM_H + (uint32_t)(grp * 128);
                uint32_t a0[32], a1[32];
                tmem_ld32_nowait(ta + (uint32_t)(half * 64), a0);
                tmem_ld32_nowait(ta + (uint32_t)(half * 64 + 32), a1);
                tmem_ld_wait();
                const float* bias = sb1 + j * 128 + half * 64;
                uint32_t w[32];
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    const float2 bb = *(reinterpret_cast<const float2*>(bias) + q);
                    const __nv_bfloat162 hv = __floats2bfloat162_rn(fmaxf(__uint_as_float(a0[2 * q]) + bb.x, 0.f),
                                                                    fmaxf(__uint_as_float(a0[2 * q + 1]) + bb.y, 0.f));
                    w[q] = *reinterpret_cast<const uint32_t*>(&hv);
                }
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    const float2 bb = *(reinterpret_cast<const float2*>(bias + 32) + q);
                    const __nv_bfloat162 hv = __floats2bfloat162_rn(fmaxf(__uint_as_float(a1[2 * q]) + bb.x, 0.f),
                                                                    fmaxf(__uint_as_float(a1[2 * q + 1]) + bb.y, 0.f));
                    w[16 + q] = *reinterpret_cast<const uint32_t*>(&hv);
                }
                // both warps of this quadrant have read their fp32 columns: the packed tile may overwrite columns [0, 64)
                asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
                tmem_st32(ta + (uint32_t)(half * 32), w);
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&hts_full[grp]);
            }
            // ---- output tile: group g handles columns [128 g, 128 g + 128) = four 32-column chunks, staged through the x region
            // (free once every GEMM1 of the tile has retired): TMA prefetches the residual half-tile into it, every thread adds
            // its row's accumulator + b2 in place, TMA stores the half-tile.  Group 1 reuses the area after group 0's stores
            // have been read out; the next x tile may only be loaded after that too (x_empty is arrived here). ----
            mbar_wait(y_full, ti & 1);
            tc_fence_after();
            const bool leader = (e & 7) == 0 && lane == 0;
            uint8_t* stage = smem + FX_OFF;
            if (leader) {
                if (grp == 1) mbar_wait(stage_free, ti & 1);
                mbar_expect_tx(&res_full[grp], 4 * FF_SLOT_BYTES);
                for (int c = 0; c < 4; ++c)
                    tma_load_2d(&map_res, stage + c * FF_SLOT_BYTES, &res_full[grp], grp * 128 + c * 32, t * 128);
            }
            uint32_t acc0[32], acc1[32];
            tmem_ld32_nowait(lane_base + TM_Y + (uint32_t)(grp * 128 + half * 64), acc0);
            tmem_ld32_nowait(lane_base + TM_Y + (uint32_t)(grp * 128 + half * 64 + 32), acc1);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(y_empty);                 // the accumulator is in registers: GEMM2 of the next tile may start
            mbar_wait(&res_full[grp], ti & 1);
            epilogue_slab<float, FF_SLOT_BYTES>(acc0, half * 2, grp * 128 + half * 64, nullptr, p.b2, true, 0, stage, r, r & 7);
            epilogue_slab<float, FF_SLOT_BYTES>(acc1, half * 2 + 1, grp * 128 + half * 64 + 32, nullptr, p.b2, true, 0, stage, r, r & 7);
            fence_async_smem();
            asm volatile("bar.sync %0, 256;" ::"r"(9 + grp) : "memory");
            if (leader) {
                for (int c = 0; c < 4; ++c)
                    tma_store_2d(&map_out, stage + c * FF_SLOT_BYTES, grp * 128 + c * 32, t * 128);
                tma_store_commit();
                tma_store_wait_read0();
                if (grp == 0) mbar_arrive(stage_free); else mbar_arrive(x_empty);
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

// The fused kernel takes ~48 us per 128-row tile and SM (49.5 us at 148 tiles, 89.5 at 248, 96.6 at 296), the two GEMMs it
// replaces ~0.39 us per tile in total (97 us at 248 tiles): it pays from one tile per SM on, unless the last round of tiles
// would leave more than half of the SMs idle (e.g. 194 tiles: ~82 vs 76 us).
bool ffn_fused_preferred(int64_t M)
{
    const int64_t tiles = ceil_div(M, 128), sms = num_sms();
    if (tiles < sms) return false;
    const int64_t rem = tiles % sms;
    return rem == 0 || 2 * rem >= sms;
}

bool ffn_fused_supported(int d, int ff, int64_t M, const void* x, const void* w1, const void* w2, const float* residual, const float* out,
                         int ld_res, int ldo)
{
    if (d != 256 || ff % 256 != 0 || ff < 256 || ff > 3072 || M < 1) return false;
    if (((uintptr_t)x & 15) || ((uintptr_t)w1 & 15) || ((uintptr_t)w2 & 15) || ((uintptr_t)residual & 15) || ((uintptr_t)out & 15)) return false;
    return ld_res % 4 == 0 && ldo % 4 == 0;
}

// out[M, 256] (fp32) = residual + relu(x W1^T + b1) W2^T + b2; x [M, 256] bf16 (row stride 256), W1 [ff, 256], W2 [256, ff] bf16
int launch_ffn_fused(const void* x, const void* w1, const float* b1, const void* w2, const float* b2, const float* residual, int ld_res,
                     float* out, int ldo, int64_t M, int ff, cudaStream_t stream)
{
    SEDT_REQUIRE(ffn_fused_supported(256, ff, M, x, w1, w2, residual, out, ld_res, ldo), "ffn_fused: unsupported shape / alignment");
    SEDT_TRY(tc_init());
    CUtensorMap mx, m1, m2, mres, mout;
    const uint32_t box[2] = {64u, 128u};
    {
        const uint64_t dims[2] = {256, (uint64_t)M}; const uint64_t strides[1] = {256 * 2};
        SEDT_TRY(encode_map(&mx, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, x, 2, dims, strides, box));
    }
    {
        const uint64_t dims[2] = {256, (uint64_t)ff}; const uint64_t strides[1] = {256 * 2};
        SEDT_TRY(encode_map(&m1, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, w1, 2, dims, strides, box));
    }
    {
        const uint64_t dims[2] = {(uint64_t)ff, 256}; const uint64_t strides[1] = {(uint64_t)ff * 2};
        SEDT_TRY(encode_map(&m2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, w2, 2, dims, strides, box));
    }
    {
        const uint32_t obox[2] = {32u, 128u};                     // 32 fp32 columns = one 128-byte swizzle row
        const uint64_t odims[2] = {256, (uint64_t)M};
        const uint64_t rstr[1] = {(uint64_t)ld_res * 4}, ostr[1] = {(uint64_t)ldo * 4};
        SEDT_TRY(encode_map(&mres, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, residual, 2, odims, rstr, obox));
        SEDT_TRY(encode_map(&mout, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, out, 2, odims, ostr, obox));
    }
    FfnParams p;
    p.b1 = b1; p.b2 = b2; p.residual = residual; p.out = out; p.ld_res = ld_res; p.ldo = ldo;
    p.M = (int)M; p.nch = ff / 128; p.tiles_m = (int)ceil_div(M, 128);
    static bool attr_set = false;
    if (!attr_set) {
        SEDT_CHECK_CUDA(cudaFuncSetAttribute(ffn_fused_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TS_SMEM));
        attr_set = true;
    }
    const int grid = std::min(p.tiles_m, num_sms());
    ProfScope _prof(PROF_GEMM_TC, stream);
    SEDT_CHECK_CUDA(launch_pdl(ffn_fused_ts_kernel, dim3((unsigned)grid), dim3(TS_THREADS), TS_SMEM, stream, 1, mx, m1, m2, mres, mout, p));
    SEDT_COUNT_KIND(KK_FFN_FUSED);
    SEDT_CHECK_CUDA(cudaGetLastError());
    return SEDT_OK;
}

}  // namespace sedt
