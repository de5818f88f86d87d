// Fused tail of a layer1 bottleneck (torchvision resnet.py:150-161 via sedt/backbone.py:68):
//
//     out = relu( bn3(conv3( relu(bn2(conv2(h1))) )) + identity )          conv2: 3x3, 64 -> 64;  conv3: 1x1, 64 -> 256
//
// as ONE persistent kernel.  The unfused path writes the 64-channel h2 tensor to HBM and reads it back (2 x 65 MB per block at
// B = 256) and pays two launches whose tails do not overlap; conv3 alone is HBM-bound (585 MB in 110 us).  Here a CTA walks over
// 128-pixel tiles (8 rows x 16 columns) and per tile
//
//   TMA      three halo boxes of h1 ((bh + 2) rows, one per horizontal tap shift: the HALO trick of gemm_tc4.cu)
//   tcgen05  conv2: 9 taps x 4 k-steps of [128 x 64 x 16] MMAs against the resident 72 KiB filter   -> TMEM acc2[buf] (64 columns)
//   rows     acc2 + bias2 -> ReLU -> bf16, packed IN PLACE in acc2's columns [0, 32)                  (the h2 tile never leaves TMEM)
//   tcgen05  conv3: 4 k-steps of [128 x 256 x 16] MMAs, A = the packed h2 tile in TMEM, B = the resident 32 KiB filter -> acc3 (256 columns)
//   rows     acc3 + bias3 + identity + ReLU -> bf16, 64 columns at a time through a ring of three 16 KiB staging chunks
//            (store warp: TMA-prefetches the identity chunk into the buffer, TMA-stores the finished chunk)
//
// Software pipeline: conv2 of tile i + 1 is issued before conv3 of tile i, so the tensor core works on the next tile while the
// row warps turn acc2 into h2; acc2 is double buffered, acc3 single (128 + 256 TMEM columns).  Shared memory: 3 x 20 KiB halo
// ring + 72 KiB W2 + 32 KiB W3 + 3 x 16 KiB chunk ring = 212 KiB.  Rounding points are those of the two separate launches
// (h2 and the output in bf16), so the results agree up to fp32 summation order.
//
// 352 threads: warp 0 TMA producer, warp 1 MMA issuer, warps 2..9 row warps (warp & 3 = TMEM lane quadrant, (warp - 2) >> 2 =
// column half), warp 10 store warp.
#include "tc_common.cuh"
#include <algorithm>
#include <cstdlib>

namespace sedt {
namespace {

using namespace tc;

constexpr int BF_THREADS = 352;
constexpr int BF_STAGES = 3;                       // one tile's three halo boxes
constexpr int BF_HALO_BYTES = 20480;               // (bh + 2) * bw pixels * 128 B <= 160 pixels
constexpr int BF_W2_BYTES = 9 * 64 * 64 * 2;       // 72 KiB: [tap][64 out][64 in]
constexpr int BF_W3_BYTES = 256 * 64 * 2;          // 32 KiB: [256 out][64 in]
constexpr int BF_CHUNK_BYTES = 16384;              // 128 pixels x 64 bf16 columns
constexpr int BF_NCHUNK = 4;                       // 256 output columns
constexpr int BF_OB = 3;                           // staging chunks in flight
constexpr int BF_A_OFF = 0;
constexpr int BF_W2_OFF = BF_STAGES * BF_HALO_BYTES;
constexpr int BF_W3_OFF = BF_W2_OFF + BF_W2_BYTES;
constexpr int BF_OUT_OFF = BF_W3_OFF + BF_W3_BYTES;
constexpr int BF_BAR_OFF = BF_OUT_OFF + BF_OB * BF_CHUNK_BYTES;
constexpr int BF_NBARS = 2 * BF_STAGES + 1 + 2 + 2 + 2 + 2 + 2 * BF_OB;
constexpr int BF_SMEM = BF_BAR_OFF + BF_NBARS * 8 + 16 + 1024;
static_assert(BF_OUT_OFF % 1024 == 0 && BF_W2_OFF % 1024 == 0 && BF_W3_OFF % 1024 == 0, "swizzled tiles need 1024-byte alignment");
static_assert(BF_SMEM <= 232448, "shared memory budget exceeded");

constexpr uint32_t TM_ACC2 = 0;        // two conv2 accumulators of 64 columns; the packed h2 tile overwrites columns [0, 32) of its buffer
constexpr uint32_t TM_ACC3 = 128;      // conv3 accumulator, 256 columns

struct BfParams {
    const float* bias2;        // [64]  folded bn2 bias
    const float* bias3;        // [256] folded bn3 bias
    const __nv_bfloat16* residual;   // identity tensor [B, H, W, ld_res] (RES_REG variant: read straight into registers)
    int ld_res, H, W;
    int bw, bh, tiles_w, tiles_h, total_tiles;
};

// RES_REG: the identity chunk is not TMA-prefetched into the staging buffer (where its HBM latency sits inside the 3-deep chunk
// ring) but loaded by the row threads straight into registers at the start of stage 2, one tile ahead of its use
template <bool RES_REG>
__global__ void __launch_bounds__(BF_THREADS, 1)
bneck_tail_kernel(const __grid_constant__ CUtensorMap map_halo, const __grid_constant__ CUtensorMap map_w2,
                  const __grid_constant__ CUtensorMap map_w3, const __grid_constant__ CUtensorMap map_out,
                  const __grid_constant__ CUtensorMap map_res, const __grid_constant__ BfParams p)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* full_bar = (uint64_t*)(smem + BF_BAR_OFF);
    uint64_t* empty_bar = full_bar + BF_STAGES;
    uint64_t* w_full = empty_bar + BF_STAGES;
    uint64_t* acc2_full = w_full + 1;            // [2] MMA -> rows: conv2 accumulator ready
    uint64_t* h2_ready = acc2_full + 2;          // [2] rows -> MMA: packed h2 tile is in TMEM
    uint64_t* acc2_free = h2_ready + 2;          // [2] conv3 has read the h2 tile: the buffer may be overwritten
    uint64_t* acc3_full = acc2_free + 2;         // MMA -> rows
    uint64_t* acc3_empty = acc3_full + 1;        // rows -> MMA
    uint64_t* buf_ready = acc3_empty + 1;        // [OB] identity chunk has landed in the staging buffer
    uint64_t* buf_full = buf_ready + BF_OB;      // [OB] rows -> store warp
    uint32_t* tmem_slot = (uint32_t*)(buf_full + BF_OB);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&map_halo); prefetch_tmap(&map_w2); prefetch_tmap(&map_w3); prefetch_tmap(&map_out); prefetch_tmap(&map_res);
        for (int s = 0; s < BF_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(w_full, 1);
        for (int s = 0; s < 2; ++s) { mbar_init(&acc2_full[s], 1); mbar_init(&h2_ready[s], 8); mbar_init(&acc2_free[s], 1); }
        mbar_init(acc3_full, 1); mbar_init(acc3_empty, 8);
        for (int s = 0; s < BF_OB; ++s) { mbar_init(&buf_ready[s], 1); mbar_init(&buf_full[s], 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    pdl_trigger();
    pdl_wait();
    const uint32_t tmem_base = *tmem_slot;
    const int first = (int)blockIdx.x, stride = (int)gridDim.x;

    auto tile_coords = [&](int t, int& w0, int& h0, int& n0) {
        w0 = (t % p.tiles_w) * p.bw;
        h0 = ((t / p.tiles_w) % p.tiles_h) * p.bh;
        n0 = t / (p.tiles_w * p.tiles_h);
    };

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            mbar_expect_tx(w_full, BF_W2_BYTES + BF_W3_BYTES);
            for (int tap = 0; tap < 9; ++tap) tma_load_2d(&map_w2, smem + BF_W2_OFF + tap * 8192, w_full, tap * 64, 0);
            tma_load_2d(&map_w3, smem + BF_W3_OFF, w_full, 0, 0);
            const uint32_t halo_bytes = (uint32_t)((p.bh + 2) * p.bw * 128);
            int stage = 0; uint32_t phase = 0;
            for (int t = first; t < p.total_tiles; t += stride) {
                int w0, h0, n0;
                tile_coords(t, w0, h0, n0);
                for (int dwi = 0; dwi < 3; ++dwi) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    mbar_expect_tx(&full_bar[stage], halo_bytes);
                    tma_load_4d(&map_halo, smem + BF_A_OFF + stage * BF_HALO_BYTES, &full_bar[stage], 0, w0 + dwi - 1, h0 - 1, n0);
                    if (++stage == BF_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc2 = make_idesc(128, 64);
            constexpr uint32_t idesc3 = make_idesc(128, 256);
            const uint32_t sW2 = smem_u32(smem + BF_W2_OFF), sW3 = smem_u32(smem + BF_W3_OFF);
            int stage = 0; uint32_t phase = 0;
            mbar_wait(w_full, 0);
            tc_fence_after();
            // conv3 of local tile j: A = the packed h2 tile in acc2[j & 1]
            auto conv3 = [&](int j) {
                const int as = j & 1;
                mbar_wait(&h2_ready[as], (uint32_t)(j >> 1) & 1);
                mbar_wait(acc3_empty, ((uint32_t)j & 1) ^ 1);          // the previous tile's accumulator has been drained
                tc_fence_after();
                const uint32_t a = tmem_base + TM_ACC2 + (uint32_t)(as * 64);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_bf16_ts(tmem_base + TM_ACC3, a + (uint32_t)(8 * k), make_smem_desc(sW3 + k * 32), idesc3, k > 0 ? 1u : 0u);
                umma_commit(&acc2_free[as]);
                umma_commit(acc3_full);
            };
            int li = 0;
            for (int t = first; t < p.total_tiles; t += stride, ++li) {
                const int as = li & 1;
                mbar_wait(&acc2_free[as], (((uint32_t)li >> 1) & 1) ^ 1);   // conv3 of tile li - 2 has consumed this buffer
                tc_fence_after();
                const uint32_t d2 = tmem_base + TM_ACC2 + (uint32_t)(as * 64);
                for (int dwi = 0; dwi < 3; ++dwi) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t base = smem_u32(smem + BF_A_OFF + stage * BF_HALO_BYTES);
#pragma unroll
                    for (int dhi = 0; dhi < 3; ++dhi) {
                        const uint32_t sa = base + (uint32_t)(dhi * p.bw * 128);            // tap (dhi, dwi): rows shifted by dhi
                        const uint32_t sb = sW2 + (uint32_t)((dhi * 3 + dwi) * 8192);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_bf16(d2, make_smem_desc(sa + k * 32), make_smem_desc(sb + k * 32), idesc2, (dwi > 0 || dhi > 0 || k > 0) ? 1u : 0u);
                    }
                    umma_commit(&empty_bar[stage]);
                    if (++stage == BF_STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit(&acc2_full[as]);
                if (li > 0) conv3(li - 1);                              // the previous tile's h2 is ready by now (or soon)
            }
            if (li > 0) conv3(li - 1);
        }
    } else if (warp == 10) {
        // ===================== store warp: identity prefetch, chunk stores, staging recycling =====================
        if (lane == 0) {
            uint8_t* out_base = smem + BF_OUT_OFF;
            int ntiles = 0;
            for (int t = first; t < p.total_tiles; t += stride) ++ntiles;
            const int nq = ntiles * BF_NCHUNK;
            auto prefetch = [&](int q) {                                 // global chunk q = local tile q / 4, columns (q % 4) * 64
                int w0, h0, n0;
                tile_coords(first + (q / BF_NCHUNK) * stride, w0, h0, n0);
                const int ob = q % BF_OB;
                if constexpr (RES_REG) {
                    mbar_arrive(&buf_ready[ob]);                         // the staging buffer is free again, nothing to load
                } else {
                    mbar_expect_tx(&buf_ready[ob], BF_CHUNK_BYTES);
                    tma_load_4d(&map_res, out_base + ob * BF_CHUNK_BYTES, &buf_ready[ob], (q % BF_NCHUNK) * 64, w0, h0, n0);
                }
            };
            for (int q = 0; q < BF_OB && q < nq; ++q) prefetch(q);
            for (int q = 0; q < nq; ++q) {
                const int ob = q % BF_OB;
                int w0, h0, n0;
                tile_coords(first + (q / BF_NCHUNK) * stride, w0, h0, n0);
                mbar_wait(&buf_full[ob], (uint32_t)(q / BF_OB) & 1);
                tma_store_4d(&map_out, out_base + ob * BF_CHUNK_BYTES, (q % BF_NCHUNK) * 64, w0, h0, n0);
                tma_store_commit();
                tma_store_wait_read0();                                  // the staging buffer has been read out
                if (q + BF_OB < nq) prefetch(q + BF_OB);
            }
        }
    } else {
        // ===================== row warps 2..9 =====================
        const int quad = warp & 3, half = (warp - 2) >> 2;
        const int r = quad * 32 + lane;
        const int sw = r & 7;
        const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
        uint8_t* out_base = smem + BF_OUT_OFF;
        // stage 1 of local tile j: conv2 accumulator -> + bias2 -> ReLU -> bf16, packed in place (this thread: row r, columns
        // [32 half, 32 half + 32) -> packed columns [16 half, 16 half + 16))
        auto stage1 = [&](int j) {
            const int as = j & 1;
            mbar_wait(&acc2_full[as], (uint32_t)(j >> 1) & 1);
            tc_fence_after();
            const uint32_t ta = lane_base + TM_ACC2 + (uint32_t)(as * 64);
            uint32_t acc[32], pk[16];
            tmem_ld32(ta + (uint32_t)(half * 32), acc);
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias2 + half * 32) + j4);
                const __nv_bfloat162 lo = __floats2bfloat162_rn(fmaxf(__uint_as_float(acc[4 * j4]) + b4.x, 0.f),
                                                                fmaxf(__uint_as_float(acc[4 * j4 + 1]) + b4.y, 0.f));
                const __nv_bfloat162 hi = __floats2bfloat162_rn(fmaxf(__uint_as_float(acc[4 * j4 + 2]) + b4.z, 0.f),
                                                                fmaxf(__uint_as_float(acc[4 * j4 + 3]) + b4.w, 0.f));
                pk[2 * j4] = *reinterpret_cast<const uint32_t*>(&lo);
                pk[2 * j4 + 1] = *reinterpret_cast<const uint32_t*>(&hi);
            }
            // both warps of this lane quadrant have read their fp32 columns: the packed tile may overwrite columns [0, 32)
            asm volatile("bar.sync %0, 64;" ::"r"(1 + quad) : "memory");
            tmem_st16(ta + (uint32_t)(half * 16), pk);
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&h2_ready[as]);
        };
        // stage 2 of local tile j: conv3 accumulator + bias3 + identity + ReLU -> bf16 staging chunks (this thread: row r, columns
        // [32 half, 32 half + 32) of every 64-column chunk)
        auto stage2 = [&](int j) {
            uint4 idn[BF_NCHUNK][4];
            if constexpr (RES_REG) {
                // identity: this thread's 32 columns of every chunk of row r, straight from global memory (64 contiguous bytes per chunk)
                int w0, h0, n0;
                tile_coords(first + j * stride, w0, h0, n0);
                const int py = h0 + r / p.bw, px = w0 + r % p.bw;
                const bool valid = py < p.H && px < p.W;
                const __nv_bfloat16* rrow = p.residual + (((size_t)n0 * p.H + py) * p.W + px) * p.ld_res + half * 32;
#pragma unroll
                for (int c = 0; c < BF_NCHUNK; ++c)
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        idn[c][k] = valid ? __ldg(reinterpret_cast<const uint4*>(rrow + c * 64) + k) : make_uint4(0, 0, 0, 0);
            }
            mbar_wait(acc3_full, (uint32_t)j & 1);
            tc_fence_after();
#pragma unroll
            for (int c = 0; c < BF_NCHUNK; ++c) {
                const int q = j * BF_NCHUNK + c, ob = q % BF_OB;
                uint32_t acc[32];
                tmem_ld32(lane_base + TM_ACC3 + (uint32_t)(c * 64 + half * 32), acc);
                if (c == BF_NCHUNK - 1) {                                // the accumulator is in registers: conv3 of the next tile may run
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(acc3_empty);
                }
                mbar_wait(&buf_ready[ob], (uint32_t)(q / BF_OB) & 1);    // staging chunk free (and, without RES_REG, the identity chunk landed)
                if constexpr (RES_REG) {
                    uint8_t* rowp = out_base + ob * BF_CHUNK_BYTES + r * 128;
                    const float* b3 = p.bias3 + c * 64 + half * 32;
#pragma unroll
                    for (int j8 = 0; j8 < 4; ++j8) {
                        const float4 ba = __ldg(reinterpret_cast<const float4*>(b3) + 2 * j8), bb = __ldg(reinterpret_cast<const float4*>(b3) + 2 * j8 + 1);
                        const float bv[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
                        const uint32_t rw[4] = {idn[c][j8].x, idn[c][j8].y, idn[c][j8].z, idn[c][j8].w};
                        uint32_t w[4];
#pragma unroll
                        for (int qq = 0; qq < 4; ++qq) {
                            const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&rw[qq]);
                            const float a = fmaxf(__uint_as_float(acc[8 * j8 + 2 * qq]) + bv[2 * qq] + __low2float(h), 0.f);
                            const float b = fmaxf(__uint_as_float(acc[8 * j8 + 2 * qq + 1]) + bv[2 * qq + 1] + __high2float(h), 0.f);
                            const __nv_bfloat162 o = __floats2bfloat162_rn(a, b);
                            w[qq] = *reinterpret_cast<const uint32_t*>(&o);
                        }
                        *reinterpret_cast<uint4*>(rowp + (((half * 4 + j8) ^ sw) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
                    }
                } else {
                    epilogue_slab<__nv_bfloat16, BF_CHUNK_BYTES>(acc, half, c * 64 + half * 32, nullptr, p.bias3, true, 1,
                                                                 out_base + ob * BF_CHUNK_BYTES, r, sw);
                }
                fence_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&buf_full[ob]);
            }
        };
        int li = 0;
        for (int t = first; t < p.total_tiles; t += stride, ++li) {
            stage1(li);
            if (li > 0) stage2(li - 1);
        }
        if (li > 0) stage2(li - 1);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

}  // namespace

// SEDT_BNECK_FUSED: unset / 1 = fuse conv2 + conv3 of the layer1 bottlenecks, 0 = two launches
bool bneck_tail_enabled()
{
    static const bool on = [] { const char* e = getenv("SEDT_BNECK_FUSED"); return e == nullptr || atoi(e) != 0; }();
    return on;
}

// g2: the 3x3 convolution (64 -> 64, stride 1, dilation 1, bias + ReLU, bf16), g3: the 1x1 convolution that follows it
// (64 -> 256, bias + residual + ReLU, bf16); g2.out is NOT written
bool bneck_tail_supported(const ConvGemm& g2, const ConvGemm& g3)
{
    if (!conv_tc_supported(g2) || !conv_tc_supported(g3)) return false;
    if (g2.in_dt != DT_BF16 || g2.out_dt != DT_BF16 || g3.out_dt != DT_BF16) return false;
    if (g2.R != 3 || g2.S != 3 || g2.stride != 1 || g2.dil != 1 || g2.pad != 1 || g2.Cin != 64 || g2.Cout != 64 || g2.relu != 1) return false;
    if (g3.R != 1 || g3.stride != 1 || g3.Cin != 64 || g3.Cout != 256 || g3.relu != 1 || g3.residual == nullptr) return false;
    if (g2.scale != nullptr || g3.scale != nullptr || g2.bias == nullptr || g3.bias == nullptr || g2.residual != nullptr) return false;
    if (g2.B != g3.B || g2.Ho != g3.H || g2.Wo != g3.W || g3.Ho != g3.H || g3.Wo != g3.W || g2.Ho != g2.H || g2.Wo != g2.W) return false;
    if (g3.lda != 64 || g2.lda != 64) return false;
    // tile geometry as build_problem() derives it: one clip per tile, 8 or 16 pixels wide, 128 pixels
    auto pow2 = [](int v) { int q = 1; while (q < v) q <<= 1; return q; };
    const int bw = pow2(g2.Wo), bh = std::min(pow2(g2.Ho), BLOCK_M / std::max(bw, 1));
    return (bw == 8 || bw == 16) && bw * bh == BLOCK_M;
}

int launch_bneck_tail(const ConvGemm& g2, const ConvGemm& g3, cudaStream_t stream)
{
    SEDT_REQUIRE(bneck_tail_supported(g2, g3), "bneck_tail: unsupported block shape");
    SEDT_TRY(tc_init());
    TcProblem pr;
    SEDT_TRY(build_problem(g2, 64, &pr));                        // tile geometry + the [64 x 576] weight map (64-row boxes per tap)
    const TcParams& q = pr.p;
    SEDT_REQUIRE(q.bn == 1 && (q.bw == 8 || q.bw == 16) && q.bw * q.bh == BLOCK_M, "bneck_tail: tile %d x %d x %d", q.bw, q.bh, q.bn);
    CUtensorMap mhalo, mw3, mo, mr;
    {
        const uint64_t dims[4] = {64, (uint64_t)g2.W, (uint64_t)g2.H, (uint64_t)g2.B};
        const uint64_t strides[3] = {(uint64_t)g2.lda * 2, (uint64_t)g2.W * g2.lda * 2, (uint64_t)g2.H * g2.W * g2.lda * 2};
        const uint32_t box[4] = {64u, (uint32_t)q.bw, (uint32_t)(q.bh + 2), 1u};
        SEDT_TRY(encode_map(&mhalo, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, g2.in, 4, dims, strides, box));
    }
    {
        const uint64_t dims[2] = {64, 256}; const uint64_t strides[1] = {64 * 2};
        const uint32_t box[2] = {64u, 256u};
        SEDT_TRY(encode_map(&mw3, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, g3.w, 2, dims, strides, box));
    }
    SEDT_TRY(encode_out_map(&mo, g3.out, g3.ldc, false, g3, q));
    SEDT_TRY(encode_out_map(&mr, g3.residual, g3.ld_res, false, g3, q));
    BfParams p;
    p.bias2 = g2.bias; p.bias3 = g3.bias; p.bw = q.bw; p.bh = q.bh; p.tiles_w = q.tiles_w; p.tiles_h = q.tiles_h;
    p.total_tiles = pr.tiles_m;
    p.residual = (const __nv_bfloat16*)g3.residual; p.ld_res = g3.ld_res; p.H = g3.Ho; p.W = g3.Wo;
    // SEDT_BNECK_RESREG: 1 = identity through registers, 0 = TMA-prefetched into the staging chunk
    static const bool res_reg = [] { const char* e = getenv("SEDT_BNECK_RESREG"); return e != nullptr && atoi(e) != 0; }();
    static bool attr_set = false;
    if (!attr_set) {
        SEDT_CHECK_CUDA(cudaFuncSetAttribute(bneck_tail_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, BF_SMEM));
        SEDT_CHECK_CUDA(cudaFuncSetAttribute(bneck_tail_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, BF_SMEM));
        attr_set = true;
    }
    const int grid = std::min(pr.tiles_m, num_sms());
    ProfScope _prof(PROF_GEMM_TC, stream);
    if (res_reg)
        SEDT_CHECK_CUDA(launch_pdl(bneck_tail_kernel<true>, dim3((unsigned)grid), dim3(BF_THREADS), BF_SMEM, stream, 1, mhalo, pr.map_b, mw3, mo, mr, p));
    else
        SEDT_CHECK_CUDA(launch_pdl(bneck_tail_kernel<false>, dim3((unsigned)grid), dim3(BF_THREADS), BF_SMEM, stream, 1, mhalo, pr.map_b, mw3, mo, mr, p));
    SEDT_COUNT_KIND(KK_BOTTLENECK_FUSED);
    SEDT_CHECK_CUDA(cudaGetLastError());
    return SEDT_OK;
}

}  // namespace sedt
