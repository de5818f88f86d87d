// Weight-stationary variant of the persistent implicit-GEMM kernel for short reductions (K <= 256: the 1x1
// convolutions of layer1-3 with 64..256 input channels, the Q/K/V/out projections and FFN linear1).
//
// With K this short a 128 x 128 output tile needs as many operand bytes as it produces, and in the streaming
// kernel every tile re-fetches its weight tile from L2 (ncu: ~9 TB/s of TMA traffic, MMA warp waiting on
// full barriers).  Here each CTA owns ONE N tile: its [128 x K] weights are loaded once and stay in shared
// memory while the CTA walks over M tiles, so only the A boxes (16 KiB per 64-wide k block) stream through
// a deeper ring.  Everything else (TMEM double buffering, 8 epilogue warps, store warp, TMA-staged residual
// and output) is as in gemm_tc2.cu.
#include "tc_common.cuh"
#include <cstdlib>

namespace sedt {
namespace {

using namespace tc;

constexpr int NUM_THREADS4 = 352;

// HALO (3x3, stride 1, dilation 1, 64 input channels, 8- or 16-pixel-wide tiles): instead of nine [128 px x 64 ch] boxes per
// tile - which re-read nearly the same pixels nine times from L2 (ncu: these launches sat at 20 % tensor pipe, 13 % DRAM,
// i.e. on the L2 -> SM path) - the producer loads three boxes of bh + 2 rows, one per horizontal tap shift.  The three
// vertical taps of a shift are the same shared-memory image read at row offsets 0, bw, 2 bw pixels: whole 1024-byte swizzle
// atoms, so the UMMA descriptor just starts further down.  A traffic per tile: 3 x (bh + 2) / bh boxes instead of 9.
constexpr int HALO_SLOT_BYTES = 20480;                                      // (bh + 2) * bw pixels * 128 B <= 160 pixels

template <int BLOCK_N, int STAGES, int OUT_BUFS, typename TO, int MAX_KB, int HALO = 0>
struct Smem4 {
    static constexpr int A_SLOT = HALO ? HALO_SLOT_BYTES : A_STAGE_BYTES;
    static constexpr int B_KB_BYTES = BLOCK_N * BLOCK_K * 2;                // one 64-wide k block of the weight tile
    static constexpr int CHUNK_COLS = 128 / (int)sizeof(TO);
    static constexpr int NCHUNK = BLOCK_N / CHUNK_COLS;
    static constexpr int CHUNK_BYTES = BLOCK_M * 128;
    static constexpr int OUT_BYTES = NCHUNK * CHUNK_BYTES;
    static constexpr int B_OFFSET = STAGES * A_SLOT;                        // A ring first, then the resident weights
    static constexpr int OUT_OFFSET = B_OFFSET + MAX_KB * B_KB_BYTES;
    static constexpr int BAR_OFFSET = OUT_OFFSET + OUT_BUFS * OUT_BYTES;
    static constexpr int NBARS = 2 * STAGES + 4 + 2 * OUT_BUFS + 1;
    static constexpr int TOTAL = BAR_OFFSET + NBARS * 8 + 16 + 1024;
};

template <int BLOCK_N, int STAGES, int OUT_BUFS, typename TO, int MAX_KB, int HALO = 0>
__global__ void __launch_bounds__(NUM_THREADS4, 1)
conv_tc4_kernel(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_a1,
                const __grid_constant__ CUtensorMap map_a2, const __grid_constant__ CUtensorMap map_a3,
                const __grid_constant__ CUtensorMap map_b, const __grid_constant__ CUtensorMap map_out,
                const __grid_constant__ CUtensorMap map_res, const __grid_constant__ TcParams p,
                const int tiles_nc, const int total_tiles)
{
    using L = Smem4<BLOCK_N, STAGES, OUT_BUFS, TO, MAX_KB, HALO>;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // keep the pointer derived from the __shared__ symbol so that staging traffic compiles to LDS/STS
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* full_bar = (uint64_t*)(smem + L::BAR_OFFSET);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* acc_full = empty_bar + STAGES;          // [2]  MMA -> epilogue
    uint64_t* acc_empty = acc_full + 2;               // [2]  epilogue -> MMA
    uint64_t* buf_ready = acc_empty + 2;              // [OUT_BUFS] store warp / residual TMA -> epilogue
    uint64_t* buf_full = buf_ready + OUT_BUFS;        // [OUT_BUFS] epilogue -> store warp
    uint64_t* b_full = buf_full + OUT_BUFS;           // resident weight tile has landed
    uint32_t* tmem_slot = (uint32_t*)(b_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cpb = p.Cin / BLOCK_K;
    const int num_kb = p.ntaps * cpb;
    const bool has_res = p.residual != nullptr;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&map_a0); prefetch_tmap(&map_b); prefetch_tmap(&map_out);
        if (has_res) prefetch_tmap(&map_res);
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(b_full, 1);
        for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 8); }
        for (int s = 0; s < OUT_BUFS; ++s) { mbar_init(&buf_ready[s], 1); mbar_init(&buf_full[s], 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc<2 * BLOCK_N>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    // PDL: the successor may be scheduled from here on (this CTA owns its TMEM columns already, so a co-resident
    // successor CTA can never starve it); everything above overlapped the predecessor's tail, whose outputs are our operands
    pdl_trigger();
    pdl_wait();
    const uint32_t tmem_base = *tmem_slot;
    // schedule: this CTA owns ONE N tile (its weights stay in shared memory) and walks over M tiles
    const int n_tile = (int)blockIdx.x % tiles_nc;
    const int first_item = (int)blockIdx.x / tiles_nc, item_stride = (int)gridDim.x / tiles_nc;
    const int col0_fixed = n_tile * BLOCK_N;
    uint8_t* sB = smem + L::B_OFFSET;

    auto tile_coords = [&](int m_tile, int& w0, int& h0, int& n0, int& col0) {
        const int tw = m_tile % p.tiles_w;
        const int th = (m_tile / p.tiles_w) % p.tiles_h;
        const int tn = m_tile / (p.tiles_w * p.tiles_h);
        w0 = tw * p.bw; h0 = th * p.bh; n0 = tn * p.bn; col0 = col0_fixed;
    };

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            // the weight tile [BLOCK_N x K] once
            mbar_expect_tx(b_full, (uint32_t)(num_kb * L::B_KB_BYTES));
            for (int kb = 0; kb < num_kb; ++kb) {
                const int tap = kb / cpb, c0 = (kb - tap * cpb) * BLOCK_K;
                tma_load_2d(&map_b, sB + kb * L::B_KB_BYTES, b_full, tap * p.Cin + c0, col0_fixed);
            }
            int stage = 0; uint32_t phase = 0;
            for (int t = first_item; t < total_tiles; t += item_stride) {
                int w0, h0, n0, col0;
                tile_coords(t, w0, h0, n0, col0);
                if constexpr (HALO) {
                    // map_a3: the same activation with a (64 ch, bw, bh + 2, 1) box; rows above / below the clip and the
                    // columns left / right of it are zero-filled by TMA = the convolution's zero padding
                    const uint32_t halo_bytes = (uint32_t)((p.bh + 2) * p.bw * 128);
                    for (int dwi = 0; dwi < 3; ++dwi) {
                        mbar_wait(&empty_bar[stage], phase ^ 1);
                        mbar_expect_tx(&full_bar[stage], halo_bytes);
                        tma_load_4d(&map_a3, smem + stage * L::A_SLOT, &full_bar[stage], 0, w0 + dwi - 1, h0 - 1, n0);
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                    continue;
                }
                for (int tap = 0; tap < p.ntaps; ++tap) {
                    const int mi = p.tap_map[tap];
                    const CUtensorMap* ma = mi == 0 ? &map_a0 : (mi == 1 ? &map_a1 : (mi == 2 ? &map_a2 : &map_a3));
                    const int cw = w0 + p.tap_dw[tap], ch = h0 + p.tap_dh[tap];
                    for (int c0 = 0; c0 < p.Cin; c0 += BLOCK_K) {
                        mbar_wait(&empty_bar[stage], phase ^ 1);
                        mbar_expect_tx(&full_bar[stage], A_STAGE_BYTES);
                        tma_load_4d(ma, smem + stage * L::A_SLOT, &full_bar[stage], c0, cw, ch, n0);
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one thread) =====
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(BLOCK_M, BLOCK_N);
            int stage = 0; uint32_t phase = 0;
            int li = 0;
            mbar_wait(b_full, 0);
            for (int t = first_item; t < total_tiles; t += item_stride, ++li) {
                const int as = li & 1;
                mbar_wait(&acc_empty[as], ((li >> 1) & 1) ^ 1);      // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(as * BLOCK_N);
                if constexpr (HALO) {
                    for (int dwi = 0; dwi < 3; ++dwi) {
                        mbar_wait(&full_bar[stage], phase);
                        tc_fence_after();
                        const uint32_t base = smem_u32(smem + stage * L::A_SLOT);
#pragma unroll
                        for (int dhi = 0; dhi < 3; ++dhi) {
                            const uint32_t sa = base + (uint32_t)(dhi * p.bw * 128);       // tap (dhi, dwi): rows shifted by dhi
                            const uint32_t sb = smem_u32(sB + (dhi * 3 + dwi) * L::B_KB_BYTES);
#pragma unroll
                            for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
                                umma_bf16(tmem_d, make_smem_desc(sa + k * UMMA_K * 2), make_smem_desc(sb + k * UMMA_K * 2), idesc,
                                          (dwi > 0 || dhi > 0 || k > 0) ? 1u : 0u);
                        }
                        umma_commit(&empty_bar[stage]);
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                    umma_commit(&acc_full[as]);
                    continue;
                }
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * L::A_SLOT);
                    const uint32_t sb = smem_u32(sB + kb * L::B_KB_BYTES);
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                        umma_bf16(tmem_d, make_smem_desc(sa + k * UMMA_K * 2), make_smem_desc(sb + k * UMMA_K * 2), idesc,
                                  (kb > 0 || k > 0) ? 1u : 0u);
                    }
                    umma_commit(&empty_bar[stage]);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit(&acc_full[as]);
            }
        }
    } else if (warp == 10) {
        // ===== store warp: TMA stores of finished tiles, residual prefetch, staging-buffer recycling =====
        if (lane == 0) {
            uint8_t* out_base = smem + L::OUT_OFFSET;
            auto make_ready = [&](int t, int buf) {      // staging buffer `buf` becomes usable for tile t
                if (has_res) {
                    int w0, h0, n0, col0;
                    tile_coords(t, w0, h0, n0, col0);
                    mbar_expect_tx(&buf_ready[buf], L::OUT_BYTES);
#pragma unroll
                    for (int c = 0; c < L::NCHUNK; ++c)
                        tma_load_4d(&map_res, out_base + buf * L::OUT_BYTES + c * L::CHUNK_BYTES, &buf_ready[buf],
                                    col0 + c * L::CHUNK_COLS, w0, h0, n0);
                } else {
                    mbar_arrive(&buf_ready[buf]);
                }
            };
            {
                int t = first_item;
                for (int k = 0; k < OUT_BUFS && t < total_tiles; ++k, t += item_stride) make_ready(t, k);
            }
            int li = 0;
            for (int t = first_item; t < total_tiles; t += item_stride, ++li) {
                const int ob = li % OUT_BUFS;
                int w0, h0, n0, col0;
                tile_coords(t, w0, h0, n0, col0);
                mbar_wait(&buf_full[ob], (li / OUT_BUFS) & 1);
#pragma unroll
                for (int c = 0; c < L::NCHUNK; ++c)
                    tma_store_4d(&map_out, out_base + ob * L::OUT_BYTES + c * L::CHUNK_BYTES, col0 + c * L::CHUNK_COLS, w0, h0, n0);
                tma_store_commit();
                tma_store_wait_read0();                  // staging buffer has been read out
                const int tn = t + OUT_BUFS * item_stride;
                if (tn < total_tiles) make_ready(tn, ob);
            }
        }
    } else {
        // ===== epilogue warps 2..9: warp & 3 = TMEM lane quadrant, (warp - 2) >> 2 = column half of the tile.
        // (Letting two groups of four warps drain alternate tiles concurrently was measured slower: 5.04 vs 4.90 ms/step.) =====
        const int quad = warp & 3;
        const int half = (warp - 2) >> 2;
        const int r = quad * 32 + lane;                          // tile row owned by this thread
        uint8_t* out_base = smem + L::OUT_OFFSET;
        const int sw = r & 7;
        constexpr int C_CNT = BLOCK_N / 64;                      // 32-column slabs per warp
        int li = 0;
        for (int t = first_item; t < total_tiles; t += item_stride, ++li) {
            const int as = li & 1, ob = li % OUT_BUFS;
            const int col0 = col0_fixed;
            mbar_wait(&buf_ready[ob], (li / OUT_BUFS) & 1);      // staging free (and residual landed)
            mbar_wait(&acc_full[as], (li >> 1) & 1);
            tc_fence_after();
            uint8_t* ostage = out_base + ob * L::OUT_BYTES;
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * BLOCK_N);
            const int c_lo = half * C_CNT;
            if constexpr (C_CNT == 1) {
                uint32_t acc0[32];
                tmem_ld32_nowait(taddr + (uint32_t)(c_lo * 32), acc0);
                tmem_ld_wait();
                epilogue_slab<TO, L::CHUNK_BYTES>(acc0, c_lo, col0 + c_lo * 32, p.scale, p.bias, has_res, p.relu, ostage, r, sw);
            } else {
#pragma unroll 1
                for (int c = c_lo; c < c_lo + C_CNT; c += 2) {
                    uint32_t acc0[32], acc1[32];
                    tmem_ld32_nowait(taddr + (uint32_t)(c * 32), acc0);
                    tmem_ld32_nowait(taddr + (uint32_t)(c * 32 + 32), acc1);
                    tmem_ld_wait();
                    epilogue_slab<TO, L::CHUNK_BYTES>(acc0, c, col0 + c * 32, p.scale, p.bias, has_res, p.relu, ostage, r, sw);
                    epilogue_slab<TO, L::CHUNK_BYTES>(acc1, c + 1, col0 + c * 32 + 32, p.scale, p.bias, has_res, p.relu, ostage, r, sw);
                }
            }
            // accumulator stage is free again; staged tile is visible to the async proxy
            tc_fence_before();
            fence_async_smem();
            __syncwarp();
            if (lane == 0) { mbar_arrive(&acc_empty[as]); mbar_arrive(&buf_full[ob]); }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<2 * BLOCK_N>(tmem_base);
    }
}

template <int BLOCK_N, int STAGES, int OUT_BUFS, typename TO, int MAX_KB, int HALO = 0>
int launch_v4(const TcProblem& pr, const CUtensorMap& mo, const CUtensorMap& mr, cudaStream_t stream)
{
    using L = Smem4<BLOCK_N, STAGES, OUT_BUFS, TO, MAX_KB, HALO>;
    static_assert(L::TOTAL <= 232448, "shared memory budget exceeded");
    auto kern = conv_tc4_kernel<BLOCK_N, STAGES, OUT_BUFS, TO, MAX_KB, HALO>;
    static bool attr_set = false;
    if (!attr_set) {
        SEDT_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
        attr_set = true;
    }
    // grid: a multiple of the number of N tiles, every CTA keeps one of them
    const int per_n = std::max(1, std::min(num_sms() / pr.tiles_nc, pr.tiles_m));
    const int grid = per_n * pr.tiles_nc;
    ProfScope _prof(PROF_GEMM_TC, stream);
    SEDT_CHECK_CUDA(launch_pdl(kern, dim3((unsigned)grid), dim3(NUM_THREADS4), L::TOTAL, stream, 1, pr.map_a[0], pr.map_a[1],
                               pr.map_a[2], pr.map_a[3], pr.map_b, mo, mr, pr.p, pr.tiles_nc, pr.tiles_m));
    SEDT_COUNT_KIND(KK_CONV_TC4_WS);
    SEDT_CHECK_CUDA(cudaGetLastError());
    return SEDT_OK;
}

}  // namespace

bool conv_tc_ws_supported(const ConvGemm& g)
{
    if (!conv_tc_supported(g)) return false;
    const int num_kb = g.R * g.S * g.Cin / BLOCK_K;
    // 128-wide N tiles with K <= 256, or the 64-channel 3x3 convolutions of layer1 (K = 576, whole 72 KiB filter resident)
    if (g.Cout == 64) return g.out_dt == DT_BF16 && num_kb <= 9;
    return num_kb <= 4 && g.Cout % 128 == 0 && g.Cout / 128 <= num_sms();
}

int launch_conv_tc_ws(const ConvGemm& g, cudaStream_t stream)
{
    SEDT_REQUIRE(conv_tc_ws_supported(g), "conv_tc_ws: unsupported shape");
    const bool f32 = g.out_dt == DT_F32;
    TcProblem pr;
    const int block_n = g.Cout == 64 ? 64 : 128;
    SEDT_TRY(build_problem(g, block_n, &pr));
    CUtensorMap mo, mr;
    SEDT_TRY(encode_out_map(&mo, g.out, g.ldc, f32, g, pr.p));
    if (g.residual != nullptr) SEDT_TRY(encode_out_map(&mr, g.residual, g.ld_res, f32, g, pr.p));
    else mr = mo;
    if (block_n == 64) {
        static const bool halo_on = [] { const char* e = getenv("SEDT_HALO"); return e == nullptr || atoi(e) != 0; }();
        const TcParams& q = pr.p;
        if (halo_on && g.R == 3 && g.S == 3 && g.stride == 1 && g.dil == 1 && g.pad == 1 && g.Cin == 64 && q.bn == 1 &&
            (q.bw == 8 || q.bw == 16) && q.bw * q.bh == BLOCK_M) {
            // the halo box: same tensor as map_a[0], (bh + 2) rows per box
            const uint64_t dims[4] = {(uint64_t)g.Cin, (uint64_t)g.W, (uint64_t)g.H, (uint64_t)g.B};
            const uint64_t strides[3] = {(uint64_t)g.lda * 2, (uint64_t)g.W * g.lda * 2, (uint64_t)g.H * g.W * g.lda * 2};
            const uint32_t box[4] = {(uint32_t)BLOCK_K, (uint32_t)q.bw, (uint32_t)(q.bh + 2), 1u};
            SEDT_TRY(encode_map(&pr.map_a[3], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, g.in, 4, dims, strides, box));
            return launch_v4<64, 5, 2, __nv_bfloat16, 9, 1>(pr, mo, mr, stream);
        }
        return launch_v4<64, 6, 2, __nv_bfloat16, 9>(pr, mo, mr, stream);
    }
    return f32 ? launch_v4<128, 6, 1, float, 4>(pr, mo, mr, stream) : launch_v4<128, 6, 2, __nv_bfloat16, 4>(pr, mo, mr, stream);
}

}  // namespace sedt
