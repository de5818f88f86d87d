// Persistent TMA + tcgen05 implicit-GEMM convolution / linear (production kernel).
//
// Same math and operand path as gemm_tc.cu (4-D TMA boxes per filter tap, zero padding by TMA
// out-of-bounds fill, stride-2 through phase views, 128B-swizzled K-major operands, fp32
// accumulation in TMEM) with the three things the one-tile-per-CTA kernel lacks:
//   * one CTA per SM loops over output tiles (static round-robin, N tiles adjacent so CTAs that
//     run together share the same A box through L2);
//   * the accumulator is double-buffered in TMEM (2 x BLOCK_N columns): the epilogue warps drain
//     tile i while the producer / MMA warps already run the main loop of tile i+1;
//   * the epilogue is staged through shared memory: the residual tile arrives by TMA (prefetched
//     one tile ahead), results are written back with TMA stores (coalesced, asynchronous, rows
//     outside the tensor are clipped by the store), so no thread ever waits on a global load.
//
// Warp roles (352 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..9 =
// epilogue (two per TMEM lane quadrant, each thread one row x half the columns), warp 10 = TMA store /
// residual prefetch.
#include "tc_common.cuh"
#include <cstdlib>

namespace sedt {
namespace {

using namespace tc;

constexpr int NUM_THREADS2 = 352;      // 11 warps: producer, MMA, 8 x epilogue (two per TMEM lane quadrant), store

template <int BLOCK_N, int STAGES, int OUT_BUFS, typename TO>
struct Smem2 {
    static constexpr int B_STAGE_BYTES = BLOCK_N * BLOCK_K * 2;
    static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
    static constexpr int CHUNK_COLS = 128 / (int)sizeof(TO);              // columns per 128-byte staging row
    static constexpr int NCHUNK = BLOCK_N / CHUNK_COLS;
    static constexpr int CHUNK_BYTES = BLOCK_M * 128;                     // 16 KiB, keeps 1024-B alignment
    static constexpr int OUT_BYTES = NCHUNK * CHUNK_BYTES;
    static constexpr int OUT_OFFSET = STAGES * STAGE_BYTES;
    static constexpr int BAR_OFFSET = OUT_OFFSET + OUT_BUFS * OUT_BYTES;
    static constexpr int NBARS = 2 * STAGES + 4 + 2 * OUT_BUFS;
    static constexpr int TOTAL = BAR_OFFSET + NBARS * 8 + 16 + 1024;
};

template <int BLOCK_N, int STAGES, int OUT_BUFS, typename TO>
__global__ void __launch_bounds__(NUM_THREADS2, 1)
conv_tc2_kernel(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_a1,
                const __grid_constant__ CUtensorMap map_a2, const __grid_constant__ CUtensorMap map_a3,
                const __grid_constant__ CUtensorMap map_b, const __grid_constant__ CUtensorMap map_out,
                const __grid_constant__ CUtensorMap map_res, const __grid_constant__ TcParams p,
                const int tiles_nc, const int total_tiles)
{
    using L = Smem2<BLOCK_N, STAGES, OUT_BUFS, TO>;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // keep the pointer derived from the __shared__ symbol so that staging traffic compiles to LDS/STS
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* full_bar = (uint64_t*)(smem + L::BAR_OFFSET);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* acc_full = empty_bar + STAGES;          // [2]  MMA -> epilogue
    uint64_t* acc_empty = acc_full + 2;               // [2]  epilogue -> MMA
    uint64_t* buf_ready = acc_empty + 2;              // [OUT_BUFS] store warp / residual TMA -> epilogue
    uint64_t* buf_full = buf_ready + OUT_BUFS;        // [OUT_BUFS] epilogue -> store warp
    uint32_t* tmem_slot = (uint32_t*)(buf_full + OUT_BUFS);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cpb = p.Cin / BLOCK_K;
    const int num_kb = p.ntaps * cpb;
    const bool has_res = p.residual != nullptr;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&map_a0); prefetch_tmap(&map_b); prefetch_tmap(&map_out);
        if (has_res) prefetch_tmap(&map_res);
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
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
    const int first_item = (int)blockIdx.x, item_stride = (int)gridDim.x;

    // item -> coordinates.  N tiles are adjacent in the schedule: t = m_item * tiles_nc + n_tile.
    auto tile_coords = [&](int t, int& w0, int& h0, int& n0, int& col0) {
        const int n_tile = t % tiles_nc, m_tile = t / tiles_nc;
        const int tw = m_tile % p.tiles_w;
        const int th = (m_tile / p.tiles_w) % p.tiles_h;
        const int tn = m_tile / (p.tiles_w * p.tiles_h);
        w0 = tw * p.bw; h0 = th * p.bh; n0 = tn * p.bn; col0 = n_tile * BLOCK_N;
    };

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int t = first_item; t < total_tiles; t += item_stride) {
                int w0, h0, n0, col0;
                tile_coords(t, w0, h0, n0, col0);
                // nested loops instead of kb / cpb: this single thread's instruction latency is the pace of the pipeline
                for (int tap = 0; tap < p.ntaps; ++tap) {
                    const int mi = p.tap_map[tap];
                    const CUtensorMap* ma = mi == 0 ? &map_a0 : (mi == 1 ? &map_a1 : (mi == 2 ? &map_a2 : &map_a3));
                    const int cw = w0 + p.tap_dw[tap], ch = h0 + p.tap_dh[tap], kbase = tap * p.Cin;
                    for (int c0 = 0; c0 < p.Cin; c0 += BLOCK_K) {
                        mbar_wait(&empty_bar[stage], phase ^ 1);
                        mbar_expect_tx(&full_bar[stage], L::STAGE_BYTES);
                        uint8_t* sa = smem + stage * L::STAGE_BYTES;
                        tma_load_4d(ma, sa, &full_bar[stage], c0, cw, ch, n0);
                        tma_load_2d(&map_b, sa + A_STAGE_BYTES, &full_bar[stage], kbase + c0, col0);
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
            for (int t = first_item; t < total_tiles; t += item_stride, ++li) {
                const int as = li & 1;
                mbar_wait(&acc_empty[as], ((li >> 1) & 1) ^ 1);      // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(as * BLOCK_N);
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * L::STAGE_BYTES);
                    const uint32_t sb = sa + A_STAGE_BYTES;
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
            const int col0 = (t % tiles_nc) * BLOCK_N;
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

template <int BLOCK_N, int STAGES, int OUT_BUFS, typename TO>
int launch_v2(const TcProblem& pr, const CUtensorMap& mo, const CUtensorMap& mr, cudaStream_t stream)
{
    using L = Smem2<BLOCK_N, STAGES, OUT_BUFS, TO>;
    static_assert(L::TOTAL <= 232448, "shared memory budget exceeded");
    auto kern = conv_tc2_kernel<BLOCK_N, STAGES, OUT_BUFS, TO>;
    static bool attr_set = false;
    if (!attr_set) {
        SEDT_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
        attr_set = true;
    }
    const int total = pr.tiles_m * pr.tiles_nc;
    const int grid = std::min(total, num_sms());
    ProfScope _prof(PROF_GEMM_TC, stream);
    SEDT_CHECK_CUDA(launch_pdl(kern, dim3((unsigned)grid), dim3(NUM_THREADS2), L::TOTAL, stream, 1, pr.map_a[0], pr.map_a[1],
                               pr.map_a[2], pr.map_a[3], pr.map_b, mo, mr, pr.p, pr.tiles_nc, total));
    SEDT_COUNT_KIND(KK_CONV_TC2);
    SEDT_CHECK_CUDA(cudaGetLastError());
    return SEDT_OK;
}

}  // namespace

int launch_conv_tc(const ConvGemm& g, cudaStream_t stream)
{
    static const bool force_v1 = [] { const char* e = getenv("SEDT_TC_V1"); return e != nullptr && e[0] == '1'; }();
    if (force_v1) return launch_conv_tc_v1(g, stream);
    SEDT_REQUIRE(conv_tc_supported(g), "conv_tc: unsupported shape");
    static const int use_2sm = [] { const char* e = getenv("SEDT_TC_2SM"); return e ? atoi(e) : 1; }();
    // two SMs per tile pay once the main loop is long (K >= 512); short-K layers are epilogue / HBM bound
    static const int min_k_2sm = [] { const char* e = getenv("SEDT_2SM_MIN_K"); return e ? atoi(e) : 512; }();
    if (use_2sm && g.R * g.S * g.Cin >= min_k_2sm && conv_tc_2sm_preferred(g)) return launch_conv_tc_2sm(g, stream);
    // short reductions: weight-stationary kernel (each CTA keeps its N tile's weights in shared memory)
    static const int use_ws = [] { const char* e = getenv("SEDT_TC_WS"); return e ? atoi(e) : 1; }();
    if (use_ws && conv_tc_ws_supported(g) && ceil_div((int64_t)g.B * g.Ho * g.Wo, BLOCK_M) >= 2 * (num_sms() / ceil_div(g.Cout, 128)))
        return launch_conv_tc_ws(g, stream);
    const bool f32 = g.out_dt == DT_F32;
    // BLOCK_N: the widest tile that divides Cout, still leaves about two tiles per SM and has a main
    // loop long enough (num_kb >= min_kb256) to hide the single-buffered 256-wide epilogue
    static const int min_kb256 = [] { const char* e = getenv("SEDT_BN256_MIN_KB"); return e ? atoi(e) : 16; }();
    static const int f32_bn = [] { const char* e = getenv("SEDT_F32_BN"); return e ? atoi(e) : 128; }();
    const int num_kb = g.R * g.S * g.Cin / BLOCK_K;
    int block_n = 64;
    if (g.Cout % 128 == 0 && !(f32 && f32_bn == 64)) block_n = 128;
    if (!f32 && g.Cout % 256 == 0 && num_kb >= min_kb256) {
        const int64_t m_tiles = ceil_div((int64_t)g.B * g.Ho * g.Wo, BLOCK_M);
        if (m_tiles * (g.Cout / 256) >= 2 * num_sms()) block_n = 256;
    }
    TcProblem pr;
    SEDT_TRY(build_problem(g, block_n, &pr));
    CUtensorMap mo, mr;
    SEDT_TRY(encode_out_map(&mo, g.out, g.ldc, f32, g, pr.p));
    if (g.residual != nullptr) SEDT_TRY(encode_out_map(&mr, g.residual, g.ld_res, f32, g, pr.p));
    else mr = mo;
#define SEDT_V2(BN, ST, OB, T) launch_v2<BN, ST, OB, T>(pr, mo, mr, stream)
    if (f32) {
        if (block_n == 128) return SEDT_V2(128, 4, 1, float);
        return SEDT_V2(64, 4, 2, float);
    }
    if (block_n == 256) return SEDT_V2(256, 3, 1, __nv_bfloat16);
    if (block_n == 128) return SEDT_V2(128, 5, 2, __nv_bfloat16);
    return SEDT_V2(64, 8, 2, __nv_bfloat16);
#undef SEDT_V2
}

}  // namespace sedt
