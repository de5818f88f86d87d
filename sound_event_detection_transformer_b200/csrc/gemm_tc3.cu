// 2-SM (cta_group::2) variant of the persistent implicit-GEMM kernel for wide layers (Cout % 256 == 0).
//
// A cluster of two CTAs computes a 256 x 256 output tile per step: CTA r owns M tile 2*pm + r (128 output
// pixels) and loads, per 64-wide k block, its own A box (16 KiB) and rows [128 r, 128 r + 128) of the B
// (weight) tile (16 KiB).  One thread of the leader CTA issues tcgen05.mma.cta_group::2 (M = 256, N = 256):
// the tensor cores of both SMs read A from their own shared memory and B from both halves, and each SM
// accumulates its 128 x 256 slice in its own TMEM.  Compared with the single-CTA kernel a stage holds half
// the operand bytes per FLOP, which is what the shared-memory port (128 B/clk, shared by TMA writes and MMA
// operand reads) needs to stop being the limiter.
//
// Synchronisation (mbarriers at identical offsets in both CTAs):
//   full[s]      leader only; expects the bytes of BOTH CTAs' loads (the peer's TMA signals it remotely)
//   empty[s]     both; arrived by the leader's tcgen05.commit.cta_group::2 multicast
//   acc_full[a]  both; multicast commit after the last k block of a tile
//   acc_empty[a] leader only, count 16: the eight epilogue warps of both CTAs (peer arrives remotely)
//   buf_ready / buf_full: per-CTA handshake between epilogue warps and the store warp (as in gemm_tc2.cu),
//   on 32 KiB sub-tiles (128 bf16 / 64 fp32 columns) so that stores overlap the next sub-tile's epilogue.
#include "tc_common.cuh"
#include <cstdlib>

namespace sedt {
namespace {

using namespace tc;

constexpr int T3_THREADS = 352;              // producer, MMA, 8 epilogue warps, store warp
constexpr int T3_BN = 256;                   // columns per cluster tile
constexpr int T3_HALF = 128;                 // B rows held per CTA
constexpr int T3_STAGE = A_STAGE_BYTES + T3_HALF * BLOCK_K * 2;     // 32 KiB
constexpr int T3_OUT_BYTES = 32768;          // one epilogue sub-tile: 128 rows x (128 bf16 | 64 fp32) columns
constexpr uint32_t kPeerMask = 0xFEFFFFFFu;  // shared::cluster address of the same offset in the leader CTA

template <int STAGES>
struct Smem3 {
    static constexpr int OUT_OFFSET = STAGES * T3_STAGE;
    static constexpr int BAR_OFFSET = OUT_OFFSET + 2 * T3_OUT_BYTES;
    static constexpr int NBARS = 2 * STAGES + 4 + 4;
    static constexpr int TOTAL = BAR_OFFSET + NBARS * 8 + 16 + 1024;
};

__device__ __forceinline__ void tma_load_4d_2sm(const CUtensorMap* map, void* smem, uint32_t leader_bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(smem)), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(const CUtensorMap* map, void* smem, uint32_t leader_bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem)), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)0x3) : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
    // arrive on the barrier at this offset in CTA 0 of the cluster (works from either CTA)
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(0));
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}

template <int STAGES, typename TO>
__global__ void __launch_bounds__(T3_THREADS, 1)
conv_tc3_kernel(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_a1,
                const __grid_constant__ CUtensorMap map_a2, const __grid_constant__ CUtensorMap map_a3,
                const __grid_constant__ CUtensorMap map_b, const __grid_constant__ CUtensorMap map_out,
                const __grid_constant__ CUtensorMap map_res, const __grid_constant__ TcParams p,
                const int tiles_nc, const int total_items)
{
    using L = Smem3<STAGES>;
    constexpr int SUBW = T3_OUT_BYTES / (BLOCK_M * (int)sizeof(TO));      // columns per epilogue sub-tile: 128 | 64
    constexpr int NSUB = T3_BN / SUBW;                                    // sub-tiles per tile: 2 | 4
    constexpr int CCOLS = 128 / (int)sizeof(TO);                          // columns per 128-byte staging row: 64 | 32
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* full_bar = (uint64_t*)(smem + L::BAR_OFFSET);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* acc_full = empty_bar + STAGES;          // [2]
    uint64_t* acc_empty = acc_full + 2;               // [2]
    uint64_t* buf_ready = acc_empty + 2;              // [2]
    uint64_t* buf_full = buf_ready + 2;               // [2]
    uint32_t* tmem_slot = (uint32_t*)(buf_full + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int crank = (int)cluster_ctarank();
    const bool leader = crank == 0;
    const int cpb = p.Cin / BLOCK_K;
    const int num_kb = p.ntaps * cpb;
    const bool has_res = p.residual != nullptr;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&map_a0); prefetch_tmap(&map_b); prefetch_tmap(&map_out);
        if (has_res) prefetch_tmap(&map_res);
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 16); }
        for (int s = 0; s < 2; ++s) { mbar_init(&buf_ready[s], 1); mbar_init(&buf_full[s], 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    // PDL: the successor may be scheduled from here on (this CTA owns its TMEM columns already, so a co-resident
    // successor CTA can never starve it); everything above overlapped the predecessor's tail, whose outputs are our operands
    pdl_trigger();
    pdl_wait();
    const uint32_t tmem_base = *tmem_slot;

    const int first_item = (int)(blockIdx.x >> 1), item_stride = (int)(gridDim.x >> 1);
    auto tile_coords = [&](int t, int& w0, int& h0, int& n0, int& col0) {
        const int n_tile = t % tiles_nc, m_tile = (t / tiles_nc) * 2 + crank;
        const int tw = m_tile % p.tiles_w;
        const int th = (m_tile / p.tiles_w) % p.tiles_h;
        const int tn = m_tile / (p.tiles_w * p.tiles_h);
        w0 = tw * p.bw; h0 = th * p.bh; n0 = tn * p.bn; col0 = n_tile * T3_BN;
    };

    if (warp == 0) {
        // ===== TMA producer (both CTAs): own A box + own half of the weight tile, signalled to the leader =====
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int t = first_item; t < total_items; t += item_stride) {
                int w0, h0, n0, col0;
                tile_coords(t, w0, h0, n0, col0);
                for (int tap = 0; tap < p.ntaps; ++tap) {
                    const int mi = p.tap_map[tap];
                    const CUtensorMap* ma = mi == 0 ? &map_a0 : (mi == 1 ? &map_a1 : (mi == 2 ? &map_a2 : &map_a3));
                    const int cw = w0 + p.tap_dw[tap], ch = h0 + p.tap_dh[tap], kbase = tap * p.Cin;
                    for (int c0 = 0; c0 < p.Cin; c0 += BLOCK_K) {
                        mbar_wait(&empty_bar[stage], phase ^ 1);
                        if (leader) mbar_expect_tx(&full_bar[stage], 2 * T3_STAGE);
                        const uint32_t lbar = smem_u32(&full_bar[stage]) & kPeerMask;
                        uint8_t* sa = smem + stage * T3_STAGE;
                        tma_load_4d_2sm(ma, sa, lbar, c0, cw, ch, n0);
                        tma_load_2d_2sm(&map_b, sa + A_STAGE_BYTES, lbar, kbase + c0, col0 + crank * T3_HALF);
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: one thread of the leader CTA drives the tensor cores of both SMs =====
        if (leader && lane == 0) {
            constexpr uint32_t idesc = make_idesc(256, T3_BN);
            int stage = 0; uint32_t phase = 0;
            int li = 0;
            for (int t = first_item; t < total_items; t += item_stride, ++li) {
                const int as = li & 1;
                mbar_wait(&acc_empty[as], ((li >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(as * T3_BN);
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * T3_STAGE);
                    const uint32_t sb = sa + A_STAGE_BYTES;
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
                        umma_bf16_2sm(tmem_d, make_smem_desc(sa + k * UMMA_K * 2), make_smem_desc(sb + k * UMMA_K * 2), idesc,
                                      (kb > 0 || k > 0) ? 1u : 0u);
                    umma_commit_2sm(&empty_bar[stage]);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit_2sm(&acc_full[as]);
            }
        }
    } else if (warp == 10) {
        // ===== store warp: half tiles of 128 columns =====
        if (lane == 0) {
            uint8_t* out_base = smem + L::OUT_OFFSET;
            // sub-tile q of this CTA's tile sequence uses staging buffer q & 1
            auto make_ready = [&](int t, int st, int buf) {
                if (has_res) {
                    int w0, h0, n0, col0;
                    tile_coords(t, w0, h0, n0, col0);
                    mbar_expect_tx(&buf_ready[buf], T3_OUT_BYTES);
#pragma unroll
                    for (int c = 0; c < 2; ++c)
                        tma_load_4d(&map_res, out_base + buf * T3_OUT_BYTES + c * 16384, &buf_ready[buf],
                                    col0 + st * SUBW + c * CCOLS, w0, h0, n0);
                } else {
                    mbar_arrive(&buf_ready[buf]);
                }
            };
            if (first_item < total_items) { make_ready(first_item, 0, 0); make_ready(first_item, 1, 1); }
            int q = 0;
            for (int t = first_item; t < total_items; t += item_stride) {
                int w0, h0, n0, col0;
                tile_coords(t, w0, h0, n0, col0);
                for (int st = 0; st < NSUB; ++st, ++q) {
                    const int buf = q & 1;
                    mbar_wait(&buf_full[buf], (q >> 1) & 1);
#pragma unroll
                    for (int c = 0; c < 2; ++c)
                        tma_store_4d(&map_out, out_base + buf * T3_OUT_BYTES + c * 16384, col0 + st * SUBW + c * CCOLS, w0, h0, n0);
                    tma_store_commit();
                    tma_store_wait_read0();
                    // the sub-tile that will use this buffer next is two ahead in the sequence
                    int t2 = t, st2 = st + 2;
                    if (st2 >= NSUB) { st2 -= NSUB; t2 += item_stride; }
                    if (t2 < total_items) make_ready(t2, st2, buf);
                }
            }
        }
    } else {
        // ===== epilogue warps 2..9 (both CTAs, each on its own TMEM slice): two warps per lane quadrant,
        // each taking 64 of the 128 columns of a half tile =====
        const int quad = warp & 3;
        const int colhalf = (warp - 2) >> 2;
        const int r = quad * 32 + lane;
        uint8_t* out_base = smem + L::OUT_OFFSET;
        const int sw = r & 7;
        int li = 0, q = 0;
        for (int t = first_item; t < total_items; t += item_stride, ++li) {
            const int as = li & 1;
            const int col0 = (t % tiles_nc) * T3_BN;
            mbar_wait(&acc_full[as], (li >> 1) & 1);
            tc_fence_after();
            for (int st = 0; st < NSUB; ++st, ++q) {
                const int buf = q & 1;
                mbar_wait(&buf_ready[buf], (q >> 1) & 1);
                uint8_t* ostage = out_base + buf * T3_OUT_BYTES;
                const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * T3_BN + st * SUBW);
                if constexpr (sizeof(TO) == 2) {
                    const int c = colhalf * 2;                   // two 32-column slabs of the 128-column sub-tile
                    uint32_t acc0[32], acc1[32];
                    tmem_ld32_nowait(taddr + (uint32_t)(c * 32), acc0);
                    tmem_ld32_nowait(taddr + (uint32_t)(c * 32 + 32), acc1);
                    tmem_ld_wait();
                    const int nb = col0 + st * SUBW + c * 32;
                    epilogue_slab<TO, 16384>(acc0, c, nb, p.scale, p.bias, has_res, p.relu, ostage, r, sw);
                    epilogue_slab<TO, 16384>(acc1, c + 1, nb + 32, p.scale, p.bias, has_res, p.relu, ostage, r, sw);
                } else {
                    const int c = colhalf;                       // one 32-column slab of the 64-column sub-tile
                    uint32_t acc0[32];
                    tmem_ld32_nowait(taddr + (uint32_t)(c * 32), acc0);
                    tmem_ld_wait();
                    epilogue_slab<TO, 16384>(acc0, c, col0 + st * SUBW + c * 32, p.scale, p.bias, has_res, p.relu, ostage, r, sw);
                }
                fence_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&buf_full[buf]);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(&acc_empty[as]);
        }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

template <int STAGES, typename TO>
int launch_v3(const TcProblem& pr, const CUtensorMap& mo, const CUtensorMap& mr, cudaStream_t stream)
{
    using L = Smem3<STAGES>;
    static_assert(L::TOTAL <= 232448, "shared memory budget exceeded");
    auto kern = conv_tc3_kernel<STAGES, TO>;
    static bool attr_set = false;
    if (!attr_set) {
        SEDT_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
        attr_set = true;
    }
    const int total = ((pr.tiles_m + 1) / 2) * pr.tiles_nc;
    const int grid = std::min(total * 2, num_sms()) & ~1;
    ProfScope _prof(PROF_GEMM_TC, stream);
    SEDT_CHECK_CUDA(launch_pdl(kern, dim3((unsigned)grid), dim3(T3_THREADS), L::TOTAL, stream, 2, pr.map_a[0], pr.map_a[1], pr.map_a[2],
                               pr.map_a[3], pr.map_b, mo, mr, pr.p, pr.tiles_nc, total));
    SEDT_COUNT_KIND(KK_CONV_TC3_2SM);
    SEDT_CHECK_CUDA(cudaGetLastError());
    return SEDT_OK;
}

}  // namespace

bool conv_tc_2sm_supported(const ConvGemm& g)
{
    if (!conv_tc_supported(g)) return false;
    return g.Cout % T3_BN == 0;
}

bool conv_tc_2sm_preferred(const ConvGemm& g)
{
    if (!conv_tc_2sm_supported(g)) return false;
    // enough 256 x 256 cluster tiles to occupy the 74 SM pairs
    const int64_t m_tiles = ceil_div((int64_t)g.B * g.Ho * g.Wo, BLOCK_M);
    return ((m_tiles + 1) / 2) * (g.Cout / T3_BN) >= num_sms() / 2;
}

int launch_conv_tc_2sm(const ConvGemm& g, cudaStream_t stream)
{
    SEDT_REQUIRE(conv_tc_2sm_supported(g), "conv_tc_2sm: unsupported shape");
    TcProblem pr;
    SEDT_TRY(build_problem(g, T3_BN, &pr, 2));          // B box = 128 rows: each CTA loads its half of the 256-wide tile
    CUtensorMap mo, mr;
    const bool f32 = g.out_dt == DT_F32;
    SEDT_TRY(encode_out_map(&mo, g.out, g.ldc, f32, g, pr.p));
    if (g.residual != nullptr) SEDT_TRY(encode_out_map(&mr, g.residual, g.ld_res, f32, g, pr.p));
    else mr = mo;
    return f32 ? launch_v3<5, float>(pr, mo, mr, stream) : launch_v3<5, __nv_bfloat16>(pr, mo, mr, stream);
}

}  // namespace sedt
