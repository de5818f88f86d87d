// Weight gradient of a convolution / linear layer on tcgen05:
//
//   dW[co, (r,s), ci] += sum over output pixels m = (b, ho, wo) of  dY[m, co] * X[pixel(m) shifted by tap (r,s), ci]
//
// i.e. a GEMM whose reduction runs over the PIXELS.  Both operands are read exactly as the forward
// kernel reads its A operand - 4-D TMA boxes of 128 output pixels x 64 channels over the NHWC tensors,
// the X box shifted by the filter tap with out-of-bounds zero fill (stride 2 through the phase views) -
// and the same shared-memory image ([128 pixels][64 channels], 128-byte swizzle) is handed to
// tcgen05.mma as an MN-MAJOR operand: channels are the M (dY) / N (X) index, pixels the K index.
// No transposed copy of an activation or a gradient is ever written.
//
// Work item = (128-wide Cout tile, BLOCK_N-wide Cin tile, filter tap, K split); one CTA per item.
// 192 threads: warp 0 TMA producer, warp 1 MMA issuer, warps 2..5 epilogue (TMEM -> red.global.add.f32,
// the caller zeroes dW).  Replaces the weight-gradient half of autograd's convolution / addmm backward
// for sedt/backbone.py (layer2-4), sedt/sedt.py:36 (input_proj) and the nn.Linear layers of
// sedt/transformer.py.
#include "tc_common.cuh"
#include <algorithm>

namespace sedt {
namespace {

using namespace tc;

constexpr int WG_THREADS = 192;
constexpr int WG_BOX_BYTES = BLOCK_M * BLOCK_K * 2;          // one [128 pixels][64 channels] box = 16 KiB

struct WgParams {
    const float* row_scale;      // optional per-output-channel factor applied in the epilogue
    float* dw;
    int ldw;                     // R*S*Cin
    int Cin;
    int bw, bh, bn, tiles_w, tiles_h, tiles_m;
    int ntaps, cin_tiles, splits;
    int8_t tap_map[9], tap_dh[9], tap_dw[9];
};

template <int BLOCK_N, int STAGES>
struct SmemWg {
    static constexpr int A_BYTES = 2 * WG_BOX_BYTES;                       // 128 output channels of dY
    static constexpr int B_BYTES = (BLOCK_N / 64) * WG_BOX_BYTES;          // BLOCK_N input channels of X
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
    static constexpr int TOTAL = BAR_OFFSET + (2 * STAGES + 1) * 8 + 16 + 1024;
};

template <int BLOCK_N, int STAGES>
__global__ void __launch_bounds__(WG_THREADS)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap map_x0, const __grid_constant__ CUtensorMap map_x1,
                  const __grid_constant__ CUtensorMap map_x2, const __grid_constant__ CUtensorMap map_x3,
                  const __grid_constant__ CUtensorMap map_dy, const __grid_constant__ WgParams p)
{
    using L = SmemWg<BLOCK_N, STAGES>;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* full_bar = (uint64_t*)(smem + L::BAR_OFFSET);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* accum_bar = empty_bar + STAGES;
    uint32_t* tmem_slot = (uint32_t*)(accum_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // work item
    int item = blockIdx.x;
    const int split = item % p.splits; item /= p.splits;
    const int tap = item % p.ntaps; item /= p.ntaps;
    const int cin_tile = item % p.cin_tiles;
    const int cout_tile = item / p.cin_tiles;
    const int co0 = cout_tile * 128, ci0 = cin_tile * BLOCK_N;
    // pixel tiles of this split: t = split, split + splits, ...
    const int nkb = (p.tiles_m - split + p.splits - 1) / p.splits;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&map_x0); prefetch_tmap(&map_dy);
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(accum_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc<BLOCK_N>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            const int mi = p.tap_map[tap];
            const CUtensorMap* mx = mi == 0 ? &map_x0 : (mi == 1 ? &map_x1 : (mi == 2 ? &map_x2 : &map_x3));
            const int dw_ = p.tap_dw[tap], dh_ = p.tap_dh[tap];
            int stage = 0; uint32_t phase = 0;
            for (int kb = 0; kb < nkb; ++kb) {
                const int t = split + kb * p.splits;
                const int tw = t % p.tiles_w, th = (t / p.tiles_w) % p.tiles_h, tn = t / (p.tiles_w * p.tiles_h);
                const int w0 = tw * p.bw, h0 = th * p.bh, n0 = tn * p.bn;
                mbar_wait(&empty_bar[stage], phase ^ 1);
                mbar_expect_tx(&full_bar[stage], L::STAGE_BYTES);
                uint8_t* sa = smem + stage * L::STAGE_BYTES;
                uint8_t* sb = sa + L::A_BYTES;
                tma_load_4d(&map_dy, sa, &full_bar[stage], co0, w0, h0, n0);
                tma_load_4d(&map_dy, sa + WG_BOX_BYTES, &full_bar[stage], co0 + 64, w0, h0, n0);
#pragma unroll
                for (int j = 0; j < BLOCK_N / 64; ++j)
                    tma_load_4d(mx, sb + j * WG_BOX_BYTES, &full_bar[stage], ci0 + j * 64, w0 + dw_, h0 + dh_, n0);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(128, BLOCK_N) | (1u << 15) | (1u << 16);     // A and B MN-major
            int stage = 0; uint32_t phase = 0;
            for (int kb = 0; kb < nkb; ++kb) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                const uint32_t sa = smem_u32(smem + stage * L::STAGE_BYTES);
                const uint32_t sb = sa + L::A_BYTES;
#pragma unroll
                for (int k = 0; k < BLOCK_M / UMMA_K; ++k) {           // 16 pixel rows = 2048 bytes per MMA
                    const uint64_t da = make_smem_desc_mn(sa + k * UMMA_K * 128, WG_BOX_BYTES);
                    const uint64_t db = make_smem_desc_mn(sb + k * UMMA_K * 128, WG_BOX_BYTES);
                    umma_bf16(tmem_base, da, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
                }
                umma_commit(&empty_bar[stage]);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
            umma_commit(accum_bar);
        }
    } else {
        // epilogue: TMEM lane = output channel, columns = input channels of this tap
        const int quad = warp & 3;
        const int r = quad * 32 + lane;
        float* drow = p.dw + (size_t)(co0 + r) * p.ldw + (size_t)tap * p.Cin + ci0;
        const float rs = p.row_scale != nullptr ? p.row_scale[co0 + r] : 1.f;
        mbar_wait(accum_bar, 0);
        tc_fence_after();
        if (nkb > 0) {
#pragma unroll 1
            for (int c = 0; c < BLOCK_N / 32; ++c) {
                uint32_t acc[32];
                tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(c * 32), acc);
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    atomicAdd(reinterpret_cast<float4*>(drow + c * 32 + j),
                              make_float4(__uint_as_float(acc[j]) * rs, __uint_as_float(acc[j + 1]) * rs, __uint_as_float(acc[j + 2]) * rs,
                                          __uint_as_float(acc[j + 3]) * rs));
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<BLOCK_N>(tmem_base);
    }
}

template <int BLOCK_N, int STAGES>
int launch_wg(const TcProblem& pr, const CUtensorMap& mdy, const WgParams& p, int grid, cudaStream_t stream)
{
    using L = SmemWg<BLOCK_N, STAGES>;
    static_assert(L::TOTAL <= 232448, "shared memory budget exceeded");
    auto kern = conv_wgrad_kernel<BLOCK_N, STAGES>;
    static bool attr_set = false;
    if (!attr_set) {
        SEDT_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
        attr_set = true;
    }
    ProfScope _prof(PROF_GEMM_TC, stream);
    kern<<<grid, WG_THREADS, L::TOTAL, stream>>>(pr.map_a[0], pr.map_a[1], pr.map_a[2], pr.map_a[3], mdy, p);
    SEDT_COUNT_KIND(KK_WGRAD_TC);
    SEDT_CHECK_CUDA(cudaGetLastError());
    return SEDT_OK;
}

}  // namespace

bool conv_wgrad_tc_supported(const WgradGemm& g)
{
    if (g.Cin % 64 != 0 || g.Cout % 128 != 0) return false;
    if (g.lda % 8 != 0 || g.ldy % 8 != 0) return false;
    if (((uintptr_t)g.x & 15) || ((uintptr_t)g.dy & 15) || ((uintptr_t)g.dw & 15)) return false;
    if (g.R != g.S || (g.R != 1 && g.R != 3)) return false;
    if (g.stride != 1 && g.stride != 2) return false;
    if (g.stride == 2 && (g.H < 2 || g.W < 2 || g.dil != 1)) return false;
    if (g.R == 3 && g.pad != g.dil) return false;
    if (g.R == 1 && g.pad != 0) return false;
    int bw = 1; while (bw < g.Wo) bw <<= 1;
    if (bw > BLOCK_M) return false;
    return (int64_t)g.B * g.Ho * g.Wo >= 1;
}

// dW (fp32, [Cout][R*S*Cin], zeroed by the caller) += dY^T * im2col(X)
int launch_conv_wgrad_tc(const WgradGemm& g, cudaStream_t stream)
{
    SEDT_REQUIRE(conv_wgrad_tc_supported(g), "conv_wgrad: unsupported shape");
    SEDT_TRY(tc_init());
    // the forward problem's pixel tiling, phase maps and tap table, with X as the activation
    ConvGemm f;
    f.in = g.x; f.w = nullptr; f.out = nullptr;
    f.in_dt = f.out_dt = DT_BF16;
    f.B = g.B; f.H = g.H; f.W = g.W; f.Cin = g.Cin; f.lda = g.lda;
    f.Ho = g.Ho; f.Wo = g.Wo; f.Cout = g.Cout; f.ldc = g.ldy;
    f.R = g.R; f.S = g.S; f.stride = g.stride; f.dil = g.dil; f.pad = g.pad;
    const int block_n = g.Cin % 128 == 0 ? 128 : 64;
    TcProblem pr;
    SEDT_TRY(build_problem(f, block_n, &pr));
    CUtensorMap mdy;
    SEDT_TRY(encode_out_map(&mdy, g.dy, g.ldy, false, f, pr.p));

    WgParams p;
    memset(&p, 0, sizeof(p));
    p.dw = g.dw; p.row_scale = g.row_scale; p.ldw = g.R * g.S * g.Cin; p.Cin = g.Cin;
    p.bw = pr.p.bw; p.bh = pr.p.bh; p.bn = pr.p.bn; p.tiles_w = pr.p.tiles_w; p.tiles_h = pr.p.tiles_h; p.tiles_m = pr.tiles_m;
    p.ntaps = g.R * g.S; p.cin_tiles = g.Cin / block_n;
    memcpy(p.tap_map, pr.p.tap_map, sizeof(p.tap_map));
    memcpy(p.tap_dh, pr.p.tap_dh, sizeof(p.tap_dh));
    memcpy(p.tap_dw, pr.p.tap_dw, sizeof(p.tap_dw));
    // K splits: about four CTAs per SM in flight, at least four pixel tiles per CTA
    const int base_items = (g.Cout / 128) * p.cin_tiles * p.ntaps;
    int splits = (int)std::max<int64_t>(1, ceil_div((int64_t)4 * num_sms(), base_items));
    splits = std::min(splits, std::max(1, pr.tiles_m / 4));
    p.splits = splits;
    const int grid = base_items * splits;
    return block_n == 128 ? launch_wg<128, 3>(pr, mdy, p, grid, stream) : launch_wg<64, 4>(pr, mdy, p, grid, stream);
}

}  // namespace sedt
