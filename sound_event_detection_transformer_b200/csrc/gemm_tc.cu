// TMA-fed tcgen05 implicit-GEMM convolution / linear for sm_100a.
//
//   out[m, n] = act((sum_k A[m,k] W[n,k]) * scale[n] + bias[n] + residual[m,n])
//
// A is never materialised: a 128-row M tile is a (bw x bh x bn) box of output
// pixels, and for every filter tap the producer warp issues ONE 4-D TMA load of
// the input box shifted by that tap; rows that fall into the zero padding are
// filled by TMA's out-of-bounds handling.  Stride-2 convolutions read from up
// to four "phase" views of the input (even/odd rows x even/odd columns), each a
// plain strided tensor map, so they need no gather either.  Operands land in
// shared memory in the 128-byte-swizzled K-major layout that tcgen05.mma
// consumes directly; the fp32 accumulator lives in TMEM and is read back with
// tcgen05.ld by four epilogue warps that fuse FrozenBatchNorm scale/bias, bias,
// residual add and ReLU.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator +
// single-thread MMA issuer, warps 2..5 = epilogue (one TMEM lane quadrant each).
// Replaces the cuDNN/cuBLAS calls under torchvision's Bottleneck
// (resnet.py:143-163), SEDT.input_proj (sedt/sedt.py:36,88) and the nn.Linear
// layers of sedt/transformer.py.
#include "kernels.h"
#include <cuda.h>
#include <cudaTypedefs.h>

namespace sedt {
namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;                 // 64 bf16 = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int NUM_THREADS = 192;
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;

struct TcParams {
    const float* scale;
    const float* bias;
    const void* residual;
    void* out;
    int ld_res, ldc;
    int B, Ho, Wo;               // output geometry (rows of the GEMM)
    int bw, bh, bn;              // M-tile box, bw*bh*bn == 128
    int tiles_w, tiles_h;        // tiles along Wo / Ho (tiles along B = gridDim.x / (tiles_w*tiles_h))
    int Cin, ntaps;
    int relu;
    int8_t tap_map[9], tap_dh[9], tap_dw[9];
};

// ---- PTX wrappers ---------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must surface as a trapped kernel, never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) __trap();
    }
}
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, void* smem, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(smem)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, void* smem, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <int NCOLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(NCOLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(NCOLS) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, bf16 x bf16 -> fp32
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// mbarrier arrives once every tcgen05.mma issued so far by this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, 128-byte swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
//   [0,14) start>>4 | [16,30) LBO>>4 (=1, unused for swizzled K-major) | [32,46) SBO>>4 (8 rows * 128 B = 1024)
//   [46,48) version = 1 | [61,64) layout = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// kind::f16 instruction descriptor (cute::UMMA::InstrDescriptor): D=f32, A=B=bf16, both K-major
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

template <int BLOCK_N, int STAGES>
struct SmemLayout {
    static constexpr int B_STAGE_BYTES = BLOCK_N * BLOCK_K * 2;
    static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
    static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
    static constexpr int TOTAL = BAR_OFFSET + (2 * STAGES + 1) * 8 + 16 + 1024;     // + alignment slack
};

template <int BLOCK_N, int STAGES, typename TO>
__global__ void __launch_bounds__(NUM_THREADS)
conv_tc_kernel(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_a1,
               const __grid_constant__ CUtensorMap map_a2, const __grid_constant__ CUtensorMap map_a3,
               const __grid_constant__ CUtensorMap map_b, const __grid_constant__ TcParams p)
{
    using L = SmemLayout<BLOCK_N, STAGES>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);     // SWIZZLE_128B needs 1024-B alignment
    uint64_t* full_bar = (uint64_t*)(smem + L::BAR_OFFSET);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* accum_bar = empty_bar + STAGES;
    uint32_t* tmem_slot = (uint32_t*)(accum_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // tile coordinates
    const int tile_m = blockIdx.x, n_col0 = blockIdx.y * BLOCK_N;
    const int tw = tile_m % p.tiles_w;
    const int th = (tile_m / p.tiles_w) % p.tiles_h;
    const int tn = tile_m / (p.tiles_w * p.tiles_h);
    const int w0 = tw * p.bw, h0 = th * p.bh, n0 = tn * p.bn;
    const int cpb = p.Cin / BLOCK_K;                 // k-blocks per filter tap
    const int num_kb = p.ntaps * cpb;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&map_a0); prefetch_tmap(&map_b);
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
        // ===== TMA producer =====
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int kb = 0; kb < num_kb; ++kb) {
                const int tap = kb / cpb, c0 = (kb - tap * cpb) * BLOCK_K;
                mbar_wait(&empty_bar[stage], phase ^ 1);
                mbar_expect_tx(&full_bar[stage], L::STAGE_BYTES);
                uint8_t* sa = smem + stage * L::STAGE_BYTES;
                uint8_t* sb = sa + A_STAGE_BYTES;
                const int mi = p.tap_map[tap];
                const CUtensorMap* ma = mi == 0 ? &map_a0 : (mi == 1 ? &map_a1 : (mi == 2 ? &map_a2 : &map_a3));
                tma_load_4d(ma, sa, &full_bar[stage], c0, w0 + p.tap_dw[tap], h0 + p.tap_dh[tap], n0);
                tma_load_2d(&map_b, sb, &full_bar[stage], tap * p.Cin + c0, n_col0);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one thread) =====
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(BLOCK_M, BLOCK_N);
            int stage = 0; uint32_t phase = 0;
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                const uint32_t sa = smem_u32(smem + stage * L::STAGE_BYTES);
                const uint32_t sb = sa + A_STAGE_BYTES;
#pragma unroll
                for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                    const uint64_t da = make_smem_desc(sa + k * UMMA_K * 2);
                    const uint64_t db = make_smem_desc(sb + k * UMMA_K * 2);
                    umma_bf16(tmem_base, da, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
                }
                umma_commit(&empty_bar[stage]);        // frees the smem stage when these MMAs retire
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
            umma_commit(accum_bar);                    // accumulator complete
        }
    } else {
        // ===== epilogue: TMEM -> registers -> global =====
        const int quad = warp & 3;                     // TMEM lane quadrant this warp may access
        const int r = quad * 32 + lane;                // row of the M tile
        const int rw = r % p.bw, rh = (r / p.bw) % p.bh, rn = r / (p.bw * p.bh);
        const int ow = w0 + rw, oh = h0 + rh, on = n0 + rn;
        const bool valid = ow < p.Wo && oh < p.Ho && on < p.B;
        const size_t m = ((size_t)on * p.Ho + oh) * p.Wo + ow;
        TO* orow = (TO*)p.out + m * p.ldc + n_col0;
        const TO* rrow = p.residual ? (const TO*)p.residual + m * p.ld_res + n_col0 : nullptr;

        mbar_wait(accum_bar, 0);
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < BLOCK_N / 32; ++c) {
            uint32_t acc[32];
            tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(c * 32), acc);
            if (!valid) continue;
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(acc[j]);
            const int nb = n_col0 + c * 32;
            if (p.scale != nullptr) {
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4) {
                    const float4 s4 = __ldg(reinterpret_cast<const float4*>(p.scale + nb) + j4);
                    v[4 * j4] *= s4.x; v[4 * j4 + 1] *= s4.y; v[4 * j4 + 2] *= s4.z; v[4 * j4 + 3] *= s4.w;
                }
            }
            if (p.bias != nullptr) {
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4) {
                    const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + nb) + j4);
                    v[4 * j4] += b4.x; v[4 * j4 + 1] += b4.y; v[4 * j4 + 2] += b4.z; v[4 * j4 + 3] += b4.w;
                }
            }
            if constexpr (sizeof(TO) == 2) {
                if (rrow != nullptr) {
#pragma unroll
                    for (int j8 = 0; j8 < 4; ++j8) {
                        const uint4 u = *(reinterpret_cast<const uint4*>(rrow + c * 32) + j8);
                        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&w[q]);
                            v[8 * j8 + 2 * q] += __low2float(h); v[8 * j8 + 2 * q + 1] += __high2float(h);
                        }
                    }
                }
#pragma unroll
                for (int j8 = 0; j8 < 4; ++j8) {
                    uint32_t w[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        float a = v[8 * j8 + 2 * q], b = v[8 * j8 + 2 * q + 1];
                        if (p.relu) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); }
                        const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
                        w[q] = *reinterpret_cast<const uint32_t*>(&h);
                    }
                    *(reinterpret_cast<uint4*>(orow + c * 32) + j8) = make_uint4(w[0], w[1], w[2], w[3]);
                }
            } else {
                if (rrow != nullptr) {
#pragma unroll
                    for (int j4 = 0; j4 < 8; ++j4) {
                        const float4 r4 = *(reinterpret_cast<const float4*>(rrow + c * 32) + j4);
                        v[4 * j4] += r4.x; v[4 * j4 + 1] += r4.y; v[4 * j4 + 2] += r4.z; v[4 * j4 + 3] += r4.w;
                    }
                }
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4) {
                    float4 o4 = make_float4(v[4 * j4], v[4 * j4 + 1], v[4 * j4 + 2], v[4 * j4 + 3]);
                    if (p.relu) { o4.x = fmaxf(o4.x, 0.f); o4.y = fmaxf(o4.y, 0.f); o4.z = fmaxf(o4.z, 0.f); o4.w = fmaxf(o4.w, 0.f); }
                    *(reinterpret_cast<float4*>(orow + c * 32) + j4) = o4;
                }
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

// ---- host side ----------------------------------------------------------------------------
PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;

int encode_map(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
               const uint32_t* box)
{
    uint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult rc = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims,
                           strides_bytes, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (CUresult %d) rank=%d dims=[%llu,%llu,%llu,%llu] box=[%u,%u,%u,%u]", (int)rc, rank,
                  (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)(rank > 2 ? dims[2] : 0),
                  (unsigned long long)(rank > 3 ? dims[3] : 0), box[0], box[1], rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0);
        return SEDT_ERR_CUDA;
    }
    return SEDT_OK;
}

static inline int pow2_ceil(int v) { int p = 1; while (p < v) p <<= 1; return p; }

template <int BLOCK_N, int STAGES, typename TO>
int launch_variant(const CUtensorMap* maps, const CUtensorMap& mb, const TcParams& p, dim3 grid, cudaStream_t stream)
{
    using L = SmemLayout<BLOCK_N, STAGES>;
    auto kern = conv_tc_kernel<BLOCK_N, STAGES, TO>;
    static bool attr_set = false;
    if (!attr_set) {
        SEDT_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
        attr_set = true;
    }
    ProfScope _prof(PROF_GEMM_TC, stream);
    kern<<<grid, NUM_THREADS, L::TOTAL, stream>>>(maps[0], maps[1], maps[2], maps[3], mb, p);
    SEDT_COUNT_LAUNCH();
    SEDT_CHECK_CUDA(cudaGetLastError());
    return SEDT_OK;
}

}  // namespace

int tc_init()
{
    if (g_encode != nullptr) return SEDT_OK;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || fn == nullptr || qres != cudaDriverEntryPointSuccess) {
        set_error("tc_init: cuTensorMapEncodeTiled is not available (%s)", e != cudaSuccess ? cudaGetErrorString(e) : "no driver entry point");
        (void)cudaGetLastError();
        return SEDT_ERR_CUDA;
    }
    g_encode = (PFN_cuTensorMapEncodeTiled_v12000)fn;
    return SEDT_OK;
}

bool conv_tc_supported(const ConvGemm& g)
{
    if (g.in_dt != DT_BF16) return false;
    if (g.Cin % BLOCK_K != 0 || g.Cout % 64 != 0) return false;
    if (g.lda % 8 != 0 || g.ldc % 8 != 0 || (g.residual && g.ld_res % 8 != 0)) return false;
    if (((uintptr_t)g.in & 15) || ((uintptr_t)g.w & 15) || ((uintptr_t)g.out & 15) || ((uintptr_t)g.residual & 15)) return false;
    if (((uintptr_t)g.scale & 15) || ((uintptr_t)g.bias & 15)) return false;
    if (g.R != g.S || (g.R != 1 && g.R != 3)) return false;
    if (g.stride != 1 && g.stride != 2) return false;
    if (g.stride == 2 && (g.H < 2 || g.W < 2 || g.dil != 1)) return false;
    if (g.R == 3 && g.pad != g.dil) return false;
    if (g.R == 1 && g.pad != 0) return false;
    const int bw = pow2_ceil(g.Wo);
    if (bw > BLOCK_M) return false;
    if ((int64_t)g.B * g.Ho * g.Wo < 1) return false;
    return true;
}

int launch_conv_tc(const ConvGemm& g, cudaStream_t stream)
{
    SEDT_REQUIRE(conv_tc_supported(g), "conv_tc: unsupported shape");
    SEDT_TRY(tc_init());
    TcParams p{};
    p.scale = g.scale; p.bias = g.bias; p.residual = g.residual; p.out = g.out;
    p.ld_res = g.ld_res; p.ldc = g.ldc; p.B = g.B; p.Ho = g.Ho; p.Wo = g.Wo; p.Cin = g.Cin; p.relu = g.relu;
    p.bw = pow2_ceil(g.Wo);
    p.bh = std::min(pow2_ceil(g.Ho), BLOCK_M / p.bw);
    p.bn = BLOCK_M / (p.bw * p.bh);
    p.tiles_w = (int)ceil_div(g.Wo, p.bw);
    p.tiles_h = (int)ceil_div(g.Ho, p.bh);
    const int tiles_n = (int)ceil_div(g.B, p.bn);
    p.ntaps = g.R * g.S;

    // A tensor maps: one per input phase that the taps touch
    CUtensorMap maps[4];
    memset(maps, 0, sizeof(maps));
    const uint32_t box[4] = {(uint32_t)BLOCK_K, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bn};
    int phase_map[2][2] = {{-1, -1}, {-1, -1}};
    int nmaps = 0;
    for (int r = 0; r < g.R; ++r) {
        for (int s = 0; s < g.S; ++s) {
            const int th = r * g.dil - g.pad, tws = s * g.dil - g.pad;     // tap offset in input pixels
            int ph = 0, pw = 0, dh = th, dw = tws;
            if (g.stride == 2) {
                ph = ((th % 2) + 2) % 2; pw = ((tws % 2) + 2) % 2;
                dh = (th - ph) / 2; dw = (tws - pw) / 2;                 // floor division
            }
            if (phase_map[ph][pw] < 0) {
                const int st = g.stride;
                const uint64_t hs = (uint64_t)(g.H - ph + st - 1) / st, wsz = (uint64_t)(g.W - pw + st - 1) / st;
                const uint64_t dims[4] = {(uint64_t)g.Cin, wsz, hs, (uint64_t)g.B};
                const uint64_t strides[3] = {(uint64_t)st * g.lda * 2, (uint64_t)st * g.W * g.lda * 2, (uint64_t)g.H * g.W * g.lda * 2};
                const char* base = (const char*)g.in + ((size_t)ph * g.W + pw) * g.lda * 2;
                SEDT_TRY(encode_map(&maps[nmaps], base, 4, dims, strides, box));
                phase_map[ph][pw] = nmaps++;
            }
            const int tap = r * g.S + s;
            p.tap_map[tap] = (int8_t)phase_map[ph][pw];
            p.tap_dh[tap] = (int8_t)dh;
            p.tap_dw[tap] = (int8_t)dw;
        }
    }
    for (int i = nmaps; i < 4; ++i) maps[i] = maps[0];

    const int block_n = g.Cout % 128 == 0 ? 128 : 64;
    CUtensorMap mb;
    {
        const uint64_t K = (uint64_t)g.R * g.S * g.Cin;
        const uint64_t dims[2] = {K, (uint64_t)g.Cout};
        const uint64_t strides[1] = {K * 2};
        const uint32_t bbox[2] = {(uint32_t)BLOCK_K, (uint32_t)block_n};
        SEDT_TRY(encode_map(&mb, g.w, 2, dims, strides, bbox));
    }
    dim3 grid((unsigned)(p.tiles_w * p.tiles_h * tiles_n), (unsigned)(g.Cout / block_n));
    const bool f32 = g.out_dt == DT_F32;
    if (block_n == 128) {
        return f32 ? launch_variant<128, 3, float>(maps, mb, p, grid, stream)
                   : launch_variant<128, 3, __nv_bfloat16>(maps, mb, p, grid, stream);
    }
    return f32 ? launch_variant<64, 4, float>(maps, mb, p, grid, stream)
               : launch_variant<64, 4, __nv_bfloat16>(maps, mb, p, grid, stream);
}

}  // namespace sedt
