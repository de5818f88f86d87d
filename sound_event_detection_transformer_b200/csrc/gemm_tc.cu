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
#include "tc_common.cuh"

namespace sedt {
namespace {

using namespace tc;

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
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // stays a __shared__ pointer (LDS/STS)
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
template <int BLOCK_N, int STAGES, typename TO>
int launch_v1(const TcProblem& pr, dim3 grid, cudaStream_t stream)
{
    using L = SmemLayout<BLOCK_N, STAGES>;
    auto kern = conv_tc_kernel<BLOCK_N, STAGES, TO>;
    static bool attr_set = false;
    if (!attr_set) {
        SEDT_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
        attr_set = true;
    }
    ProfScope _prof(PROF_GEMM_TC, stream);
    kern<<<grid, NUM_THREADS, L::TOTAL, stream>>>(pr.map_a[0], pr.map_a[1], pr.map_a[2], pr.map_a[3], pr.map_b, pr.p);
    SEDT_COUNT_KIND(KK_CONV_TC_V1);
    SEDT_CHECK_CUDA(cudaGetLastError());
    return SEDT_OK;
}

PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;
int g_num_sms = 0;

static inline int pow2_ceil(int v) { int p = 1; while (p < v) p <<= 1; return p; }

}  // namespace

namespace tc {

int num_sms()
{
    if (g_num_sms == 0) {
        int dev = 0, n = 148;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        g_num_sms = n > 0 ? n : 148;
    }
    return g_num_sms;
}

int encode_map(CUtensorMap* map, CUtensorMapDataType dt, const void* base, int rank, const uint64_t* dims,
               const uint64_t* strides_bytes, const uint32_t* box)
{
    SEDT_TRY(tc_init());
    uint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult rc = g_encode(map, dt, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (CUresult %d) rank=%d dims=[%llu,%llu,%llu,%llu] box=[%u,%u,%u,%u]", (int)rc, rank,
                  (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)(rank > 2 ? dims[2] : 0),
                  (unsigned long long)(rank > 3 ? dims[3] : 0), box[0], box[1], rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0);
        return SEDT_ERR_CUDA;
    }
    return SEDT_OK;
}

int build_problem(const ConvGemm& g, int block_n, TcProblem* out, int b_split)
{
    TcParams& p = out->p;
    memset(out, 0, sizeof(*out));
    p.scale = g.scale; p.bias = g.bias; p.residual = g.residual; p.out = g.out;
    p.ld_res = g.ld_res; p.ldc = g.ldc; p.B = g.B; p.Ho = g.Ho; p.Wo = g.Wo; p.Cin = g.Cin; p.relu = g.relu;
    p.bw = pow2_ceil(g.Wo);
    p.bh = std::min(pow2_ceil(g.Ho), BLOCK_M / p.bw);
    p.bn = BLOCK_M / (p.bw * p.bh);
    p.tiles_w = (int)ceil_div(g.Wo, p.bw);
    p.tiles_h = (int)ceil_div(g.Ho, p.bh);
    p.ntaps = g.R * g.S;
    out->tiles_m = p.tiles_w * p.tiles_h * (int)ceil_div(g.B, p.bn);
    out->tiles_nc = g.Cout / block_n;
    out->block_n = block_n;

    // A tensor maps: one per input phase that the taps touch (stride 2 => even/odd rows x even/odd columns)
    const uint32_t box[4] = {(uint32_t)BLOCK_K, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bn};
    int phase_map[2][2] = {{-1, -1}, {-1, -1}};
    int nmaps = 0;
    for (int r = 0; r < g.R; ++r) {
        for (int s = 0; s < g.S; ++s) {
            const int th = r * g.dil - g.pad, tws = s * g.dil - g.pad;     // tap offset in input pixels
            int ph = 0, pw = 0, dh = th, dw = tws;
            if (g.stride == 2) {
                ph = ((th % 2) + 2) % 2; pw = ((tws % 2) + 2) % 2;
                dh = (th - ph) / 2; dw = (tws - pw) / 2;                 // exact: th - ph is even
            }
            if (phase_map[ph][pw] < 0) {
                const int st = g.stride;
                const uint64_t hs = (uint64_t)(g.H - ph + st - 1) / st, wsz = (uint64_t)(g.W - pw + st - 1) / st;
                const uint64_t dims[4] = {(uint64_t)g.Cin, wsz, hs, (uint64_t)g.B};
                const uint64_t strides[3] = {(uint64_t)st * g.lda * 2, (uint64_t)st * g.W * g.lda * 2, (uint64_t)g.H * g.W * g.lda * 2};
                const char* base = (const char*)g.in + ((size_t)ph * g.W + pw) * g.lda * 2;
                SEDT_TRY(encode_map(&out->map_a[nmaps], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, base, 4, dims, strides, box));
                phase_map[ph][pw] = nmaps++;
            }
            const int tap = r * g.S + s;
            p.tap_map[tap] = (int8_t)phase_map[ph][pw];
            p.tap_dh[tap] = (int8_t)dh;
            p.tap_dw[tap] = (int8_t)dw;
        }
    }
    for (int i = nmaps; i < 4; ++i) out->map_a[i] = out->map_a[0];
    if (g.w == nullptr) return SEDT_OK;                    // activation-side maps only (gemm_wgrad.cu)
    const uint64_t K = (uint64_t)g.R * g.S * g.Cin;
    const uint64_t bdims[2] = {K, (uint64_t)g.Cout};
    const uint64_t bstrides[1] = {K * 2};
    const uint32_t bbox[2] = {(uint32_t)BLOCK_K, (uint32_t)(block_n / b_split)};     // b_split = 2: each CTA of a pair loads half
    return encode_map(&out->map_b, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, g.w, 2, bdims, bstrides, bbox);
}

int encode_out_map(CUtensorMap* m, const void* base, int ld, bool f32, const ConvGemm& g, const TcParams& p)
{
    const int es = f32 ? 4 : 2;
    const uint64_t dims[4] = {(uint64_t)g.Cout, (uint64_t)g.Wo, (uint64_t)g.Ho, (uint64_t)g.B};
    const uint64_t strides[3] = {(uint64_t)ld * es, (uint64_t)g.Wo * ld * es, (uint64_t)g.Ho * g.Wo * ld * es};
    const uint32_t box[4] = {(uint32_t)(128 / es), (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bn};
    return encode_map(m, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, base, 4, dims, strides, box);
}


}  // namespace tc

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
    if (pow2_ceil(g.Wo) > BLOCK_M) return false;
    if ((int64_t)g.B * g.Ho * g.Wo < 1) return false;
    return true;
}

// v1: one tile per CTA, direct-global epilogue.  Kept as the simple reference implementation of the
// tcgen05 path (SEDT_TC_V1=1 selects it); the persistent kernel in gemm_tc2.cu is the production one.
int launch_conv_tc_v1(const ConvGemm& g, cudaStream_t stream)
{
    SEDT_REQUIRE(conv_tc_supported(g), "conv_tc: unsupported shape");
    SEDT_REQUIRE(g.relu != 2, "conv_tc v1: the ReLU-mask epilogue is only implemented in the persistent kernels");
    const int block_n = g.Cout % 128 == 0 ? 128 : 64;
    TcProblem pr;
    SEDT_TRY(build_problem(g, block_n, &pr));
    dim3 grid((unsigned)pr.tiles_m, (unsigned)pr.tiles_nc);
    const bool f32 = g.out_dt == DT_F32;
    if (block_n == 128) return f32 ? launch_v1<128, 3, float>(pr, grid, stream) : launch_v1<128, 3, __nv_bfloat16>(pr, grid, stream);
    return f32 ? launch_v1<64, 4, float>(pr, grid, stream) : launch_v1<64, 4, __nv_bfloat16>(pr, grid, stream);
}

}  // namespace sedt
