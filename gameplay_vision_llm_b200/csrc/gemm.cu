// gemm.cu — K2: persistent, warp-specialised tcgen05 GEMM with a fused epilogue.
//
//   out[M,N] = act(A[M,K] . W[N,K]^T + bias) + residual
//
// Both operands are K-major bf16, so a [rows x 64] box lands in shared memory as rows of 128 bytes —
// exactly one SWIZZLE_128B atom wide — and the same bytes are addressed by the UMMA shared-memory
// descriptor.  One CTA per SM loops over output tiles (m-major order: the CTAs running at the same
// time share one A panel and the whole weight stays in the 126 MB L2):
//
//   warp 0    TMA producer: A box 128x64, W box BNx64 per stage, mbarrier expect_tx
//   warp 1    tcgen05.mma issuer (one elected lane): 4 x (128 x BN x 16) per stage, accumulating in TMEM;
//             tcgen05.commit releases the smem stage / publishes the accumulator
//   warps 2-5 epilogue: tcgen05.ld 32 lanes x 32 columns -> bias, GELU, residual -> bf16/fp32 stores
//
// The accumulator is double-buffered in TMEM (2 x BN columns), so the epilogue of tile i overlaps the
// MMAs of tile i+1.  Ragged M / N / K edges are handled by TMA (out-of-bounds reads are zero) and by
// row / column guards on the stores.
#include "common.cuh"

namespace gvl {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int kGemmThreads = 192;
constexpr int A_STAGE_BYTES = BM * BK * 2;

struct GemmParams {
    const float* bias;
    const __nv_bfloat16* residual;
    int ldr;
    int res_row_mod;
    void* out;
    int ldo;
    int out_f32;
    int M, N, K;
    int act;
    int m_tiles, n_tiles, k_blocks;
};

template <int BN>
struct GemmCfg {
    static constexpr int B_STAGE_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
    static constexpr int STAGES = BN == 256 ? 4 : (BN == 192 ? 5 : 6);
    static constexpr int TMEM_COLS = (2 * BN <= 256) ? 256 : 512;
    static constexpr int BIAS_BYTES = 4 * BN * 4;
    static constexpr int BAR_BYTES = (2 * STAGES + 4) * 8 + 16;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BIAS_BYTES + BAR_BYTES + 1024 /*alignment slack*/;
};

template <int ACT>
__device__ __forceinline__ float apply_act(float x) {
    if (ACT == GVL_ACT_GELU_TANH) return gelu_tanh_f(x);
    if (ACT == GVL_ACT_GELU_ERF) return gelu_erf_f(x);
    return x;
}

template <int BN>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                         const GemmParams p) {
    using Cfg = GemmCfg<BN>;
    constexpr int STAGES = Cfg::STAGES;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;
    uint8_t* sB = smem + STAGES * A_STAGE_BYTES;
    float* sBias = reinterpret_cast<float*>(smem + STAGES * Cfg::STAGE_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES + Cfg::BIAS_BYTES);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + STAGES;
    uint64_t* tmem_full_bar = bars + 2 * STAGES;
    uint64_t* tmem_empty_bar = bars + 2 * STAGES + 2;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tmem_full_bar[s], 1);
            mbar_init(&tmem_empty_bar[s], 4);
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<Cfg::TMEM_COLS>(tmem_ptr_smem);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    const int num_tiles = p.m_tiles * p.n_tiles;

    if (warp == 0) {
        // ===== TMA producer =====
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int m_blk = tile / p.n_tiles, n_blk = tile % p.n_tiles;
                for (int kb = 0; kb < p.k_blocks; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
                    tma_load_2d(sA + stage * A_STAGE_BYTES, &tmA, &full_bar[stage], kb * BK, m_blk * BM);
                    tma_load_2d(sB + stage * Cfg::B_STAGE_BYTES, &tmB, &full_bar[stage], kb * BK, n_blk * BN);
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (elect_one()) {
            constexpr uint32_t idesc = umma_idesc_bf16(BM, BN);
            int stage = 0;
            uint32_t phase = 0;
            int as = 0;
            uint32_t aphase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                mbar_wait(&tmem_empty_bar[as], aphase ^ 1);
                tcgen05_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(as * BN);
                for (int kb = 0; kb < p.k_blocks; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tcgen05_fence_after();
                    const uint32_t a_addr = smem_u32(sA + stage * A_STAGE_BYTES);
                    const uint32_t b_addr = smem_u32(sB + stage * Cfg::B_STAGE_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        umma_bf16_ss(d_tmem, umma_desc_sw128(a_addr + k * 32), umma_desc_sw128(b_addr + k * 32), idesc,
                                     (uint32_t)((kb | k) != 0));
                    }
                    umma_commit(&empty_bar[stage]);
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                umma_commit(&tmem_full_bar[as]);
                as ^= 1;
                if (as == 0) aphase ^= 1;
            }
        }
    } else {
        // ===== epilogue warps =====
        const int q = warp & 3;  // TMEM lane quadrant this warp may read
        float* myBias = sBias + q * BN;
        int as = 0;
        uint32_t aphase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int m_blk = tile / p.n_tiles, n_blk = tile % p.n_tiles;
            const int n0 = n_blk * BN;
            for (int c = lane; c < BN; c += 32) myBias[c] = (p.bias != nullptr && n0 + c < p.N) ? p.bias[n0 + c] : 0.0f;
            __syncwarp();
            mbar_wait(&tmem_full_bar[as], aphase);
            tcgen05_fence_after();
            const int row = m_blk * BM + q * 32 + lane;
            const bool row_ok = row < p.M;
            const int rrow = p.res_row_mod > 0 ? (row % p.res_row_mod) : row;
            const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN);
#pragma unroll 1
            for (int chunk = 0; chunk < BN / 32; ++chunk) {
                const int c0 = n0 + chunk * 32;
                if (c0 >= p.N) break;  // warp-uniform
                uint32_t r[32];
                tmem_ld_32x32(t_addr + (uint32_t)(chunk * 32), r);
                tmem_ld_wait();
                if (row_ok) {
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const int col = c0 + g * 8;
                        if (col < p.N) {
                            float v[8];
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                float x = __uint_as_float(r[g * 8 + j]) + myBias[chunk * 32 + g * 8 + j];
                                if (p.act == GVL_ACT_GELU_TANH)
                                    x = gelu_tanh_f(x);
                                else if (p.act == GVL_ACT_GELU_ERF)
                                    x = gelu_erf_f(x);
                                v[j] = x;
                            }
                            if (p.residual != nullptr) {
                                const uint4 rv =
                                    *reinterpret_cast<const uint4*>(p.residual + (size_t)rrow * p.ldr + col);
                                v[0] += bf16_lo(rv.x); v[1] += bf16_hi(rv.x);
                                v[2] += bf16_lo(rv.y); v[3] += bf16_hi(rv.y);
                                v[4] += bf16_lo(rv.z); v[5] += bf16_hi(rv.z);
                                v[6] += bf16_lo(rv.w); v[7] += bf16_hi(rv.w);
                            }
                            if (p.out_f32) {
                                float* o = reinterpret_cast<float*>(p.out) + (size_t)row * p.ldo + col;
                                *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
                                *reinterpret_cast<float4*>(o + 4) = make_float4(v[4], v[5], v[6], v[7]);
                            } else {
                                __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + (size_t)row * p.ldo + col;
                                uint4 ov;
                                ov.x = pack_bf16x2(v[0], v[1]);
                                ov.y = pack_bf16x2(v[2], v[3]);
                                ov.z = pack_bf16x2(v[4], v[5]);
                                ov.w = pack_bf16x2(v[6], v[7]);
                                *reinterpret_cast<uint4*>(o) = ov;
                            }
                        }
                    }
                }
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty_bar[as]);
            as ^= 1;
            if (as == 0) aphase ^= 1;
        }
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
    }
}

template <int BN>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, cudaStream_t stream) {
    using Cfg = GemmCfg<BN>;
    GVL_CUDA(cudaFuncSetAttribute(gemm_bf16_tcgen05_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  Cfg::SMEM_BYTES));
    const int tiles = p.m_tiles * p.n_tiles;
    const int grid = tiles < sm_count() ? tiles : sm_count();
    ProfScope prof(GVL_K_GEMM, 2.0 * p.M * (double)p.N * p.K, stream);
    gemm_bf16_tcgen05_kernel<BN><<<grid, kGemmThreads, Cfg::SMEM_BYTES, stream>>>(tmA, tmB, p);
    GVL_LAUNCH_CHECK("gemm_bf16_tcgen05_kernel");
    return 0;
}

static int pick_bn(int N) {
    // smallest padded width wins; ties go to the wider tile (fewer A re-reads, more MMA per barrier).
    const int cands[3] = {256, 192, 128};
    int best = 256;
    long best_pad = -1;
    for (int i = 0; i < 3; ++i) {
        const int bn = cands[i];
        const long padded = (long)((N + bn - 1) / bn) * bn;
        if (best_pad < 0 || padded < best_pad) {
            best_pad = padded;
            best = bn;
        }
    }
    return best;
}

}  // namespace gvl

extern "C" int gvl_gemm_bf16(const void* A, int lda, const void* W, int ldw, const float* bias, const void* residual,
                             int ldr, int res_row_mod, void* out, int ldo, int out_f32, int M, int N, int K, int act,
                             void* stream) {
    using namespace gvl;
    GVL_CHECK_ARG(A && W && out, "gvl_gemm_bf16: null pointer");
    GVL_CHECK_ARG(M > 0 && N > 0 && K > 0, "gvl_gemm_bf16: bad shape M=%d N=%d K=%d", M, N, K);
    GVL_CHECK_ARG(N % 8 == 0 && K % 8 == 0, "gvl_gemm_bf16: N and K must be multiples of 8 (N=%d K=%d)", N, K);
    GVL_CHECK_ARG(lda % 8 == 0 && ldw % 8 == 0 && lda >= K && ldw >= K, "gvl_gemm_bf16: bad lda/ldw %d/%d", lda, ldw);
    GVL_CHECK_ARG(ldo >= N && ldo % (out_f32 ? 4 : 8) == 0, "gvl_gemm_bf16: bad ldo %d", ldo);
    GVL_CHECK_ARG(((uintptr_t)A % 16 == 0) && ((uintptr_t)W % 16 == 0) && ((uintptr_t)out % 16 == 0),
                  "gvl_gemm_bf16: pointers must be 16-byte aligned");
    GVL_CHECK_ARG(residual == nullptr || (ldr % 8 == 0 && ldr >= N && (uintptr_t)residual % 16 == 0),
                  "gvl_gemm_bf16: bad residual ld/alignment");
    GVL_CHECK_ARG(act >= 0 && act <= 2, "gvl_gemm_bf16: bad act %d", act);

    const int bn = pick_bn(N);
    GemmParams p;
    p.bias = bias;
    p.residual = reinterpret_cast<const __nv_bfloat16*>(residual);
    p.ldr = ldr;
    p.res_row_mod = res_row_mod;
    p.out = out;
    p.ldo = ldo;
    p.out_f32 = out_f32;
    p.M = M;
    p.N = N;
    p.K = K;
    p.act = act;
    p.m_tiles = (M + BM - 1) / BM;
    p.n_tiles = (N + bn - 1) / bn;
    p.k_blocks = (K + BK - 1) / BK;

    CUtensorMap tmA, tmB;
    int rc = make_tmap_2d_bf16(&tmA, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, BM, BK, true);
    if (rc) return rc;
    rc = make_tmap_2d_bf16(&tmB, W, (uint64_t)N, (uint64_t)K, (uint64_t)ldw, (uint32_t)bn, BK, true);
    if (rc) return rc;

    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    switch (bn) {
        case 256: return launch_gemm<256>(tmA, tmB, p, s);
        case 192: return launch_gemm<192>(tmA, tmB, p, s);
        default: return launch_gemm<128>(tmA, tmB, p, s);
    }
}
