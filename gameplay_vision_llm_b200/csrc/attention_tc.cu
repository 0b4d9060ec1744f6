// attention_tc.cu — K4 on the 5th-generation tensor cores: fused non-causal self-attention with both
// GEMMs (S = Q K^T and O += P V) issued as tcgen05.mma, accumulators in tensor memory.
//
// One CTA = 128 query rows of one (image, head); keys/values stream in blocks of 128 through a 2-stage
// TMA pipeline.  Roles (192 threads, two CTAs per SM so one CTA's softmax overlaps the other's MMAs):
//
//   warp 0     TMA producer: 4-D tensor maps over qkv viewed as [image][token][3*heads][head_dim]; the head dim
//              72 is fetched as a 64-wide SWIZZLE_128B panel plus a 16-wide SWIZZLE_32B panel whose upper
//              8 columns are out of bounds and therefore zero — the k-padding 72 -> 80 costs no memory
//   warp 1     MMA issuer: S[128x128] = Q.K^T (SS, both K-major), then O[128 x hd] += P.V with P read from
//              TMEM (TS form) and V as an MN-major shared-memory operand (64-wide and 16-wide N panels)
//   warps 2-5  softmax, one thread per query row: tcgen05.ld the S row, running max with lazy rescaling
//              (O is only rescaled when the max grows by more than 2^8), p = exp2(s*c - m), row sum in fp32,
//              P packed to bf16 and written back over the S columns with tcgen05.st
//
// TMEM columns: [0,128) S (fp32) aliased by P (bf16 pairs, columns [0,64)); [128, 128+80) O.
#include "common.cuh"

namespace gvl {

constexpr int ATC_BQ = 128;
constexpr int ATC_BKV = 128;
constexpr int ATC_THREADS = 192;
constexpr int ATC_TMEM_COLS = 256;
constexpr float ATC_RESCALE_THRESHOLD = 8.0f;  // log2 units

template <int HD>
struct AtcCfg {
    static constexpr bool TAIL = HD > 64;  // second, 16-wide panel for d in [64, 80)
    static constexpr int DPAD = TAIL ? 80 : 64;
    static constexpr int P0_BYTES = 128 * 128;            // 128 rows x 64 bf16, SWIZZLE_128B
    static constexpr int P1_BYTES = TAIL ? 128 * 32 : 0;  // 128 rows x 16 bf16, SWIZZLE_32B
    static constexpr int TILE_BYTES = P0_BYTES + P1_BYTES;
    static constexpr int KV_BYTES = 2 * TILE_BYTES;
    static constexpr int STAGES = 2;
    static constexpr int SMEM_BYTES = TILE_BYTES * (1 + 2 * STAGES) + 256 /*barriers*/ + 1024 /*alignment*/;
};

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <int HD>
__global__ void __launch_bounds__(ATC_THREADS, 2)
attention_tc_kernel(const __grid_constant__ CUtensorMap tm64, const __grid_constant__ CUtensorMap tm16,
                    __nv_bfloat16* __restrict__ out, int T, int H, float scale_log2) {
    using Cfg = AtcCfg<HD>;
    extern __shared__ uint8_t atc_smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(atc_smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;                                   // [P0 | P1]
    uint8_t* sK = smem + Cfg::TILE_BYTES;                 // [stage][P0 | P1]
    uint8_t* sV = sK + Cfg::STAGES * Cfg::TILE_BYTES;     // [stage][P0 | P1]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sV + Cfg::STAGES * Cfg::TILE_BYTES);
    uint64_t* q_full = bars;
    uint64_t* kv_full = bars + 1;   // [2]
    uint64_t* kv_empty = bars + 3;  // [2]
    uint64_t* s_full = bars + 5;
    uint64_t* p_full = bars + 6;
    uint64_t* pv_done = bars + 7;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * ATC_BQ;
    const int h = blockIdx.y, b = blockIdx.z;
    const int nblk = (T + ATC_BKV - 1) / ATC_BKV;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm64);
        if (Cfg::TAIL) tma_prefetch_desc(&tm16);
        mbar_init(q_full, 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(&kv_full[s], 1);
            mbar_init(&kv_empty[s], 1);
        }
        mbar_init(s_full, 1);
        mbar_init(p_full, 4);
        mbar_init(pv_done, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<ATC_TMEM_COLS>(tmem_ptr_smem);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    const uint32_t tS = tmem_base;        // S / P
    const uint32_t tO = tmem_base + 128;  // O

    if (warp == 0) {
        // ===== TMA producer =====
        if (elect_one()) {
            mbar_arrive_expect_tx(q_full, Cfg::TILE_BYTES);
            tma_load_4d(sQ, &tm64, q_full, 0, h, q0, b);
            if (Cfg::TAIL) tma_load_4d(sQ + Cfg::P0_BYTES, &tm16, q_full, 64, h, q0, b);
            for (int j = 0; j < nblk; ++j) {
                const int st = j & 1;
                const uint32_t ph = (uint32_t)(j >> 1) & 1u;
                mbar_wait(&kv_empty[st], ph ^ 1);
                mbar_arrive_expect_tx(&kv_full[st], Cfg::KV_BYTES);
                uint8_t* k = sK + st * Cfg::TILE_BYTES;
                uint8_t* v = sV + st * Cfg::TILE_BYTES;
                tma_load_4d(k, &tm64, &kv_full[st], 0, H + h, j * ATC_BKV, b);
                tma_load_4d(v, &tm64, &kv_full[st], 0, 2 * H + h, j * ATC_BKV, b);
                if (Cfg::TAIL) {
                    tma_load_4d(k + Cfg::P0_BYTES, &tm16, &kv_full[st], 64, H + h, j * ATC_BKV, b);
                    tma_load_4d(v + Cfg::P0_BYTES, &tm16, &kv_full[st], 64, 2 * H + h, j * ATC_BKV, b);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (elect_one()) {
            constexpr uint32_t idescS = umma_idesc_bf16_major(128, ATC_BKV, 0, 0);  // Q, K both K-major
            constexpr uint32_t idescV64 = umma_idesc_bf16_major(128, 64, 0, 1);     // V: MN-major B
            constexpr uint32_t idescV16 = umma_idesc_bf16_major(128, 16, 0, 1);
            mbar_wait(q_full, 0);
            const uint32_t q_addr = smem_u32(sQ);
            for (int j = 0; j < nblk; ++j) {
                const int st = j & 1;
                const uint32_t ph = (uint32_t)(j >> 1) & 1u;
                mbar_wait(&kv_full[st], ph);
                if (j > 0) mbar_wait(pv_done, (uint32_t)(j - 1) & 1u);  // P_{j-1} (aliases S) fully consumed
                tcgen05_fence_after();
                const uint32_t k_addr = smem_u32(sK + st * Cfg::TILE_BYTES);
                const uint32_t v_addr = smem_u32(sV + st * Cfg::TILE_BYTES);
                // S = Q K^T : 4 k-steps over the 64-wide panel (+1 over the zero-padded 16-wide panel)
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_bf16_ss(tS, umma_desc_sw128(q_addr + k * 32), umma_desc_sw128(k_addr + k * 32), idescS,
                                 (uint32_t)(k > 0));
                if (Cfg::TAIL)
                    umma_bf16_ss(tS, umma_desc(q_addr + Cfg::P0_BYTES, 0, 256, 6),
                                 umma_desc(k_addr + Cfg::P0_BYTES, 0, 256, 6), idescS, 1u);
                umma_commit(s_full);
                // O += P V : k runs over the keys of this block, 16 per MMA
                mbar_wait(p_full, (uint32_t)j & 1u);
                tcgen05_fence_after();
                const int valid = min(ATC_BKV, T - j * ATC_BKV);
                const int ksteps = (valid + 15) >> 4;
                for (int kk = 0; kk < ksteps; ++kk) {
                    const uint32_t acc = (uint32_t)((j | kk) != 0);
                    umma_bf16_ts(tO, tS + (uint32_t)(kk * 8), umma_desc(v_addr + kk * 2048, 0, 1024, 2), idescV64, acc);
                    if (Cfg::TAIL)
                        umma_bf16_ts(tO + 64, tS + (uint32_t)(kk * 8),
                                     umma_desc(v_addr + Cfg::P0_BYTES + kk * 512, 0, 256, 6), idescV16, acc);
                }
                umma_commit(&kv_empty[st]);
                umma_commit(pv_done);
            }
        }
    } else {
        // ===== softmax: one thread per query row =====
        const int q = warp & 3;
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        const int row = q0 + q * 32 + lane;
        float m_used = -INFINITY, l = 0.f;
        for (int j = 0; j < nblk; ++j) {
            mbar_wait(s_full, (uint32_t)j & 1u);
            tcgen05_fence_after();
            uint32_t s[4][32];
#pragma unroll
            for (int c = 0; c < 4; ++c) tmem_ld_32x32(tS + lane_off + (uint32_t)(c * 32), s[c]);
            tmem_ld_wait();
            const int valid = min(ATC_BKV, T - j * ATC_BKV);
            if (valid < ATC_BKV) {
#pragma unroll
                for (int c = 0; c < 4; ++c)
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (c * 32 + i >= valid) s[c][i] = 0xff800000u;  // -inf
            }
            float mx = -INFINITY;
#pragma unroll
            for (int c = 0; c < 4; ++c)
#pragma unroll
                for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(s[c][i]));
            const float mt = mx * scale_log2;
            float factor = 1.0f;
            bool need = false;
            if (j == 0) {
                m_used = mt;
            } else if (mt > m_used + ATC_RESCALE_THRESHOLD) {
                need = true;
                factor = ex2_approx(m_used - mt);
                m_used = mt;
                l *= factor;
            }
            const float nm = -m_used;
            float rs = 0.f;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint32_t pk[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float p0 = ex2_approx(fmaf(__uint_as_float(s[c][2 * i]), scale_log2, nm));
                    const float p1 = ex2_approx(fmaf(__uint_as_float(s[c][2 * i + 1]), scale_log2, nm));
                    rs += p0 + p1;
                    pk[i] = pack_bf16x2(p0, p1);
                }
                tmem_st_32x16(tS + lane_off + (uint32_t)(c * 16), pk);
            }
            l += rs;
            if (__any_sync(0xffffffffu, need)) {
                // rare: the running max grew by more than the threshold -> rescale this warp's O rows
                uint32_t o[32];
                tmem_ld_32x32(tO + lane_off, o);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * factor);
                tmem_st_32x32(tO + lane_off, o);
                tmem_ld_32x32(tO + lane_off + 32, o);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * factor);
                tmem_st_32x32(tO + lane_off + 32, o);
                if (Cfg::TAIL) {
                    uint32_t o2[16];
                    tmem_ld_32x16(tO + lane_off + 64, o2);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; ++i) o2[i] = __float_as_uint(__uint_as_float(o2[i]) * factor);
                    tmem_st_32x16(tO + lane_off + 64, o2);
                }
            }
            tmem_st_wait();
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(p_full);
        }
        // ---- finalise: O / l -> bf16 -> global ----
        mbar_wait(pv_done, (uint32_t)(nblk - 1) & 1u);
        tcgen05_fence_after();
        const float inv = 1.0f / l;
        const int D = H * HD;
        __nv_bfloat16* orow = out + ((size_t)b * T + row) * D + (size_t)h * HD;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            uint32_t o[32];
            tmem_ld_32x32(tO + lane_off + (uint32_t)(c * 32), o);
            tmem_ld_wait();
            if (row < T) {
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    uint4 v;
                    v.x = pack_bf16x2(__uint_as_float(o[g * 8 + 0]) * inv, __uint_as_float(o[g * 8 + 1]) * inv);
                    v.y = pack_bf16x2(__uint_as_float(o[g * 8 + 2]) * inv, __uint_as_float(o[g * 8 + 3]) * inv);
                    v.z = pack_bf16x2(__uint_as_float(o[g * 8 + 4]) * inv, __uint_as_float(o[g * 8 + 5]) * inv);
                    v.w = pack_bf16x2(__uint_as_float(o[g * 8 + 6]) * inv, __uint_as_float(o[g * 8 + 7]) * inv);
                    *reinterpret_cast<uint4*>(orow + c * 32 + g * 8) = v;
                }
            }
        }
        if (Cfg::TAIL) {
            uint32_t o2[16];
            tmem_ld_32x16(tO + lane_off + 64, o2);
            tmem_ld_wait();
            if (row < T) {
                uint4 v;  // d = 64..71 (columns 72..79 are the zero padding)
                v.x = pack_bf16x2(__uint_as_float(o2[0]) * inv, __uint_as_float(o2[1]) * inv);
                v.y = pack_bf16x2(__uint_as_float(o2[2]) * inv, __uint_as_float(o2[3]) * inv);
                v.z = pack_bf16x2(__uint_as_float(o2[4]) * inv, __uint_as_float(o2[5]) * inv);
                v.w = pack_bf16x2(__uint_as_float(o2[6]) * inv, __uint_as_float(o2[7]) * inv);
                *reinterpret_cast<uint4*>(orow + 64) = v;
            }
        }
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        tmem_dealloc<ATC_TMEM_COLS>(tmem_base);
    }
}

template <int HD>
int launch_attention_tc(const void* qkv, void* out, int B, int T, int H, float scale, cudaStream_t s) {
    using Cfg = AtcCfg<HD>;
    // qkv viewed as [B][T][3H][HD], innermost first
    const uint64_t dims[4] = {(uint64_t)HD, (uint64_t)3 * H, (uint64_t)T, (uint64_t)B};
    const uint64_t strides[3] = {(uint64_t)HD * 2, (uint64_t)3 * H * HD * 2, (uint64_t)T * 3 * H * HD * 2};
    const uint32_t box64[4] = {64, 1, 128, 1};
    const uint32_t box16[4] = {16, 1, 128, 1};
    CUtensorMap tm64, tm16;
    int rc = make_tmap_nd_bf16(&tm64, qkv, 4, dims, strides, box64, 128);
    if (rc) return rc;
    rc = make_tmap_nd_bf16(&tm16, qkv, 4, dims, strides, Cfg::TAIL ? box16 : box64, Cfg::TAIL ? 32 : 128);
    if (rc) return rc;
    GVL_CUDA(cudaFuncSetAttribute(attention_tc_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    dim3 grid((T + ATC_BQ - 1) / ATC_BQ, H, B);
    ProfScope prof(GVL_K_ATTENTION, 4.0 * B * (double)H * T * (double)T * HD, s);
    attention_tc_kernel<HD><<<grid, ATC_THREADS, Cfg::SMEM_BYTES, s>>>(tm64, tm16, reinterpret_cast<__nv_bfloat16*>(out), T,
                                                                      H, scale * 1.4426950408889634f);
    GVL_LAUNCH_CHECK("attention_tc_kernel");
    return 0;
}

template int launch_attention_tc<72>(const void*, void*, int, int, int, float, cudaStream_t);
template int launch_attention_tc<64>(const void*, void*, int, int, int, float, cudaStream_t);

}  // namespace gvl
