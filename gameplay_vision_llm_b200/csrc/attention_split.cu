// attention_split.cu — K4 variant: the S-double-buffered tcgen05 attention of attention_sdb.cu with every query
// row's softmax SPLIT between two threads of different warps.
//
// Why: the softmax of attention_sdb.cu is a single-warp-latency-bound instruction stream (tcgen05.ld latency,
// dependent max chain, exp2 at 8 issue cycles per warp instruction, tcgen05.st + wait) — with one thread per row a
// CTA has four softmax warps, two CTAs per SM give two warps per scheduler, and the MUFU pipe idles more than half
// of the time (ncu: 45 % busy).  Tensor memory caps the SM at two CTAs (2 x (2 x 64 S + 80 O) columns), so the way
// to more warps per scheduler is more warps per row tile: warps w and w + 4 own the same TMEM lane quadrant
// (a warp may only touch lanes 32 (w % 4) .. +31) and split each 64-key block into keys 0-31 / 32-63.
//
//   warps 0-7   softmax: thread (w, l) handles row 32 (w & 3) + l, key half w >> 2.  Per block: tcgen05.ld of its 32
//               scores, local max, exchange of the two half maxima through shared memory (double-buffered slot, one
//               64-thread named barrier per pair), identical lazy-rescale decision in both threads,
//               p = exp2(s c - m), partial row sum, 16 packed bf16 columns written over its half of the S buffer.
//               Rare rescale of O and the final O / l read-out are split by columns (0-31 | 32-79).
//   warp 8      TMA producer, warp 9 MMA issuer — unchanged from attention_sdb.cu (p_full now counts 8 warps).
#include "common.cuh"

#include <cstdlib>

namespace gvl {

constexpr int SPL_BQ = 128;
constexpr int SPL_BKV = 64;
constexpr float SPL_RESCALE_THRESHOLD = 8.0f;  // log2 units

template <int HD>
struct SplCfg {
    static constexpr int SOFTMAX_WARPS = 8;
    static constexpr int THREADS = (SOFTMAX_WARPS + 2) * 32;
    static constexpr int TMEM_COLS = 256;
    static constexpr int STAGES = 4;
    static constexpr int O_COL = 128;
    static constexpr bool TAIL = HD > 64;
    static constexpr int DPAD = TAIL ? 80 : 64;
    static constexpr int Q_P0 = 128 * 128;
    static constexpr int Q_P1 = TAIL ? 128 * 32 : 0;
    static constexpr int Q_BYTES = Q_P0 + Q_P1;
    static constexpr int KV_P0 = SPL_BKV * 128;
    static constexpr int KV_P1 = TAIL ? SPL_BKV * 32 : 0;
    static constexpr int KV_BYTES = KV_P0 + KV_P1;
    static constexpr int XCH_BYTES = 2 * 128 * 2 * 4;  // [slot][row][half] floats
    static constexpr int SMEM_BYTES = Q_BYTES + 2 * STAGES * KV_BYTES + XCH_BYTES + 256 /*barriers*/ + 1024 /*alignment*/;
};

__device__ __forceinline__ float spl_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <int HD>
__global__ void __launch_bounds__(SplCfg<HD>::THREADS, 2)
attention_split_kernel(const __grid_constant__ CUtensorMap tmq64, const __grid_constant__ CUtensorMap tmq16,
                       const __grid_constant__ CUtensorMap tmk64, const __grid_constant__ CUtensorMap tmk16,
                       __nv_bfloat16* __restrict__ out, int T, int H, float scale_log2) {
    using Cfg = SplCfg<HD>;
    constexpr int STAGES = Cfg::STAGES;
    constexpr int W_TMA = Cfg::SOFTMAX_WARPS, W_MMA = Cfg::SOFTMAX_WARPS + 1;
    extern __shared__ uint8_t spl_smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(spl_smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;
    uint8_t* sK = smem + Cfg::Q_BYTES;
    uint8_t* sV = sK + STAGES * Cfg::KV_BYTES;
    float* sX = reinterpret_cast<float*>(sV + STAGES * Cfg::KV_BYTES);  // [2][128][2]
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sX) + Cfg::XCH_BYTES);
    uint64_t* q_full = bars;
    uint64_t* kv_full = bars + 1;
    uint64_t* kv_empty = kv_full + STAGES;
    uint64_t* s_full = kv_empty + STAGES;   // [buffer]
    uint64_t* p_full = s_full + 2;          // [buffer]  P(j) written (8 warps)
    uint64_t* pv_done = p_full + 2;         // PV(j) complete, one phase per block (rare rescale path)
    uint64_t* o_done = pv_done + 1;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(o_done + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * SPL_BQ;
    const int h = blockIdx.y, b = blockIdx.z;
    const int nblk = (T + SPL_BKV - 1) / SPL_BKV;

    if (warp == W_TMA && lane == 0) {
        tma_prefetch_desc(&tmq64);
        tma_prefetch_desc(&tmk64);
        if (Cfg::TAIL) {
            tma_prefetch_desc(&tmq16);
            tma_prefetch_desc(&tmk16);
        }
        mbar_init(q_full, 1);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&kv_full[s], 1);
            mbar_init(&kv_empty[s], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&s_full[i], 1);
            mbar_init(&p_full[i], Cfg::SOFTMAX_WARPS);
        }
        mbar_init(pv_done, 1);
        mbar_init(o_done, 1);
        fence_barrier_init();
    }
    if (warp == W_MMA) tmem_alloc<Cfg::TMEM_COLS>(tmem_ptr_smem);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp == W_TMA) {
        // ===== TMA producer =====
        if (elect_one()) {
            mbar_arrive_expect_tx(q_full, Cfg::Q_BYTES);
            tma_load_4d(sQ, &tmq64, q_full, 0, h, q0, b);
            if (Cfg::TAIL) tma_load_4d(sQ + Cfg::Q_P0, &tmq16, q_full, 64, h, q0, b);
            int st = 0;
            uint32_t ph = 0;
            for (int j = 0; j < nblk; ++j) {
                mbar_wait(&kv_empty[st], ph ^ 1);
                mbar_arrive_expect_tx(&kv_full[st], 2 * Cfg::KV_BYTES);
                uint8_t* k = sK + st * Cfg::KV_BYTES;
                uint8_t* v = sV + st * Cfg::KV_BYTES;
                tma_load_4d(k, &tmk64, &kv_full[st], 0, H + h, j * SPL_BKV, b);
                tma_load_4d(v, &tmk64, &kv_full[st], 0, 2 * H + h, j * SPL_BKV, b);
                if (Cfg::TAIL) {
                    tma_load_4d(k + Cfg::KV_P0, &tmk16, &kv_full[st], 64, H + h, j * SPL_BKV, b);
                    tma_load_4d(v + Cfg::KV_P0, &tmk16, &kv_full[st], 64, 2 * H + h, j * SPL_BKV, b);
                }
                if (++st == STAGES) {
                    st = 0;
                    ph ^= 1;
                }
            }
        }
    } else if (warp == W_MMA) {
        // ===== MMA issuer =====
        if (elect_one()) {
            constexpr uint32_t idescS = umma_idesc_bf16_major(128, SPL_BKV, 0, 0);  // Q, K both K-major
            constexpr uint32_t idescV64 = umma_idesc_bf16_major(128, 64, 0, 1);     // V: MN-major B
            constexpr uint32_t idescV16 = umma_idesc_bf16_major(128, 16, 0, 1);
            const uint32_t q_addr = smem_u32(sQ);
            const uint32_t tO = tmem_base + Cfg::O_COL;
            auto issue_s = [&](int j) {  // S(j) -> buffer j & 1; K(j) sits in stage j % STAGES
                const int st = j % STAGES;
                mbar_wait(&kv_full[st], (uint32_t)(j / STAGES) & 1u);
                tcgen05_fence_after();
                const uint32_t k_addr = smem_u32(sK + st * Cfg::KV_BYTES);
                const uint32_t tS = tmem_base + (uint32_t)((j & 1) * 64);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_bf16_ss(tS, umma_desc_sw128(q_addr + k * 32), umma_desc_sw128(k_addr + k * 32), idescS,
                                 (uint32_t)(k > 0));
                if (Cfg::TAIL)
                    umma_bf16_ss(tS, umma_desc(q_addr + Cfg::Q_P0, 0, 256, 6), umma_desc(k_addr + Cfg::KV_P0, 0, 256, 6),
                                 idescS, 1u);
                umma_commit(&s_full[j & 1]);
            };
            mbar_wait(q_full, 0);
            issue_s(0);
            if (nblk > 1) issue_s(1);
            for (int j = 0; j < nblk; ++j) {
                const int st = j % STAGES;
                const int valid = min(SPL_BKV, T - j * SPL_BKV);
                const int ksteps = (valid + 15) >> 4;
                const uint32_t v_addr = smem_u32(sV + st * Cfg::KV_BYTES);
                const uint32_t tP = tmem_base + (uint32_t)((j & 1) * 64);
                mbar_wait(&p_full[j & 1], (uint32_t)(j >> 1) & 1u);
                tcgen05_fence_after();
                for (int kk = 0; kk < ksteps; ++kk) {
                    const uint32_t acc = (uint32_t)((j | kk) != 0);
                    umma_bf16_ts(tO, tP + (uint32_t)(kk * 8), umma_desc(v_addr + kk * 2048, 0, 1024, 2), idescV64, acc);
                    if (Cfg::TAIL)
                        umma_bf16_ts(tO + 64, tP + (uint32_t)(kk * 8),
                                     umma_desc(v_addr + Cfg::KV_P0 + kk * 512, 0, 256, 6), idescV16, acc);
                }
                umma_commit(&kv_empty[st]);
                umma_commit(pv_done);
                if (j == nblk - 1) umma_commit(o_done);
                // S(j+2) reuses the buffer of P(j): queued behind PV(j), in-order execution protects it
                if (j + 2 < nblk) issue_s(j + 2);
            }
        }
    } else {
        // ===== softmax: thread = (row, key half) =====
        const int q = warp & 3;       // TMEM lane quadrant
        const int half = warp >> 2;   // keys [32 half, 32 half + 32) of every block
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        const uint32_t tO = tmem_base + Cfg::O_COL + lane_off;
        const int r_in = q * 32 + lane;  // row inside the tile
        const int row = q0 + r_in;
        const bool rows_live = q0 + q * 32 < T;  // warp-uniform
        float m_used = -INFINITY, l = 0.f;
        auto pair_sync = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(1 + q), "n"(64) : "memory"); };
        for (int j = 0; j < nblk; ++j) {
            const int buf = j & 1;
            const uint32_t tS = tmem_base + (uint32_t)(buf * 64) + lane_off;
            mbar_wait(&s_full[buf], (uint32_t)(j >> 1) & 1u);
            tcgen05_fence_after();
            if (rows_live) {
                uint32_t s[32];
                tmem_ld_32x32(tS + (uint32_t)(half * 32), s);
                tmem_ld_wait();
                const int valid = min(SPL_BKV, T - j * SPL_BKV) - half * 32;  // keys of this half that exist (may be <= 0)
                if (valid < 32) {
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (i >= valid) s[i] = 0xff800000u;  // -inf
                }
                float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    mx0 = fmaxf(mx0, fmaxf(__uint_as_float(s[4 * i]), __uint_as_float(s[4 * i + 1])));
                    mx1 = fmaxf(mx1, fmaxf(__uint_as_float(s[4 * i + 2]), __uint_as_float(s[4 * i + 3])));
                }
                const float mine = fmaxf(mx0, mx1);
                // exchange the two half maxima (slot j & 1: the partner may still be reading the other slot)
                float* xs = sX + (buf * 128 + r_in) * 2;
                xs[half] = mine;
                pair_sync();
                const float mt = fmaxf(mine, xs[half ^ 1]) * scale_log2;
                float factor = 1.0f;
                bool need = false;
                if (j == 0) {
                    m_used = mt;
                } else if (mt > m_used + SPL_RESCALE_THRESHOLD) {  // same inputs, same decision in both threads
                    need = true;
                    factor = spl_ex2(m_used - mt);
                    m_used = mt;
                    l *= factor;
                }
                const float nm = -m_used;
                float rs0 = 0.f, rs1 = 0.f;
                if (valid > 0) {  // warp-uniform: a half without a single key is never read by PV
                    uint32_t pk[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const float p0 = spl_ex2(fmaf(__uint_as_float(s[2 * i]), scale_log2, nm));
                        const float p1 = spl_ex2(fmaf(__uint_as_float(s[2 * i + 1]), scale_log2, nm));
                        rs0 += p0;
                        rs1 += p1;
                        pk[i] = pack_bf16x2(p0, p1);
                    }
                    tmem_st_32x16(tS + (uint32_t)(half * 16), pk);
                }
                l += rs0 + rs1;
                if (__any_sync(0xffffffffu, need)) {
                    // rare: rescale this thread's share of the O row once PV(j-1) has completed
                    // (S(j) complete implies PV(j-2) complete: parity unambiguous)
                    mbar_wait(pv_done, (uint32_t)(j - 1) & 1u);
                    tcgen05_fence_after();
                    uint32_t o[16];
                    const int c_lo = half == 0 ? 0 : 2, c_hi = half == 0 ? 2 : Cfg::DPAD / 16;
#pragma unroll 1
                    for (int c = c_lo; c < c_hi; ++c) {
                        tmem_ld_32x16(tO + (uint32_t)(c * 16), o);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * factor);
                        tmem_st_32x16(tO + (uint32_t)(c * 16), o);
                    }
                }
                tmem_st_wait();
            }
            // (a warp whose rows all lie beyond the sequence only keeps the barrier protocol going: its P rows
            // are garbage, they feed O rows that are never stored)
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_full[buf]);
        }
        // ---- finalise: O / l -> bf16 -> global; columns 0-31 by half 0, 32-.. by half 1 ----
        if (rows_live) {
            float* xs = sX + r_in * 2;  // slot 0 again; its last use (block nblk-1 or nblk-2) is behind a pair_sync
            pair_sync();                // both threads are done with the exchange slots
            xs[half] = l;
            pair_sync();
            l += xs[half ^ 1];
        }
        mbar_wait(o_done, 0);
        tcgen05_fence_after();
        if (rows_live) {
            const float inv = 1.0f / l;
            const int D = H * HD;
            __nv_bfloat16* orow = out + ((size_t)b * T + row) * D + (size_t)h * HD;
            {
                uint32_t o[32];
                tmem_ld_32x32(tO + (uint32_t)(half * 32), o);
                tmem_ld_wait();
                if (row < T) {
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        uint4 v;
                        v.x = pack_bf16x2(__uint_as_float(o[g * 8 + 0]) * inv, __uint_as_float(o[g * 8 + 1]) * inv);
                        v.y = pack_bf16x2(__uint_as_float(o[g * 8 + 2]) * inv, __uint_as_float(o[g * 8 + 3]) * inv);
                        v.z = pack_bf16x2(__uint_as_float(o[g * 8 + 4]) * inv, __uint_as_float(o[g * 8 + 5]) * inv);
                        v.w = pack_bf16x2(__uint_as_float(o[g * 8 + 6]) * inv, __uint_as_float(o[g * 8 + 7]) * inv);
                        *reinterpret_cast<uint4*>(orow + half * 32 + g * 8) = v;
                    }
                }
            }
            if (Cfg::TAIL && half == 1) {
                uint32_t o2[16];
                tmem_ld_32x16(tO + 64, o2);
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
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == W_MMA) {
        tcgen05_fence_after();
        tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
    }
}

template <int HD>
int launch_attention_split(const void* qkv, void* out, int B, int T, int H, float scale, cudaStream_t s) {
    using Cfg = SplCfg<HD>;
    // qkv viewed as [B][T][3H][HD], innermost first; Q boxes hold 128 rows, K/V boxes 64
    const uint64_t dims[4] = {(uint64_t)HD, (uint64_t)3 * H, (uint64_t)T, (uint64_t)B};
    const uint64_t strides[3] = {(uint64_t)HD * 2, (uint64_t)3 * H * HD * 2, (uint64_t)T * 3 * H * HD * 2};
    const uint32_t bq64[4] = {64, 1, SPL_BQ, 1}, bq16[4] = {16, 1, SPL_BQ, 1};
    const uint32_t bk64[4] = {64, 1, SPL_BKV, 1}, bk16[4] = {16, 1, SPL_BKV, 1};
    CUtensorMap tq64, tq16, tk64, tk16;
    int rc = make_tmap_nd_bf16(&tq64, qkv, 4, dims, strides, bq64, 128);
    if (rc) return rc;
    rc = make_tmap_nd_bf16(&tk64, qkv, 4, dims, strides, bk64, 128);
    if (rc) return rc;
    rc = make_tmap_nd_bf16(&tq16, qkv, 4, dims, strides, Cfg::TAIL ? bq16 : bq64, Cfg::TAIL ? 32 : 128);
    if (rc) return rc;
    rc = make_tmap_nd_bf16(&tk16, qkv, 4, dims, strides, Cfg::TAIL ? bk16 : bk64, Cfg::TAIL ? 32 : 128);
    if (rc) return rc;
    GVL_CUDA(cudaFuncSetAttribute(attention_split_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    dim3 grid((T + SPL_BQ - 1) / SPL_BQ, H, B);
    ProfScope prof(GVL_K_ATTENTION, 4.0 * B * (double)H * T * (double)T * HD, s);
    attention_split_kernel<HD><<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, s>>>(
        tq64, tq16, tk64, tk16, reinterpret_cast<__nv_bfloat16*>(out), T, H, scale * 1.4426950408889634f);
    GVL_LAUNCH_CHECK("attention_split_kernel");
    return 0;
}

template int launch_attention_split<72>(const void*, void*, int, int, int, float, cudaStream_t);
template int launch_attention_split<64>(const void*, void*, int, int, int, float, cudaStream_t);

}  // namespace gvl
