// attention_sdb.cu — K4, production variant: fused non-causal self-attention on tcgen05, two query tiles per CTA,
// each with its score tile DOUBLE-BUFFERED in tensor memory.
//
// For head dims this small the kernel is bound by the MUFU (exp2) pipe, not by the tensor cores: the B200 SM
// retires 16 exp2 per clock, and a softmax-like instruction mix only gets there with two independent warps per
// scheduler (tools/scratch/mufu_bench.cu: 9.5 exp/clk/SM with one warp per scheduler, 14.9 with two).  So the
// design goal is that the softmax warps never wait and never run in lockstep:
//   - S(j+1) is computed into the other buffer while softmax(j) runs; PV(j) is issued as soon as P(j) is written
//     and S(j+2) is queued right behind it (tcgen05.mma executes in issue order, which protects the P(j) columns);
//   - a CTA owns 256 query rows of one (image, head) as two independent 128-row tiles, so each scheduler holds
//     two softmax warps (one per tile) whose load / max / store phases overlap the other's exp phase;
//   - keys are processed in blocks of 64 so that 2 tiles x 2 buffers x 64 score columns + 2 x 80 output columns
//     fit the 512 TMEM columns; both tiles share every K/V stage.
// The single-buffer kernel (attention_tc.cu, S -> softmax -> PV -> S serialised because P overwrites S) is kept
// for A/B runs (GVL_ATTN_TC1=1).
//
//   warps 0-3   softmax of tile 0, warps 4-7 softmax of tile 1: one thread = one query row x 64 keys; tcgen05.ld
//               the S row, running max with lazy rescaling of O (only when the max grows by more than 2^8),
//               p = exp2(s*c - m), fp32 row sum, P packed to bf16 pairs and written over the first 32 columns of
//               the S buffer (the thread's own row, already in registers) with tcgen05.st.
//   warp 8      TMA producer: Q tiles once, then K/V blocks of 64 keys through a 6-stage ring.  Head dim 72 is
//               fetched as a 64-wide SWIZZLE_128B panel plus a 16-wide SWIZZLE_32B panel whose upper 8 columns are
//               out of bounds (zero) — the k-padding 72 -> 80 costs no memory.
//   warps 9,10  MMA issuers, one per tile (a single issuer would make each tile wait for the other's P):
//               S = Q K^T (SS, both K-major), O += P V with P read from TMEM (TS form) and V as an MN-major
//               shared-memory operand (64-wide and 16-wide N panels).
//
// TMEM columns: tile t, buffer u: S/P at [128 t + 64 u, +64); O_t at [256 + 80 t, +80).
#include "common.cuh"

#include <cstdlib>

namespace gvl {

constexpr int SDB_BQ = 128;            // rows per query tile (two tiles per CTA)
constexpr int SDB_BKV = 64;            // keys per block
// NT = query tiles per CTA.  NT = 2: one CTA per SM, 512 TMEM columns, K/V stages shared by both tiles.
// NT = 1: two CTAs per SM (256 TMEM columns each), which also overlaps one CTA's prologue / epilogue (Q and first
// K/V loads, TMEM allocation, final O read-out) with the other CTA's steady state.
constexpr float SDB_RESCALE_THRESHOLD = 8.0f;  // log2 units

template <int HD, int NT>
struct SdbCfg {
    static constexpr int THREADS = (4 * NT + 1 + NT) * 32;  // 4 softmax warps per tile, TMA warp, one MMA warp per tile
    static constexpr int TMEM_COLS = NT == 2 ? 512 : 256;
    static constexpr int STAGES = NT == 2 ? 6 : 4;
    static constexpr int O_COL = NT * 128;                   // first output-accumulator column
    static constexpr bool TAIL = HD > 64;  // second, 16-wide panel for d in [64, 80)
    static constexpr int DPAD = TAIL ? 80 : 64;
    static constexpr int Q_P0 = 128 * 128;             // 128 rows x 64 bf16, SWIZZLE_128B
    static constexpr int Q_P1 = TAIL ? 128 * 32 : 0;   // 128 rows x 16 bf16, SWIZZLE_32B
    static constexpr int Q_BYTES = Q_P0 + Q_P1;
    static constexpr int KV_P0 = SDB_BKV * 128;
    static constexpr int KV_P1 = TAIL ? SDB_BKV * 32 : 0;
    static constexpr int KV_BYTES = KV_P0 + KV_P1;     // one K or V block
    static constexpr int SMEM_BYTES = NT * Q_BYTES + 2 * STAGES * KV_BYTES + 256 /*barriers*/ + 1024 /*alignment*/;
};

// Two non-negative fp32 probabilities -> packed bf16x2, rounded to nearest (ties up) with integer adds + one byte
// permute.  F2FP.BF16.PACK_AB executes on the XU pipe — the same quarter-rate pipe as MUFU.EX2, which ncu shows is
// this kernel's busiest unit (60 %) — so packing on the ALU pipe takes a third of the XU work away.
__device__ __forceinline__ uint32_t sdb_pack_bf16x2(float lo, float hi) {
    return __byte_perm(__float_as_uint(lo) + 0x8000u, __float_as_uint(hi) + 0x8000u, 0x7632);
}
__device__ __forceinline__ float sdb_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// Timeline instrumentation (tools/attn_trace.py): when `trace` is set, the CTA (1, 3, 17) records SM-clock stamps per
// key block: softmax warp 0 / lane 0 -> slots 0-5, the MMA thread -> slots 8-11, the TMA thread -> slot 12.
#define SDB_TRACE(jj, slot) \
    do { if (trace != nullptr && traced) trace[(jj) * 16 + (slot)] = clock64(); } while (0)

template <int HD, int NT>
__global__ void __launch_bounds__(SdbCfg<HD, NT>::THREADS, NT == 2 ? 1 : 2)
attention_sdb_kernel(const __grid_constant__ CUtensorMap tmq64, const __grid_constant__ CUtensorMap tmq16,
                     const __grid_constant__ CUtensorMap tmk64, const __grid_constant__ CUtensorMap tmk16,
                     __nv_bfloat16* __restrict__ out, int T, int H, float scale_log2, int dbg, long long* __restrict__ trace) {
    using Cfg = SdbCfg<HD, NT>;
    constexpr int SDB_STAGES = Cfg::STAGES;
    constexpr int W_TMA = 4 * NT, W_MMA = 4 * NT + 1;  // warp roles: [0, 4 NT) softmax, then TMA, then NT MMA warps
    extern __shared__ uint8_t sdb_smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(sdb_smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;                                  // [tile][P0 | P1]
    uint8_t* sK = smem + NT * Cfg::Q_BYTES;              // [stage][P0 | P1]
    uint8_t* sV = sK + SDB_STAGES * Cfg::KV_BYTES;       // [stage][P0 | P1]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sV + SDB_STAGES * Cfg::KV_BYTES);
    uint64_t* q_full = bars;                      // [1]
    uint64_t* kv_full = bars + 1;                 // [STAGES]
    uint64_t* kv_empty = kv_full + SDB_STAGES;    // [STAGES]
    uint64_t* s_full = kv_empty + SDB_STAGES;     // [tile][buffer]  S_t(j) complete in buffer j & 1
    uint64_t* p_full = s_full + 4;                // [tile][buffer]  P_t(j) written (4 warps)
    uint64_t* pv_done = p_full + 4;               // [tile]  PV_t(j) complete, one phase per block (rare rescale path
                                                  //         only: a parity wait is valid at most one phase behind)
    uint64_t* o_done = pv_done + 2;               // [tile]  last PV_t complete
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(o_done + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * (NT * SDB_BQ);
    const int h = blockIdx.y, b = blockIdx.z;
    const int nblk = (T + SDB_BKV - 1) / SDB_BKV;
    const bool traced = blockIdx.x == 1 && blockIdx.y == 3 && blockIdx.z == 17 && (threadIdx.x & 31) == 0;
    const int ntile = (NT == 2 && q0 + SDB_BQ < T) ? 2 : 1;  // the second tile may lie entirely beyond the sequence

    if (warp == W_TMA && lane == 0) {
        tma_prefetch_desc(&tmq64);
        tma_prefetch_desc(&tmk64);
        if (Cfg::TAIL) {
            tma_prefetch_desc(&tmq16);
            tma_prefetch_desc(&tmk16);
        }
        mbar_init(q_full, 1);
        for (int s = 0; s < SDB_STAGES; ++s) {
            mbar_init(&kv_full[s], 1);
            mbar_init(&kv_empty[s], ntile);  // one tcgen05.commit per tile's MMA warp
        }
        for (int i = 0; i < 4; ++i) {
            mbar_init(&s_full[i], 1);
            mbar_init(&p_full[i], 4);
        }
        for (int t = 0; t < 2; ++t) {
            mbar_init(&pv_done[t], 1);
            mbar_init(&o_done[t], 1);
        }
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
            mbar_arrive_expect_tx(q_full, ntile * Cfg::Q_BYTES);
            for (int t = 0; t < ntile; ++t) {
                uint8_t* q = sQ + t * Cfg::Q_BYTES;
                tma_load_4d(q, &tmq64, q_full, 0, h, q0 + t * SDB_BQ, b);
                if (Cfg::TAIL) tma_load_4d(q + Cfg::Q_P0, &tmq16, q_full, 64, h, q0 + t * SDB_BQ, b);
            }
            int st = 0;
            uint32_t ph = 0;
            for (int j = 0; j < nblk; ++j) {
                mbar_wait(&kv_empty[st], ph ^ 1);
                SDB_TRACE(j, 12);
                mbar_arrive_expect_tx(&kv_full[st], 2 * Cfg::KV_BYTES);
                uint8_t* k = sK + st * Cfg::KV_BYTES;
                uint8_t* v = sV + st * Cfg::KV_BYTES;
                tma_load_4d(k, &tmk64, &kv_full[st], 0, H + h, j * SDB_BKV, b);
                tma_load_4d(v, &tmk64, &kv_full[st], 0, 2 * H + h, j * SDB_BKV, b);
                if (Cfg::TAIL) {
                    tma_load_4d(k + Cfg::KV_P0, &tmk16, &kv_full[st], 64, H + h, j * SDB_BKV, b);
                    tma_load_4d(v + Cfg::KV_P0, &tmk16, &kv_full[st], 64, 2 * H + h, j * SDB_BKV, b);
                }
                if (++st == SDB_STAGES) {
                    st = 0;
                    ph ^= 1;
                }
            }
        }
    } else if (warp >= W_MMA) {
        // ===== MMA issuer of tile (warp - W_MMA) =====
        const int t = warp - W_MMA;
        if (t < ntile && elect_one()) {
            constexpr uint32_t idescS = umma_idesc_bf16_major(128, SDB_BKV, 0, 0);  // Q, K both K-major
            constexpr uint32_t idescV64 = umma_idesc_bf16_major(128, 64, 0, 1);     // V: MN-major B
            constexpr uint32_t idescV16 = umma_idesc_bf16_major(128, 16, 0, 1);
            const uint32_t q_addr = smem_u32(sQ + t * Cfg::Q_BYTES);
            const uint32_t tO = tmem_base + Cfg::O_COL + (uint32_t)(t * Cfg::DPAD);
            auto issue_s = [&](int j) {  // S_t(j) -> buffer j & 1; K(j) sits in stage j % STAGES
                const int st = j % SDB_STAGES;
                mbar_wait(&kv_full[st], (uint32_t)(j / SDB_STAGES) & 1u);
                tcgen05_fence_after();
                const uint32_t k_addr = smem_u32(sK + st * Cfg::KV_BYTES);
                const uint32_t tS = tmem_base + (uint32_t)(t * 128 + (j & 1) * 64);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_bf16_ss(tS, umma_desc_sw128(q_addr + k * 32), umma_desc_sw128(k_addr + k * 32), idescS,
                                 (uint32_t)(k > 0));
                if (Cfg::TAIL && !(dbg & 2))
                    umma_bf16_ss(tS, umma_desc(q_addr + Cfg::Q_P0, 0, 256, 6), umma_desc(k_addr + Cfg::KV_P0, 0, 256, 6),
                                 idescS, 1u);
                umma_commit(&s_full[t * 2 + (j & 1)]);
            };
            mbar_wait(q_full, 0);
            issue_s(0);
            if (nblk > 1) issue_s(1);
            for (int j = 0; j < nblk; ++j) {
                const int st = j % SDB_STAGES;
                const int valid = min(SDB_BKV, T - j * SDB_BKV);
                const int ksteps = (valid + 15) >> 4;
                const uint32_t v_addr = smem_u32(sV + st * Cfg::KV_BYTES);
                // O_t += P_t(j) V(j): k runs over the keys of this block, 16 per MMA; P is read from TMEM
                const uint32_t tP = tmem_base + (uint32_t)(t * 128 + (j & 1) * 64);
                SDB_TRACE(j, 8);
                mbar_wait(&p_full[t * 2 + (j & 1)], (uint32_t)(j >> 1) & 1u);
                SDB_TRACE(j, 9);
                tcgen05_fence_after();
                for (int kk = 0; kk < ksteps; ++kk) {
                    const uint32_t acc = (uint32_t)((j | kk) != 0);
                    umma_bf16_ts(tO, tP + (uint32_t)(kk * 8), umma_desc(v_addr + kk * 2048, 0, 1024, 2), idescV64, acc);
                    if (Cfg::TAIL && !(dbg & 1))
                        umma_bf16_ts(tO + 64, tP + (uint32_t)(kk * 8),
                                     umma_desc(v_addr + Cfg::KV_P0 + kk * 512, 0, 256, 6), idescV16, acc);
                }
                umma_commit(&kv_empty[st]);  // this tile is done with K(j) (read by S_t(j), issued earlier) and V(j)
                umma_commit(&pv_done[t]);
                if (j == nblk - 1) umma_commit(&o_done[t]);
                // S_t(j+2) reuses the buffer of P_t(j): queued behind PV_t(j), in-order execution protects it
                SDB_TRACE(j, 10);
                if (j + 2 < nblk) issue_s(j + 2);
                SDB_TRACE(j, 11);
            }
        }
    } else {
        // ===== softmax: one thread = one query row x 64 keys per block; warps 0-3 -> tile 0, warps 4-7 -> tile 1 =====
        const int t = warp >> 2;
        if (t < ntile) {
            const int q = warp & 3;  // TMEM lane quadrant this warp may access
            const uint32_t lane_off = (uint32_t)(q * 32) << 16;
            const uint32_t tO = tmem_base + Cfg::O_COL + (uint32_t)(t * Cfg::DPAD) + lane_off;
            const int row = q0 + t * SDB_BQ + q * 32 + lane;
            float m_used = -INFINITY, l = 0.f;
            const bool rows_live = q0 + t * SDB_BQ + q * 32 < T;  // warp-uniform: at least one of the 32 rows exists
            auto load_s = [&](int j, uint32_t (&s)[2][32]) {     // wait for S_t(j), start its TMEM -> register loads
                const uint32_t tS = tmem_base + (uint32_t)(t * 128 + (j & 1) * 64) + lane_off;
                if (warp == 0) SDB_TRACE(j, 0);
                mbar_wait(&s_full[t * 2 + (j & 1)], (uint32_t)(j >> 1) & 1u);
                if (warp == 0) SDB_TRACE(j, 1);
                tcgen05_fence_after();
#pragma unroll
                for (int c = 0; c < 2; ++c) tmem_ld_32x32(tS + (uint32_t)(c * 32), s[c]);
            };
            // (prefetching S_t(j+1) into a second register buffer while block j is exponentiated was tried: 168
            // registers do not hold both rows, ptxas spills one and the kernel gets 2x slower)
            for (int j = 0; j < nblk; ++j) {
                const int buf = j & 1;
                const uint32_t tS = tmem_base + (uint32_t)(t * 128 + buf * 64) + lane_off;
                uint32_t s[2][32];
                load_s(j, s);
                tmem_ld_wait();
                if (warp == 0) SDB_TRACE(j, 2);
                if (rows_live) {
                    const int valid = min(SDB_BKV, T - j * SDB_BKV);
                    if (valid < SDB_BKV) {
#pragma unroll
                        for (int c = 0; c < 2; ++c)
#pragma unroll
                            for (int i = 0; i < 32; ++i)
                                if (c * 32 + i >= valid) s[c][i] = 0xff800000u;  // -inf
                    }
                    // four independent max chains (one chain of 32 dependent FMNMX3 is ~130 cycles of pure latency)
                    float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        mx0 = fmaxf(mx0, fmaxf(__uint_as_float(s[0][4 * i]), __uint_as_float(s[0][4 * i + 1])));
                        mx1 = fmaxf(mx1, fmaxf(__uint_as_float(s[0][4 * i + 2]), __uint_as_float(s[0][4 * i + 3])));
                        mx2 = fmaxf(mx2, fmaxf(__uint_as_float(s[1][4 * i]), __uint_as_float(s[1][4 * i + 1])));
                        mx3 = fmaxf(mx3, fmaxf(__uint_as_float(s[1][4 * i + 2]), __uint_as_float(s[1][4 * i + 3])));
                    }
                    const float mt = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * scale_log2;
                    float factor = 1.0f;
                    bool need = false;
                    if (j == 0) {
                        m_used = mt;
                    } else if (mt > m_used + SDB_RESCALE_THRESHOLD) {
                        need = true;
                        factor = sdb_ex2(m_used - mt);
                        m_used = mt;
                        l *= factor;
                    }
                    const float nm = -m_used;
                    float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        if (c * 32 < valid) {  // warp-uniform: chunks without a single key are never read by PV
                            uint32_t pk[16];
#pragma unroll
                            for (int i = 0; i < 16; ++i) {
                                const float p0 = sdb_ex2(fmaf(__uint_as_float(s[c][2 * i]), scale_log2, nm));
                                const float p1 = sdb_ex2(fmaf(__uint_as_float(s[c][2 * i + 1]), scale_log2, nm));
                                rs0 += p0;
                                rs1 += p1;
                                pk[i] = sdb_pack_bf16x2(p0, p1);
                            }
                            tmem_st_32x16(tS + (uint32_t)(c * 16), pk);
                        }
                    }
                    l += rs0 + rs1;
                    if (warp == 0) SDB_TRACE(j, 3);
                    if (__any_sync(0xffffffffu, need)) {
                        // rare: the running max grew by more than the threshold -> rescale this warp's O rows once
                        // PV_t(j-1) has completed (S_t(j) complete implies PV_t(j-2) complete: parity unambiguous)
                        mbar_wait(&pv_done[t], (uint32_t)(j - 1) & 1u);
                        tcgen05_fence_after();
                        uint32_t o[32];
#pragma unroll
                        for (int c = 0; c < 2; ++c) {
                            tmem_ld_32x32(tO + (uint32_t)(c * 32), o);
                            tmem_ld_wait();
#pragma unroll
                            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * factor);
                            tmem_st_32x32(tO + (uint32_t)(c * 32), o);
                        }
                        if (Cfg::TAIL) {
                            uint32_t o2[16];
                            tmem_ld_32x16(tO + 64, o2);
                            tmem_ld_wait();
#pragma unroll
                            for (int i = 0; i < 16; ++i) o2[i] = __float_as_uint(__uint_as_float(o2[i]) * factor);
                            tmem_st_32x16(tO + 64, o2);
                        }
                    }
                    tmem_st_wait();
                    if (warp == 0) SDB_TRACE(j, 4);
                }
                // (a warp whose rows all lie beyond the sequence only keeps the barrier protocol going: its P rows
                // are garbage, they feed O rows that are never stored)
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&p_full[t * 2 + buf]);
                if (warp == 0) SDB_TRACE(j, 5);
            }
            // ---- finalise: O / l -> bf16 -> global ----
            mbar_wait(&o_done[t], 0);
            tcgen05_fence_after();
            const float inv = 1.0f / l;
            const int D = H * HD;
            __nv_bfloat16* orow = out + ((size_t)b * T + row) * D + (size_t)h * HD;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                uint32_t o[32];
                tmem_ld_32x32(tO + (uint32_t)(c * 32), o);
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

long long* g_attn_trace = nullptr;  // set by gvl_debug_set_attn_trace (tuning aid)

template <int HD, int NT>
static int launch_attention_sdb_nt(const void* qkv, void* out, int B, int T, int H, float scale, cudaStream_t s) {
    using Cfg = SdbCfg<HD, NT>;
    // qkv viewed as [B][T][3H][HD], innermost first; Q boxes hold 128 rows, K/V boxes 64
    const uint64_t dims[4] = {(uint64_t)HD, (uint64_t)3 * H, (uint64_t)T, (uint64_t)B};
    const uint64_t strides[3] = {(uint64_t)HD * 2, (uint64_t)3 * H * HD * 2, (uint64_t)T * 3 * H * HD * 2};
    const uint32_t bq64[4] = {64, 1, SDB_BQ, 1}, bq16[4] = {16, 1, SDB_BQ, 1};
    const uint32_t bk64[4] = {64, 1, SDB_BKV, 1}, bk16[4] = {16, 1, SDB_BKV, 1};
    CUtensorMap tq64, tq16, tk64, tk16;
    int rc = make_tmap_nd_bf16(&tq64, qkv, 4, dims, strides, bq64, 128);
    if (rc) return rc;
    rc = make_tmap_nd_bf16(&tk64, qkv, 4, dims, strides, bk64, 128);
    if (rc) return rc;
    rc = make_tmap_nd_bf16(&tq16, qkv, 4, dims, strides, Cfg::TAIL ? bq16 : bq64, Cfg::TAIL ? 32 : 128);
    if (rc) return rc;
    rc = make_tmap_nd_bf16(&tk16, qkv, 4, dims, strides, Cfg::TAIL ? bk16 : bk64, Cfg::TAIL ? 32 : 128);
    if (rc) return rc;
    GVL_CUDA(cudaFuncSetAttribute(attention_sdb_kernel<HD, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  Cfg::SMEM_BYTES));
    static const int dbg = [] { const char* e = getenv("GVL_ATTN_DEBUG"); return e ? atoi(e) : 0; }();  // timing experiments
    dim3 grid((T + NT * SDB_BQ - 1) / (NT * SDB_BQ), H, B);
    ProfScope prof(GVL_K_ATTENTION, 4.0 * B * (double)H * T * (double)T * HD, s);
    attention_sdb_kernel<HD, NT><<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, s>>>(
        tq64, tq16, tk64, tk16, reinterpret_cast<__nv_bfloat16*>(out), T, H, scale * 1.4426950408889634f, dbg, g_attn_trace);
    GVL_LAUNCH_CHECK("attention_sdb_kernel");
    return 0;
}

template <int HD>
int launch_attention_sdb(const void* qkv, void* out, int B, int T, int H, float scale, cudaStream_t s) {
    static const int nt = [] {
        const char* e = getenv("GVL_ATTN_NT");  // tuning switch: query tiles per CTA
        return (e && e[0] == '2') ? 2 : 1;
    }();
    return nt == 2 ? launch_attention_sdb_nt<HD, 2>(qkv, out, B, T, H, scale, s)
                   : launch_attention_sdb_nt<HD, 1>(qkv, out, B, T, H, scale, s);
}

template int launch_attention_sdb<72>(const void*, void*, int, int, int, float, cudaStream_t);
template int launch_attention_sdb<64>(const void*, void*, int, int, int, float, cudaStream_t);

}  // namespace gvl

// Tuning aid, not part of the product ABI surface in include/gvl.h: device buffer of >= 16 x 16 int64 that receives
// the timeline stamps of one CTA (see SDB_TRACE); NULL turns tracing off.
extern "C" __attribute__((visibility("default"))) void gvl_debug_set_attn_trace(void* device_buffer) {
    gvl::g_attn_trace = reinterpret_cast<long long*>(device_buffer);
}
