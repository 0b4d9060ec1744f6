// attention_sdb.cu — K4, production variant: fused non-causal self-attention on tcgen05 with the score tile
// DOUBLE-BUFFERED in tensor memory.
//
// Production configuration (NT = 1): one CTA = one 128-row query tile of one (image, head), 6 warps, two CTAs per SM
// (2 x (2 x 64 S + 80 O + 16 L) TMEM columns).  NT = 2 (two tiles per CTA, one CTA per SM, K/V stages shared) and
// NT = 3 (three tiles, S single-buffered) exist only in -DGVL_EXPERIMENTS builds (GVL_ATTN_NT); both measure slower.
//   - S(j+1) is computed into the other buffer while softmax(j) runs; PV(j) is issued as soon as P(j) is written
//     and S(j+2) is queued right behind it (tcgen05.mma executes in issue order, which protects the P(j) columns);
//   - keys are processed in blocks of 64;
//   - the softmax inner loop is FFMA2 (two scores per instruction), MUFU.EX2 and F2FP only: the row sums are
//     accumulated by the tensor cores (L += P . 1 with a constant tile of ones), so l is the sum of exactly the bf16
//     P values the PV MMAs consume.
// What bounds it (tools/attn_trace.py): a softmax warp alone issues one MUFU every ~13.5 cycles,
// two per scheduler ~8.6; with ~400 cycles of fixed per-block hand-shake (mbarrier round trip, tcgen05.ld / st waits)
// and a 6400-cycle CTA start-up the exp pipe is ~60 % busy.  Neither fewer softmax instructions, nor fewer MMAs, nor
// more softmax warps per scheduler (NT = 3, split rows) moved the 282 us per layer.
//
//   warps [0, 4 NT)  softmax, four warps per tile: one thread = one query row x 64 keys; tcgen05.ld the S row,
//               running max with lazy rescaling of O and L (only when the max grows by more than 2^8),
//               p = exp2(s*c - m), P packed to bf16 pairs and written over the first 32 columns of the S buffer
//               (the thread's own row, already in registers) with tcgen05.st.
//   next warp   TMA producer: Q tiles once, then K/V blocks of 64 keys through a ring.  Head dim 72 is fetched as a
//               64-wide SWIZZLE_128B panel plus a 16-wide SWIZZLE_32B panel whose upper 8 columns are out of bounds
//               (zero) — the k-padding 72 -> 80 costs no memory.
//   NT warps    MMA issuers, one per tile: S = Q K^T (SS, both K-major), O += P V and L += P 1 with P read from TMEM
//               (TS form) and V as an MN-major shared-memory operand (64-wide and 16-wide N panels).
//
// TMEM columns: tile t, buffer u: S/P at [S_COLS t + 64 u, +64); O_t at [O_COL + DPAD t, +DPAD); L_t at [L_COL + 16 t, +16).
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
    static constexpr int NBUF = NT == 3 ? 1 : 2;             // S buffers per tile (NT = 3: single, P overwrites S)
    static constexpr int THREADS = (4 * NT + 1 + NT) * 32;  // 4 softmax warps per tile, TMA warp, one MMA warp per tile
    static constexpr int TMEM_COLS = NT >= 2 ? 512 : 256;
    static constexpr int STAGES = NT == 2 ? 6 : 4;
    static constexpr int S_COLS = NBUF * 64;                 // score columns per tile
    static constexpr int O_COL = NT * S_COLS;                // first output-accumulator column
    static constexpr int L_COL = O_COL + NT * (HD > 64 ? 80 : 64);  // row-sum accumulators, 16 columns per tile
    static constexpr bool TAIL = HD > 64;  // second, 16-wide panel for d in [64, 80)
    static constexpr int DPAD = TAIL ? 80 : 64;
    static constexpr int Q_P0 = 128 * 128;             // 128 rows x 64 bf16, SWIZZLE_128B
    static constexpr int Q_P1 = TAIL ? 128 * 32 : 0;   // 128 rows x 16 bf16, SWIZZLE_32B
    static constexpr int Q_BYTES = Q_P0 + Q_P1;
    static constexpr int KV_P0 = SDB_BKV * 128;
    static constexpr int KV_P1 = TAIL ? SDB_BKV * 32 : 0;
    static constexpr int KV_BYTES = KV_P0 + KV_P1;     // one K or V block
    static constexpr int ONES_BYTES = 512;             // 16 keys x 16 columns of bf16 1.0
    static constexpr int SMEM_BYTES = NT * Q_BYTES + 2 * STAGES * KV_BYTES + ONES_BYTES + 256 /*barriers*/ + 1024 /*alignment*/;
};

// (s0, s1) * c + nm for two scores in one FFMA2 (fma.rn.f32x2, sm_100).  cc / nn = the constants packed twice.
__device__ __forceinline__ uint64_t sdb_pack2(float lo, float hi) {
    uint64_t d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi));
    return d;
}
__device__ __forceinline__ void sdb_scale2(uint32_t s0, uint32_t s1, uint64_t cc, uint64_t nn, float& x0, float& x1) {
    uint64_t a, d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "r"(s0), "r"(s1));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(cc), "l"(nn));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(x0), "=f"(x1) : "l"(d));
}
__device__ __forceinline__ float sdb_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// Timeline instrumentation (tools/attn_trace.py): when `trace` is set, the CTA (1, 3, 17) records SM-clock stamps per
// key block: softmax warp 0 / lane 0 -> slots 0-5, the MMA thread -> slots 8-11, the TMA thread -> slot 12.
// (compiled in only with -DGVL_ATTN_TRACE: the stamps cost 10 % of the kernel — its softmax warps are bound by their
// own instruction stream, every extra instruction per key block shows)
#ifdef GVL_ATTN_TRACE
#define SDB_TRACE(jj, slot) \
    do { if (trace != nullptr && traced) trace[(jj) * 16 + (slot)] = clock64(); } while (0)
#else
#define SDB_TRACE(jj, slot) do { } while (0)
#endif

// VARLEN (the ragged batch of the masked-region route, gvl_attention_varlen_bf16): items of different lengths are
// concatenated along the token axis (the tensor maps see ONE batch of `T` = all tokens); blockIdx.x indexes a table of
// query tiles {first token of the item, item length, first query row of the tile}.  A Q box that runs past its item's
// end reads the next item's rows (finite; those output rows are never stored), K / V boxes likewise (those keys are
// masked to -inf, so the next item's finite V rows are multiplied by exactly zero).
template <int HD, int NT, bool VARLEN = false>
__global__ void __launch_bounds__(SdbCfg<HD, NT>::THREADS, NT >= 2 ? 1 : 2)
attention_sdb_kernel(const __grid_constant__ CUtensorMap tmq64, const __grid_constant__ CUtensorMap tmq16,
                     const __grid_constant__ CUtensorMap tmk64, const __grid_constant__ CUtensorMap tmk16,
                     __nv_bfloat16* __restrict__ out, int T_arg, int H, float scale_log2, int dbg,
                     long long* __restrict__ trace, const int4* __restrict__ tiles) {
    using Cfg = SdbCfg<HD, NT>;
    constexpr int SDB_STAGES = Cfg::STAGES;
    constexpr int W_TMA = 4 * NT, W_MMA = 4 * NT + 1;  // warp roles: [0, 4 NT) softmax, then TMA, then NT MMA warps
    extern __shared__ uint8_t sdb_smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(sdb_smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;                                  // [tile][P0 | P1]
    uint8_t* sK = smem + NT * Cfg::Q_BYTES;              // [stage][P0 | P1]
    uint8_t* sV = sK + SDB_STAGES * Cfg::KV_BYTES;       // [stage][P0 | P1]
    uint8_t* sOnes = sV + SDB_STAGES * Cfg::KV_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sOnes + Cfg::ONES_BYTES);
    uint64_t* q_full = bars;                      // [1]
    uint64_t* kv_full = bars + 1;                 // [STAGES]
    uint64_t* kv_empty = kv_full + SDB_STAGES;    // [STAGES]
    uint64_t* s_full = kv_empty + SDB_STAGES;     // [tile][buffer]  S_t(j) complete in buffer j & 1
    uint64_t* p_full = s_full + 2 * NT;           // [tile][buffer]  P_t(j) written (4 warps)
    uint64_t* pv_done = p_full + 2 * NT;          // [tile]  PV_t(j) complete, one phase per block (rare rescale path
                                                  //         only: a parity wait is valid at most one phase behind)
    uint64_t* o_done = pv_done + NT;              // [tile]  last PV_t complete
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(o_done + NT);
    constexpr int NBUF = Cfg::NBUF;
    // buffer and mbarrier phase parity of key block j (double-buffered: two blocks per phase pair)
    auto sbuf = [](int j) { return NBUF == 2 ? (j & 1) : 0; };
    auto spar = [](int j) { return (uint32_t)(NBUF == 2 ? (j >> 1) : j) & 1u; };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    pdl_trigger();
    int q0 = blockIdx.x * (NT * SDB_BQ), T = T_arg, tok0 = 0;  // tok0: first token of the item on the token axis
    if (VARLEN) {
        const int4 e = tiles[blockIdx.x];
        tok0 = e.x;
        T = e.y;
        q0 = e.z;
    }
    const int h = blockIdx.y, b = blockIdx.z;
    const int nblk = (T + SDB_BKV - 1) / SDB_BKV;
    const bool traced = blockIdx.x == 1 && blockIdx.y == 3 && blockIdx.z == 17 && (threadIdx.x & 31) == 0;
    const int ntile = min(NT, (T - q0 + SDB_BQ - 1) / SDB_BQ);  // trailing tiles may lie entirely beyond the sequence

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
        for (int i = 0; i < 2 * NT; ++i) {
            mbar_init(&s_full[i], 1);
            mbar_init(&p_full[i], 4);
        }
        for (int t = 0; t < NT; ++t) {
            mbar_init(&pv_done[t], 1);
            mbar_init(&o_done[t], 1);
        }
        fence_barrier_init();
    }
    if (warp == 0) {
        for (int i = lane; i < Cfg::ONES_BYTES / 4; i += 32) reinterpret_cast<uint32_t*>(sOnes)[i] = 0x3F803F80u;
        fence_proxy_async_smem();  // read by tcgen05.mma (async proxy)
    }
    if (warp == W_MMA) tmem_alloc<Cfg::TMEM_COLS>(tmem_ptr_smem);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    pdl_wait();  // the QKV GEMM's output is visible from here on

    if (warp == W_TMA) {
        // ===== TMA producer =====
        if (elect_one()) {
            mbar_arrive_expect_tx(q_full, ntile * Cfg::Q_BYTES);
            for (int t = 0; t < ntile; ++t) {
                uint8_t* q = sQ + t * Cfg::Q_BYTES;
                tma_load_4d(q, &tmq64, q_full, 0, h, tok0 + q0 + t * SDB_BQ, b);
                if (Cfg::TAIL) tma_load_4d(q + Cfg::Q_P0, &tmq16, q_full, 64, h, tok0 + q0 + t * SDB_BQ, b);
            }
            int st = 0;
            uint32_t ph = 0;
            for (int j = 0; j < nblk; ++j) {
                mbar_wait(&kv_empty[st], ph ^ 1);
                SDB_TRACE(j, 12);
                mbar_arrive_expect_tx(&kv_full[st], 2 * Cfg::KV_BYTES);
                uint8_t* k = sK + st * Cfg::KV_BYTES;
                uint8_t* v = sV + st * Cfg::KV_BYTES;
                tma_load_4d(k, &tmk64, &kv_full[st], 0, H + h, tok0 + j * SDB_BKV, b);
                tma_load_4d(v, &tmk64, &kv_full[st], 0, 2 * H + h, tok0 + j * SDB_BKV, b);
                if (Cfg::TAIL) {
                    tma_load_4d(k + Cfg::KV_P0, &tmk16, &kv_full[st], 64, H + h, tok0 + j * SDB_BKV, b);
                    tma_load_4d(v + Cfg::KV_P0, &tmk16, &kv_full[st], 64, 2 * H + h, tok0 + j * SDB_BKV, b);
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
            const uint32_t tL = tmem_base + Cfg::L_COL + (uint32_t)(t * 16);
            const uint32_t ones_addr = smem_u32(sOnes);
            auto issue_s = [&](int j) {  // S_t(j) -> buffer j & 1; K(j) sits in stage j % STAGES
                const int st = j % SDB_STAGES;
                mbar_wait(&kv_full[st], (uint32_t)(j / SDB_STAGES) & 1u);
                tcgen05_fence_after();
                const uint32_t k_addr = smem_u32(sK + st * Cfg::KV_BYTES);
                const uint32_t tS = tmem_base + (uint32_t)(t * Cfg::S_COLS + sbuf(j) * 64);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_bf16_ss(tS, umma_desc_sw128(q_addr + k * 32), umma_desc_sw128(k_addr + k * 32), idescS,
                                 (uint32_t)(k > 0));
                if (Cfg::TAIL && !(dbg & 2))
                    umma_bf16_ss(tS, umma_desc(q_addr + Cfg::Q_P0, 0, 256, 6), umma_desc(k_addr + Cfg::KV_P0, 0, 256, 6),
                                 idescS, 1u);
                umma_commit(&s_full[t * 2 + sbuf(j)]);
            };
            mbar_wait(q_full, 0);
            issue_s(0);
            if (NBUF == 2 && nblk > 1) issue_s(1);
            for (int j = 0; j < nblk; ++j) {
                const int st = j % SDB_STAGES;
                const int valid = min(SDB_BKV, T - j * SDB_BKV);
                const int ksteps = (valid + 15) >> 4;
                const uint32_t v_addr = smem_u32(sV + st * Cfg::KV_BYTES);
                // O_t += P_t(j) V(j): k runs over the keys of this block, 16 per MMA; P is read from TMEM
                const uint32_t tP = tmem_base + (uint32_t)(t * Cfg::S_COLS + sbuf(j) * 64);
                SDB_TRACE(j, 8);
                mbar_wait(&p_full[t * 2 + sbuf(j)], spar(j));
                SDB_TRACE(j, 9);
                tcgen05_fence_after();
                for (int kk = 0; kk < ksteps; ++kk) {
                    const uint32_t acc = (uint32_t)((j | kk) != 0);
                    umma_bf16_ts(tO, tP + (uint32_t)(kk * 8), umma_desc(v_addr + kk * 2048, 0, 1024, 2), idescV64, acc);
                    if (Cfg::TAIL && !(dbg & 1))
                        umma_bf16_ts(tO + 64, tP + (uint32_t)(kk * 8),
                                     umma_desc(v_addr + Cfg::KV_P0 + kk * 512, 0, 256, 6), idescV16, acc);
                    // row sums on the tensor cores: L += P . 1 (a constant 16 x 16 tile of ones, any layout): the tensor
                    // pipe has slack, and l is then the sum of exactly the bf16 P values the PV MMAs consumed
                    umma_bf16_ts(tL, tP + (uint32_t)(kk * 8), umma_desc(ones_addr, 0, 256, 6), idescV16, acc);
                }
                umma_commit(&kv_empty[st]);  // this tile is done with K(j) (read by S_t(j), issued earlier) and V(j)
                umma_commit(&pv_done[t]);
                if (j == nblk - 1) umma_commit(&o_done[t]);
                // S_t(j+2) reuses the buffer of P_t(j): queued behind PV_t(j), in-order execution protects it
                SDB_TRACE(j, 10);
                if (j + NBUF < nblk) issue_s(j + NBUF);  // the next S into the buffer PV_t(j) has just been queued to read
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
            const uint32_t tL = tmem_base + Cfg::L_COL + (uint32_t)(t * 16) + lane_off;
            const int row = q0 + t * SDB_BQ + q * 32 + lane;
            float m_used = -INFINITY, l = 0.f;
            const bool rows_live = q0 + t * SDB_BQ + q * 32 < T;  // warp-uniform: at least one of the 32 rows exists
            auto load_s = [&](int j, uint32_t (&s)[2][32]) {     // wait for S_t(j), start its TMEM -> register loads
                const uint32_t tS = tmem_base + (uint32_t)(t * Cfg::S_COLS + sbuf(j) * 64) + lane_off;
                if (warp == 0) SDB_TRACE(j, 0);
                mbar_wait(&s_full[t * 2 + sbuf(j)], spar(j));
                if (warp == 0) SDB_TRACE(j, 1);
                tcgen05_fence_after();
#pragma unroll
                for (int c = 0; c < 2; ++c) tmem_ld_32x32(tS + (uint32_t)(c * 32), s[c]);
            };
            // (prefetching S_t(j+1) into a second register buffer while block j is exponentiated was tried: 168
            // registers do not hold both rows, ptxas spills one and the kernel gets 2x slower)
            for (int j = 0; j < nblk; ++j) {
                const int buf = sbuf(j);
                const uint32_t tS = tmem_base + (uint32_t)(t * Cfg::S_COLS + buf * 64) + lane_off;
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
                    }
                    const float nm = -m_used;
                    const uint64_t cc = sdb_pack2(scale_log2, scale_log2), nn = sdb_pack2(nm, nm);
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        if (c * 32 < valid) {  // warp-uniform: chunks without a single key are never read by PV
                            uint32_t pk[16];
#pragma unroll
                            for (int i = 0; i < 16; ++i) {
                                float x0, x1;
                                sdb_scale2(s[c][2 * i], s[c][2 * i + 1], cc, nn, x0, x1);
                                const float p0 = sdb_ex2(x0);
                                const float p1 = sdb_ex2(x1);
                                pk[i] = pack_bf16x2(p0, p1);
                            }
                            tmem_st_32x16(tS + (uint32_t)(c * 16), pk);
                        }
                    }
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
                        {  // the row-sum accumulator scales with O
                            uint32_t o2[16];
                            tmem_ld_32x16(tL, o2);
                            tmem_ld_wait();
#pragma unroll
                            for (int i = 0; i < 16; ++i) o2[i] = __float_as_uint(__uint_as_float(o2[i]) * factor);
                            tmem_st_32x16(tL, o2);
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
            {  // the row sum: any column of the L accumulator (sum of the bf16 P the PV MMAs consumed)
                uint32_t o2[16];
                tmem_ld_32x16(tL, o2);
                tmem_ld_wait();
                l = __uint_as_float(o2[0]);
            }
            const float inv = 1.0f / l;
            const int D = H * HD;
            __nv_bfloat16* orow = out + ((size_t)b * T_arg + tok0 + row) * D + (size_t)h * HD;
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
#ifdef GVL_EXPERIMENTS
    static const int dbg = [] { const char* e = getenv("GVL_ATTN_DEBUG"); return e ? atoi(e) : 0; }();  // timing experiments
#else
    const int dbg = 0;
#endif
    dim3 grid((T + NT * SDB_BQ - 1) / (NT * SDB_BQ), H, B);
    ProfScope prof(GVL_K_ATTENTION, 4.0 * B * (double)H * T * (double)T * HD, s);
    GVL_CUDA(launch_pdl(attention_sdb_kernel<HD, NT>, grid, dim3(Cfg::THREADS), Cfg::SMEM_BYTES, s, tq64, tq16, tk64, tk16,
                        reinterpret_cast<__nv_bfloat16*>(out), T, H, scale * 1.4426950408889634f, dbg, g_attn_trace,
                        (const int4*)nullptr));
    GVL_LAUNCH_CHECK("attention_sdb_kernel");
    return 0;
}

// Ragged batch: qkv / out hold M_total token rows (items back to back); tiles: device int4 [n_tiles] (see the kernel).
template <int HD>
int launch_attention_sdb_varlen(const void* qkv, void* out, int M_total, const void* tiles, int n_tiles, double score_elems,
                                int H, float scale, cudaStream_t s) {
    using Cfg = SdbCfg<HD, 1>;
    const uint64_t dims[4] = {(uint64_t)HD, (uint64_t)3 * H, (uint64_t)M_total, 1};
    const uint64_t strides[3] = {(uint64_t)HD * 2, (uint64_t)3 * H * HD * 2, (uint64_t)M_total * 3 * H * HD * 2};
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
    GVL_CUDA(cudaFuncSetAttribute(attention_sdb_kernel<HD, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  Cfg::SMEM_BYTES));
    ProfScope prof(GVL_K_ATTENTION, 4.0 * H * score_elems * HD, s);
    GVL_CUDA(launch_pdl(attention_sdb_kernel<HD, 1, true>, dim3(n_tiles, H, 1), dim3(Cfg::THREADS), Cfg::SMEM_BYTES, s, tq64,
                        tq16, tk64, tk16, reinterpret_cast<__nv_bfloat16*>(out), M_total, H, scale * 1.4426950408889634f, 0,
                        (long long*)nullptr, reinterpret_cast<const int4*>(tiles)));
    GVL_LAUNCH_CHECK("attention_sdb_kernel<varlen>");
    return 0;
}
template int launch_attention_sdb_varlen<72>(const void*, void*, int, const void*, int, double, int, float, cudaStream_t);
template int launch_attention_sdb_varlen<64>(const void*, void*, int, const void*, int, double, int, float, cudaStream_t);

template <int HD>
int launch_attention_sdb(const void* qkv, void* out, int B, int T, int H, float scale, cudaStream_t s) {
#ifdef GVL_EXPERIMENTS  // A/B builds: query tiles per CTA (2: one CTA per SM; 3: single-buffered S) — both measure slower
    static const int nt = [] {
        const char* e = getenv("GVL_ATTN_NT");
        return (e && (e[0] == '2' || e[0] == '3')) ? e[0] - '0' : 1;
    }();
    if (nt == 3) return launch_attention_sdb_nt<HD, 3>(qkv, out, B, T, H, scale, s);
    if (nt == 2) return launch_attention_sdb_nt<HD, 2>(qkv, out, B, T, H, scale, s);
#endif
    return launch_attention_sdb_nt<HD, 1>(qkv, out, B, T, H, scale, s);
}

template int launch_attention_sdb<72>(const void*, void*, int, int, int, float, cudaStream_t);
template int launch_attention_sdb<64>(const void*, void*, int, int, int, float, cudaStream_t);

}  // namespace gvl

// Tuning aid, not part of the product ABI surface in include/gvl.h: device buffer of >= 16 x 16 int64 that receives
// the timeline stamps of one CTA (see SDB_TRACE); NULL turns tracing off.
extern "C" __attribute__((visibility("default"))) void gvl_debug_set_attn_trace(void* device_buffer) {
    gvl::g_attn_trace = reinterpret_cast<long long*>(device_buffer);
}
