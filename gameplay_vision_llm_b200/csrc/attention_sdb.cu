// attention_sdb.cu — K4, production variant: fused non-causal self-attention on tcgen05 with the score tile
// DOUBLE-BUFFERED in tensor memory.
//
// PERSISTENT CTAs: two CTAs per SM (2 x (2 x 64 S + 80 O + 16 L) TMEM columns), 6 warps each; a CTA walks the work
// items (one item = one 128-row query tile of one (image, head)) blockIdx.x, blockIdx.x + gridDim.x, ... with every
// mbarrier phase, the K/V ring and the S buffer parity running on across items, so that barrier initialisation, TMEM
// allocation, the CTA launch and — above all — the DRAM latency of the next item's Q / first K blocks (the producer
// runs ahead through the ring) and its S(0), S(1) MMAs overlap the previous item's last blocks and its O read-out.
// What is left exposed per item is the O read-out itself.  (One item per CTA measured 7.2 k of 27 k cycles per item
// in start-up + tail, profiles/r02_attention_experiments.md section 3.  Round 1's two- and three-tiles-per-CTA
// variants measured slower and are gone.)
//   - S(j+1) is computed into the other buffer while softmax(j) runs; PV(j) is issued as soon as P(j) is written
//     and S(j+2) is queued right behind it (tcgen05.mma executes in issue order, which protects the P(j) columns);
//   - keys are processed in blocks of 64;
//   - the softmax inner loop is FFMA2 (two scores per instruction), MUFU.EX2 and F2FP only: the row sums are
//     accumulated by the tensor cores (L += P . 1 with a constant tile of ones), so l is the sum of exactly the bf16
//     P values the PV MMAs consume.
// What bounds it (tools/attn_trace.py, profiles/r02_attention_persistent_ncu_summary.txt): no pipe
// is saturated (ncu: MUFU 52 %, tensor 39 %, issue slots 37 %).  A scheduler holds two softmax warps, one per resident
// CTA; one alone issues a MUFU every ~10-13 cycles (8 is the pipe's rate), and the two drift into phase — exponentiate
// together at half rate each, then do their ~450 cycles of non-MUFU work (barrier round trip, tcgen05.ld, max,
// tcgen05.st, arrive) together with the pipe idle: ~1.65 k cycles per 64-key block per CTA against 1.02 k of MUFU work
// for the pair.  Measured and not kept: more softmax warps per scheduler, fewer softmax instructions, a polynomial exp2
// share, fewer MMAs, a leaner MMA issue path (SLOWER: bursts from one CTA delay the other's S on the shared tensor
// pipe), dedicated read-out warps, register prefetch of S(j+1), an overlapped non-blocking s_full test
// (profiles/r02_attention_experiments.md).  252-259 us per layer (B 64, T 729, 16 heads of 72) at 1.965 GHz.
//
//   warps 0-3   softmax: one thread = one query row x 64 keys; tcgen05.ld the S row,
//               running max with lazy rescaling of O and L (only when the max grows by more than 2^8),
//               p = exp2(s*c - m), P packed to bf16 pairs and written over the first 32 columns of the S buffer
//               (the thread's own row, already in registers) with tcgen05.st.
//   warp 4      TMA producer: per item the Q tile, then K/V blocks of 64 keys through a ring.  Head dim 72 is fetched as a
//               64-wide SWIZZLE_128B panel plus a 16-wide SWIZZLE_32B panel whose upper 8 columns are out of bounds
//               (zero) — the k-padding 72 -> 80 costs no memory.
//   warp 5      MMA issuer: S = Q K^T (SS, both K-major), O += P V and L += P 1 with P read from TMEM
//               (TS form) and V as an MN-major shared-memory operand (64-wide and 16-wide N panels).
//
// TMEM columns: S/P buffer u at [64 u, +64); O at [128, +DPAD); L at [128 + DPAD, +16).
#include "common.cuh"

#include <algorithm>
#include <cstdlib>
#include <type_traits>

namespace gvl {

constexpr int SDB_BQ = 128;            // rows per query tile
constexpr int SDB_BKV = 64;            // keys per block
constexpr float SDB_RESCALE_THRESHOLD = 8.0f;  // log2 units

template <int HD>
struct SdbCfg {
    static constexpr int THREADS = 6 * 32;   // 4 softmax warps, TMA warp, MMA warp
    static constexpr int TMEM_COLS = 256;
    static constexpr int STAGES = 4;
    static constexpr int S_COLS = 128;       // two score buffers
    static constexpr int O_COL = S_COLS;     // first output-accumulator column
    static constexpr bool TAIL = HD > 64;    // second, 16-wide panel for d in [64, 80)
    static constexpr int DPAD = TAIL ? 80 : 64;
    static constexpr int L_COL = O_COL + DPAD;         // row-sum accumulator, 16 columns
    static constexpr int Q_P0 = 128 * 128;             // 128 rows x 64 bf16, SWIZZLE_128B
    static constexpr int Q_P1 = TAIL ? 128 * 32 : 0;   // 128 rows x 16 bf16, SWIZZLE_32B
    static constexpr int Q_BYTES = Q_P0 + Q_P1;
    static constexpr int KV_P0 = SDB_BKV * 128;
    static constexpr int KV_P1 = TAIL ? SDB_BKV * 32 : 0;
    static constexpr int KV_BYTES = KV_P0 + KV_P1;     // one K or V block
    static constexpr int ONES_BYTES = 512;             // 16 keys x 16 columns of bf16 1.0
    static constexpr int STAGE_BYTES = 32 * 64;        // per softmax warp: 32 output rows x 32 bf16 on their way to global
    static constexpr int SMEM_BYTES =
        Q_BYTES + 2 * STAGES * KV_BYTES + ONES_BYTES + 256 /*barriers*/ + 4 * STAGE_BYTES + 1024 /*alignment*/;
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

// Timeline instrumentation (tools/attn_trace.py): when `trace` is set, the CTA that owns item (x 1, head 3, image 17)
// records SM-clock stamps per key block of that item (rows 0-11 of the trace buffer, row 12 = its finalisation) and of
// the item it runs next (rows 16-28): softmax warp 0 / lane 0 -> slots 0-5, the MMA thread -> slots 8-11, the TMA
// thread -> slot 12.  (compiled in only with -DGVL_ATTN_TRACE: the stamps cost 10 % of the kernel — its softmax warps
// are bound by their own instruction stream, every extra instruction per key block shows)
#ifdef GVL_ATTN_TRACE
#define SDB_TRACE(jj, slot) \
    do { if (traced) trace[(trow + (jj)) * 16 + (slot)] = clock64(); } while (0)
#else
#define SDB_TRACE(jj, slot) do { } while (0)
#endif

// One work item: a 128-row query tile of one (image, head).  Items are numbered x fastest (the query tiles of one
// (image, head) are neighbours, so the CTAs that run at the same time share K / V in L2), then head, then image.
struct SdbItem {
    int q0, T, tok0, h, b, nblk;
};

// VARLEN (the ragged batch of the masked-region route, gvl_attention_varlen_bf16): items of different lengths are
// concatenated along the token axis (the tensor maps see ONE batch of `T` = all tokens); x indexes a table of
// query tiles {first token of the item, item length, first query row of the tile}.  A Q box that runs past its item's
// end reads the next item's rows (finite; those output rows are never stored), K / V boxes likewise (those keys are
// masked to -inf, so the next item's finite V rows are multiplied by exactly zero).
template <bool VARLEN>
__device__ __forceinline__ SdbItem sdb_item(int w, int NX, int H, int T_arg, const int4* __restrict__ tiles) {
    SdbItem it;
    const int x = w % NX, r = w / NX;
    it.h = r % H;
    it.b = r / H;
    it.q0 = x * SDB_BQ;
    it.T = T_arg;
    it.tok0 = 0;  // first token of the item on the token axis
    if (VARLEN) {
        const int4 e = tiles[x];
        it.tok0 = e.x;
        it.T = e.y;
        it.q0 = e.z;
    }
    it.nblk = (it.T + SDB_BKV - 1) / SDB_BKV;
    return it;
}

template <int HD, bool VARLEN = false>
__global__ void __launch_bounds__(SdbCfg<HD>::THREADS, 2)
attention_sdb_kernel(const __grid_constant__ CUtensorMap tmq64, const __grid_constant__ CUtensorMap tmq16,
                     const __grid_constant__ CUtensorMap tmk64, const __grid_constant__ CUtensorMap tmk16,
                     __nv_bfloat16* __restrict__ out, int T_arg, int H, int NX, int n_items, float scale_log2,
                     long long* __restrict__ trace, const int4* __restrict__ tiles) {
    using Cfg = SdbCfg<HD>;
    constexpr int SDB_STAGES = Cfg::STAGES;
    constexpr int W_TMA = 4, W_MMA = 5;  // warp roles: [0, 4) softmax, then TMA, then MMA
    extern __shared__ uint8_t sdb_smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(sdb_smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;                                  // [P0 | P1]
    uint8_t* sK = smem + Cfg::Q_BYTES;                   // [stage][P0 | P1]
    uint8_t* sV = sK + SDB_STAGES * Cfg::KV_BYTES;       // [stage][P0 | P1]
    uint8_t* sOnes = sV + SDB_STAGES * Cfg::KV_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sOnes + Cfg::ONES_BYTES);
    uint64_t* q_full = bars;                      // Q of the item loaded, one phase per item
    uint64_t* q_empty = bars + 1;                 // every S MMA of the item complete (Q may be overwritten), one phase per item
    uint64_t* kv_full = bars + 2;                 // [STAGES]
    uint64_t* kv_empty = kv_full + SDB_STAGES;    // [STAGES]
    uint64_t* s_full = kv_empty + SDB_STAGES;     // [buffer]  S(g) complete in buffer g & 1
    uint64_t* p_full = s_full + 2;                // [buffer]  P(g) written (4 warps)
    uint64_t* pv_done = p_full + 2;               // PV(g) complete, one phase per block (rare rescale path only: a
                                                  // parity wait is valid at most one phase behind)
    uint64_t* o_done = pv_done + 1;               // last PV of the item complete, one phase per item
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(o_done + 1);
    uint8_t* sStage = sOnes + Cfg::ONES_BYTES + 256;  // [softmax warp][32 rows x 64 B]
    // g = key blocks this CTA has processed before + j: the block counter that runs on across items.  Buffer and
    // mbarrier phase parity of block g (double-buffered: two blocks per phase pair); K/V stage and its phase.
    auto sbuf = [](int g) { return g & 1; };
    auto spar = [](int g) { return (uint32_t)(g >> 1) & 1u; };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    pdl_trigger();

    if (warp == W_TMA && lane == 0) {
        tma_prefetch_desc(&tmq64);
        tma_prefetch_desc(&tmk64);
        if (Cfg::TAIL) {
            tma_prefetch_desc(&tmq16);
            tma_prefetch_desc(&tmk16);
        }
        mbar_init(q_full, 1);
        mbar_init(q_empty, 1);
        for (int s = 0; s < SDB_STAGES; ++s) {
            mbar_init(&kv_full[s], 1);
            mbar_init(&kv_empty[s], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&s_full[i], 1);
            mbar_init(&p_full[i], 4);
        }
        mbar_init(pv_done, 1);
        mbar_init(o_done, 1);
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
        // ===== TMA producer: runs ahead of the consumers through the K/V ring, across item boundaries =====
        if (elect_one()) {
            int st = 0, n = 0;
            uint32_t ph = 0;
            for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++n) {
                const SdbItem it = sdb_item<VARLEN>(w, NX, H, T_arg, tiles);
                const int tw = 1 + NX * (3 + H * 17), trow = w == tw ? 0 : 16;
                const bool traced = trace != nullptr && (w == tw || w == tw + (int)gridDim.x);
                if (n > 0) mbar_wait(q_empty, (uint32_t)(n - 1) & 1u);  // the previous item's S MMAs have read Q
                mbar_arrive_expect_tx(q_full, Cfg::Q_BYTES);
                tma_load_4d(sQ, &tmq64, q_full, 0, it.h, it.tok0 + it.q0, it.b);
                if (Cfg::TAIL) tma_load_4d(sQ + Cfg::Q_P0, &tmq16, q_full, 64, it.h, it.tok0 + it.q0, it.b);
                for (int j = 0; j < it.nblk; ++j) {
                    mbar_wait(&kv_empty[st], ph ^ 1);
                    SDB_TRACE(j, 12);
                    mbar_arrive_expect_tx(&kv_full[st], 2 * Cfg::KV_BYTES);
                    uint8_t* k = sK + st * Cfg::KV_BYTES;
                    uint8_t* v = sV + st * Cfg::KV_BYTES;
                    tma_load_4d(k, &tmk64, &kv_full[st], 0, H + it.h, it.tok0 + j * SDB_BKV, it.b);
                    tma_load_4d(v, &tmk64, &kv_full[st], 0, 2 * H + it.h, it.tok0 + j * SDB_BKV, it.b);
                    if (Cfg::TAIL) {
                        tma_load_4d(k + Cfg::KV_P0, &tmk16, &kv_full[st], 64, H + it.h, it.tok0 + j * SDB_BKV, it.b);
                        tma_load_4d(v + Cfg::KV_P0, &tmk16, &kv_full[st], 64, 2 * H + it.h, it.tok0 + j * SDB_BKV, it.b);
                    }
                    if (++st == SDB_STAGES) {
                        st = 0;
                        ph ^= 1;
                    }
                }
            }
        }
    } else if (warp == W_MMA) {
        // ===== MMA issuer =====
        if (elect_one()) {
            constexpr uint32_t idescS = umma_idesc_bf16_major(128, SDB_BKV, 0, 0);  // Q, K both K-major
            constexpr uint32_t idescV64 = umma_idesc_bf16_major(128, 64, 0, 1);     // V: MN-major B
            constexpr uint32_t idescV16 = umma_idesc_bf16_major(128, 16, 0, 1);
            const uint32_t q_addr = smem_u32(sQ);
            const uint32_t tO = tmem_base + Cfg::O_COL;
            const uint32_t tL = tmem_base + Cfg::L_COL;
            const uint32_t ones_addr = smem_u32(sOnes);
            int g0 = 0, n = 0;
            for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++n) {
                const SdbItem it = sdb_item<VARLEN>(w, NX, H, T_arg, tiles);
                const int tw = 1 + NX * (3 + H * 17), trow = w == tw ? 0 : 16;
                const bool traced = trace != nullptr && (w == tw || w == tw + (int)gridDim.x);
                const int nblk = it.nblk;
                auto issue_s = [&](int j) {  // S(g) -> buffer g & 1; K(g) sits in stage g % STAGES
                    const int g = g0 + j, st = g % SDB_STAGES;
                    mbar_wait(&kv_full[st], (uint32_t)(g / SDB_STAGES) & 1u);
                    tcgen05_fence_after();
                    const uint32_t k_addr = smem_u32(sK + st * Cfg::KV_BYTES);
                    const uint32_t tS = tmem_base + (uint32_t)(sbuf(g) * 64);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16_ss(tS, umma_desc_sw128(q_addr + k * 32), umma_desc_sw128(k_addr + k * 32), idescS,
                                     (uint32_t)(k > 0));
                    if (Cfg::TAIL)
                        umma_bf16_ss(tS, umma_desc(q_addr + Cfg::Q_P0, 0, 256, 6), umma_desc(k_addr + Cfg::KV_P0, 0, 256, 6),
                                     idescS, 1u);
                    umma_commit(&s_full[sbuf(g)]);
                    if (j == nblk - 1) umma_commit(q_empty);  // the item's last read of Q
                };
                // S(0), S(1) of this item: their buffers held P(g0 - 2), P(g0 - 1) of the previous item, whose PV MMAs
                // are already queued (tcgen05.mma executes in issue order) — so these run while the softmax warps
                // drain the previous item's O
                mbar_wait(q_full, (uint32_t)n & 1u);
                issue_s(0);
                if (nblk > 1) issue_s(1);
                for (int j = 0; j < nblk; ++j) {
                    const int g = g0 + j, st = g % SDB_STAGES;
                    const int valid = min(SDB_BKV, it.T - j * SDB_BKV);
                    const int ksteps = (valid + 15) >> 4;
                    const uint32_t v_addr = smem_u32(sV + st * Cfg::KV_BYTES);
                    // O += P(g) V(g): k runs over the keys of this block, 16 per MMA; P is read from TMEM.  The first
                    // block of an item overwrites O and L: its p_full phase completes only after all four softmax
                    // warps have read the previous item's O and L out (they do that before they touch S(g0)).
                    const uint32_t tP = tmem_base + (uint32_t)(sbuf(g) * 64);
                    SDB_TRACE(j, 8);
                    mbar_wait(&p_full[sbuf(g)], spar(g));
                    SDB_TRACE(j, 9);
                    tcgen05_fence_after();
                    for (int kk = 0; kk < ksteps; ++kk) {
                        const uint32_t acc = (uint32_t)((j | kk) != 0);
                        umma_bf16_ts(tO, tP + (uint32_t)(kk * 8), umma_desc(v_addr + kk * 2048, 0, 1024, 2), idescV64, acc);
                        if (Cfg::TAIL)
                            umma_bf16_ts(tO + 64, tP + (uint32_t)(kk * 8),
                                         umma_desc(v_addr + Cfg::KV_P0 + kk * 512, 0, 256, 6), idescV16, acc);
                        // row sums on the tensor cores: L += P . 1 (a constant 16 x 16 tile of ones, any layout): the
                        // tensor pipe has slack, and l is then the sum of exactly the bf16 P values the PV MMAs consumed
                        umma_bf16_ts(tL, tP + (uint32_t)(kk * 8), umma_desc(ones_addr, 0, 256, 6), idescV16, acc);
                    }
                    umma_commit(&kv_empty[st]);  // done with K(g) (read by S(g), issued earlier) and V(g)
                    umma_commit(pv_done);
                    if (j == nblk - 1) umma_commit(o_done);
                    // S(g+2) reuses the buffer of P(g): queued behind PV(g), in-order execution protects it
                    SDB_TRACE(j, 10);
                    if (j + 2 < nblk) issue_s(j + 2);
                    SDB_TRACE(j, 11);
                }
                g0 += nblk;
            }
        }
    } else {
        // ===== softmax: one thread = one query row x 64 keys per block =====
        const int q = warp & 3;  // TMEM lane quadrant this warp may access
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        const uint32_t tO = tmem_base + Cfg::O_COL + lane_off;
        const uint32_t tL = tmem_base + Cfg::L_COL + lane_off;
        int g0 = 0, n = 0;
        for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++n) {
            const SdbItem it = sdb_item<VARLEN>(w, NX, H, T_arg, tiles);
            const int tw = 1 + NX * (3 + H * 17), trow = w == tw ? 0 : 16;
            const bool traced = trace != nullptr && (w == tw || w == tw + (int)gridDim.x) && lane == 0;
            const int T = it.T, nblk = it.nblk;
            const int row = it.q0 + q * 32 + lane;
            float m_used = -INFINITY;
            const bool rows_live = it.q0 + q * 32 < T;  // warp-uniform: at least one of the 32 rows exists
            // (prefetching S(g+1) into a second register buffer while block g is exponentiated was tried: 168
            // registers do not hold both rows, ptxas spills one and the kernel gets 2x slower)
            // one key block.  PARTIAL (compile-time) = the block holds fewer than 64 keys: only an item's last block can,
            // and only that copy of the body carries the -inf masking — as a run-time `if` inside one body the compiler
            // if-converts it into 63 ISETP + 63 SEL per block (40 % of the block's instructions, ahead of the row max on
            // the critical path) for every block of every item
            auto block = [&](const int j, auto partial_tag) {
                constexpr bool PARTIAL = decltype(partial_tag)::value;
                const int g = g0 + j, buf = sbuf(g);
                const uint32_t tS = tmem_base + (uint32_t)(buf * 64) + lane_off;
                uint32_t s[2][32];
                if (warp == 0) SDB_TRACE(j, 0);
                mbar_wait(&s_full[buf], spar(g));  // S(g) complete; start its TMEM -> register loads
                if (warp == 0) SDB_TRACE(j, 1);
                tcgen05_fence_after();
#pragma unroll
                for (int c = 0; c < 2; ++c) tmem_ld_32x32(tS + (uint32_t)(c * 32), s[c]);
                tmem_ld_wait();
                if (warp == 0) SDB_TRACE(j, 2);
                if (rows_live) {
                    const int valid = PARTIAL ? T - j * SDB_BKV : SDB_BKV;
                    if (PARTIAL) {
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
                        // PV(g-1) has completed (S(g) complete implies PV(g-2) complete: parity unambiguous)
                        mbar_wait(pv_done, (uint32_t)(g - 1) & 1u);
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
                if (lane == 0) mbar_arrive(&p_full[buf]);
                if (warp == 0) SDB_TRACE(j, 5);
            };
            for (int j = 0; j + 1 < nblk; ++j) block(j, std::false_type{});
            if (T - (nblk - 1) * SDB_BKV < SDB_BKV)
                block(nblk - 1, std::true_type{});
            else
                block(nblk - 1, std::false_type{});
            g0 += nblk;
            // ---- finalise: O / l -> bf16 -> global.  Every softmax warp observes every o_done phase (also a warp
            // without live rows: that is what keeps it from running more than one item ahead); the MMA issuer has
            // meanwhile queued S(0), S(1) of the next item, and its first PV waits for this read-out through p_full.
            if (warp == 0) SDB_TRACE(12, 0);
            mbar_wait(o_done, (uint32_t)n & 1u);
            if (warp == 0) SDB_TRACE(12, 1);
            tcgen05_fence_after();
            if (rows_live) {
                // all of the item's accumulator rows in flight at once (one wait instead of four round trips), then
                // through a per-warp staging tile so that the global stores are row-contiguous: a thread owns a row,
                // and 32 rows x 16 bytes per store instruction would touch 32 lines (288 line writes per warp and
                // item; measured ~2-3.7 k cycles of read-out per item) — staged, four lanes write one row's 64 bytes
                uint32_t o0[32], o1[32], o2[16];
                const uint32_t lraw = tmem_ld_32x1(tL);  // the row sum: any column of the L accumulator
                tmem_ld_32x32(tO, o0);
                tmem_ld_32x32(tO + 32, o1);
                if (Cfg::TAIL) tmem_ld_32x16(tO + 64, o2);
                tmem_ld_wait();
                const float inv = 1.0f / __uint_as_float(lraw);
                const int D = H * HD;
                const int row0 = it.q0 + q * 32;  // the warp's first row within the item
                __nv_bfloat16* obase = out + ((size_t)it.b * T_arg + it.tok0 + row0) * D + (size_t)it.h * HD;
                uint8_t* stage = sStage + q * Cfg::STAGE_BYTES;
                const int sw = (lane >> 1) & 3;  // 16-byte slot k of row r sits at slot k ^ ((r >> 1) & 3): conflict-free
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const uint32_t(&o)[32] = c == 0 ? o0 : o1;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        uint4 v;
                        v.x = pack_bf16x2(__uint_as_float(o[k * 8 + 0]) * inv, __uint_as_float(o[k * 8 + 1]) * inv);
                        v.y = pack_bf16x2(__uint_as_float(o[k * 8 + 2]) * inv, __uint_as_float(o[k * 8 + 3]) * inv);
                        v.z = pack_bf16x2(__uint_as_float(o[k * 8 + 4]) * inv, __uint_as_float(o[k * 8 + 5]) * inv);
                        v.w = pack_bf16x2(__uint_as_float(o[k * 8 + 6]) * inv, __uint_as_float(o[k * 8 + 7]) * inv);
                        *reinterpret_cast<uint4*>(stage + lane * 64 + ((k ^ sw) << 4)) = v;
                    }
                    __syncwarp();
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int r = i * 8 + (lane >> 2), k = lane & 3;
                        const uint4 v = *reinterpret_cast<const uint4*>(stage + r * 64 + ((k ^ ((r >> 1) & 3)) << 4));
                        if (row0 + r < T) *reinterpret_cast<uint4*>(obase + (size_t)r * D + c * 32 + k * 8) = v;
                    }
                    __syncwarp();
                }
                if (Cfg::TAIL && row < T) {
                    uint4 v;  // d = 64..71 (columns 72..79 are the zero padding)
                    v.x = pack_bf16x2(__uint_as_float(o2[0]) * inv, __uint_as_float(o2[1]) * inv);
                    v.y = pack_bf16x2(__uint_as_float(o2[2]) * inv, __uint_as_float(o2[3]) * inv);
                    v.z = pack_bf16x2(__uint_as_float(o2[4]) * inv, __uint_as_float(o2[5]) * inv);
                    v.w = pack_bf16x2(__uint_as_float(o2[6]) * inv, __uint_as_float(o2[7]) * inv);
                    *reinterpret_cast<uint4*>(obase + (size_t)lane * D + 64) = v;
                }
            }
            if (warp == 0) SDB_TRACE(12, 2);
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
int g_attn_ctas_per_sm = 2;         // set by gvl_debug_set_attn_ctas_per_sm (tuning aid): 0 = one item per CTA

// CTAs of a launch over n_items work items: the resident set (two per SM) walks the items; fewer items than that get
// one CTA each.
static int sdb_grid(int n_items) {
    if (g_attn_ctas_per_sm <= 0) return n_items;
    return std::min(n_items, g_attn_ctas_per_sm * sm_count());
}

template <int HD, bool VARLEN>
static int launch_sdb(const void* qkv, void* out, uint64_t tokens_per_image, int B, int T_arg, int H, int NX,
                      const void* tiles, double flops, float scale, cudaStream_t s) {
    using Cfg = SdbCfg<HD>;
    // qkv viewed as [B][tokens][3H][HD], innermost first; Q boxes hold 128 rows, K/V boxes 64
    const uint64_t dims[4] = {(uint64_t)HD, (uint64_t)3 * H, tokens_per_image, (uint64_t)B};
    const uint64_t strides[3] = {(uint64_t)HD * 2, (uint64_t)3 * H * HD * 2, tokens_per_image * 3 * H * HD * 2};
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
    const long long n_items = (long long)NX * H * B;
    if (n_items > 2147483647LL) {
        set_error("attention: %lld work items exceed the 32-bit item counter", n_items);
        return 1;
    }
    GVL_CUDA(cudaFuncSetAttribute(attention_sdb_kernel<HD, VARLEN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  Cfg::SMEM_BYTES));
    ProfScope prof(GVL_K_ATTENTION, flops, s);
    GVL_CUDA(launch_pdl(attention_sdb_kernel<HD, VARLEN>, dim3(sdb_grid((int)n_items)), dim3(Cfg::THREADS), Cfg::SMEM_BYTES,
                        s, tq64, tq16, tk64, tk16, reinterpret_cast<__nv_bfloat16*>(out), T_arg, H, NX, (int)n_items,
                        scale * 1.4426950408889634f, g_attn_trace, reinterpret_cast<const int4*>(tiles)));
    GVL_LAUNCH_CHECK(VARLEN ? "attention_sdb_kernel<varlen>" : "attention_sdb_kernel");
    return 0;
}

// Ragged batch: qkv / out hold M_total token rows (items back to back); tiles: device int4 [n_tiles] (see sdb_item).
template <int HD>
int launch_attention_sdb_varlen(const void* qkv, void* out, int M_total, const void* tiles, int n_tiles, double score_elems,
                                int H, float scale, cudaStream_t s) {
    return launch_sdb<HD, true>(qkv, out, (uint64_t)M_total, 1, M_total, H, n_tiles, tiles, 4.0 * H * score_elems * HD, scale,
                                s);
}
template int launch_attention_sdb_varlen<72>(const void*, void*, int, const void*, int, double, int, float, cudaStream_t);
template int launch_attention_sdb_varlen<64>(const void*, void*, int, const void*, int, double, int, float, cudaStream_t);

template <int HD>
int launch_attention_sdb(const void* qkv, void* out, int B, int T, int H, float scale, cudaStream_t s) {
    return launch_sdb<HD, false>(qkv, out, (uint64_t)T, B, T, H, (T + SDB_BQ - 1) / SDB_BQ, nullptr,
                                 4.0 * B * (double)H * T * (double)T * HD, scale, s);
}

template int launch_attention_sdb<72>(const void*, void*, int, int, int, float, cudaStream_t);
template int launch_attention_sdb<64>(const void*, void*, int, int, int, float, cudaStream_t);

}  // namespace gvl

// Tuning aids, not part of the product ABI surface in include/gvl.h.
// Device buffer of >= 16 x 16 int64 that receives the timeline stamps of one CTA (see SDB_TRACE); NULL = tracing off.
extern "C" __attribute__((visibility("default"))) void gvl_debug_set_attn_trace(void* device_buffer) {
    gvl::g_attn_trace = reinterpret_cast<long long*>(device_buffer);
}
// Resident CTAs per SM the persistent attention grid is sized for (default 2); 0 = one work item per CTA (the
// non-persistent launch, for A/B measurements of the same kernel in one process).
extern "C" __attribute__((visibility("default"))) void gvl_debug_set_attn_ctas_per_sm(int n) {
    gvl::g_attn_ctas_per_sm = n;
}
