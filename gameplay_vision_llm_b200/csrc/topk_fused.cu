// topk_fused.cu — K8, query-batch path: cosine scores on tcgen05 with the top-k selection FUSED into the epilogue.
//
// The previous tensor path wrote the whole fp32 score matrix (Q x N: 36.9 MB for 128 queries over 72 000 rows) and read
// it back three times (scale, segmented select, re-score): 0.35 ms against the 0.09 ms the index read itself needs.
// Here the scores never leave the SM:
//
//   * persistent CTAs walk tiles of 256 index rows; one tile = S[128 queries x 256 rows] = Q . E_tile^T accumulated over
//     D in tensor memory (TMA -> separate shared-memory rings for the query tile and the index chunks -> tcgen05.mma
//     128x64x16 into four 64-column groups, two accumulator buffers so the epilogue of tile i overlaps the main loop of
//     tile i + 1);
//   * epilogue: one thread = one query (TMEM lane).  It scales each score by the row's 1/|e_n| (the query's own 1/|q|
//     does not change its ranking and is applied at the end), applies the query's row window, and appends every row
//     that reaches a running lower bound of the k-th best score (the 16th largest of 32 group maxima, minus the re-scoring
//     margin) to a private list of up to KL = 64 candidates (in global memory: appends are rare) — no per-element list
//     maintenance; each tile is read from TMEM twice (bound, then collection);
//   * gvl::topk_refine_kernel (topk.cu) merges them per query and re-scores
//     everything within a margin of the provisional k-th score EXACTLY (float64 dot products and norms), so the final
//     ranking is the float64 one, ties by ascending index — identical to what the scan path produces after the same
//     refinement.
//
// Index traffic: N * D * 2 bytes once per batch of <= 128 queries (the query tile, 1 MB at D = 4096, is re-read from L2).
#include "common.cuh"

namespace gvl {

constexpr int TF_BM = 128;       // queries per batch (TMEM lanes)
constexpr int TF_BN = 256;       // index rows per tile (TMEM columns per accumulator buffer)
constexpr int TF_BK = 64;        // bf16 elements per k-step (one 128-byte swizzled row)
// Loop nest and rings.  The index is consumed in chunks of TF_NG = 64 rows x TF_KA k-atoms of 64 columns, each chunk
// accumulating into its own 64 TMEM columns of the 256-row tile (MMA 128 x 64 x 16); the query tile (L2-resident) is
// staged per K block of TF_KA atoms and shared by the tile's four chunks.  Each ring has its own producer warp, and
// CTAs start their walk over K at different blocks (krot).
// Measured on 72k x 4096, 128 queries (profiles/r02_topk_variants.txt): the scoring kernel streams the index at
// ~4.1 TB/s (143-147 us) with 256-row boxes and 3+3, 4+4 stages, with 64-row chunks and 4+16 stages (shipped), with
// index boxes issued in bursts of 2 or 4 k-steps, and with the K rotation; a deeper index ring at the cost of the
// query ring is slower (3+5: 155 us, 2+6: 169 us), chunks of 2 / 4 atoms (256 / 512 contiguous bytes per row) are slower
// (153 / 156 us), and multicasting the query tile over clusters of 2 / 4 CTAs changes nothing / costs 60 % (lock-step).
// A row-contiguous reader (gvl_row_inv_norm, 512 bytes per warp instruction) gets 6.4 TB/s from the same matrix, so
// the remaining gap is in how 128-byte box rows reach DRAM, not in ring depth — what is left to try is a cp.async
// producer that fetches whole 512-byte row segments and swizzles them into place itself.
#ifndef GVL_TF_KA
#define GVL_TF_KA 1
#define GVL_TF_SA 4
#define GVL_TF_SB 16
#endif
#ifndef GVL_TF_ROT
#define GVL_TF_ROT 5
#endif
constexpr int TF_KA = GVL_TF_KA;  // 64-column atoms per K block
constexpr int TF_SA = GVL_TF_SA;  // query-tile stages (TF_KA x 16 KB each)
constexpr int TF_SB = GVL_TF_SB;  // index-chunk stages (TF_KA x 8 KB each)
constexpr int TF_NG = 64;         // index rows per chunk
constexpr int TF_KL = 64;        // candidates kept per (query, CTA), in global memory (appends are rare)
constexpr int TF_GROUPS = 32;    // running group maxima per query thread; the bound is their 16th largest (k <= 16)
constexpr int TF_THREADS = 224;  // warps 0-3 epilogue, 4 index-tile TMA, 5 MMA, 6 query-tile TMA
constexpr int TF_A_ATOM = TF_BM * TF_BK * 2;  // one [128 x 64] swizzled box
constexpr int TF_B_ATOM = TF_NG * TF_BK * 2;  // one [64 x 64] swizzled box
constexpr int TF_A_BYTES = TF_KA * TF_A_ATOM;
constexpr int TF_B_BYTES = TF_KA * TF_B_ATOM;
constexpr int TF_SMEM_BYTES = TF_SA * TF_A_BYTES + TF_SB * TF_B_BYTES + 256 + 1024;
static_assert(TF_SMEM_BYTES <= 227 * 1024, "rings exceed shared memory");

// List full (rare): keep the best TF_KL — replace the worst entry if this row beats it.  Rows arrive in ascending
// order, so an equal score never displaces an earlier row.
static __device__ __noinline__ void tf_replace_worst(float* my_s, int32_t* my_i, float s, int n) {
    float ws = my_s[0];
    int wi = my_i[0], wp = 0;
    for (int e = 1; e < TF_KL; ++e) {
        const float es = my_s[e];
        const int ei = my_i[e];
        if (es < ws || (es == ws && ei > wi)) {
            ws = es;
            wi = ei;
            wp = e;
        }
    }
    if (s > ws) {
        my_s[wp] = s;
        my_i[wp] = n;
    }
}

__global__ void __launch_bounds__(TF_THREADS, 1)
topk_fused_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_e, int D, int nq,
                  int span_lo, int span_hi, const float* __restrict__ inv_e, const float* __restrict__ inv_q, float margin,
                  int k, const int32_t* __restrict__ row_lo, const int32_t* __restrict__ row_hi, int N,
                  float* __restrict__ cand_s, int32_t* __restrict__ cand_i) {
    extern __shared__ uint8_t tf_smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tf_smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sB = smem;                                   // [SB][KA][64 x 64 bf16]
    uint8_t* sA = smem + TF_SB * TF_B_BYTES;              // [SA][KA][128 x 64 bf16]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sA + TF_SA * TF_A_BYTES);
    uint64_t* full_a = bars;                 // [SA]
    uint64_t* empty_a = full_a + TF_SA;      // [SA]
    uint64_t* full_b = empty_a + TF_SA;      // [SB]
    uint64_t* empty_b = full_b + TF_SB;      // [SB]
    uint64_t* acc_full = empty_b + TF_SB;    // [2]
    uint64_t* acc_empty = acc_full + 2;      // [2]
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ntiles = (span_hi - span_lo + TF_BN - 1) / TF_BN;
    const int ksteps = (D + TF_BK - 1) / TF_BK;

    if (warp == 4 && lane == 0) {
        tma_prefetch_desc(&tm_q);
        tma_prefetch_desc(&tm_e);
        for (int s = 0; s < TF_SA; ++s) {
            mbar_init(&full_a[s], 1);
            mbar_init(&empty_a[s], 1);
        }
        for (int s = 0; s < TF_SB; ++s) {
            mbar_init(&full_b[s], 1);
            mbar_init(&empty_b[s], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&acc_full[i], 1);
            mbar_init(&acc_empty[i], 4);
        }
        fence_barrier_init();
    }
    if (warp == 5) tmem_alloc<512>(tmem_ptr_smem);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    const int kblocks = (ksteps + TF_KA - 1) / TF_KA;  // atoms past D are out of bounds: zero fill
    // CTAs walk K from different starting blocks so that, at any moment, the 148 index streams sit at different column
    // offsets of their rows (the sum is order-independent up to fp32 rounding, which the exact re-scoring absorbs)
    const int krot = (int)((blockIdx.x * (unsigned)GVL_TF_ROT) % (unsigned)kblocks);

    if (warp == 4) {
        // ===== TMA producer, index chunks (DRAM) =====
        if (elect_one()) {
            int st = 0;
            uint32_t ph = 0;
            for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
                const int n0 = span_lo + t * TF_BN;
                for (int kb = 0; kb < kblocks; ++kb)
                    for (int ng = 0; ng < TF_BN / TF_NG; ++ng) {
                        mbar_wait(&empty_b[st], ph ^ 1u);
                        mbar_arrive_expect_tx(&full_b[st], TF_B_BYTES);
#pragma unroll
                        for (int a = 0; a < TF_KA; ++a)
                            tma_load_2d(sB + st * TF_B_BYTES + a * TF_B_ATOM, &tm_e, &full_b[st],
                                        (((kb + krot) % kblocks) * TF_KA + a) * TF_BK, n0 + ng * TF_NG);
                        if (++st == TF_SB) {
                            st = 0;
                            ph ^= 1u;
                        }
                    }
            }
        }
    } else if (warp == 6) {
        // ===== TMA producer, query tile (L2) =====
        if (elect_one()) {
            int st = 0;
            uint32_t ph = 0;
            for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
                for (int kb = 0; kb < kblocks; ++kb) {
                    mbar_wait(&empty_a[st], ph ^ 1u);
                    mbar_arrive_expect_tx(&full_a[st], TF_A_BYTES);
#pragma unroll
                    for (int a = 0; a < TF_KA; ++a)
                        tma_load_2d(sA + st * TF_A_BYTES + a * TF_A_ATOM, &tm_q, &full_a[st],
                                    (((kb + krot) % kblocks) * TF_KA + a) * TF_BK, 0);
                    if (++st == TF_SA) {
                        st = 0;
                        ph ^= 1u;
                    }
                }
            }
        }
    } else if (warp == 5) {
        // ===== MMA issuer =====
        if (elect_one()) {
            constexpr uint32_t idesc = umma_idesc_bf16(TF_BM, TF_NG);
            int sa = 0, sb = 0;
            uint32_t pha = 0, phb = 0;
            int it = 0;
            for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
                const int buf = it & 1;
                mbar_wait(&acc_empty[buf], (((uint32_t)(it >> 1)) & 1u) ^ 1u);  // epilogue of tile it - 2 has drained it
                tcgen05_fence_after();
                const uint32_t tD = tmem_base + (uint32_t)(buf * TF_BN);
                for (int kb = 0; kb < kblocks; ++kb) {
                    mbar_wait(&full_a[sa], pha);
                    const uint32_t a_addr = smem_u32(sA + sa * TF_A_BYTES);
                    for (int ng = 0; ng < TF_BN / TF_NG; ++ng) {
                        mbar_wait(&full_b[sb], phb);
                        tcgen05_fence_after();
                        const uint32_t b_addr = smem_u32(sB + sb * TF_B_BYTES);
#pragma unroll
                        for (int kk = 0; kk < TF_KA * 4; ++kk)
                            umma_bf16_ss(tD + (uint32_t)(ng * TF_NG),
                                         umma_desc_sw128(a_addr + (kk >> 2) * TF_A_ATOM + (kk & 3) * 32),
                                         umma_desc_sw128(b_addr + (kk >> 2) * TF_B_ATOM + (kk & 3) * 32), idesc,
                                         (uint32_t)((kb | kk) != 0));
                        umma_commit(&empty_b[sb]);
                        if (++sb == TF_SB) {
                            sb = 0;
                            phb ^= 1u;
                        }
                    }
                    umma_commit(&empty_a[sa]);
                    if (++sa == TF_SA) {
                        sa = 0;
                        pha ^= 1u;
                    }
                }
                umma_commit(&acc_full[buf]);
            }
        }
    } else if (warp < 4) {
        // ===== epilogue: thread = query =====
        // Candidate collection without per-element list maintenance: 32 running group maxima (element i of every
        // 32-column chunk belongs to group i); their 16th largest is a score that at least 16 distinct rows of this CTA
        // reach — a valid lower bound of the CTA's (hence the global) k-th best score for k <= 16 — and every row
        // scoring >= that bound minus the re-scoring margin is appended to the thread's list.  A full list falls back
        // to replace-the-worst (exact top-KL, slow, rare).
        const int q = warp * 32 + lane;
        const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
        const bool q_on = q < nq;
        const int lo = q_on ? (row_lo ? max(row_lo[q], span_lo) : span_lo) : 0;
        const int hi = q_on ? (row_hi ? min(row_hi[q], span_hi) : span_hi) : 0;
        const float margin_raw = q_on ? margin / inv_q[q] : 0.f;  // candidate scores lack the 1/|q| factor
        // this thread's candidate list lives in global memory (written rarely, re-read only when the bound tightens)
        float* my_s = cand_s + ((size_t)(q_on ? q : 0) * gridDim.x + blockIdx.x) * TF_KL;
        int32_t* my_i = cand_i + ((size_t)(q_on ? q : 0) * gridDim.x + blockIdx.x) * TF_KL;
        int count = 0;
        float thr = -INFINITY;
        float gmax[TF_GROUPS];
#pragma unroll
        for (int g = 0; g < TF_GROUPS; ++g) gmax[g] = -INFINITY;
        auto append = [&](float s, int n) {
            if (count < TF_KL) {
                my_s[count] = s;
                my_i[count] = n;
                ++count;
            } else {
                tf_replace_worst(my_s, my_i, s, n);
            }
        };
        // 16th largest of the 32 group maxima (bitonic sorting network on registers, descending) minus the margin:
        // at least 16 distinct rows of this CTA score >= it, so it never exceeds the k-th best (k <= 16)
        auto group_bound = [&]() {
            float v[TF_GROUPS];
#pragma unroll
            for (int g = 0; g < TF_GROUPS; ++g) v[g] = gmax[g];
#pragma unroll
            for (int k2 = 2; k2 <= TF_GROUPS; k2 <<= 1) {
#pragma unroll
                for (int j = k2 >> 1; j > 0; j >>= 1) {
#pragma unroll
                    for (int i = 0; i < TF_GROUPS; ++i) {
                        const int l = i ^ j;
                        if (l > i) {
                            const bool desc = (i & k2) == 0;
                            const float a = v[i], c = v[l];
                            const float hi_ = fmaxf(a, c), lo_ = fminf(a, c);
                            v[i] = desc ? hi_ : lo_;
                            v[l] = desc ? lo_ : hi_;
                        }
                    }
                }
            }
            return v[15] - margin_raw;  // -inf while fewer than 16 groups have seen an eligible row
        };
        // drop list entries that fell below the tightened bound (keeps the list near 16-25 entries)
        auto compact = [&]() {
            int w = 0;
            for (int e = 0; e < count; ++e) {
                const float es = my_s[e];
                const int ei = my_i[e];
                if (es >= thr) {
                    my_s[w] = es;
                    my_i[w] = ei;
                    ++w;
                }
            }
            count = w;
        };
        int it = 0;
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
            const int buf = it & 1;
            const int n0 = span_lo + t * TF_BN;
            mbar_wait(&acc_full[buf], ((uint32_t)(it >> 1)) & 1u);
            tcgen05_fence_after();
            const uint32_t tD = tmem_base + (uint32_t)(buf * TF_BN) + lane_off;
            // every tile is read from TMEM twice: first its scores tighten the bound (a later tile may hold a whole
            // scene of near-duplicates above the old bound — appending them all would flood the list), then the rows
            // that reach the new bound are collected
            for (int pass = 0; pass < 2; ++pass) {
                for (int c = 0; c < TF_BN / 32; ++c) {
                    uint32_t v[32];
                    tmem_ld_32x32(tD + (uint32_t)(c * 32), v);
                    tmem_ld_wait();
                    const int nb = n0 + c * 32;
                    const float my_ie = (nb + lane < N) ? __ldg(inv_e + nb + lane) : 0.f;  // 1/|e| of this chunk's rows
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const float ie = __shfl_sync(0xffffffffu, my_ie, i);
                        const int n = nb + i;
                        const bool in = n >= lo && n < hi;
                        const float s = __uint_as_float(v[i]) * ie;
                        if (pass == 0) gmax[i % TF_GROUPS] = fmaxf(gmax[i % TF_GROUPS], in ? s : -INFINITY);
                        else if (in && s >= thr) append(s, n);
                    }
                }
                if (pass == 0) {
                    const float nb_ = group_bound();
                    if (nb_ > thr) {
                        thr = nb_;
                        if (count > 0) compact();
                    }
                }
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[buf]);
        }
        // ---- this CTA's candidates of query q ----
        if (q_on)
            for (int e = count; e < TF_KL; ++e) my_i[e] = -1;  // empty slots (lists fill front to back)
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 5) {
        tcgen05_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

int topk_fused_grid(int span) {
    const int ntiles = (span + TF_BN - 1) / TF_BN;
    int grid = ntiles < sm_count() ? ntiles : sm_count();
    if (grid < 1) grid = 1;
    return grid;
}

// Launches the fused scoring kernel for one batch of nq <= 128 queries.  cand_s / cand_i: [nq][grid][TF_KL].
// Returns the grid size through *grid_out (the number of candidate lists per query).
int launch_topk_fused(const void* index, int N, int D, const void* queries, int nq, int span_lo, int span_hi,
                      const float* inv_e, const float* inv_q, float margin, int k, const int32_t* row_lo,
                      const int32_t* row_hi, float* cand_s, int32_t* cand_i, int* grid_out, cudaStream_t s) {
    CUtensorMap tq, te;
    int rc = make_tmap_2d_bf16(&tq, queries, (uint64_t)nq, (uint64_t)D, (uint64_t)D, TF_BM, TF_BK, true);
    if (rc) return rc;
    rc = make_tmap_2d_bf16(&te, index, (uint64_t)N, (uint64_t)D, (uint64_t)D, TF_NG, TF_BK, true);
    if (rc) return rc;
    const int grid = topk_fused_grid(span_hi - span_lo);
    GVL_CUDA(cudaFuncSetAttribute(topk_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TF_SMEM_BYTES));
    ProfScope prof(GVL_K_TOPK_SCORES, (double)(span_hi - span_lo) * D * 2, s);
    topk_fused_kernel<<<grid, TF_THREADS, TF_SMEM_BYTES, s>>>(tq, te, D, nq, span_lo, span_hi, inv_e, inv_q, margin, k, row_lo,
                                                              row_hi, N, cand_s, cand_i);
    GVL_LAUNCH_CHECK("topk_fused_kernel");
    *grid_out = grid;
    return 0;
}

int topk_fused_list_len() { return TF_KL; }

}  // namespace gvl
