// topk_fused.cu — K8, query-batch path: cosine scores on tcgen05 with the top-k selection FUSED into the epilogue.
//
// The previous tensor path wrote the whole fp32 score matrix (Q x N: 36.9 MB for 128 queries over 72 000 rows) and read
// it back three times (scale, segmented select, re-score): 0.35 ms against the 0.09 ms the index read itself needs.
// Here the scores never leave the SM:
//
//   * persistent CTAs walk tiles of 256 index rows; one tile = S[128 queries x 256 rows] = Q . E_tile^T accumulated over
//     D in tensor memory (TMA -> 3-stage smem ring -> tcgen05.mma 128x256x16, two accumulator buffers so the epilogue of
//     tile i overlaps the main loop of tile i + 1);
//   * epilogue: one thread = one query (TMEM lane).  It scales each score by the row's 1/|e_n| (the query's own 1/|q|
//     does not change its ranking and is applied at the end), applies the query's row window, and keeps the best
//     KL = 32 candidates it has seen in a private list in shared memory (threshold = the list's worst entry);
//   * every CTA writes its 128 x KL candidates; gvl::topk_refine_kernel (topk.cu) merges them per query and re-scores
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
constexpr int TF_STAGES = 3;
constexpr int TF_KL = 32;        // candidates kept per (query, CTA)
constexpr int TF_THREADS = 192;  // warps 0-3 epilogue, 4 TMA, 5 MMA
constexpr int TF_A_BYTES = TF_BM * TF_BK * 2;
constexpr int TF_B_BYTES = TF_BN * TF_BK * 2;
constexpr int TF_STAGE_BYTES = TF_A_BYTES + TF_B_BYTES;
constexpr int TF_LIST_BYTES = TF_KL * TF_BM * 8;
constexpr int TF_SMEM_BYTES = TF_STAGES * TF_STAGE_BYTES + TF_LIST_BYTES + 256 + 1024;

__global__ void __launch_bounds__(TF_THREADS, 1)
topk_fused_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_e, int D, int nq,
                  int span_lo, int span_hi, const float* __restrict__ inv_e, const int32_t* __restrict__ row_lo,
                  const int32_t* __restrict__ row_hi, int N, float* __restrict__ cand_s, int32_t* __restrict__ cand_i) {
    extern __shared__ uint8_t tf_smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tf_smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;                                   // [STAGES][128 x 64 bf16]
    uint8_t* sB = smem + TF_STAGES * TF_A_BYTES;          // [STAGES][256 x 64 bf16]
    float* ls = reinterpret_cast<float*>(smem + TF_STAGES * TF_STAGE_BYTES);   // [KL][128]
    int32_t* li = reinterpret_cast<int32_t*>(ls + TF_KL * TF_BM);              // [KL][128]
    uint64_t* bars = reinterpret_cast<uint64_t*>(li + TF_KL * TF_BM);
    uint64_t* full = bars;                   // [STAGES]
    uint64_t* empty = full + TF_STAGES;      // [STAGES]
    uint64_t* acc_full = empty + TF_STAGES;  // [2]
    uint64_t* acc_empty = acc_full + 2;      // [2]
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ntiles = (span_hi - span_lo + TF_BN - 1) / TF_BN;
    const int ksteps = (D + TF_BK - 1) / TF_BK;

    if (warp == 4 && lane == 0) {
        tma_prefetch_desc(&tm_q);
        tma_prefetch_desc(&tm_e);
        for (int s = 0; s < TF_STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
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

    if (warp == 4) {
        // ===== TMA producer =====
        if (elect_one()) {
            int st = 0;
            uint32_t ph = 0;
            for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
                const int n0 = span_lo + t * TF_BN;
                for (int k = 0; k < ksteps; ++k) {
                    mbar_wait(&empty[st], ph ^ 1u);
                    mbar_arrive_expect_tx(&full[st], TF_STAGE_BYTES);
                    tma_load_2d(sA + st * TF_A_BYTES, &tm_q, &full[st], k * TF_BK, 0);
                    tma_load_2d(sB + st * TF_B_BYTES, &tm_e, &full[st], k * TF_BK, n0);
                    if (++st == TF_STAGES) {
                        st = 0;
                        ph ^= 1u;
                    }
                }
            }
        }
    } else if (warp == 5) {
        // ===== MMA issuer =====
        if (elect_one()) {
            constexpr uint32_t idesc = umma_idesc_bf16(TF_BM, TF_BN);
            int st = 0;
            uint32_t ph = 0;
            int it = 0;
            for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
                const int buf = it & 1;
                mbar_wait(&acc_empty[buf], (((uint32_t)(it >> 1)) & 1u) ^ 1u);  // epilogue of tile it - 2 has drained it
                tcgen05_fence_after();
                const uint32_t tD = tmem_base + (uint32_t)(buf * TF_BN);
                for (int k = 0; k < ksteps; ++k) {
                    mbar_wait(&full[st], ph);
                    tcgen05_fence_after();
                    const uint32_t a_addr = smem_u32(sA + st * TF_A_BYTES), b_addr = smem_u32(sB + st * TF_B_BYTES);
#pragma unroll
                    for (int kk = 0; kk < TF_BK / 16; ++kk)
                        umma_bf16_ss(tD, umma_desc_sw128(a_addr + kk * 32), umma_desc_sw128(b_addr + kk * 32), idesc,
                                     (uint32_t)((k | kk) != 0));
                    umma_commit(&empty[st]);
                    if (++st == TF_STAGES) {
                        st = 0;
                        ph ^= 1u;
                    }
                }
                umma_commit(&acc_full[buf]);
            }
        }
    } else {
        // ===== epilogue: thread = query =====
        const int q = warp * 32 + lane;
        const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
        const bool q_on = q < nq;
        const int lo = q_on ? (row_lo ? max(row_lo[q], span_lo) : span_lo) : 0;
        const int hi = q_on ? (row_hi ? min(row_hi[q], span_hi) : span_hi) : 0;
        float* my_s = ls + q;       // entry e at my_s[e * 128]
        int32_t* my_i = li + q;
        int count = 0, minpos = 0;
        float thr = -INFINITY;      // list's worst score once it is full
        int it = 0;
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
            const int buf = it & 1;
            const int n0 = span_lo + t * TF_BN;
            mbar_wait(&acc_full[buf], ((uint32_t)(it >> 1)) & 1u);
            tcgen05_fence_after();
            const uint32_t tD = tmem_base + (uint32_t)(buf * TF_BN) + lane_off;
            for (int c = 0; c < TF_BN / 32; ++c) {
                uint32_t v[32];
                tmem_ld_32x32(tD + (uint32_t)(c * 32), v);
                tmem_ld_wait();
                const int nb = n0 + c * 32;
                if (nb >= hi || nb + 32 <= lo) continue;  // (thread-local skip; the TMEM load above is warp-collective)
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const int n = nb + i;
                    if (n < lo || n >= hi) continue;
                    const float s = __uint_as_float(v[i]) * __ldg(inv_e + n);
                    // rows arrive in ascending order: an equal score never displaces an earlier row
                    if (!(s > thr)) continue;  // also drops NaN
                    if (count < TF_KL) {
                        my_s[count * TF_BM] = s;
                        my_i[count * TF_BM] = n;
                        if (++count < TF_KL) continue;
                    } else {
                        my_s[minpos * TF_BM] = s;
                        my_i[minpos * TF_BM] = n;
                    }
                    // list full: find its worst entry (lowest score, latest row among equals)
                    float ws = my_s[0];
                    int wi = my_i[0], wp = 0;
#pragma unroll 8
                    for (int e = 1; e < TF_KL; ++e) {
                        const float es = my_s[e * TF_BM];
                        const int ei = my_i[e * TF_BM];
                        if (es < ws || (es == ws && ei > wi)) {
                            ws = es;
                            wi = ei;
                            wp = e;
                        }
                    }
                    thr = ws;
                    minpos = wp;
                }
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[buf]);
        }
        // ---- this CTA's candidates of query q ----
        if (q_on) {
            float* os = cand_s + ((size_t)q * gridDim.x + blockIdx.x) * TF_KL;
            int32_t* oi = cand_i + ((size_t)q * gridDim.x + blockIdx.x) * TF_KL;
            for (int e = 0; e < TF_KL; ++e) {
                os[e] = e < count ? my_s[e * TF_BM] : -INFINITY;
                oi[e] = e < count ? my_i[e * TF_BM] : -1;
            }
        }
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 5) {
        tcgen05_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

// Launches the fused scoring kernel for one batch of nq <= 128 queries.  cand_s / cand_i: [nq][grid][TF_KL].
// Returns the grid size through *grid_out (the number of candidate lists per query).
int launch_topk_fused(const void* index, int N, int D, const void* queries, int nq, int span_lo, int span_hi,
                      const float* inv_e, const int32_t* row_lo, const int32_t* row_hi, float* cand_s, int32_t* cand_i,
                      int* grid_out, cudaStream_t s) {
    CUtensorMap tq, te;
    int rc = make_tmap_2d_bf16(&tq, queries, (uint64_t)nq, (uint64_t)D, (uint64_t)D, TF_BM, TF_BK, true);
    if (rc) return rc;
    rc = make_tmap_2d_bf16(&te, index, (uint64_t)N, (uint64_t)D, (uint64_t)D, TF_BN, TF_BK, true);
    if (rc) return rc;
    const int ntiles = (span_hi - span_lo + TF_BN - 1) / TF_BN;
    int grid = ntiles < sm_count() ? ntiles : sm_count();
    if (grid < 1) grid = 1;
    GVL_CUDA(cudaFuncSetAttribute(topk_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TF_SMEM_BYTES));
    ProfScope prof(GVL_K_TOPK_SCORES, (double)(span_hi - span_lo) * D * 2, s);
    topk_fused_kernel<<<grid, TF_THREADS, TF_SMEM_BYTES, s>>>(tq, te, D, nq, span_lo, span_hi, inv_e, row_lo, row_hi, N,
                                                              cand_s, cand_i);
    GVL_LAUNCH_CHECK("topk_fused_kernel");
    *grid_out = grid;
    return 0;
}

int topk_fused_list_len() { return TF_KL; }

}  // namespace gvl
