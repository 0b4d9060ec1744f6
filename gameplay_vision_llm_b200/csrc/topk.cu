// topk.cu — K8: cosine top-k over the timeline index (HBM-bound scan + exact selection).
//
// Pass 1 streams the bf16 index once per batch of 8 queries: one warp per index row, 16-byte
// coalesced loads, fp32 dot products against the queries held in shared memory, the row norm computed
// from the same registers.  The per-lane accumulation order and the xor-shuffle reduction tree do not
// depend on the row's position, so identical rows get bit-identical scores (ties are real ties).
// Pass 2 selects the k best per query under the total order (score descending, index ascending) by k
// rounds of a block-wide arg-max over the elements that come after the previous winner in that order —
// no mutation, no sort network, deterministic.
//
// gvl_topk_cosine_ex adds (a) a per-query row range [row_lo, row_hi) — the timeline index is in timestamp
// order, so the reference's "events within +-window of t" filter (scripts/realtime_inference.py:988-998,
// src/agent_core/qwen_reasoning_core.py:1462-1490) is a contiguous range, and only that range is scored and
// ranked — and (b) a tensor-core scoring path for query batches: scores = Q . E^T is one skinny tcgen05 GEMM
// (M = queries, N = index rows, fp32 out) that reads the index ONCE for the whole batch instead of once per
// 8 queries, followed by a scale pass with the cached 1/|e_n| and 1/|q|.
#include "common.cuh"

#include <algorithm>

namespace gvl {

constexpr int TOPK_QB = 8;        // queries per scan pass
constexpr int TOPK_THREADS = 256;
constexpr int TOPK_REFINE_THREADS = 1024;  // one CTA per query: its float64 re-scoring is a chain of DRAM round trips, one warp per row
constexpr int TOPK_CAND = 512;    // candidates per query the tensor path re-scores exactly
constexpr int TOPK_MAX_LISTS = 512;       // candidate lists per query the refinement's pre-filter handles
constexpr int TOPK_RANK_CAP = 1024;       // candidates ranked by counting instead of k arg-max rounds
constexpr int TOPK_REFINE_STAGE = 4096;  // non-empty candidates of one query staged in shared memory by the refinement
constexpr float TOPK_MARGIN = 6e-5f;  // > 2 x the tensor path's score error (measured 9e-6 at D = 4096)
constexpr float TOPK_MARGIN_SCAN = 8e-6f;  // > 2 x the fp32 scan's accumulation error (measured < 2e-6 at D = 4096)

// The per-lane arithmetic of one 8-element chunk, shared by the scan and the rescoring kernel so that both
// produce bit-identical scores for the same (row, query) pair.
__device__ __forceinline__ void tk_unpack8(const uint4& u, float (&e)[8]) {
    e[0] = bf16_lo(u.x); e[1] = bf16_hi(u.x); e[2] = bf16_lo(u.y); e[3] = bf16_hi(u.y);
    e[4] = bf16_lo(u.z); e[5] = bf16_hi(u.z); e[6] = bf16_lo(u.w); e[7] = bf16_hi(u.w);
}
__device__ __forceinline__ float tk_sq8(float acc, const float (&e)[8]) {
#pragma unroll
    for (int j = 0; j < 8; ++j) acc = fmaf(e[j], e[j], acc);
    return acc;
}
__device__ __forceinline__ float tk_dot8(float acc, const float (&e)[8], const uint4& v) {
    float q[8];
    tk_unpack8(v, q);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc = fmaf(e[j], q[j], acc);
    return acc;
}
__device__ __forceinline__ float tk_warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void __launch_bounds__(TOPK_THREADS)
cos_scores_kernel(const __nv_bfloat16* __restrict__ index, int r0, int r1, int D,
                  const __nv_bfloat16* __restrict__ queries, int nq, float eps, float* __restrict__ scores /* [nq, ld] */,
                  size_t ld) {
    extern __shared__ __align__(16) uint8_t tk_smem[];
    uint4* sQ = reinterpret_cast<uint4*>(tk_smem);                                    // [nq][D/8] bf16 chunks
    float* sQn = reinterpret_cast<float*>(tk_smem + (size_t)TOPK_QB * D * 2);         // [TOPK_QB] 1/max(|q|,eps)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int chunks = D >> 3;
    for (int i = tid; i < nq * chunks; i += TOPK_THREADS)
        sQ[i] = reinterpret_cast<const uint4*>(queries)[i];
    __syncthreads();
    if (warp < nq) {
        float acc = 0.f;
        for (int c = lane; c < chunks; c += 32) {
            float f[8];
            tk_unpack8(sQ[warp * chunks + c], f);
            acc = tk_sq8(acc, f);
        }
        acc = tk_warp_sum(acc);
        if (lane == 0) sQn[warp] = 1.0f / fmaxf(sqrtf(acc), eps);
    }
    __syncthreads();

    const int warps_total = gridDim.x * (TOPK_THREADS / 32);
    for (int row = r0 + blockIdx.x * (TOPK_THREADS / 32) + warp; row < r1; row += warps_total) {
        const uint4* er = reinterpret_cast<const uint4*>(index + (size_t)row * D);
        float dot[TOPK_QB];
#pragma unroll
        for (int q = 0; q < TOPK_QB; ++q) dot[q] = 0.f;
        float nrm = 0.f;
        // four 16-byte loads in flight per lane (the arithmetic below consumes them in the same order as a plain loop)
        for (int c0 = lane; c0 < chunks; c0 += 128) {
            uint4 u[4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
                u[i] = (c0 + 32 * i < chunks) ? __ldg(er + c0 + 32 * i) : make_uint4(0, 0, 0, 0);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int c = c0 + 32 * i;
                if (c < chunks) {
                    float e[8];
                    tk_unpack8(u[i], e);
                    nrm = tk_sq8(nrm, e);
#pragma unroll
                    for (int q = 0; q < TOPK_QB; ++q)
                        if (q < nq) dot[q] = tk_dot8(dot[q], e, sQ[q * chunks + c]);
                }
            }
        }
        nrm = tk_warp_sum(nrm);
        const float inv_e = 1.0f / fmaxf(sqrtf(nrm), eps);
#pragma unroll
        for (int q = 0; q < TOPK_QB; ++q) {
            if (q < nq) {
                const float d = tk_warp_sum(dot[q]);
                if (lane == 0) scores[(size_t)q * ld + row] = d * sQn[q] * inv_e;
            }
        }
    }
}

// fp32 inputs (find_similar_regions / compute_similarity on fp32 embeddings, e.g. VideoMAE clip vectors or fp32
// projections — the reference upcasts with .float(), src/perception/siglip_semantic_encoder.py:604-638): the same
// warp-per-row scan with fp32 rows and queries, 4 queries per pass held in shared memory.
constexpr int TOPK_QB_F32 = 4;

__global__ void __launch_bounds__(TOPK_THREADS)
cos_scores_f32_kernel(const float* __restrict__ index, int N, int D, const float* __restrict__ queries, int nq, float eps,
                      float* __restrict__ scores /* [nq, ld] */, size_t ld) {
    extern __shared__ __align__(16) uint8_t tk_smem[];
    float4* sQ = reinterpret_cast<float4*>(tk_smem);                                       // [nq][D/4]
    float* sQn = reinterpret_cast<float*>(tk_smem + (size_t)TOPK_QB_F32 * D * 4);          // [TOPK_QB_F32]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int chunks = D >> 2;
    for (int i = tid; i < nq * chunks; i += TOPK_THREADS) sQ[i] = reinterpret_cast<const float4*>(queries)[i];
    __syncthreads();
    if (warp < nq) {
        float acc = 0.f;
        for (int c = lane; c < chunks; c += 32) {
            const float4 f = sQ[warp * chunks + c];
            acc = fmaf(f.x, f.x, acc); acc = fmaf(f.y, f.y, acc); acc = fmaf(f.z, f.z, acc); acc = fmaf(f.w, f.w, acc);
        }
        acc = tk_warp_sum(acc);
        if (lane == 0) sQn[warp] = 1.0f / fmaxf(sqrtf(acc), eps);
    }
    __syncthreads();
    const int warps_total = gridDim.x * (TOPK_THREADS / 32);
    for (int row = blockIdx.x * (TOPK_THREADS / 32) + warp; row < N; row += warps_total) {
        const float4* er = reinterpret_cast<const float4*>(index + (size_t)row * D);
        float dot[TOPK_QB_F32];
#pragma unroll
        for (int q = 0; q < TOPK_QB_F32; ++q) dot[q] = 0.f;
        float nrm = 0.f;
        for (int c = lane; c < chunks; c += 32) {
            const float4 e = __ldg(er + c);
            nrm = fmaf(e.x, e.x, nrm); nrm = fmaf(e.y, e.y, nrm); nrm = fmaf(e.z, e.z, nrm); nrm = fmaf(e.w, e.w, nrm);
#pragma unroll
            for (int q = 0; q < TOPK_QB_F32; ++q)
                if (q < nq) {
                    const float4 v = sQ[q * chunks + c];
                    dot[q] = fmaf(e.x, v.x, dot[q]); dot[q] = fmaf(e.y, v.y, dot[q]);
                    dot[q] = fmaf(e.z, v.z, dot[q]); dot[q] = fmaf(e.w, v.w, dot[q]);
                }
        }
        nrm = tk_warp_sum(nrm);
        const float inv_e = 1.0f / fmaxf(sqrtf(nrm), eps);
#pragma unroll
        for (int q = 0; q < TOPK_QB_F32; ++q)
            if (q < nq) {
                const float d = tk_warp_sum(dot[q]);
                if (lane == 0) scores[(size_t)q * ld + row] = d * sQn[q] * inv_e;
            }
    }
}

// (score, idx) a "better" than b under (score desc, idx asc)
__device__ __forceinline__ bool tk_better(float sa, int ia, float sb, int ib) {
    return sa > sb || (sa == sb && ia < ib);
}

// One round-based arg-max selection step shared by the kernels below: the block agrees on the best (score, index)
// among the per-thread bests under (score desc, index asc).
struct TkBlockBest {
    float ws[32];
    int wi[32];
    float best_s;
    int best_i;
};
__device__ __forceinline__ void tk_block_argmax(TkBlockBest& sh, float bs, int bi, float& out_s, int& out_i) {
    const int n_warps = (int)(blockDim.x >> 5);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float os = __shfl_xor_sync(0xffffffffu, bs, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (tk_better(os, oi, bs, bi)) {
            bs = os;
            bi = oi;
        }
    }
    if (lane == 0) {
        sh.ws[warp] = bs;
        sh.wi[warp] = bi;
    }
    __syncthreads();
    if (tid == 0) {
        float fs = sh.ws[0];
        int fi = sh.wi[0];
        for (int w = 1; w < n_warps; ++w)
            if (tk_better(sh.ws[w], sh.wi[w], fs, fi)) {
                fs = sh.ws[w];
                fi = sh.wi[w];
            }
        sh.best_s = fs;
        sh.best_i = fi;
    }
    __syncthreads();
    out_s = sh.best_s;
    out_i = sh.best_i;
}

// Selection, stage 1: the scored span is cut into gridDim.x segments; block (segment, query) stages its piece of the
// query's score row in shared memory and selects the segment's k best by k rounds of a block-wide arg-max over the
// elements that come after the previous winner in the total order — no mutation, no sort network, deterministic.
// (One block per query over the whole row — the first version — left 128 blocks crawling over 72 000 scores 16 times:
// 0.9 ms of the 1.1 ms retrieval.)
constexpr int TOPK_SEG = 4096;      // scores per segment (16 KB of shared memory)
constexpr int TOPK_MAX_SEGS = 64;

__global__ void __launch_bounds__(TOPK_THREADS)
topk_select_seg_kernel(const float* __restrict__ scores, int N, size_t ld, int k, int span_lo, int seg_len,
                       const int32_t* __restrict__ row_lo, const int32_t* __restrict__ row_hi,
                       float* __restrict__ cand_s, int32_t* __restrict__ cand_i) {
    extern __shared__ float tk_seg[];  // [seg_len] when it fits, else unused
    __shared__ TkBlockBest sh;
    const int tid = threadIdx.x;
    const int qi = blockIdx.y, seg = blockIdx.x;
    const float* sc = scores + (size_t)qi * ld;
    const int qlo = row_lo ? max(0, row_lo[qi]) : 0;
    const int qhi = row_hi ? min(N, row_hi[qi]) : N;
    const int lo = max(qlo, span_lo + seg * seg_len);
    const int hi = min(qhi, span_lo + (seg + 1) * seg_len);
    const bool staged = seg_len <= TOPK_SEG;
    if (staged)
        for (int t = lo + tid; t < hi; t += TOPK_THREADS) tk_seg[t - lo] = sc[t];
    __syncthreads();
    float* os = cand_s + ((size_t)qi * gridDim.x + seg) * k;
    int32_t* oi = cand_i + ((size_t)qi * gridDim.x + seg) * k;
    float prev_s = INFINITY;
    int prev_i = -1;
    for (int r = 0; r < k; ++r) {
        float bs = -INFINITY;
        int bi = 0x7fffffff;
        for (int t = lo + tid; t < hi; t += TOPK_THREADS) {
            const float v = staged ? tk_seg[t - lo] : sc[t];
            // eligible = strictly after the previous winner in the total order (NaN never is)
            const bool elig = (r == 0) ? (v == v) : (v < prev_s || (v == prev_s && t > prev_i));
            if (elig && tk_better(v, t, bs, bi)) {
                bs = v;
                bi = t;
            }
        }
        tk_block_argmax(sh, bs, bi, prev_s, prev_i);
        if (tid == 0) {
            os[r] = prev_i == 0x7fffffff ? -INFINITY : prev_s;
            oi[r] = prev_i == 0x7fffffff ? -1 : prev_i;
        }
        if (prev_i == 0x7fffffff) {  // segment exhausted: the remaining slots are empty
            for (int r2 = r + 1 + tid; r2 < k; r2 += TOPK_THREADS) {
                os[r2] = -INFINITY;
                oi[r2] = -1;
            }
            break;
        }
    }
}

// Selection, stage 2: the k best of the segs x k segment winners of one query.
__global__ void __launch_bounds__(TOPK_THREADS)
topk_merge_kernel(const float* __restrict__ cand_s, const int32_t* __restrict__ cand_i, int segs, int k,
                  float* __restrict__ out_scores, int32_t* __restrict__ out_idx) {
    __shared__ TkBlockBest sh;
    const int tid = threadIdx.x, qi = blockIdx.x;
    const float* cs = cand_s + (size_t)qi * segs * k;
    const int32_t* ci = cand_i + (size_t)qi * segs * k;
    const int n = segs * k;
    float prev_s = INFINITY;
    int prev_i = -1;
    for (int r = 0; r < k; ++r) {
        float bs = -INFINITY;
        int bi = 0x7fffffff;
        for (int j = tid; j < n; j += TOPK_THREADS) {
            const int t = ci[j];
            const float v = cs[j];
            const bool elig = t >= 0 && ((r == 0) || (v < prev_s || (v == prev_s && t > prev_i)));
            if (elig && tk_better(v, t, bs, bi)) {
                bs = v;
                bi = t;
            }
        }
        tk_block_argmax(sh, bs, bi, prev_s, prev_i);
        if (tid == 0) {
            out_scores[(size_t)qi * k + r] = prev_i == 0x7fffffff ? -INFINITY : prev_s;
            out_idx[(size_t)qi * k + r] = prev_i == 0x7fffffff ? -1 : prev_i;
        }
    }
}

// Final stage of BOTH scoring paths: exact re-ranking of the near-top candidates.
//
// Approximate scores (fp32 scan: accumulation rounding ~1e-6; tcgen05 scores: ~1e-5) decide WHICH rows can be among
// the k best; their ORDER is decided here in float64: every candidate whose approximate score is within `margin` of
// the provisional k-th best is re-scored as  <q, e> / (max(|q|, eps) max(|e|, eps))  with float64 dot products and
// norms (bf16 inputs are exact in float64, products exact, 4096-term sums good to ~1e-15) and the k best under
// (score descending, row ascending) are returned, scores rounded to fp32.  The result is the float64 ranking — what
// `cos_sim` + a stable `argsort(descending)` give on float64 data — and it is the same for the scan and the tensor path.
//
// Candidates of query qi: cand_s / cand_i [qi][lists][list_len] in any order (row -1 = empty slot); `raw` = scores
// still lack the query's 1/|q| factor (tensor path).  A list that is full and whose worst entry is still within the
// margin may have dropped rows that matter (hundreds of near-duplicates of the query inside one list's span), and
// more than TOPK_CAND candidates do not fit the re-scoring buffer: both are counted in *overflow; the query then keeps
// the ranking by approximate scores.
__device__ __forceinline__ double tk_warp_sum_f64(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// order-preserving float <-> int key (signed integer comparison == float comparison, NaN excluded by the caller)
__device__ __forceinline__ int tk_order_key(float v) {
    const int i = __float_as_int(v);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float tk_order_unkey(int k) { return __int_as_float(k >= 0 ? k : k ^ 0x7fffffff); }
__device__ __forceinline__ bool tk_better_f64(double sa, int ia, double sb, int ib) {
    return sa > sb || (sa == sb && ia < ib);
}

__global__ void __launch_bounds__(TOPK_REFINE_THREADS)
topk_refine_kernel(const __nv_bfloat16* __restrict__ index, int D, const __nv_bfloat16* __restrict__ queries, float eps,
                   const float* __restrict__ cand_s, const int32_t* __restrict__ cand_i, int lists, int list_len, int k,
                   float margin, int raw, float* __restrict__ out_scores, int32_t* __restrict__ out_idx,
                   int* __restrict__ overflow) {
    __shared__ int s_cand[TOPK_CAND];
    __shared__ double s_exact[TOPK_CAND];
    __shared__ float s_vs[TOPK_REFINE_STAGE];   // the query's non-empty candidates, staged once (the k selection rounds
    __shared__ int s_vi[TOPK_REFINE_STAGE];     // below would otherwise re-read the sparse global lists k times)
    __shared__ int s_lmax[TOPK_MAX_LISTS];      // per-list maximum score (order-preserving integer key)
    __shared__ int s_n, s_trunc, s_nv, s_valid;
    __shared__ unsigned s_thr0, s_kth;
    __shared__ double s_qn;
    __shared__ TkBlockBest sh;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int qi = blockIdx.x;
    int L = lists * list_len;
    const float* cs = cand_s + (size_t)qi * L;
    const int32_t* ci = cand_i + (size_t)qi * L;
    const float* gcs = cs;   // the list structure (for the truncation check) stays in global memory
    const int32_t* gci = ci;
    const int chunks = D >> 3;
    const uint4* qr = reinterpret_cast<const uint4*>(queries + (size_t)qi * D);
    // |q| in float64
    if (warp == 0) {
        double acc = 0.0;
        for (int c = lane; c < chunks; c += 32) {
            float f[8];
            tk_unpack8(__ldg(qr + c), f);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc += (double)f[j] * (double)f[j];
        }
        acc = tk_warp_sum_f64(acc);
        if (lane == 0) s_qn = fmax(sqrt(acc), (double)eps);
    }
    if (tid == 0) s_n = 0, s_trunc = 0, s_nv = 0, s_thr0 = 0xff800000u;  // -inf
    for (int j = tid; j < lists && j < TOPK_MAX_LISTS; j += TOPK_REFINE_THREADS) s_lmax[j] = tk_order_key(-INFINITY);
    __syncthreads();
    const float iq = raw ? (float)(1.0 / s_qn) : 1.0f;
    const float margin_raw = raw ? margin / iq : margin;  // the margin is in cosine units
    // Pre-filter (several lists): the k-th largest of the per-list maxima is reached by k distinct rows, so it is a
    // lower bound of the provisional k-th best score; rows below it minus the margin cannot matter.
    float pre = -INFINITY;
    if (lists >= 2 && lists >= k && lists <= TOPK_MAX_LISTS) {
        for (int j = tid; j < L; j += TOPK_REFINE_THREADS) {
            const float v = cs[j];
            if (ci[j] >= 0 && v == v) atomicMax(&s_lmax[j / list_len], tk_order_key(v));
        }
        __syncthreads();
        for (int t = tid; t < lists; t += TOPK_REFINE_THREADS) {
            const int mine = s_lmax[t];
            int rank = 0;
            for (int j = 0; j < lists; ++j) {
                const int o = s_lmax[j];
                rank += (o > mine || (o == mine && j < t)) ? 1 : 0;
            }
            if (rank == k - 1) s_thr0 = __float_as_uint(tk_order_unkey(mine));
        }
        __syncthreads();
        pre = __uint_as_float(s_thr0) - margin_raw;  // (-inf when fewer than k lists hold a row)
    }
    for (int j = tid; j < L; j += TOPK_REFINE_THREADS) {
        const int t = ci[j];
        if (t >= 0) {
            const float v = cs[j];
            if (!(v < pre)) {  // keeps NaN scores, like the unfiltered path
                const int slot = atomicAdd(&s_nv, 1);
                if (slot < TOPK_REFINE_STAGE) {
                    s_vs[slot] = v;
                    s_vi[slot] = t;
                }
            }
        }
    }
    __syncthreads();
    const bool staged = s_nv <= TOPK_REFINE_STAGE;
    if (staged) {  // (else: more candidates than the stage holds — work on the global lists)
        cs = s_vs;
        ci = s_vi;
        L = s_nv;
    }
    // provisional top-k by approximate score (also the fall-back result)
    float kth = -INFINITY;
    int found = 0;
    if (staged && L <= TOPK_RANK_CAP) {
        // few candidates: every thread ranks its own among all of them — no block-wide rounds
        if (tid == 0) s_kth = __float_as_uint(-INFINITY), s_valid = 0;
        __syncthreads();
        for (int j = tid; j < L; j += TOPK_REFINE_THREADS) {
            const float v = cs[j];
            const int t = ci[j];
            if (!(v == v)) continue;
            int rank = 0;
            for (int i = 0; i < L; ++i) {
                const float o = cs[i];
                rank += (o == o && tk_better(o, ci[i], v, t)) ? 1 : 0;
            }
            atomicAdd(&s_valid, 1);
            if (rank < k) {
                out_scores[(size_t)qi * k + rank] = v * iq;
                out_idx[(size_t)qi * k + rank] = t;
                if (rank == k - 1) s_kth = __float_as_uint(v);
            }
        }
        __syncthreads();
        found = s_valid < k ? s_valid : k;
        kth = __uint_as_float(s_kth);
    } else {
        // k rounds of arg-max over the candidates
        float prev_s = INFINITY;
        int prev_i = -1;
        for (int r = 0; r < k; ++r) {
            float bs = -INFINITY;
            int bi = 0x7fffffff;
            for (int j = tid; j < L; j += TOPK_REFINE_THREADS) {
                const int t = ci[j];
                const float v = cs[j];
                const bool elig = t >= 0 && ((r == 0) ? (v == v) : (v < prev_s || (v == prev_s && t > prev_i)));
                if (elig && tk_better(v, t, bs, bi)) {
                    bs = v;
                    bi = t;
                }
            }
            tk_block_argmax(sh, bs, bi, prev_s, prev_i);
            if (prev_i == 0x7fffffff) break;  // fewer than k candidates (block-uniform)
            if (tid == 0) {
                out_scores[(size_t)qi * k + r] = prev_s * iq;
                out_idx[(size_t)qi * k + r] = prev_i;
            }
            kth = prev_s;
            found = r + 1;
        }
    }
    if (found < k) {
        for (int r2 = found + tid; r2 < k; r2 += TOPK_REFINE_THREADS) {
            out_scores[(size_t)qi * k + r2] = -INFINITY;
            out_idx[(size_t)qi * k + r2] = -1;
        }
        kth = -INFINITY;  // every candidate is re-scored
    }
    // candidates within the margin of the provisional k-th score (margin is in cosine units)
    const float thr = kth - margin_raw;
    for (int j = tid; j < L; j += TOPK_REFINE_THREADS) {
        const int t = ci[j];
        if (t >= 0 && cs[j] >= thr) {
            const int slot = atomicAdd(&s_n, 1);
            if (slot < TOPK_CAND) s_cand[slot] = t;
        }
    }
    // a full list whose worst entry is still inside the margin may have dropped qualifying rows
    for (int li_ = tid; li_ < lists; li_ += TOPK_REFINE_THREADS) {
        float worst = INFINITY;
        bool full = gci[li_ * list_len + list_len - 1] >= 0;  // lists fill front to back
        for (int e = 0; full && e < list_len; ++e) {
            if (gci[li_ * list_len + e] < 0) full = false;
            else worst = fminf(worst, gcs[li_ * list_len + e]);
        }
        if (full && worst >= thr && lists * list_len > k) atomicOr(&s_trunc, 1);
    }
    __syncthreads();
    const int n = s_n;
    if (n > TOPK_CAND || s_trunc) {
        if (tid == 0 && overflow) atomicAdd(overflow, 1);
        if (n > TOPK_CAND) return;  // provisional (approximate-score) result stays
    }
    // exact float64 scores: one warp per candidate
    for (int j = warp; j < n; j += TOPK_REFINE_THREADS / 32) {
        const uint4* er = reinterpret_cast<const uint4*>(index + (size_t)s_cand[j] * D);
        double dot = 0.0, nrm = 0.0;
        for (int c = lane; c < chunks; c += 32) {
            float e[8], f[8];
            tk_unpack8(__ldg(er + c), e);
            tk_unpack8(__ldg(qr + c), f);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                nrm += (double)e[u] * (double)e[u];
                dot += (double)e[u] * (double)f[u];
            }
        }
        nrm = tk_warp_sum_f64(nrm);
        dot = tk_warp_sum_f64(dot);
        if (lane == 0) s_exact[j] = dot / (s_qn * fmax(sqrt(nrm), (double)eps));
    }
    __syncthreads();
    // final selection under (score desc, row asc) on the float64 scores: n <= TOPK_CAND, every thread ranks its own
    // candidates; slots beyond the number of non-NaN scores keep the provisional entries
    for (int j = tid; j < n; j += TOPK_REFINE_THREADS) {
        const double v = s_exact[j];
        const int t = s_cand[j];
        if (!(v == v)) continue;
        int rank = 0;
        for (int i = 0; i < n; ++i) {
            const double o = s_exact[i];
            rank += (o == o && tk_better_f64(o, s_cand[i], v, t)) ? 1 : 0;
        }
        if (rank < k) {
            out_scores[(size_t)qi * k + rank] = (float)v;
            out_idx[(size_t)qi * k + rank] = t;
        }
    }
}

// 1 / max(|row|, eps) of bf16 rows: one warp per row (index rows or queries)
__global__ void __launch_bounds__(TOPK_THREADS)
row_inv_norm_kernel(const __nv_bfloat16* __restrict__ x, int rows, int D, float eps, float* __restrict__ inv) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int chunks = D >> 3;
    const int warps_total = gridDim.x * (TOPK_THREADS / 32);
    for (int row = blockIdx.x * (TOPK_THREADS / 32) + warp; row < rows; row += warps_total) {
        const uint4* er = reinterpret_cast<const uint4*>(x + (size_t)row * D);
        float nrm = 0.f;
        for (int c = lane; c < chunks; c += 32) {
            float e[8];
            tk_unpack8(__ldg(er + c), e);
            nrm = tk_sq8(nrm, e);
        }
        nrm = tk_warp_sum(nrm);
        if (lane == 0) inv[row] = 1.0f / fmaxf(sqrtf(nrm), eps);
    }
}

int launch_topk_fused(const void* index, int N, int D, const void* queries, int nq, int span_lo, int span_hi,
                      const float* inv_e, const float* inv_q, float margin, int k, const int32_t* row_lo,
                      const int32_t* row_hi, float* cand_s, int32_t* cand_i, int* grid_out, cudaStream_t s);  // topk_fused.cu
int topk_fused_list_len();
int topk_fused_grid(int span);  // CTAs (= candidate lists per query) the fused kernel runs for a span of rows

}  // namespace gvl

// scratch layout (floats), ld = max(roundup4(N), TOPK_MIN_LD):
//   [Q][ld]  fp32 scores of the scan path  /  candidate lists of the fused tensor path ([Q][CTAs][32] scores, then rows)
//   [ld] 1/|e_n| | [roundup4(Q)] reserved | 4: overflow counter (int) |
//   [Q][TOPK_MAX_SEGS][64] segment-winner scores | same, int32 rows | [Q][64] pre-selected scores | same, int32 rows
constexpr size_t TOPK_MIN_LD = 2 * 160 * 64;  // room for the fused path's lists of up to 160 CTAs
static size_t topk_ld(int N) {
    const size_t ld = ((size_t)N + 3) & ~(size_t)3;
    return ld < TOPK_MIN_LD ? TOPK_MIN_LD : ld;
}
static size_t topk_cand_offset(int N, int Q) {
    const size_t ld = topk_ld(N);
    return (size_t)Q * ld + ld + (((size_t)Q + 3) & ~(size_t)3) + 4;
}
static size_t topk_presel_offset(int N, int Q) { return topk_cand_offset(N, Q) + 2 * (size_t)Q * gvl::TOPK_MAX_SEGS * 64; }

// two-stage selection over scratch[Q][ld] scores: segment winners (grid = segments x queries), then their merge
static int topk_select(float* scratch, int N, int Q, int k, int span_lo, int span, const int32_t* row_lo,
                       const int32_t* row_hi, float* out_scores, int32_t* out_idx, cudaStream_t s) {
    using namespace gvl;
    const size_t ld = topk_ld(N);
    int segs = (span + TOPK_SEG - 1) / TOPK_SEG;
    segs = segs < 1 ? 1 : (segs > TOPK_MAX_SEGS ? TOPK_MAX_SEGS : segs);
    int seg_len = ((span + segs - 1) / segs + 3) & ~3;
    if (seg_len < 4) seg_len = 4;
    float* cand_s = scratch + topk_cand_offset(N, Q);
    int32_t* cand_i = reinterpret_cast<int32_t*>(cand_s + (size_t)Q * TOPK_MAX_SEGS * 64);
    const size_t seg_smem = seg_len <= TOPK_SEG ? (size_t)seg_len * sizeof(float) : 0;
    ProfScope prof(GVL_K_TOPK_SELECT, (double)Q * span * 4, s);
    topk_select_seg_kernel<<<dim3((unsigned)segs, (unsigned)Q), TOPK_THREADS, seg_smem, s>>>(
        scratch, N, ld, k, span_lo, seg_len, row_lo, row_hi, cand_s, cand_i);
    GVL_LAUNCH_CHECK("topk_select_seg_kernel");
    topk_merge_kernel<<<Q, TOPK_THREADS, 0, s>>>(cand_s, cand_i, segs, k, out_scores, out_idx);
    GVL_LAUNCH_CHECK("topk_merge_kernel");
    return 0;
}

extern "C" size_t gvl_topk_scratch_floats(int N, int Q) {
    if (N <= 0 || Q <= 0) return 0;
    return topk_presel_offset(N, Q) + 2 * (size_t)Q * 64;
}

extern "C" int gvl_row_inv_norm(const void* rows, int N, int D, float eps, float* inv_norm, void* stream) {
    using namespace gvl;
    GVL_CHECK_ARG(rows && inv_norm, "gvl_row_inv_norm: null pointer");
    GVL_CHECK_ARG(N > 0 && D > 0 && D % 8 == 0 && (uintptr_t)rows % 16 == 0, "gvl_row_inv_norm: bad shape N=%d D=%d", N, D);
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    int grid = (N + TOPK_THREADS / 32 - 1) / (TOPK_THREADS / 32);
    if (grid > sm_count() * 8) grid = sm_count() * 8;
    ProfScope prof(GVL_K_TOPK_SCORES, (double)N * D * 2, s);
    row_inv_norm_kernel<<<grid, TOPK_THREADS, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(rows), N, D, eps, inv_norm);
    GVL_LAUNCH_CHECK("row_inv_norm_kernel");
    return 0;
}

extern "C" int gvl_topk_cosine_ex(const void* index, int N, int D, const void* queries, int Q, int k, float eps,
                                  const int32_t* row_lo, const int32_t* row_hi, int span_lo, int span_hi, int mode,
                                  const float* inv_norm, float* scratch, float* out_scores, int32_t* out_idx,
                                  void* stream) {
    using namespace gvl;
    GVL_CHECK_ARG(index && queries && scratch && out_scores && out_idx, "gvl_topk_cosine: null pointer");
    GVL_CHECK_ARG(N > 0 && Q > 0 && D > 0 && D % 8 == 0 && D <= 8192, "gvl_topk_cosine: bad shape N=%d Q=%d D=%d", N, Q, D);
    GVL_CHECK_ARG(k > 0 && k <= 64, "gvl_topk_cosine: k=%d out of range [1,64]", k);
    GVL_CHECK_ARG((uintptr_t)index % 16 == 0 && (uintptr_t)queries % 16 == 0 && (uintptr_t)scratch % 16 == 0,
                  "gvl_topk_cosine: misaligned pointer");
    GVL_CHECK_ARG((row_lo == nullptr) == (row_hi == nullptr), "gvl_topk_cosine: row_lo and row_hi come together");
    GVL_CHECK_ARG(mode >= GVL_TOPK_AUTO && mode <= GVL_TOPK_TENSOR, "gvl_topk_cosine: bad mode %d", mode);
    // [span_lo, span_hi): rows any query may rank (the union of the per-query ranges, known to the host that built them)
    if (row_lo == nullptr) span_lo = 0, span_hi = N;
    GVL_CHECK_ARG(span_lo >= 0 && span_hi <= N && span_lo <= span_hi, "gvl_topk_cosine: bad span [%d, %d)", span_lo, span_hi);
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const size_t ld = topk_ld(N);
    const int rows_per_cta = TOPK_THREADS / 32;
    const int max_grid = sm_count() * 8;
    const int span = span_hi - span_lo;
    if (mode == GVL_TOPK_AUTO) mode = (Q > TOPK_QB && span >= 4096 && k <= 16) ? GVL_TOPK_TENSOR : GVL_TOPK_SCAN;
    const __nv_bfloat16* idx_bf = reinterpret_cast<const __nv_bfloat16*>(index);
    const __nv_bfloat16* q_bf = reinterpret_cast<const __nv_bfloat16*>(queries);
    int* overflow = reinterpret_cast<int*>(scratch + (size_t)Q * ld + ld + (((size_t)Q + 3) & ~(size_t)3));
    GVL_CUDA(cudaMemsetAsync(overflow, 0, sizeof(int), s));
    if (span <= 0) {  // nothing is eligible: every slot empty
        ProfScope prof(GVL_K_TOPK_SELECT, 0.0, s);
        topk_refine_kernel<<<Q, TOPK_REFINE_THREADS, 0, s>>>(idx_bf, D, q_bf, eps, scratch, reinterpret_cast<int32_t*>(scratch), 0, 0, k,
                                                      0.f, 0, out_scores, out_idx, overflow);
        GVL_LAUNCH_CHECK("topk_refine_kernel");
        return 0;
    }

    if (mode == GVL_TOPK_TENSOR) {
        // fused tcgen05 scoring + per-CTA candidate lists (topk_fused.cu), batches of up to 128 queries
        float* inv_e = scratch + (size_t)Q * ld;
        if (inv_norm == nullptr) {
            int rc = gvl_row_inv_norm(idx_bf + (size_t)span_lo * D, span, D, eps, inv_e + span_lo, stream);
            if (rc) return rc;
        }
        const int KL = topk_fused_list_len();
        GVL_CHECK_ARG(k <= 16, "gvl_topk_cosine: the tensor path serves k <= 16 (k = %d): use GVL_TOPK_SCAN", k);
        float* inv_q = inv_e + ld;  // [roundup4(Q)]
        {
            int rc = gvl_row_inv_norm(queries, Q, D, eps, inv_q, stream);
            if (rc) return rc;
        }
        for (int q0 = 0; q0 < Q; q0 += 128) {
            const int nq = Q - q0 < 128 ? Q - q0 : 128;
            // lists of this batch live in the batch's own [nq][ld] slice of the score region
            float* cand_s = scratch + (size_t)q0 * ld;
            int grid = 0;
            // cand_i directly behind cand_s: nq * grid * KL floats each, grid <= 160 (TOPK_MIN_LD guarantees the room)
            const int g = topk_fused_grid(span);
            GVL_CHECK_ARG((size_t)2 * g * KL <= ld, "gvl_topk_cosine: %d CTAs exceed the scratch layout", g);
            int32_t* cand_i = reinterpret_cast<int32_t*>(cand_s + (size_t)nq * g * KL);
            int rc = launch_topk_fused(index, N, D, q_bf + (size_t)q0 * D, nq, span_lo, span_hi, inv_norm ? inv_norm : inv_e,
                                       inv_q + q0, TOPK_MARGIN, k, row_lo ? row_lo + q0 : nullptr,
                                       row_hi ? row_hi + q0 : nullptr, cand_s, cand_i, &grid, s);
            if (rc) return rc;
            ProfScope prof(GVL_K_TOPK_SELECT, (double)nq * grid * KL * 8, s);
            topk_refine_kernel<<<nq, TOPK_REFINE_THREADS, 0, s>>>(idx_bf, D, q_bf + (size_t)q0 * D, eps, cand_s, cand_i, grid, KL, k,
                                                           TOPK_MARGIN, 1, out_scores + (size_t)q0 * k, out_idx + (size_t)q0 * k,
                                                           overflow);
            GVL_LAUNCH_CHECK("topk_refine_kernel");
        }
        return 0;
    }

    // ---- scan path: fp32 scores of every row, pre-selection of the best k + 16, float64 refinement ----
    {
        const size_t smem = (size_t)TOPK_QB * D * 2 + TOPK_QB * sizeof(float);
        GVL_CUDA(cudaFuncSetAttribute(cos_scores_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int grid = (span + rows_per_cta - 1) / rows_per_cta;
        if (grid > max_grid) grid = max_grid;
        for (int q0 = 0; q0 < Q; q0 += TOPK_QB) {
            const int nq = Q - q0 < TOPK_QB ? Q - q0 : TOPK_QB;
            ProfScope prof(GVL_K_TOPK_SCORES, (double)span * D * 2, s);
            cos_scores_kernel<<<grid, TOPK_THREADS, smem, s>>>(idx_bf, span_lo, span_hi, D, q_bf + (size_t)q0 * D, nq, eps,
                                                               scratch + (size_t)q0 * ld, ld);
            GVL_LAUNCH_CHECK("cos_scores_kernel");
        }
    }
    const int kpre = k + 16 < 64 ? k + 16 : 64;
    float* pre_s = scratch + topk_presel_offset(N, Q);
    int32_t* pre_i = reinterpret_cast<int32_t*>(pre_s + (size_t)Q * 64);
    {
        int rc = topk_select(scratch, N, Q, kpre, span_lo, span, row_lo, row_hi, pre_s, pre_i, s);
        if (rc) return rc;
    }
    {
        ProfScope prof(GVL_K_TOPK_SELECT, (double)Q * kpre * 8, s);
        topk_refine_kernel<<<Q, TOPK_REFINE_THREADS, 0, s>>>(idx_bf, D, q_bf, eps, pre_s, pre_i, 1, kpre, k, TOPK_MARGIN_SCAN, 0,
                                                      out_scores, out_idx, overflow);
        GVL_LAUNCH_CHECK("topk_refine_kernel");
    }
    return 0;
}

extern "C" int gvl_topk_cosine(const void* index, int N, int D, const void* queries, int Q, int k, float eps,
                               float* scratch, float* out_scores, int32_t* out_idx, void* stream) {
    // scratch: gvl_topk_scratch_floats(N, Q) floats
    return gvl_topk_cosine_ex(index, N, D, queries, Q, k, eps, nullptr, nullptr, 0, N, GVL_TOPK_AUTO, nullptr, scratch,
                              out_scores, out_idx, stream);
}

extern "C" int gvl_topk_cosine_f32(const float* index, int N, int D, const float* queries, int Q, int k, float eps,
                                   float* scratch, float* out_scores, int32_t* out_idx, void* stream) {
    using namespace gvl;
    GVL_CHECK_ARG(index && queries && scratch && out_scores && out_idx, "gvl_topk_cosine_f32: null pointer");
    GVL_CHECK_ARG(N > 0 && Q > 0 && D > 0 && D % 4 == 0 && D <= 8192, "gvl_topk_cosine_f32: bad shape N=%d Q=%d D=%d", N, Q, D);
    GVL_CHECK_ARG(k > 0 && k <= 64, "gvl_topk_cosine_f32: k=%d out of range [1,64]", k);
    GVL_CHECK_ARG((uintptr_t)index % 16 == 0 && (uintptr_t)queries % 16 == 0 && (uintptr_t)scratch % 16 == 0,
                  "gvl_topk_cosine_f32: misaligned pointer");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const size_t ld = topk_ld(N);
    const size_t smem = (size_t)TOPK_QB_F32 * D * 4 + TOPK_QB_F32 * sizeof(float);
    GVL_CUDA(cudaFuncSetAttribute(cos_scores_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int grid = (N + TOPK_THREADS / 32 - 1) / (TOPK_THREADS / 32);
    if (grid > sm_count() * 4) grid = sm_count() * 4;
    for (int q0 = 0; q0 < Q; q0 += TOPK_QB_F32) {
        const int nq = Q - q0 < TOPK_QB_F32 ? Q - q0 : TOPK_QB_F32;
        ProfScope prof(GVL_K_TOPK_SCORES, (double)N * D * 4, s);
        cos_scores_f32_kernel<<<grid, TOPK_THREADS, smem, s>>>(index, N, D, queries + (size_t)q0 * D, nq, eps,
                                                               scratch + (size_t)q0 * ld, ld);
        GVL_LAUNCH_CHECK("cos_scores_f32_kernel");
    }
    return topk_select(scratch, N, Q, k, 0, N, nullptr, nullptr, out_scores, out_idx, s);
}
