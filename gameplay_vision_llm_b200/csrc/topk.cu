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
            const uint4 u = sQ[warp * chunks + c];
            const float f[8] = {bf16_lo(u.x), bf16_hi(u.x), bf16_lo(u.y), bf16_hi(u.y),
                                bf16_lo(u.z), bf16_hi(u.z), bf16_lo(u.w), bf16_hi(u.w)};
#pragma unroll
            for (int j = 0; j < 8; ++j) acc += f[j] * f[j];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
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
        for (int c = lane; c < chunks; c += 32) {
            const uint4 u = __ldg(er + c);
            const float e[8] = {bf16_lo(u.x), bf16_hi(u.x), bf16_lo(u.y), bf16_hi(u.y),
                                bf16_lo(u.z), bf16_hi(u.z), bf16_lo(u.w), bf16_hi(u.w)};
#pragma unroll
            for (int j = 0; j < 8; ++j) nrm += e[j] * e[j];
#pragma unroll
            for (int q = 0; q < TOPK_QB; ++q) {
                if (q < nq) {
                    const uint4 v = sQ[q * chunks + c];
                    dot[q] += e[0] * bf16_lo(v.x) + e[1] * bf16_hi(v.x) + e[2] * bf16_lo(v.y) + e[3] * bf16_hi(v.y) +
                              e[4] * bf16_lo(v.z) + e[5] * bf16_hi(v.z) + e[6] * bf16_lo(v.w) + e[7] * bf16_hi(v.w);
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) nrm += __shfl_xor_sync(0xffffffffu, nrm, o);
        const float inv_e = 1.0f / fmaxf(sqrtf(nrm), eps);
#pragma unroll
        for (int q = 0; q < TOPK_QB; ++q) {
            if (q < nq) {
                float d = dot[q];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
                if (lane == 0) scores[(size_t)q * ld + row] = d * sQn[q] * inv_e;
            }
        }
    }
}

// (score, idx) a "better" than b under (score desc, idx asc)
__device__ __forceinline__ bool tk_better(float sa, int ia, float sb, int ib) {
    return sa > sb || (sa == sb && ia < ib);
}

__global__ void __launch_bounds__(TOPK_THREADS)
topk_select_kernel(const float* __restrict__ scores, int N, size_t ld, int k, const int32_t* __restrict__ row_lo,
                   const int32_t* __restrict__ row_hi, float* __restrict__ out_scores, int32_t* __restrict__ out_idx) {
    __shared__ float s_s[TOPK_THREADS / 32];
    __shared__ int s_i[TOPK_THREADS / 32];
    __shared__ float s_best_s;
    __shared__ int s_best_i;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const float* sc = scores + (size_t)blockIdx.x * ld;
    const int lo = row_lo ? max(0, row_lo[blockIdx.x]) : 0;
    const int hi = row_hi ? min(N, row_hi[blockIdx.x]) : N;
    float prev_s = INFINITY;
    int prev_i = -1;
    for (int r = 0; r < k; ++r) {
        float bs = -INFINITY;
        int bi = 0x7fffffff;
        for (int t = lo + tid; t < hi; t += TOPK_THREADS) {
            const float v = sc[t];
            // eligible = strictly after the previous winner in the total order
            const bool elig = (r == 0) ? (v == v) : (v < prev_s || (v == prev_s && t > prev_i));
            if (elig && tk_better(v, t, bs, bi)) {
                bs = v;
                bi = t;
            }
        }
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
            s_s[warp] = bs;
            s_i[warp] = bi;
        }
        __syncthreads();
        if (tid == 0) {
            float fs = s_s[0];
            int fi = s_i[0];
            for (int w = 1; w < TOPK_THREADS / 32; ++w)
                if (tk_better(s_s[w], s_i[w], fs, fi)) {
                    fs = s_s[w];
                    fi = s_i[w];
                }
            s_best_s = fs;
            s_best_i = fi;
            out_scores[(size_t)blockIdx.x * k + r] = fi == 0x7fffffff ? -INFINITY : fs;
            out_idx[(size_t)blockIdx.x * k + r] = fi == 0x7fffffff ? -1 : fi;
        }
        __syncthreads();
        prev_s = s_best_s;
        prev_i = s_best_i;
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
            const uint4 u = __ldg(er + c);
            const float e[8] = {bf16_lo(u.x), bf16_hi(u.x), bf16_lo(u.y), bf16_hi(u.y),
                                bf16_lo(u.z), bf16_hi(u.z), bf16_lo(u.w), bf16_hi(u.w)};
#pragma unroll
            for (int j = 0; j < 8; ++j) nrm += e[j] * e[j];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) nrm += __shfl_xor_sync(0xffffffffu, nrm, o);
        if (lane == 0) inv[row] = 1.0f / fmaxf(sqrtf(nrm), eps);
    }
}

// raw dot products -> cosines, in place: scores[q][n] *= inv_q[q] * inv_e[n] for n in [0, n_cols)
__global__ void __launch_bounds__(TOPK_THREADS)
scale_scores_kernel(float* __restrict__ scores, size_t ld, int n_cols, const float* __restrict__ inv_q,
                    const float* __restrict__ inv_e) {
    float* row = scores + (size_t)blockIdx.y * ld;
    const float iq = inv_q[blockIdx.y];
    for (int n = (blockIdx.x * TOPK_THREADS + threadIdx.x) * 4; n < n_cols; n += gridDim.x * TOPK_THREADS * 4) {
        float4 v = *reinterpret_cast<float4*>(row + n);
        const float4 e = *reinterpret_cast<const float4*>(inv_e + n);
        v.x *= iq * e.x;
        v.y *= iq * e.y;
        v.z *= iq * e.z;
        v.w *= iq * e.w;
        *reinterpret_cast<float4*>(row + n) = v;
    }
}

}  // namespace gvl

extern "C" size_t gvl_topk_scratch_floats(int N, int Q) {
    if (N <= 0 || Q <= 0) return 0;
    const size_t ld = ((size_t)N + 3) & ~(size_t)3;
    return (size_t)Q * ld + ld /* 1/|e_n| */ + (((size_t)Q + 3) & ~(size_t)3) /* 1/|q| */;
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
    const size_t ld = ((size_t)N + 3) & ~(size_t)3;
    const int rows_per_cta = TOPK_THREADS / 32;
    const int max_grid = sm_count() * 8;
    const int span = span_hi - span_lo;
    if (mode == GVL_TOPK_AUTO) mode = (Q >= 16 && span >= 4096) ? GVL_TOPK_TENSOR : GVL_TOPK_SCAN;

    int scan_lo = span_lo, scan_hi = span_hi;  // rows scored by the CUDA-core scan
    if (mode == GVL_TOPK_TENSOR && span > 0) {
        // GEMM over rows [g0, g1): g0 rounded down to the GEMM's N granularity (8 rows; the extra rows are harmless),
        // g1 rounded down — the <= 7 rows left over go through the scan kernel
        const int g0 = span_lo & ~7, g1 = g0 + ((span_hi - g0) & ~7);
        float* inv_e = scratch + (size_t)Q * ld;
        float* inv_q = inv_e + ld;
        if (g1 > g0) {
            if (inv_norm == nullptr) {
                int rc = gvl_row_inv_norm(reinterpret_cast<const __nv_bfloat16*>(index) + (size_t)g0 * D, g1 - g0, D, eps,
                                          inv_e + g0, stream);
                if (rc) return rc;
            }
            int rc = gvl_row_inv_norm(queries, Q, D, eps, inv_q, stream);
            if (rc) return rc;
            rc = gvl_gemm_bf16(queries, D, reinterpret_cast<const __nv_bfloat16*>(index) + (size_t)g0 * D, D, nullptr, nullptr,
                               0, 0, scratch + g0, (int)ld, 1, Q, g1 - g0, D, GVL_ACT_NONE, stream);
            if (rc) return rc;
            dim3 grid((unsigned)std::min<size_t>(((size_t)(g1 - g0) / 4 + TOPK_THREADS - 1) / TOPK_THREADS, 64), (unsigned)Q);
            ProfScope prof(GVL_K_TOPK_SELECT, (double)Q * (g1 - g0) * 8, s);
            scale_scores_kernel<<<grid, TOPK_THREADS, 0, s>>>(scratch + g0, ld, g1 - g0, inv_q,
                                                              (inv_norm ? inv_norm : inv_e) + g0);
            GVL_LAUNCH_CHECK("scale_scores_kernel");
        }
        scan_lo = g1;
    }
    if (scan_hi > scan_lo) {
        const size_t smem = (size_t)TOPK_QB * D * 2 + TOPK_QB * sizeof(float);
        GVL_CUDA(cudaFuncSetAttribute(cos_scores_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int grid = (scan_hi - scan_lo + rows_per_cta - 1) / rows_per_cta;
        if (grid > max_grid) grid = max_grid;
        for (int q0 = 0; q0 < Q; q0 += TOPK_QB) {
            const int nq = Q - q0 < TOPK_QB ? Q - q0 : TOPK_QB;
            ProfScope prof(GVL_K_TOPK_SCORES, (double)(scan_hi - scan_lo) * D * 2, s);
            cos_scores_kernel<<<grid, TOPK_THREADS, smem, s>>>(
                reinterpret_cast<const __nv_bfloat16*>(index), scan_lo, scan_hi, D,
                reinterpret_cast<const __nv_bfloat16*>(queries) + (size_t)q0 * D, nq, eps, scratch + (size_t)q0 * ld, ld);
            GVL_LAUNCH_CHECK("cos_scores_kernel");
        }
    }
    ProfScope prof(GVL_K_TOPK_SELECT, (double)Q * span * 4 * k, s);
    topk_select_kernel<<<Q, TOPK_THREADS, 0, s>>>(scratch, N, ld, k, row_lo, row_hi, out_scores, out_idx);
    GVL_LAUNCH_CHECK("topk_select_kernel");
    return 0;
}

extern "C" int gvl_topk_cosine(const void* index, int N, int D, const void* queries, int Q, int k, float eps,
                               float* scratch, float* out_scores, int32_t* out_idx, void* stream) {
    // scratch: gvl_topk_scratch_floats(N, Q) floats
    return gvl_topk_cosine_ex(index, N, D, queries, Q, k, eps, nullptr, nullptr, 0, N, GVL_TOPK_AUTO, nullptr, scratch,
                              out_scores, out_idx, stream);
}
