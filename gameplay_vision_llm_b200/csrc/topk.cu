// topk.cu — K8: cosine top-k over the timeline index (HBM-bound scan + exact selection).
//
// Pass 1 streams the bf16 index once per batch of 8 queries: one warp per index row, 16-byte
// coalesced loads, fp32 dot products against the queries held in shared memory, the row norm computed
// from the same registers.  The per-lane accumulation order and the xor-shuffle reduction tree do not
// depend on the row's position, so identical rows get bit-identical scores (ties are real ties).
// Pass 2 selects the k best per query under the total order (score descending, index ascending) by k
// rounds of a block-wide arg-max over the elements that come after the previous winner in that order —
// no mutation, no sort network, deterministic.
#include "common.cuh"

namespace gvl {

constexpr int TOPK_QB = 8;        // queries per scan pass
constexpr int TOPK_THREADS = 256;

__global__ void __launch_bounds__(TOPK_THREADS)
cos_scores_kernel(const __nv_bfloat16* __restrict__ index, int N, int D, const __nv_bfloat16* __restrict__ queries,
                  int nq, float eps, float* __restrict__ scores /* [nq, N] */) {
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
    for (int row = blockIdx.x * (TOPK_THREADS / 32) + warp; row < N; row += warps_total) {
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
                if (lane == 0) scores[(size_t)q * N + row] = d * sQn[q] * inv_e;
            }
        }
    }
}

// (score, idx) a "better" than b under (score desc, idx asc)
__device__ __forceinline__ bool tk_better(float sa, int ia, float sb, int ib) {
    return sa > sb || (sa == sb && ia < ib);
}

__global__ void __launch_bounds__(TOPK_THREADS)
topk_select_kernel(const float* __restrict__ scores, int N, int k, float* __restrict__ out_scores,
                   int32_t* __restrict__ out_idx) {
    __shared__ float s_s[TOPK_THREADS / 32];
    __shared__ int s_i[TOPK_THREADS / 32];
    __shared__ float s_best_s;
    __shared__ int s_best_i;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const float* sc = scores + (size_t)blockIdx.x * N;
    float prev_s = INFINITY;
    int prev_i = -1;
    for (int r = 0; r < k; ++r) {
        float bs = -INFINITY;
        int bi = 0x7fffffff;
        for (int t = tid; t < N; t += TOPK_THREADS) {
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

}  // namespace gvl

extern "C" int gvl_topk_cosine(const void* index, int N, int D, const void* queries, int Q, int k, float eps,
                               float* scratch, float* out_scores, int32_t* out_idx, void* stream) {
    using namespace gvl;
    GVL_CHECK_ARG(index && queries && scratch && out_scores && out_idx, "gvl_topk_cosine: null pointer");
    GVL_CHECK_ARG(N > 0 && Q > 0 && D > 0 && D % 8 == 0 && D <= 8192, "gvl_topk_cosine: bad shape N=%d Q=%d D=%d", N, Q, D);
    GVL_CHECK_ARG(k > 0 && k <= 64, "gvl_topk_cosine: k=%d out of range [1,64]", k);
    GVL_CHECK_ARG((uintptr_t)index % 16 == 0 && (uintptr_t)queries % 16 == 0, "gvl_topk_cosine: misaligned pointer");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const size_t smem = (size_t)TOPK_QB * D * 2 + TOPK_QB * sizeof(float);
    GVL_CUDA(cudaFuncSetAttribute(cos_scores_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int rows_per_cta = TOPK_THREADS / 32;
    int grid = (N + rows_per_cta - 1) / rows_per_cta;
    const int max_grid = sm_count() * 8;
    if (grid > max_grid) grid = max_grid;
    for (int q0 = 0; q0 < Q; q0 += TOPK_QB) {
        const int nq = Q - q0 < TOPK_QB ? Q - q0 : TOPK_QB;
        ProfScope prof(GVL_K_TOPK_SCORES, (double)N * D * 2, s);
        cos_scores_kernel<<<grid, TOPK_THREADS, smem, s>>>(
            reinterpret_cast<const __nv_bfloat16*>(index), N, D,
            reinterpret_cast<const __nv_bfloat16*>(queries) + (size_t)q0 * D, nq, eps, scratch + (size_t)q0 * N);
        GVL_LAUNCH_CHECK("cos_scores_kernel");
    }
    ProfScope prof(GVL_K_TOPK_SELECT, (double)Q * N * 4 * k, s);
    topk_select_kernel<<<Q, TOPK_THREADS, 0, s>>>(scratch, N, k, out_scores, out_idx);
    GVL_LAUNCH_CHECK("topk_select_kernel");
    return 0;
}
