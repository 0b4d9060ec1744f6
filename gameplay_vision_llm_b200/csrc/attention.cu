// attention.cu — entry points of K4 (fused non-causal self-attention, T = 729 for SigLIP / 1568 for VideoMAE; the
// kernel itself is the tcgen05 / TMEM flash attention in attention_sdb.cu — the only attention path in the library)
// and K5, the MAP-head probe attention.
#include "common.cuh"

#include <algorithm>
#include <vector>

#include <cstdlib>

namespace gvl {

// K5: probe attention of the MAP head.  One CTA per (image, head): scores over T keys with the
// constant, pre-scaled query, fp32 softmax, weighted sum of V.  Memory-bound (reads K and V once).
constexpr int PROBE_THREADS = 256;
constexpr int PROBE_MAXT = 2048;

__global__ void __launch_bounds__(PROBE_THREADS)
probe_attention_kernel(const float* __restrict__ q, const __nv_bfloat16* __restrict__ kv,
                       __nv_bfloat16* __restrict__ out, int T, int H, int hd) {
    __shared__ float s_score[PROBE_MAXT];
    __shared__ float s_q[128];
    __shared__ float s_red[PROBE_THREADS / 32];
    __shared__ float s_acc[PROBE_THREADS / 32][128];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int h = blockIdx.x, b = blockIdx.y;
    const int D = H * hd;
    const size_t ld = (size_t)2 * D;
    const __nv_bfloat16* gK = kv + (size_t)b * T * ld + (size_t)h * hd;
    const __nv_bfloat16* gV = gK + D;
    if (tid < hd) s_q[tid] = q[h * hd + tid];
    __syncthreads();

    // scores: one key per thread per pass
    float lmax = -INFINITY;
    const int chunks = hd / 8;
    for (int t = tid; t < T; t += PROBE_THREADS) {
        const uint4* kr = reinterpret_cast<const uint4*>(gK + (size_t)t * ld);
        float acc = 0.f;
        for (int c = 0; c < chunks; ++c) {
            const uint4 u = __ldg(kr + c);
            const float* qq = s_q + c * 8;
            acc += bf16_lo(u.x) * qq[0] + bf16_hi(u.x) * qq[1] + bf16_lo(u.y) * qq[2] + bf16_hi(u.y) * qq[3] +
                   bf16_lo(u.z) * qq[4] + bf16_hi(u.z) * qq[5] + bf16_lo(u.w) * qq[6] + bf16_hi(u.w) * qq[7];
        }
        s_score[t] = acc;
        lmax = fmaxf(lmax, acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
    if (lane == 0) s_red[warp] = lmax;
    __syncthreads();
    float gmax = s_red[0];
    for (int w = 1; w < PROBE_THREADS / 32; ++w) gmax = fmaxf(gmax, s_red[w]);
    __syncthreads();
    float lsum = 0.f;
    for (int t = tid; t < T; t += PROBE_THREADS) {
        const float p = __expf(s_score[t] - gmax);
        s_score[t] = p;
        lsum += p;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
    if (lane == 0) s_red[warp] = lsum;
    __syncthreads();
    float gsum = 0.f;
    for (int w = 0; w < PROBE_THREADS / 32; ++w) gsum += s_red[w];

    // weighted V sum: each warp takes keys warp, warp+8, ...; lane owns dims lane, lane+32, lane+64(, +96)
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int t = warp; t < T; t += PROBE_THREADS / 32) {
        const float p = s_score[t];
        const __nv_bfloat16* vr = gV + (size_t)t * ld;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int d = lane + j * 32;
            if (d < hd) acc[j] += p * __bfloat162float(vr[d]);
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int d = lane + j * 32;
        if (d < hd) s_acc[warp][d] = acc[j];
    }
    __syncthreads();
    if (tid < hd) {
        float v = 0.f;
        for (int w = 0; w < PROBE_THREADS / 32; ++w) v += s_acc[w][tid];
        out[(size_t)b * D + h * hd + tid] = __float2bfloat16(v / gsum);
    }
}

template <int HD>
int launch_attention_sdb(const void* qkv, void* out, int B, int T, int H, float scale, cudaStream_t s);  // attention_sdb.cu
template <int HD>
int launch_attention_sdb_varlen(const void* qkv, void* out, int M_total, const void* tiles, int n_tiles, double score_elems,
                                int H, float scale, cudaStream_t s);
#ifdef GVL_EXPERIMENTS
// round-2 experiment (two softmax warpgroups, fixed-reference softmax, polynomial exp2): measured equal or slower than
// attention_sdb.cu on every shape (profiles/r02_attention_experiments.md); built only with EXTRA=-DGVL_EXPERIMENTS
template <int HD>
int launch_attention_p2(const void* qkv, void* out, int B, int T, int H, float scale, int poly, int force_exact,
                        cudaStream_t s);  // experiments/attention_p2.cu
#endif

}  // namespace gvl

extern "C" int gvl_attention_bf16(const void* qkv, void* out, int B, int T, int H, int hd, float scale, void* stream) {
    using namespace gvl;
    GVL_CHECK_ARG(qkv && out, "gvl_attention_bf16: null pointer");
    GVL_CHECK_ARG(B > 0 && T > 0 && H > 0, "gvl_attention_bf16: bad shape B=%d T=%d H=%d", B, T, H);
    GVL_CHECK_ARG(B <= 65535 && H <= 65535, "gvl_attention_bf16: B/H exceed grid limits");
    GVL_CHECK_ARG((uintptr_t)qkv % 16 == 0 && (uintptr_t)out % 16 == 0, "gvl_attention_bf16: misaligned pointer");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
#ifdef GVL_EXPERIMENTS  // A/B builds only (make EXTRA=-DGVL_EXPERIMENTS): the shipped library has one attention path
    static const int impl = [] { const char* e = getenv("GVL_ATTN_IMPL"); return (e && e[0] == 'p') ? 1 : 0; }();
    static const int poly = [] { const char* e = getenv("GVL_ATTN_POLY"); return e ? atoi(e) : 0; }();
    if (impl == 1) {
        if (hd == 72) return launch_attention_p2<72>(qkv, out, B, T, H, scale, poly, 0, s);
        if (hd == 64) return launch_attention_p2<64>(qkv, out, B, T, H, scale, poly, 0, s);
    }
#endif
    if (hd == 72) return launch_attention_sdb<72>(qkv, out, B, T, H, scale, s);
    if (hd == 64) return launch_attention_sdb<64>(qkv, out, B, T, H, scale, s);
    set_error("gvl_attention_bf16: unsupported head dim %d (built for 72 and 64)", hd);
    return 1;
}

extern "C" int gvl_attention_varlen_tiles(int n_items, const int32_t* h_item_tokens, int32_t* h_tiles, int* n_tiles) {
    using namespace gvl;
    GVL_CHECK_ARG(n_items > 0 && h_item_tokens && n_tiles, "gvl_attention_varlen_tiles: bad arguments");
    long long count = 0, tok = 0;
    for (int i = 0; i < n_items; ++i) {
        GVL_CHECK_ARG(h_item_tokens[i] > 0, "gvl_attention_varlen_tiles: item %d has %d tokens", i, h_item_tokens[i]);
        count += (h_item_tokens[i] + 127) / 128;
        tok += h_item_tokens[i];
    }
    GVL_CHECK_ARG(count <= 2147483647LL && tok <= 2147483647LL, "gvl_attention_varlen_tiles: batch too large");
    *n_tiles = (int)count;
    if (!h_tiles) return 0;  // size query
    // longest items first: a tile's cost grows with its item's key count, so the tail of the launch is short tiles
    std::vector<int> order(n_items);
    for (int i = 0; i < n_items; ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return h_item_tokens[a] > h_item_tokens[b]; });
    std::vector<long long> start(n_items);
    tok = 0;
    for (int i = 0; i < n_items; ++i) {
        start[i] = tok;
        tok += h_item_tokens[i];
    }
    int32_t* o = h_tiles;
    for (int i : order)
        for (int q0 = 0; q0 < h_item_tokens[i]; q0 += 128) {
            o[0] = (int32_t)start[i];
            o[1] = h_item_tokens[i];
            o[2] = q0;
            o[3] = 0;
            o += 4;
        }
    return 0;
}

extern "C" int gvl_attention_varlen_bf16(const void* qkv, void* out, int M_total, const void* tiles, int n_tiles,
                                         double score_elems, int H, int hd, float scale, void* stream) {
    using namespace gvl;
    GVL_CHECK_ARG(qkv && out && tiles, "gvl_attention_varlen_bf16: null pointer");
    GVL_CHECK_ARG(M_total > 0 && n_tiles > 0 && H > 0 && H <= 65535, "gvl_attention_varlen_bf16: bad shape M=%d tiles=%d H=%d",
                  M_total, n_tiles, H);
    GVL_CHECK_ARG((uintptr_t)qkv % 16 == 0 && (uintptr_t)out % 16 == 0 && (uintptr_t)tiles % 16 == 0,
                  "gvl_attention_varlen_bf16: misaligned pointer");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (hd == 72) return launch_attention_sdb_varlen<72>(qkv, out, M_total, tiles, n_tiles, score_elems, H, scale, s);
    if (hd == 64) return launch_attention_sdb_varlen<64>(qkv, out, M_total, tiles, n_tiles, score_elems, H, scale, s);
    set_error("gvl_attention_varlen_bf16: unsupported head dim %d (built for 72 and 64)", hd);
    return 1;
}

extern "C" int gvl_probe_attention_bf16(const float* q, const void* kv, void* out, int B, int T, int H, int hd,
                                        void* stream) {
    using namespace gvl;
    GVL_CHECK_ARG(q && kv && out, "gvl_probe_attention_bf16: null pointer");
    GVL_CHECK_ARG(B > 0 && T > 0 && T <= PROBE_MAXT && H > 0 && hd % 8 == 0 && hd <= 128,
                  "gvl_probe_attention_bf16: bad shape B=%d T=%d H=%d hd=%d", B, T, H, hd);
    GVL_CHECK_ARG((uintptr_t)kv % 16 == 0, "gvl_probe_attention_bf16: misaligned pointer");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    dim3 grid(H, B);
    ProfScope prof(GVL_K_PROBE_ATTENTION, 4.0 * B * (double)T * H * hd, s);
    probe_attention_kernel<<<grid, PROBE_THREADS, 0, s>>>(q, reinterpret_cast<const __nv_bfloat16*>(kv),
                                                          reinterpret_cast<__nv_bfloat16*>(out), T, H, hd);
    GVL_LAUNCH_CHECK("probe_attention_kernel");
    return 0;
}
