// attention.cu — K4: fused non-causal self-attention for short fixed sequences (T = 729 for SigLIP,
// 1568 for VideoMAE) and K5: the MAP-head probe attention.
//
// K4 is a flash-style kernel: one CTA owns 128 query rows of one (image, head); K/V stream through a
// double-buffered cp.async pipeline in blocks of 64 keys; S = QK^T and O += PV run on the tensor cores
// via mma.sync.m16n8k16 (bf16 in, fp32 accumulate) with the online softmax in fp32 registers, so the
// T x T score matrix never touches memory.  Head dim 72 is handled by zero-padding the k-dimension of
// QK^T to 80 in shared memory only (global tensors stay dense).
// TODO(round 2): move S/O accumulators to TMEM with tcgen05.mma (this version uses the legacy HMMA path).
#include "common.cuh"

#include <cstdlib>

namespace gvl {

constexpr int ATT_BQ = 128;  // query rows per CTA (8 warps x 16)
constexpr int ATT_BKV = 64;  // keys per pipeline stage
constexpr int ATT_THREADS = 256;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(src_bytes)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ldmatrix_x2_trans(uint32_t (&r)[2], const void* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];"
                 : "=r"(r[0]), "=r"(r[1])
                 : "r"(smem_u32(p)));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int HD>
struct AttCfg {
    static constexpr int HDP = (HD + 15) / 16 * 16;  // k-dim of QK^T, zero padded
    static constexpr int SR = HDP + 8;               // smem row stride (elements): 16-B rows, conflict-free ldmatrix
    static constexpr int KSTEPS = HDP / 16;
    static constexpr int DTILES = HD / 8;            // n-tiles of the PV product
    static constexpr int CHUNKS = HD / 8;            // 16-byte chunks per global row
    static constexpr int Q_BYTES = ATT_BQ * SR * 2;
    static constexpr int KV_BYTES = ATT_BKV * SR * 2;
    static constexpr int SMEM_BYTES = Q_BYTES + 4 * KV_BYTES;
};

template <int HD>
__global__ void __launch_bounds__(ATT_THREADS, 2)
attention_bf16_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out, int T, int H,
                      float scale_log2) {
    using Cfg = AttCfg<HD>;
    constexpr int SR = Cfg::SR;
    extern __shared__ __align__(16) uint8_t att_smem[];
    __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(att_smem);
    __nv_bfloat16* sK = reinterpret_cast<__nv_bfloat16*>(att_smem + Cfg::Q_BYTES);                      // [2][BKV][SR]
    __nv_bfloat16* sV = reinterpret_cast<__nv_bfloat16*>(att_smem + Cfg::Q_BYTES + 2 * Cfg::KV_BYTES);  // [2][BKV][SR]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q0 = blockIdx.x * ATT_BQ;
    const int h = blockIdx.y, b = blockIdx.z;
    const int D = H * HD;
    const size_t ldq = (size_t)3 * D;
    const __nv_bfloat16* gQ = qkv + (size_t)b * T * ldq + (size_t)h * HD;
    const __nv_bfloat16* gK = gQ + D;
    const __nv_bfloat16* gV = gQ + 2 * D;

    // zero the k-padding columns [HD, HDP) of Q and both K buffers (cp.async never writes them)
    if constexpr (Cfg::HDP > HD) {
        constexpr int PADC = Cfg::HDP - HD;  // 8 for HD = 72
        for (int i = tid; i < (ATT_BQ + 2 * ATT_BKV) * PADC; i += ATT_THREADS) {
            const int r = i / PADC, c = HD + i % PADC;
            if (r < ATT_BQ)
                sQ[r * SR + c] = __float2bfloat16(0.f);
            else
                sK[(r - ATT_BQ) * SR + c] = __float2bfloat16(0.f);
        }
    }

    // Q tile
    for (int i = tid; i < ATT_BQ * Cfg::CHUNKS; i += ATT_THREADS) {
        const int r = i / Cfg::CHUNKS, c = i % Cfg::CHUNKS;
        const int qr = q0 + r;
        const int ok = qr < T;
        cp_async16(sQ + r * SR + c * 8, gQ + (size_t)(ok ? qr : T - 1) * ldq + c * 8, ok ? 16 : 0);
    }
    auto load_kv = [&](int blk, int buf) {
        const int k0 = blk * ATT_BKV;
        for (int i = tid; i < ATT_BKV * Cfg::CHUNKS; i += ATT_THREADS) {
            const int r = i / Cfg::CHUNKS, c = i % Cfg::CHUNKS;
            const int kr = k0 + r;
            const int ok = kr < T;
            const size_t goff = (size_t)(ok ? kr : T - 1) * ldq + c * 8;
            cp_async16(sK + (buf * ATT_BKV + r) * SR + c * 8, gK + goff, ok ? 16 : 0);
            cp_async16(sV + (buf * ATT_BKV + r) * SR + c * 8, gV + goff, ok ? 16 : 0);
        }
    };
    const int nblk = (T + ATT_BKV - 1) / ATT_BKV;
    load_kv(0, 0);
    cp_async_commit();

    float o[Cfg::DTILES][4];
#pragma unroll
    for (int i = 0; i < Cfg::DTILES; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
    float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
    uint32_t qf[Cfg::KSTEPS][4];

    for (int blk = 0; blk < nblk; ++blk) {
        const int buf = blk & 1;
        if (blk + 1 < nblk) load_kv(blk + 1, buf ^ 1);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();

        if (blk == 0) {
            // Q fragments stay in registers for the whole kernel
            const int r = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
            for (int ks = 0; ks < Cfg::KSTEPS; ++ks) ldmatrix_x4(qf[ks], sQ + r * SR + ks * 16 + (lane >> 4) * 8);
        }

        // ---- S = Q K^T (16 x 64 per warp) ----
        float s[ATT_BKV / 8][4];
#pragma unroll
        for (int i = 0; i < ATT_BKV / 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
        const __nv_bfloat16* kb = sK + buf * ATT_BKV * SR;
#pragma unroll
        for (int ks = 0; ks < Cfg::KSTEPS; ++ks) {
#pragma unroll
            for (int np = 0; np < ATT_BKV / 16; ++np) {
                uint32_t bf[4];
                const int key = np * 16 + (lane & 7) + (lane >> 4) * 8;
                const int dcol = ks * 16 + ((lane >> 3) & 1) * 8;
                ldmatrix_x4(bf, kb + key * SR + dcol);
                mma_bf16_16816(s[2 * np], qf[ks], bf[0], bf[1]);
                mma_bf16_16816(s[2 * np + 1], qf[ks], bf[2], bf[3]);
            }
        }

        // ---- online softmax (rows g and g+8 of this warp's 16) ----
        const int kbase = blk * ATT_BKV + (lane & 3) * 2;
        float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int nt = 0; nt < ATT_BKV / 8; ++nt) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int key = kbase + nt * 8 + (e & 1);
                const float v = key < T ? s[nt][e] * scale_log2 : -INFINITY;
                s[nt][e] = v;
                mx[e >> 1] = fmaxf(mx[e >> 1], v);
            }
        }
        float corr[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
            const float m_new = fmaxf(m_run[r], mx[r]);
            corr[r] = exp2f(m_run[r] - m_new);
            m_run[r] = m_new;
        }
        float rs[2] = {0.f, 0.f};
        uint32_t pf[ATT_BKV / 16][4];
#pragma unroll
        for (int nt = 0; nt < ATT_BKV / 8; ++nt) {
            const float p0 = exp2f(s[nt][0] - m_run[0]), p1 = exp2f(s[nt][1] - m_run[0]);
            const float p2 = exp2f(s[nt][2] - m_run[1]), p3 = exp2f(s[nt][3] - m_run[1]);
            rs[0] += p0 + p1;
            rs[1] += p2 + p3;
            pf[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16x2(p0, p1);
            pf[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16x2(p2, p3);
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) l_run[r] = l_run[r] * corr[r] + rs[r];
#pragma unroll
        for (int dt = 0; dt < Cfg::DTILES; ++dt) {
            o[dt][0] *= corr[0];
            o[dt][1] *= corr[0];
            o[dt][2] *= corr[1];
            o[dt][3] *= corr[1];
        }

        // ---- O += P V ----
        const __nv_bfloat16* vb = sV + buf * ATT_BKV * SR;
#pragma unroll
        for (int kk = 0; kk < ATT_BKV / 16; ++kk) {
            const int key = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
            for (int dp = 0; dp < Cfg::DTILES / 2; ++dp) {
                uint32_t bf[4];
                ldmatrix_x4_trans(bf, vb + key * SR + dp * 16 + (lane >> 4) * 8);
                mma_bf16_16816(o[2 * dp], pf[kk], bf[0], bf[1]);
                mma_bf16_16816(o[2 * dp + 1], pf[kk], bf[2], bf[3]);
            }
            if constexpr ((Cfg::DTILES & 1) != 0) {
                uint32_t bf2[2];
                ldmatrix_x2_trans(bf2, vb + key * SR + (Cfg::DTILES - 1) * 8);
                mma_bf16_16816(o[Cfg::DTILES - 1], pf[kk], bf2[0], bf2[1]);
            }
        }
        __syncthreads();  // everyone done with `buf` before the next iteration's prefetch overwrites it
    }
    cp_async_wait<0>();

    // ---- finalise: divide by the row sums, stage through this warp's Q rows, 16-byte stores ----
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
    }
    const float inv0 = 1.f / l_run[0], inv1 = 1.f / l_run[1];
    __nv_bfloat16* stage = sQ + warp * 16 * SR;
    {
        const int g = lane >> 2, t2 = (lane & 3) * 2;
#pragma unroll
        for (int dt = 0; dt < Cfg::DTILES; ++dt) {
            *reinterpret_cast<uint32_t*>(stage + g * SR + dt * 8 + t2) = pack_bf16x2(o[dt][0] * inv0, o[dt][1] * inv0);
            *reinterpret_cast<uint32_t*>(stage + (g + 8) * SR + dt * 8 + t2) =
                pack_bf16x2(o[dt][2] * inv1, o[dt][3] * inv1);
        }
    }
    __syncwarp();
    __nv_bfloat16* gO = out + (size_t)b * T * D + (size_t)h * HD;
    for (int i = lane; i < 16 * Cfg::CHUNKS; i += 32) {
        const int r = i / Cfg::CHUNKS, c = i % Cfg::CHUNKS;
        const int qr = q0 + warp * 16 + r;
        if (qr < T)
            *reinterpret_cast<uint4*>(gO + (size_t)qr * D + c * 8) = *reinterpret_cast<const uint4*>(stage + r * SR + c * 8);
    }
}

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
static int launch_attention(const void* qkv, void* out, int B, int T, int H, float scale, cudaStream_t s) {
    using Cfg = AttCfg<HD>;
    GVL_CUDA(cudaFuncSetAttribute(attention_bf16_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  Cfg::SMEM_BYTES));
    dim3 grid((T + ATT_BQ - 1) / ATT_BQ, H, B);
    ProfScope prof(GVL_K_ATTENTION, 4.0 * B * (double)H * T * (double)T * HD, s);
    attention_bf16_kernel<HD><<<grid, ATT_THREADS, Cfg::SMEM_BYTES, s>>>(
        reinterpret_cast<const __nv_bfloat16*>(qkv), reinterpret_cast<__nv_bfloat16*>(out), T, H,
        scale * 1.4426950408889634f);
    GVL_LAUNCH_CHECK("attention_bf16_kernel");
    return 0;
}

template <int HD>
int launch_attention_tc(const void* qkv, void* out, int B, int T, int H, float scale, cudaStream_t s);  // attention_tc.cu
template <int HD>
int launch_attention_sdb(const void* qkv, void* out, int B, int T, int H, float scale, cudaStream_t s);  // attention_sdb.cu
template <int HD>
int launch_attention_split(const void* qkv, void* out, int B, int T, int H, float scale, cudaStream_t s);  // attention_split.cu

}  // namespace gvl

extern "C" int gvl_attention_bf16(const void* qkv, void* out, int B, int T, int H, int hd, float scale, void* stream) {
    using namespace gvl;
    GVL_CHECK_ARG(qkv && out, "gvl_attention_bf16: null pointer");
    GVL_CHECK_ARG(B > 0 && T > 0 && H > 0, "gvl_attention_bf16: bad shape B=%d T=%d H=%d", B, T, H);
    GVL_CHECK_ARG(B <= 65535 && H <= 65535, "gvl_attention_bf16: B/H exceed grid limits");
    GVL_CHECK_ARG((uintptr_t)qkv % 16 == 0 && (uintptr_t)out % 16 == 0, "gvl_attention_bf16: misaligned pointer");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    // GVL_ATTN_LEGACY=1 selects the round-1 mma.sync kernel (kept for A/B measurements only)
    static const bool legacy = [] {
        const char* e = getenv("GVL_ATTN_LEGACY");
        return e && e[0] == '1';
    }();
    // GVL_ATTN_TC1=1 selects the single-tile tcgen05 kernel (two CTAs per SM), also A/B only
    static const bool single_tile = [] {
        const char* e = getenv("GVL_ATTN_TC1");
        return e && e[0] == '1';
    }();
    // GVL_ATTN_SPLIT=1 selects the split-row softmax variant (two threads per query row)
    static const bool split_rows = [] {
        const char* e = getenv("GVL_ATTN_SPLIT");
        return e && e[0] == '1';
    }();
    if (!legacy && !single_tile && split_rows) {
        if (hd == 72) return launch_attention_split<72>(qkv, out, B, T, H, scale, s);
        if (hd == 64) return launch_attention_split<64>(qkv, out, B, T, H, scale, s);
    } else if (!legacy && !single_tile) {
        if (hd == 72) return launch_attention_sdb<72>(qkv, out, B, T, H, scale, s);
        if (hd == 64) return launch_attention_sdb<64>(qkv, out, B, T, H, scale, s);
    } else if (!legacy) {
        if (hd == 72) return launch_attention_tc<72>(qkv, out, B, T, H, scale, s);
        if (hd == 64) return launch_attention_tc<64>(qkv, out, B, T, H, scale, s);
    } else {
        if (hd == 72) return launch_attention<72>(qkv, out, B, T, H, scale, s);
        if (hd == 64) return launch_attention<64>(qkv, out, B, T, H, scale, s);
    }
    set_error("gvl_attention_bf16: unsupported head dim %d (built for 72 and 64)", hd);
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
