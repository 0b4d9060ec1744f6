// layernorm.cu — K3: row LayerNorm, one warp per row, the row held in registers, two-pass fp32
// statistics (mean, then centred variance) reduced with warp shuffles.  HBM-bound: 2 B read + 2 B
// written per element, 8-byte coalesced accesses.
#include "common.cuh"

namespace gvl {

constexpr int kLnWarps = 8;

// NV = number of 4-element vectors per lane (row length D <= 128 * NV, D % 4 == 0).
template <int NV>
__global__ void __launch_bounds__(kLnWarps * 32)
layernorm_bf16_kernel(const __nv_bfloat16* __restrict__ x, int ldx, const float* __restrict__ gamma,
                      const float* __restrict__ beta, __nv_bfloat16* __restrict__ y, int ldy, int rows, int D,
                      float eps) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = blockIdx.x * kLnWarps + warp;
    if (row >= rows) return;
    const int nvec = D >> 2;
    const uint2* xr = reinterpret_cast<const uint2*>(x + (size_t)row * ldx);

    float v[NV][4];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int vi = lane + i * 32;
        if (vi < nvec) {
            const uint2 u = __ldg(xr + vi);
            v[i][0] = bf16_lo(u.x);
            v[i][1] = bf16_hi(u.x);
            v[i][2] = bf16_lo(u.y);
            v[i][3] = bf16_hi(u.y);
        } else {
            v[i][0] = v[i][1] = v[i][2] = v[i][3] = 0.f;
        }
        sum += (v[i][0] + v[i][1]) + (v[i][2] + v[i][3]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum / (float)D;

    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int vi = lane + i * 32;
        if (vi < nvec) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float d = v[i][j] - mean;
                sq += d * d;
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float rstd = rsqrtf(sq / (float)D + eps);

    uint2* yr = reinterpret_cast<uint2*>(y + (size_t)row * ldy);
    const float4* g4 = reinterpret_cast<const float4*>(gamma);
    const float4* b4 = reinterpret_cast<const float4*>(beta);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int vi = lane + i * 32;
        if (vi < nvec) {
            const float4 g = __ldg(g4 + vi), b = __ldg(b4 + vi);
            uint2 o;
            o.x = pack_bf16x2((v[i][0] - mean) * rstd * g.x + b.x, (v[i][1] - mean) * rstd * g.y + b.y);
            o.y = pack_bf16x2((v[i][2] - mean) * rstd * g.z + b.z, (v[i][3] - mean) * rstd * g.w + b.w);
            yr[vi] = o;
        }
    }
}

}  // namespace gvl

namespace gvl {

// LayerNorm-fusion statistics, finalised once per row: the producer GEMM's per-slab partial sums (sum x, sum x^2) ->
// (rstd, mean * rstd).  Same summation order as the consumer epilogue uses when it reads the slabs itself, so both
// routes give identical bits; the consumer then needs ONE 8-byte load per row and tile instead of slots / 2 scattered
// 16-byte loads.
constexpr int kLnFinalizeMaxPairs = 10;  // 20 slots of 64 columns = rows of up to 1280 elements in registers

__global__ void __launch_bounds__(256)
ln_finalize_kernel(const float* __restrict__ stats, int rows, int slots, float inv_d, float eps, float2* __restrict__ out) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= rows) return;
    const float4* sp = reinterpret_cast<const float4*>(stats + (size_t)row * slots * 2);
    // all loads issued before the first use (one memory round trip per row instead of slots / 2); the sums keep the
    // slot order, so the bits equal what a consumer reading the slots itself would get
    const int pairs = slots >> 1;
    float4 t[kLnFinalizeMaxPairs];
#pragma unroll
    for (int i = 0; i < kLnFinalizeMaxPairs; ++i) t[i] = i < pairs ? sp[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < kLnFinalizeMaxPairs; ++i) {
        if (i < pairs) {
            s1 += t[i].x;
            s2 += t[i].y;
            s1 += t[i].z;
            s2 += t[i].w;
        }
    }
    for (int i = kLnFinalizeMaxPairs; i < pairs; ++i) {  // wider rows than the tower's: plain loop
        const float4 u = sp[i];
        s1 += u.x;
        s2 += u.y;
        s1 += u.z;
        s2 += u.w;
    }
    const float mean = s1 * inv_d;
    const float var = fmaxf(s2 * inv_d - mean * mean, 0.0f);
    const float rstd = rsqrtf(var + eps);
    out[row] = make_float2(rstd, mean * rstd);
}

}  // namespace gvl

extern "C" int gvl_ln_finalize(const float* stats, int rows, int slots, int dim, float eps, float* out, void* stream) {
    using namespace gvl;
    GVL_CHECK_ARG(stats && out, "gvl_ln_finalize: null pointer");
    GVL_CHECK_ARG(rows > 0 && slots > 0 && slots % 2 == 0 && dim > 0, "gvl_ln_finalize: bad shape rows=%d slots=%d dim=%d",
                  rows, slots, dim);
    GVL_CHECK_ARG((uintptr_t)stats % 16 == 0 && (uintptr_t)out % 8 == 0, "gvl_ln_finalize: misaligned pointer");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    ProfScope prof(GVL_K_LAYERNORM, (double)rows * (slots * 8 + 8), s);
    ln_finalize_kernel<<<(rows + 255) / 256, 256, 0, s>>>(stats, rows, slots, 1.0f / (float)dim, eps,
                                                          reinterpret_cast<float2*>(out));
    GVL_LAUNCH_CHECK("ln_finalize_kernel");
    return 0;
}

extern "C" int gvl_layernorm_bf16(const void* x, int ldx, const float* gamma, const float* beta, void* y, int ldy,
                                  int rows, int D, float eps, void* stream) {
    using namespace gvl;
    GVL_CHECK_ARG(x && y && gamma && beta, "gvl_layernorm_bf16: null pointer");
    GVL_CHECK_ARG(rows > 0 && D > 0 && D % 4 == 0 && D <= 4096, "gvl_layernorm_bf16: bad shape rows=%d D=%d", rows, D);
    GVL_CHECK_ARG(ldx % 4 == 0 && ldy % 4 == 0 && ldx >= D && ldy >= D, "gvl_layernorm_bf16: bad ld %d/%d", ldx, ldy);
    GVL_CHECK_ARG((uintptr_t)x % 8 == 0 && (uintptr_t)y % 8 == 0 && (uintptr_t)gamma % 16 == 0 &&
                      (uintptr_t)beta % 16 == 0,
                  "gvl_layernorm_bf16: misaligned pointer");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const int grid = (rows + kLnWarps - 1) / kLnWarps;
    const auto* xp = reinterpret_cast<const __nv_bfloat16*>(x);
    auto* yp = reinterpret_cast<__nv_bfloat16*>(y);
    const int nv = (D / 4 + 31) / 32;
    ProfScope prof(GVL_K_LAYERNORM, 4.0 * rows * (double)D, s);
    if (nv <= 6)
        layernorm_bf16_kernel<6><<<grid, kLnWarps * 32, 0, s>>>(xp, ldx, gamma, beta, yp, ldy, rows, D, eps);
    else if (nv <= 9)
        layernorm_bf16_kernel<9><<<grid, kLnWarps * 32, 0, s>>>(xp, ldx, gamma, beta, yp, ldy, rows, D, eps);
    else if (nv <= 16)
        layernorm_bf16_kernel<16><<<grid, kLnWarps * 32, 0, s>>>(xp, ldx, gamma, beta, yp, ldy, rows, D, eps);
    else
        layernorm_bf16_kernel<32><<<grid, kLnWarps * 32, 0, s>>>(xp, ldx, gamma, beta, yp, ldy, rows, D, eps);
    GVL_LAUNCH_CHECK("layernorm_bf16_kernel");
    return 0;
}
