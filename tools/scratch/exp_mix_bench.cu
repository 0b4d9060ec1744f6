// Microbenchmark (tuning aid, not shipped): exp2 throughput per SM for the attention softmax inner loop when a fraction
// of the elements is computed on the FMA / ALU pipes (Cody-Waite split + cubic polynomial, packed f32x2 arithmetic)
// instead of MUFU.EX2.  Prints exps per clock per SM for 4 / 8 / 16 warps per SM and the max relative error of the
// polynomial path.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o exp_mix_bench exp_mix_bench.cu
#include <cstdint>
#include <cstdio>
#include <cmath>
#include <cuda_bf16.h>

__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint64_t pack2(float lo, float hi) { uint64_t d; asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi)); return d; }
__device__ __forceinline__ void unpack2(uint64_t d, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(d)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) { uint64_t d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) { __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi); return *reinterpret_cast<uint32_t*>(&v); }

// 2^x for two values, x <= ~100: n = rint(x) by the magic-number add, f = x - n in [-0.5, 0.5], cubic minimax of 2^f,
// exponent inserted with an integer shift-add.  x is clamped at -126 (2^-126: no denormal garbage).
constexpr float kMagic = 12582912.0f;  // 1.5 * 2^23
// minimax cubic for 2^f on [-0.5, 0.5] (relative error 7.6e-5 — far below bf16's 2^-9)
constexpr float kC0 = 0.9999277f, kC1 = 0.69325477f, kC2 = 0.24261397f, kC3 = 0.055205505f;
__device__ __forceinline__ void ex2_poly2(float x0, float x1, float& p0, float& p1) {
    x0 = fmaxf(x0, -126.0f);
    x1 = fmaxf(x1, -126.0f);
    const uint64_t x = pack2(x0, x1);
    const uint64_t t = add2(x, pack2(kMagic, kMagic));
    const uint64_t n = add2(t, pack2(-kMagic, -kMagic));
    const uint64_t f = fma2(n, pack2(-1.0f, -1.0f), x);
    uint64_t p = fma2(f, pack2(kC3, kC3), pack2(kC2, kC2));
    p = fma2(p, f, pack2(kC1, kC1));
    p = fma2(p, f, pack2(kC0, kC0));
    float t0, t1, q0, q1;
    unpack2(t, t0, t1);
    unpack2(p, q0, q1);
    p0 = __uint_as_float(__float_as_uint(q0) + (__float_as_uint(t0) << 23));
    p1 = __uint_as_float(__float_as_uint(q1) + (__float_as_uint(t1) << 23));
}

// packed half-precision MUFU: two exps per MUFU lane-op
__device__ __forceinline__ uint32_t ex2_h2(float x0, float x1) {  // f32 pair -> f16x2 -> ex2 -> f16x2
    uint32_t h, y;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(x1), "f"(x0));
    asm("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(h));
    return y;
}
__device__ __forceinline__ uint32_t ex2_b2(float x0, float x1) {  // f32 pair -> bf16x2 -> ex2 -> bf16x2
    uint32_t h, y;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(x1), "f"(x0));
    asm("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(y) : "r"(h));
    return y;
}
template <int KIND>
__global__ void k2(float* out, const float* in, int iters, float c, float nm, long long* cyc) {
    float s[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) s[i] = in[(threadIdx.x * 64 + i) & 1023];
    uint32_t acc = 0;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        const uint64_t cc = pack2(c, c), nn = pack2(nm, nm);
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            float x0, x1;
            unpack2(fma2(pack2(s[2 * i], s[2 * i + 1]), cc, nn), x0, x1);
            acc ^= KIND == 0 ? ex2_h2(x0, x1) : ex2_b2(x0, x1);
        }
        nm += 1e-7f;
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float(acc);
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int KIND>
void run2(float* out, float* in, long long* cyc) {
    for (int warps : {4, 8, 16}) {
        const int iters = 2000;
        k2<KIND><<<148, warps * 32>>>(out, in, iters, 1.f, -3.f, cyc);
        cudaDeviceSynchronize();
        long long h[148];
        cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        double c = 0;
        for (int i = 0; i < 148; ++i) c += h[i];
        c /= 148;
        printf("%s warps/SM %2d: %.0f cycles, %.2f exp/clk/SM\n", KIND == 0 ? "ex2.f16x2" : "ex2.bf16x2", warps, c,
               (double)warps * 32 * 64 * iters / c);
    }
}

// Variant with the clamp folded into a saturating FMA: u = sat(s*c/252 + (nm + 126)/252) in [0, 1]  <->  x in [-126, 126];
// t = 252 u + (magic - 126) = magic + rint(x); f = 252 u - (t - (magic - 126)); no FMNMX at all.
__device__ __forceinline__ float fma_sat(float a, float b, float c) { float d; asm("fma.rn.sat.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ void ex2_poly2_sat(float s0, float s1, float cs, float ns, float& p0, float& p1) {
    const uint64_t u = pack2(fma_sat(s0, cs, ns), fma_sat(s1, cs, ns));
    const uint64_t k252 = pack2(252.0f, 252.0f);
    const uint64_t t = fma2(u, k252, pack2(kMagic - 126.0f, kMagic - 126.0f));
    const uint64_t negw = fma2(t, pack2(-1.0f, -1.0f), pack2(kMagic - 126.0f, kMagic - 126.0f));  // -(n + 126), exact
    const uint64_t f = fma2(u, k252, negw);
    uint64_t p = fma2(f, pack2(kC3, kC3), pack2(kC2, kC2));
    p = fma2(p, f, pack2(kC1, kC1));
    p = fma2(p, f, pack2(kC0, kC0));
    float t0, t1, q0, q1;
    unpack2(t, t0, t1);
    unpack2(p, q0, q1);
    p0 = __uint_as_float(__float_as_uint(q0) + (__float_as_uint(t0) << 23));
    p1 = __uint_as_float(__float_as_uint(q1) + (__float_as_uint(t1) << 23));
}
template <int POLY>
__global__ void k3(float* out, const float* in, int iters, float c, float nm, long long* cyc) {
    float s[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) s[i] = in[(threadIdx.x * 64 + i) & 1023];
    uint32_t acc = 0;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        const uint64_t cc = pack2(c, c), nn = pack2(nm, nm);
        const float cs = c * (1.0f / 252.0f), ns = (nm + 126.0f) * (1.0f / 252.0f);
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            float x0, x1, p0, p1;
            const bool poly = ((i & 7) * POLY / 8) != (((i & 7) + 1) * POLY / 8);
            if (poly) ex2_poly2_sat(s[2 * i], s[2 * i + 1], cs, ns, p0, p1);
            else { unpack2(fma2(pack2(s[2 * i], s[2 * i + 1]), cc, nn), x0, x1); p0 = ex2(x0); p1 = ex2(x1); }
            acc ^= pack_bf16x2(p0, p1);
        }
        nm += 1e-7f;
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float(acc);
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void err_kernel_sat(float* maxrel, int n) {
    float worst = 0.f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float s = -140.0f + 280.0f * (float)i / (float)n;   // x = s * 1 + 0
        float p0, p1;
        ex2_poly2_sat(s, s + 0.37f, 1.0f / 252.0f, 126.0f / 252.0f, p0, p1);
        const float r0 = exp2f(fminf(fmaxf(s, -126.f), 126.f)), r1 = exp2f(fminf(fmaxf(s + 0.37f, -126.f), 126.f));
        worst = fmaxf(worst, fmaxf(fabsf(p0 - r0) / r0, fabsf(p1 - r1) / r1));
    }
    atomicMax(reinterpret_cast<int*>(maxrel), __float_as_int(worst));
}
template <int POLY>
void run3(float* out, float* in, long long* cyc) {
    for (int warps : {4, 8, 16}) {
        const int iters = 2000;
        k3<POLY><<<148, warps * 32>>>(out, in, iters, 1.f, -3.f, cyc);
        cudaDeviceSynchronize();
        long long h[148];
        cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        double c = 0;
        for (int i = 0; i < 148; ++i) c += h[i];
        c /= 148;
        printf("sat-poly %d/8 warps/SM %2d: %.0f cycles, %.2f exp/clk/SM\n", POLY, warps, c, (double)warps * 32 * 64 * iters / c);
    }
}

// MODE: of every 8 pairs, POLY pairs go through the polynomial (0 = all MUFU, 8 = all polynomial)
template <int POLY>
__global__ void k(float* out, const float* in, int iters, float c, float nm, long long* cyc) {
    float s[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) s[i] = in[(threadIdx.x * 64 + i) & 1023];
    uint32_t acc = 0;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        const uint64_t cc = pack2(c, c), nn = pack2(nm, nm);
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            float x0, x1, p0, p1;
            unpack2(fma2(pack2(s[2 * i], s[2 * i + 1]), cc, nn), x0, x1);
            // spread the polynomial pairs evenly through the 8-pair group
            const bool poly = ((i & 7) * POLY / 8) != (((i & 7) + 1) * POLY / 8);
            if (poly) ex2_poly2(x0, x1, p0, p1);
            else { p0 = ex2(x0); p1 = ex2(x1); }
            acc ^= pack_bf16x2(p0, p1);
        }
        nm += 1e-7f;
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float(acc);
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void err_kernel(float* maxrel, int n) {
    float worst = 0.f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float x = -130.0f + 140.0f * (float)i / (float)n;
        float p0, p1;
        ex2_poly2(x, x + 0.37f, p0, p1);
        const float r0 = exp2f(fmaxf(x, -126.f)), r1 = exp2f(fmaxf(x + 0.37f, -126.f));
        worst = fmaxf(worst, fmaxf(fabsf(p0 - r0) / r0, fabsf(p1 - r1) / r1));
    }
    atomicMax(reinterpret_cast<int*>(maxrel), __float_as_int(worst));
}

template <int POLY>
void run(float* out, float* in, long long* cyc) {
    for (int warps : {4, 8, 16}) {
        const int iters = 2000;
        k<POLY><<<148, warps * 32>>>(out, in, iters, 1.f, -3.f, cyc);
        cudaDeviceSynchronize();
        long long h[148];
        cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        double c = 0;
        for (int i = 0; i < 148; ++i) c += h[i];
        c /= 148;
        printf("poly %d/8 warps/SM %2d: %.0f cycles, %.2f exp/clk/SM\n", POLY, warps, c, (double)warps * 32 * 64 * iters / c);
    }
}

int main() {
    float *in, *out, *mr;
    long long* cyc;
    cudaMalloc(&in, 4096); cudaMalloc(&out, 1 << 22); cudaMalloc(&cyc, 148 * 8); cudaMalloc(&mr, 4);
    cudaMemset(in, 0, 4096); cudaMemset(mr, 0, 4);
    err_kernel<<<148, 256>>>(mr, 1 << 24);
    float h; cudaMemcpy(&h, mr, 4, cudaMemcpyDeviceToHost);
    printf("polynomial exp2 max relative error over [-130, 10]: %.3e\n", h);
    run<0>(out, in, cyc); run<2>(out, in, cyc); run<3>(out, in, cyc); run<4>(out, in, cyc); run<5>(out, in, cyc); run<6>(out, in, cyc); run<8>(out, in, cyc);
    run2<0>(out, in, cyc); run2<1>(out, in, cyc);
    cudaMemset(mr, 0, 4);
    err_kernel_sat<<<148, 256>>>(mr, 1 << 24);
    cudaMemcpy(&h, mr, 4, cudaMemcpyDeviceToHost);
    printf("saturating-FMA polynomial exp2 max relative error over x in [-140, 140] (clamped at +-126): %.3e\n", h);
    run3<2>(out, in, cyc); run3<3>(out, in, cyc); run3<4>(out, in, cyc); run3<5>(out, in, cyc); run3<6>(out, in, cyc); run3<8>(out, in, cyc);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
