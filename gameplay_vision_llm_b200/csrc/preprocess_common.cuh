// preprocess_common.cuh — host tap tables and the IDP.2A helpers shared by the K1 kernels (preprocess.cu,
// preprocess_stream.cu).
#pragma once
#include "common.cuh"

#include <mutex>
#include <shared_mutex>
#include <vector>

namespace gvl {

// Integer tap table of one axis (ATen:native/cpu/UpSampleKernel.cpp, _compute_index_ranges_int16_weights).
struct AxisTaps {
    std::vector<int32_t> xmin, xsize;
    std::vector<int16_t> w;  // [out][taps]
    int taps = 0, precision = 0;
};
int compute_axis_taps(int in_size, int out_size, int resample, AxisTaps& t);

// (u8 - sub[c]) / div[c] in fp32 (IEEE division on the host), [3][256] on the device, cached per (device, sub, div);
// h_copy (optional) receives the host values
int get_lut(const float* sub, const float* div, float** out, float* h_copy = nullptr);

extern std::mutex g_tab_mu;  // guards every table cache of K1

template <typename T>
int upload(const std::vector<T>& v, T** dptr) {
    GVL_CUDA(cudaMalloc((void**)dptr, v.size() * sizeof(T)));
    GVL_CUDA(cudaMemcpy(*dptr, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return 0;
}

// Table caches are bounded: a stream of differently sized images (crops, mixed-resolution videos) would otherwise
// grow them without limit.  Every K1 entry point first calls trim_table_caches_if_full() and then holds g_tab_rw
// shared while it fetches tables and launches; the trim takes it exclusively, waits for the device and frees every
// table, so no launch can be left holding a freed pointer.
constexpr size_t kTableCacheCap = 64;
extern std::shared_mutex g_tab_rw;
void trim_table_caches_if_full();
size_t stream_table_cache_size();   // preprocess_stream.cu's cache (callers hold g_tab_mu)
void stream_table_cache_clear();

// returns -1 when the geometry is outside the kernel's limits (the caller falls back to the planar kernel)
int launch_stream5(const uint8_t* frames, int B, int H, int W, int out_h, int out_w, int resample, const float* h_sub,
                   const float* h_div, void* out, int patch, int ld, cudaStream_t s);

#ifdef __CUDACC__
// d = acc + w2.lo16 * px4.byte0 + w2.hi16 * px4.byte1   (.hi: bytes 2, 3); weights signed, pixels unsigned
__device__ __forceinline__ int dp2a_lo(uint32_t w2, uint32_t px4, int acc) {
    int d;
    asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(w2), "r"(px4), "r"(acc));
    return d;
}
__device__ __forceinline__ int dp2a_hi(uint32_t w2, uint32_t px4, int acc) {
    int d;
    asm("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(w2), "r"(px4), "r"(acc));
    return d;
}
// {sat_u8(v3), sat_u8(v2), sat_u8(v1), sat_u8(v0)} with v0 in the low byte
__device__ __forceinline__ uint32_t pack4_sat_u8(int v0, int v1, int v2, int v3) {
    uint32_t hi, d;
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(hi) : "r"(v3), "r"(v2), "r"(0));
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(v1), "r"(v0), "r"(hi));
    return d;
}
#endif

}  // namespace gvl
