// region.cu — K9: the masked-region variant of the path (reference
// src/perception/siglip_semantic_encoder.py:485-562).  Bounding-box crops of ONE resident frame are resized with
// Pillow's bicubic resampler (integer two-pass, 22-bit fixed-point coefficients, uint8 intermediate — bit-exact),
// normalised through a 3 x 256 table, zero-padded to the batch canvas and written directly as bf16 im2col patch rows;
// the learned position table is re-sampled to the canvas grid the way HF's `interpolate_pos_encoding` does; the
// reference's `mean` / `max` pooling over the tokens.  The tower itself is gvl_siglip_forward on a pack whose T / pos
// describe the canvas grid.
#include "common.cuh"

#include <cmath>
#include <vector>

namespace gvl {

constexpr int PIL_PRECISION_BITS = 32 - 8 - 2;  // Pillow src/libImaging/Resample.c
constexpr int DESC_INTS = GVL_REGION_DESC_INTS;  // caller fields
constexpr int DEV_DESC_INTS = 12;                // + [10] tmp offset (bytes / 4), [11] first patch row of the region

// Pillow `bicubic_filter` (a = -0.5)
static inline double pil_bicubic(double x) {
    const double a = -0.5;
    if (x < 0.0) x = -x;
    if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
    if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
    return 0.0;
}

__device__ __forceinline__ int clip8(int v, int precision) {
    v >>= precision;
    return v < 0 ? 0 : (v > 255 ? 255 : v);
}

// Horizontal pass: tmp[row, xx, c] = clip8(2^21 + sum_t frame[y1 + row, x1 + xmin[xx] + t, c] * k[xx, t]).
// grid (blocks over ch * ow, R); one thread per output pixel (3 channels).
__global__ void __launch_bounds__(256)
region_h_kernel(const uint8_t* __restrict__ frame, int W, const int32_t* __restrict__ desc, const int32_t* __restrict__ tabs,
                uint8_t* __restrict__ scratch, int precision) {
    const int32_t* d = desc + blockIdx.y * DEV_DESC_INTS;
    const int x1 = d[0], y1 = d[1], cw = d[2], ch = d[3], ow = d[4], kh = d[6];
    const int32_t* xmin = tabs + d[8];
    const int32_t* cnt = xmin + ow;
    const int32_t* kk = cnt + ow;
    uint8_t* tmp = scratch + (size_t)d[10] * 4;
    const int total = ch * ow;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int row = i / ow, xx = i - row * ow;
        // the tables live in device memory the host cannot inspect: a window is clipped to the crop, never trusted
        const int x0 = min(max(xmin[xx], 0), cw - 1);
        const int n = min(min(cnt[xx], kh), cw - x0);
        const uint8_t* src = frame + ((size_t)(y1 + row) * W + x1 + x0) * 3;
        const int32_t* k = kk + (size_t)xx * kh;
        int s0 = 1 << (precision - 1), s1 = s0, s2 = s0;
        for (int t = 0; t < n; ++t) {
            const int w = k[t];
            s0 += (int)src[3 * t + 0] * w;
            s1 += (int)src[3 * t + 1] * w;
            s2 += (int)src[3 * t + 2] * w;
        }
        uint8_t* o = tmp + (size_t)i * 3;
        o[0] = (uint8_t)clip8(s0, precision);
        o[1] = (uint8_t)clip8(s1, precision);
        o[2] = (uint8_t)clip8(s2, precision);
    }
}

// Vertical pass + normalisation table + zero padding + im2col.  One thread per canvas pixel; the patch buffer was
// zeroed beforehand, so threads outside the region's oh x ow rectangle have nothing to write.
__global__ void __launch_bounds__(256)
region_v_kernel(const int32_t* __restrict__ desc, const int32_t* __restrict__ tabs, const uint8_t* __restrict__ scratch,
                const uint16_t* __restrict__ lut, int canvas_h, int canvas_w, int patch, int ld,
                uint16_t* __restrict__ patches, uint8_t* __restrict__ resized, int tokens_per_region, int precision,
                const float* __restrict__ lut_f32, float* __restrict__ f32_chw) {
    const int32_t* d = desc + blockIdx.y * DEV_DESC_INTS;
    const int ch = d[3], ow = d[4], oh = d[5], kv = d[7];
    const int32_t* ymin = tabs + d[9];
    const int32_t* cnt = ymin + oh;
    const int32_t* kk = cnt + oh;
    const uint8_t* tmp = scratch + (size_t)d[10] * 4;
    const int total = oh * ow;
    // canvas_w == 0: ragged output, every region on its own canvas (its out_w x out_h), rows from d[11]
    const int gw = (canvas_w > 0 ? canvas_w : ow) / patch, pp = patch * patch;
    const size_t row0 = (size_t)d[11];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int y = i / ow, x = i - y * ow;
        const int y0 = min(max(ymin[y], 0), ch - 1);
        const int n = min(min(cnt[y], kv), ch - y0);
        const uint8_t* src = tmp + ((size_t)y0 * ow + x) * 3;
        const int32_t* k = kk + (size_t)y * kv;
        int s0 = 1 << (precision - 1), s1 = s0, s2 = s0;
        for (int t = 0; t < n; ++t) {
            const int w = k[t];
            const uint8_t* p = src + (size_t)t * ow * 3;
            s0 += (int)p[0] * w;
            s1 += (int)p[1] * w;
            s2 += (int)p[2] * w;
        }
        const int v0 = clip8(s0, precision), v1 = clip8(s1, precision), v2 = clip8(s2, precision);
        if (f32_chw) {  // HF `pixel_values` layout [R, 3, canvas_h, canvas_w]
            float* f = f32_chw + ((size_t)blockIdx.y * 3 * canvas_h + y) * canvas_w + x;
            f[0] = lut_f32[v0];
            f[(size_t)canvas_h * canvas_w] = lut_f32[256 + v1];
            f[2 * (size_t)canvas_h * canvas_w] = lut_f32[512 + v2];
        }
        if (resized) {
            uint8_t* u = resized + (((size_t)blockIdx.y * canvas_h + y) * canvas_w + x) * 3;
            u[0] = (uint8_t)v0;
            u[1] = (uint8_t)v1;
            u[2] = (uint8_t)v2;
        }
        const int py = y / patch, ky = y - py * patch, px = x / patch, kx = x - px * patch;
        // a canvas that is not a whole number of patches loses its last rows / columns (a valid stride-p convolution)
        if (patches && px < gw && (canvas_w == 0 || py < canvas_h / patch)) {
            uint16_t* o = patches + (row0 + (size_t)py * gw + px) * ld + ky * patch + kx;
            o[0] = lut[v0];
            o[pp] = lut[256 + v1];
            o[2 * pp] = lut[512 + v2];
        }
    }
}

// ATen native/UpSample.h `get_cubic_upsample_coefficients` (A = -0.75), fp32 like upsample_bicubic2d on float
__device__ __forceinline__ void cubic_coeffs(float t, float c[4]) {
    const float A = -0.75f;
    const float x0 = t + 1.0f, x1 = t, x2 = 1.0f - t, x3 = 2.0f - t;
    c[0] = ((A * x0 - 5.0f * A) * x0 + 8.0f * A) * x0 - 4.0f * A;
    c[1] = ((A + 2.0f) * x1 - (A + 3.0f)) * x1 * x1 + 1.0f;
    c[2] = ((A + 2.0f) * x2 - (A + 3.0f)) * x2 * x2 + 1.0f;
    c[3] = ((A * x3 - 5.0f * A) * x3 + 8.0f * A) * x3 - 4.0f * A;
}

// out[(oy * gw + ox), d] = sum_i wy[i] * sum_j wx[j] * pos[(clamp(iy - 1 + i) * g + clamp(ix - 1 + j)), d]
// grid (gh * gw tokens), block over D in pairs.
__global__ void __launch_bounds__(256)
pos_interp_kernel(const __nv_bfloat16* __restrict__ pos, int g, int D, int gh, int gw, __nv_bfloat16* __restrict__ out) {
    const int oy = blockIdx.x / gw, ox = blockIdx.x - oy * gw;
    const float sy = (float)g / (float)gh, sx = (float)g / (float)gw;
    const float fy = sy * ((float)oy + 0.5f) - 0.5f, fx = sx * ((float)ox + 0.5f) - 0.5f;
    const float fy0 = floorf(fy), fx0 = floorf(fx);
    float wy[4], wx[4];
    cubic_coeffs(fy - fy0, wy);
    cubic_coeffs(fx - fx0, wx);
    int iy[4], ix[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        iy[i] = min(max((int)fy0 - 1 + i, 0), g - 1);
        ix[i] = min(max((int)fx0 - 1 + i, 0), g - 1);
    }
    for (int dcol = threadIdx.x; dcol < D; dcol += blockDim.x) {
        float acc = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float r = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) r += wx[j] * __bfloat162float(pos[((size_t)iy[i] * g + ix[j]) * D + dcol]);
            acc += wy[i] * r;
        }
        out[(size_t)blockIdx.x * D + dcol] = __float2bfloat16_rn(acc);
    }
}

// x: bf16 [B, T, D] -> out [B, D] = max over T (the reference's `sequence.max(dim=1)[0]`).  Same slab layout as
// mean_tokens_kernel (videomae.cu).
__global__ void __launch_bounds__(256)
max_tokens_kernel(const __nv_bfloat16* __restrict__ x, int T, int D, void* __restrict__ out, int out_f32) {
    __shared__ float part[8][256];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c0 = blockIdx.x * 256 + lane * 8;
    const int b = blockIdx.y;
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = -INFINITY;
    if (c0 < D) {
        const __nv_bfloat16* base = x + (size_t)b * T * D + c0;
        for (int t = warp; t < T; t += 8) {
            const uint4 v = *reinterpret_cast<const uint4*>(base + (size_t)t * D);
            acc[0] = fmaxf(acc[0], bf16_lo(v.x)); acc[1] = fmaxf(acc[1], bf16_hi(v.x));
            acc[2] = fmaxf(acc[2], bf16_lo(v.y)); acc[3] = fmaxf(acc[3], bf16_hi(v.y));
            acc[4] = fmaxf(acc[4], bf16_lo(v.z)); acc[5] = fmaxf(acc[5], bf16_hi(v.z));
            acc[6] = fmaxf(acc[6], bf16_lo(v.w)); acc[7] = fmaxf(acc[7], bf16_hi(v.w));
        }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) part[warp][lane * 8 + e] = acc[e];
    __syncthreads();
    const int c = blockIdx.x * 256 + threadIdx.x;
    if (c < D) {
        float s = -INFINITY;
#pragma unroll
        for (int w = 0; w < 8; ++w) s = fmaxf(s, part[w][threadIdx.x]);
        if (out_f32)
            reinterpret_cast<float*>(out)[(size_t)b * D + c] = s;
        else
            reinterpret_cast<__nv_bfloat16*>(out)[(size_t)b * D + c] = __float2bfloat16_rn(s);
    }
}

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// scratch layout: device descriptors, then one uint8 [ch, ow, 3] intermediate per region
static size_t region_layout(int R, const int32_t* h_desc, std::vector<int32_t>* dev_desc, int tokens_per_region = 0,
                            int patch = 0) {
    size_t off = align_up((size_t)R * DEV_DESC_INTS * sizeof(int32_t), 256);
    long long first_row = 0;  // uniform canvas: r * tokens_per_region; ragged: running sum of the regions' own grids
    if (dev_desc) dev_desc->assign((size_t)R * DEV_DESC_INTS, 0);
    for (int r = 0; r < R; ++r) {
        const int32_t* d = h_desc + (size_t)r * DESC_INTS;
        if (dev_desc) {
            int32_t* o = dev_desc->data() + (size_t)r * DEV_DESC_INTS;
            for (int i = 0; i < DESC_INTS; ++i) o[i] = d[i];
            o[10] = (int32_t)(off / 4);
            o[11] = (int32_t)first_row;
        }
        first_row += tokens_per_region > 0 ? tokens_per_region : (patch > 0 ? (long long)(d[4] / patch) * (d[5] / patch) : 0);
        off = align_up(off + (size_t)d[3] * d[4] * 3, 256);
    }
    return off;
}

}  // namespace gvl

extern "C" int gvl_pil_bicubic_taps(int in_size, int out_size, int max_taps, int32_t* h_xmin, int32_t* h_count,
                                    int32_t* h_coeffs, int* h_ksize) {
    using namespace gvl;
    GVL_CHECK_ARG(in_size > 0 && out_size > 0, "gvl_pil_bicubic_taps: bad sizes %d -> %d", in_size, out_size);
    // Resample.c `precompute_coeffs` with box = (0, in_size)
    double scale = (double)in_size / out_size, filterscale = scale;
    if (filterscale < 1.0) filterscale = 1.0;
    const double support = 2.0 * filterscale;
    const int ksize = (int)ceil(support) * 2 + 1;
    if (h_ksize) *h_ksize = ksize;
    if (!h_xmin && !h_count && !h_coeffs) return 0;  // size query
    GVL_CHECK_ARG(h_xmin && h_count && h_coeffs, "gvl_pil_bicubic_taps: null table pointer");
    GVL_CHECK_ARG(max_taps >= ksize, "gvl_pil_bicubic_taps: max_taps %d < ksize %d", max_taps, ksize);
    const double ss = 1.0 / filterscale;
    std::vector<double> k((size_t)ksize);
    for (int xx = 0; xx < out_size; ++xx) {
        const double center = 0.0 + (xx + 0.5) * scale;
        double ww = 0.0;
        int xmin = (int)(center - support + 0.5);
        if (xmin < 0) xmin = 0;
        int xmax = (int)(center + support + 0.5);
        if (xmax > in_size) xmax = in_size;
        xmax -= xmin;
        for (int x = 0; x < xmax; ++x) {
            const double w = pil_bicubic((x + xmin - center + 0.5) * ss);
            k[x] = w;
            ww += w;
        }
        for (int x = 0; x < xmax; ++x)
            if (ww != 0.0) k[x] /= ww;
        int32_t* out = h_coeffs + (size_t)xx * max_taps;
        for (int x = 0; x < max_taps; ++x) {
            if (x >= xmax) {
                out[x] = 0;
            } else if (k[x] < 0) {  // `normalize_coeffs_8bpc`
                out[x] = (int)(-0.5 + k[x] * (1 << PIL_PRECISION_BITS));
            } else {
                out[x] = (int)(0.5 + k[x] * (1 << PIL_PRECISION_BITS));
            }
        }
        h_xmin[xx] = xmin;
        h_count[xx] = xmax;
    }
    return 0;
}

extern "C" size_t gvl_region_scratch_bytes(int R, const int32_t* h_desc) {
    if (R <= 0 || !h_desc) return 0;
    return gvl::region_layout(R, h_desc, nullptr);
}

namespace gvl {

// Integer two-pass resize of R crops of one frame with caller-supplied tap tables:
//   out = clip8((2^(p-1) + sum px * k) >> p), horizontal pass (precision ph) into a uint8 intermediate, then vertical (pv).
static int two_pass_launch(const char* who, const uint8_t* frame, int H, int W, int R, const int32_t* h_desc,
                           const int32_t* tabs, long long tabs_ints, int ph, int pv, bool pil_tables, const uint16_t* lut,
                           const float* lut_f32, int canvas_h, int canvas_w, int patch, int ld, void* patches,
                           uint8_t* resized_u8, float* f32_chw, void* scratch, size_t scratch_bytes, void* stream) {
    GVL_CHECK_ARG(frame && h_desc && tabs && scratch && (patches || resized_u8 || f32_chw) && (lut || !patches) &&
                      (lut_f32 || !f32_chw), "%s: null pointer", who);
    GVL_CHECK_ARG(H > 0 && W > 0 && R > 0 && R <= 65535, "%s: bad shape H=%d W=%d R=%d", who, H, W, R);
    GVL_CHECK_ARG(ph >= 1 && ph <= 22 && pv >= 1 && pv <= 22, "%s: bad precision %d / %d", who, ph, pv);
    const bool ragged = canvas_h == 0 && canvas_w == 0;  // every region on its own canvas, patch rows back to back
    GVL_CHECK_ARG(patch > 0 && (ragged || (canvas_h > 0 && canvas_w > 0 && (!patches || (canvas_h >= patch && canvas_w >= patch)))),
                  "%s: canvas %dx%d smaller than one %d-pixel patch", who, canvas_h, canvas_w, patch);
    GVL_CHECK_ARG(!ragged || (patches && !resized_u8 && !f32_chw), "%s: the ragged form writes patch rows only", who);
    GVL_CHECK_ARG(!patches || (ld >= 3 * patch * patch && ld % 8 == 0), "%s: bad ld %d", who, ld);
    GVL_CHECK_ARG((uintptr_t)scratch % 256 == 0 && (uintptr_t)patches % 16 == 0, "%s: misaligned buffer", who);
    long long total_rows = 0;
    for (int r = 0; r < R; ++r) {
        const int32_t* d = h_desc + (size_t)r * DESC_INTS;
        const int x1 = d[0], y1 = d[1], cw = d[2], ch = d[3], ow = d[4], oh = d[5], kh = d[6], kv = d[7];
        GVL_CHECK_ARG(cw > 0 && ch > 0 && x1 >= 0 && y1 >= 0 && x1 + cw <= W && y1 + ch <= H,
                      "%s: region %d box (%d,%d)+(%dx%d) leaves the %dx%d frame", who, r, x1, y1, cw, ch, W, H);
        GVL_CHECK_ARG(ow > 0 && oh > 0 && (ragged ? (ow % patch == 0 && oh % patch == 0) : (ow <= canvas_w && oh <= canvas_h)),
                      "%s: region %d target %dx%d exceeds the %dx%d canvas (ragged form: must be a multiple of the patch "
                      "size)", who, r, ow, oh, canvas_w, canvas_h);
        total_rows += ragged ? (long long)(ow / patch) * (oh / patch) : (long long)(canvas_h / patch) * (canvas_w / patch);
        int need_kh = 1, need_kv = 1;
        if (pil_tables) {
            gvl_pil_bicubic_taps(cw, ow, 0, nullptr, nullptr, nullptr, &need_kh);
            gvl_pil_bicubic_taps(ch, oh, 0, nullptr, nullptr, nullptr, &need_kv);
        }
        GVL_CHECK_ARG(kh >= need_kh && kv >= need_kv, "%s: region %d tap strides %d/%d < %d/%d", who, r, kh, kv, need_kh, need_kv);
        GVL_CHECK_ARG(d[8] >= 0 && d[9] >= 0 && (long long)d[8] + (long long)ow * (2 + kh) <= tabs_ints &&
                          (long long)d[9] + (long long)oh * (2 + kv) <= tabs_ints,
                      "%s: region %d tables leave the %lld-int table buffer", who, r, tabs_ints);
    }
    GVL_CHECK_ARG(total_rows <= 2147483647LL, "%s: %lld patch rows", who, total_rows);
    const int gh = ragged ? 0 : canvas_h / patch, gw = ragged ? 0 : canvas_w / patch;
    std::vector<int32_t> dev_desc;
    const size_t need = region_layout(R, h_desc, &dev_desc, gh * gw, patch);
    GVL_CHECK_ARG(scratch_bytes >= need, "%s: scratch %zu < required %zu bytes", who, scratch_bytes, need);
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    // pageable source: the copy is staged before the call returns, so dev_desc may go out of scope
    GVL_CUDA(cudaMemcpyAsync(scratch, dev_desc.data(), dev_desc.size() * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    if (patches) GVL_CUDA(cudaMemsetAsync(patches, 0, (size_t)total_rows * ld * 2, s));
    if (resized_u8) GVL_CUDA(cudaMemsetAsync(resized_u8, 0, (size_t)R * canvas_h * canvas_w * 3, s));
    if (f32_chw) GVL_CUDA(cudaMemsetAsync(f32_chw, 0, (size_t)R * canvas_h * canvas_w * 3 * sizeof(float), s));
    long long max_h = 0, max_v = 0, src_bytes = 0;
    for (int r = 0; r < R; ++r) {
        const int32_t* d = h_desc + (size_t)r * DESC_INTS;
        max_h = std::max(max_h, (long long)d[3] * d[4]);
        max_v = std::max(max_v, (long long)d[4] * d[5]);
        src_bytes += (long long)d[2] * d[3] * 3;
    }
    ProfScope prof(GVL_K_PREPROCESS, (double)src_bytes + (double)total_rows * ld * 2, s);
    const int32_t* desc = reinterpret_cast<const int32_t*>(scratch);
    uint8_t* sc = reinterpret_cast<uint8_t*>(scratch);
    const int bh = (int)std::min<long long>((max_h + 255) / 256, 4096), bv = (int)std::min<long long>((max_v + 255) / 256, 4096);
    region_h_kernel<<<dim3(bh, R), 256, 0, s>>>(frame, W, desc, tabs, sc, ph);
    GVL_LAUNCH_CHECK("region_h_kernel");
    region_v_kernel<<<dim3(bv, R), 256, 0, s>>>(desc, tabs, sc, lut, canvas_h, canvas_w, patch, ld,
                                               reinterpret_cast<uint16_t*>(patches), resized_u8, gh * gw, pv, lut_f32, f32_chw);
    GVL_LAUNCH_CHECK("region_v_kernel");
    return 0;
}

}  // namespace gvl

extern "C" int gvl_region_patches_pil_u8(const uint8_t* frame, int H, int W, int R, const int32_t* h_desc,
                                         const int32_t* tabs, long long tabs_ints, const uint16_t* lut, int canvas_h,
                                         int canvas_w, int patch, int ld, void* patches, uint8_t* resized_u8,
                                         void* scratch, size_t scratch_bytes, void* stream) {
    return gvl::two_pass_launch("gvl_region_patches_pil_u8", frame, H, W, R, h_desc, tabs, tabs_ints, gvl::PIL_PRECISION_BITS,
                                gvl::PIL_PRECISION_BITS, true, lut, nullptr, canvas_h, canvas_w, patch, ld, patches, resized_u8,
                                nullptr, scratch, scratch_bytes, stream);
}

extern "C" int gvl_resize_two_pass_u8(const uint8_t* frame, int H, int W, int R, const int32_t* h_desc, const int32_t* tabs,
                                      long long tabs_ints, int precision_h, int precision_v, const uint16_t* lut_bf16,
                                      const float* lut_f32, int canvas_h, int canvas_w, int patch, int ld, void* patches,
                                      uint8_t* resized_u8, float* f32_chw, void* scratch, size_t scratch_bytes,
                                      void* stream) {
    return gvl::two_pass_launch("gvl_resize_two_pass_u8", frame, H, W, R, h_desc, tabs, tabs_ints, precision_h, precision_v,
                                false, lut_bf16, lut_f32, canvas_h, canvas_w, patch, ld, patches, resized_u8, f32_chw,
                                scratch, scratch_bytes, stream);
}

extern "C" int gvl_pos_interp_bicubic_bf16(const void* pos, int g, int D, int gh, int gw, void* out, void* stream) {
    using namespace gvl;
    GVL_CHECK_ARG(pos && out, "gvl_pos_interp_bicubic_bf16: null pointer");
    GVL_CHECK_ARG(g > 0 && D > 0 && gh > 0 && gw > 0 && (long long)gh * gw <= 2147483647LL / D,
                  "gvl_pos_interp_bicubic_bf16: bad shape g=%d D=%d gh=%d gw=%d", g, D, gh, gw);
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    pos_interp_kernel<<<gh * gw, 256, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(pos), g, D, gh, gw,
                                              reinterpret_cast<__nv_bfloat16*>(out));
    GVL_LAUNCH_CHECK("pos_interp_kernel");
    return 0;
}

extern "C" int gvl_max_tokens_bf16(const void* x, int B, int T, int D, void* out, int out_f32, void* stream) {
    using namespace gvl;
    GVL_CHECK_ARG(x && out, "gvl_max_tokens_bf16: null pointer");
    GVL_CHECK_ARG(B > 0 && B <= 65535 && T > 0 && D > 0 && D % 8 == 0, "gvl_max_tokens_bf16: bad shape B=%d T=%d D=%d", B, T,
                  D);
    GVL_CHECK_ARG((uintptr_t)x % 16 == 0, "gvl_max_tokens_bf16: x must be 16-byte aligned");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    ProfScope prof(GVL_K_LAYERNORM, (double)B * T * D * 2, s);
    max_tokens_kernel<<<dim3((D + 255) / 256, B), 256, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(x), T, D, out, out_f32);
    GVL_LAUNCH_CHECK("max_tokens_kernel");
    return 0;
}
