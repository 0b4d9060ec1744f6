// preprocess.cu — K1: fused uint8 frame resize (antialiased, integer two-pass) + normalize + patchify.
//
// Bit-exact restatement of what HF's SiglipImageProcessor (torchvision backend) does on CPU
// (ATen:native/cpu/UpSampleKernelAVXAntialias.h:304-437): a horizontal pass with int16 fixed-point
// weights into a uint8 intermediate, then a vertical pass, each `clip((2^(p-1) + sum px*w) >> p)`;
// then `(float(u8) - sub) / div` in fp32 (IEEE division) and, for the patch layout, a round-to-nearest
// bf16 cast scattered into im2col order for the stride-P patch-embedding GEMM.
//
// One CTA produces a tile of TY output rows x TX output columns.  The source rows it needs stream
// through a double-buffered cp.async staging ring in groups of RG rows (16-byte coalesced loads; each
// frame byte is fetched from HBM once, the ~7 % halo between vertically adjacent tiles hits L2); the
// horizontal pass writes the uint8 intermediate to shared memory, the vertical pass reads it back
// and the finished tile is staged in its final element type so the global stores are full
// contiguous segments (whole 592-element patch rows / whole CHW row pieces).
#include "common.cuh"

#include <cmath>
#include <map>
#include <mutex>
#include <tuple>
#include <vector>

namespace gvl {

// ---- host: tap tables (ATen:native/cpu/UpSampleKernel.cpp, _compute_index_ranges_int16_weights) ----

static double aa_filter(double x, int resample) {
    x = std::fabs(x);
    if (resample == GVL_RESAMPLE_BILINEAR) return x < 1.0 ? 1.0 - x : 0.0;
    const double a = -0.5;
    if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1.0;
    if (x < 2.0) return (((x - 5.0) * x + 8.0) * x - 4.0) * a;
    return 0.0;
}

struct AxisTaps {
    std::vector<int32_t> xmin, xsize;
    std::vector<int16_t> w;  // [out][taps]
    int taps = 0, precision = 0;
};

static int compute_axis_taps(int in_size, int out_size, int resample, AxisTaps& t) {
    if (in_size <= 0 || out_size <= 0) return 1;
    if (resample != GVL_RESAMPLE_BILINEAR && resample != GVL_RESAMPLE_BICUBIC) return 1;
    const int interp_size = resample == GVL_RESAMPLE_BILINEAR ? 2 : 4;
    const double scale = (double)in_size / (double)out_size;
    const double support = scale >= 1.0 ? (interp_size * 0.5) * scale : interp_size * 0.5;
    const int max_interp = (int)std::ceil(support) * 2 + 1;
    const double invscale = scale >= 1.0 ? 1.0 / scale : 1.0;
    std::vector<double> wf((size_t)out_size * max_interp, 0.0);
    t.xmin.assign(out_size, 0);
    t.xsize.assign(out_size, 0);
    double wt_max = 0.0;
    for (int i = 0; i < out_size; ++i) {
        const double center = scale * (i + 0.5);
        long long xmin = (long long)(center - support + 0.5);
        if (xmin < 0) xmin = 0;
        long long xend = (long long)(center + support + 0.5);
        if (xend > in_size) xend = in_size;
        long long xsize = xend - xmin;
        if (xsize < 0) xsize = 0;
        if (xsize > max_interp) xsize = max_interp;
        double* wp = wf.data() + (size_t)i * max_interp;
        double total = 0.0;
        for (long long j = 0; j < xsize; ++j) {
            const double w = aa_filter((double)(j + xmin - center + 0.5) * invscale, resample);
            wp[j] = w;
            total += w;
        }
        if (total != 0.0) {
            for (long long j = 0; j < xsize; ++j) {
                wp[j] /= total;
                if (wp[j] > wt_max) wt_max = wp[j];
            }
        }
        t.xmin[i] = (int32_t)xmin;
        t.xsize[i] = (int32_t)xsize;
    }
    int precision = 0;
    for (precision = 0; precision < 22; ++precision) {
        const int next_value = (int)(0.5 + wt_max * (double)(1 << (precision + 1)));
        if (next_value >= (1 << 15)) break;
    }
    t.precision = precision;
    t.taps = max_interp;
    t.w.assign((size_t)out_size * max_interp, 0);
    for (size_t k = 0; k < wf.size(); ++k) {
        const double v = wf[k] * (double)(1 << precision);
        t.w[k] = (int16_t)(v < 0 ? (int)(-0.5 + v) : (int)(0.5 + v));
    }
    return 0;
}

// ---- device tables, cached per geometry ----

struct DevTables {
    int32_t *h_min = nullptr, *h_size = nullptr, *v_min = nullptr, *v_size = nullptr;
    int16_t *h_w = nullptr, *v_w = nullptr;
    int h_taps = 0, v_taps = 0, h_prec = 0, v_prec = 0;
    int max_seg_bytes = 0;  // staging bytes per source row (16-byte aligned span) over all x tiles
    int max_rows = 0;       // source rows per y tile
    int TX = 0, TY = 0;
};

static std::mutex g_tab_mu;
typedef std::tuple<int, int, int, int, int, int, int, int, int, int> TabKey;
static std::map<TabKey, DevTables> g_tabs;

template <typename T>
static int upload(const std::vector<T>& v, T** dptr) {
    GVL_CUDA(cudaMalloc((void**)dptr, v.size() * sizeof(T)));
    GVL_CUDA(cudaMemcpy(*dptr, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return 0;
}

static int get_tables(int H, int W, int eff_h, int eff_w, int out_h, int out_w, int resample, int TX, int TY,
                      DevTables& out) {
    int dev = 0;
    GVL_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_tab_mu);
    TabKey key = std::make_tuple(dev, H, W, out_h, out_w, resample, TX, TY, eff_h, eff_w);
    auto it = g_tabs.find(key);
    if (it != g_tabs.end()) {
        out = it->second;
        return 0;
    }
    AxisTaps th, tv;
    if (compute_axis_taps(W, out_w, resample, th) || compute_axis_taps(H, out_h, resample, tv)) {
        set_error("gvl_preprocess_u8: bad resize geometry %dx%d -> %dx%d resample %d", H, W, out_h, out_w, resample);
        return 1;
    }
    DevTables d;
    d.h_taps = th.taps;
    d.v_taps = tv.taps;
    d.h_prec = th.precision;
    d.v_prec = tv.precision;
    d.TX = TX;
    d.TY = TY;
    for (int x0 = 0; x0 < eff_w; x0 += TX) {
        const int x1 = std::min(x0 + TX, eff_w);
        const int c_lo = th.xmin[x0], c_hi = th.xmin[x1 - 1] + th.xsize[x1 - 1];
        const int a_lo = (3 * c_lo) & ~15, a_hi = (3 * c_hi + 15) & ~15;
        d.max_seg_bytes = std::max(d.max_seg_bytes, a_hi - a_lo);
    }
    for (int y0 = 0; y0 < eff_h; y0 += TY) {
        const int y1 = std::min(y0 + TY, eff_h);
        d.max_rows = std::max(d.max_rows, tv.xmin[y1 - 1] + tv.xsize[y1 - 1] - tv.xmin[y0]);
    }
    int rc = upload(th.xmin, &d.h_min) || upload(th.xsize, &d.h_size) || upload(th.w, &d.h_w) ||
             upload(tv.xmin, &d.v_min) || upload(tv.xsize, &d.v_size) || upload(tv.w, &d.v_w);
    if (rc) return 2;
    g_tabs[key] = d;
    out = d;
    return 0;
}

// ---- kernel ----

constexpr int PRE_THREADS = 256;
constexpr int PRE_RG = 8;        // source rows per staging group
constexpr int PRE_MAXT_H = 24;   // max horizontal taps held in registers

struct PreParams {
    const uint8_t* frames;
    int B, H, W;
    int out_h, out_w;  // resize target
    int eff_h, eff_w;  // rows / cols actually produced (patch layout drops the remainder)
    const int32_t *h_min, *h_size, *v_min, *v_size;
    const int16_t *h_w, *v_w;
    int h_taps, v_taps, h_prec, v_prec;
    int TX, TY;
    int seg_stride;  // bytes per staged source row (multiple of 16, >= max_seg_bytes + 16*4)
    int hs_stride;   // bytes per intermediate row (multiple of 4)
    int max_rows;
    float sub[3], div[3];
    void* out;
    int layout, patch, ld, gh, gw;
    int aligned16;  // frame rows are 16-byte aligned -> cp.async path
};

__device__ __forceinline__ void pre_cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}

template <int MAXT>
__global__ void __launch_bounds__(PRE_THREADS)
preprocess_kernel(const PreParams p) {
    extern __shared__ __align__(16) uint8_t pre_smem[];
    // layout: [2][RG][seg_stride] staging | [max_rows][hs_stride] intermediate | output tile
    uint8_t* sRaw = pre_smem;
    uint8_t* sH = sRaw + 2 * PRE_RG * p.seg_stride;
    uint8_t* sOut = sH + ((p.max_rows * p.hs_stride + 15) & ~15);

    const int tid = threadIdx.x;
    const int b = blockIdx.z;
    const int x0 = blockIdx.x * p.TX, x1 = min(x0 + p.TX, p.eff_w);
    const int y0 = blockIdx.y * p.TY, y1 = min(y0 + p.TY, p.eff_h);
    const int ntx = x1 - x0, nty = y1 - y0;
    const int r_lo = p.v_min[y0];
    const int r_hi = p.v_min[y1 - 1] + p.v_size[y1 - 1];
    const int nrows = r_hi - r_lo;
    const int c_lo = p.h_min[x0];
    const int c_hi = p.h_min[x1 - 1] + p.h_size[x1 - 1];
    const int a_lo = (3 * c_lo) & ~15;
    const int a_hi = min((3 * c_hi + 15) & ~15, (p.W * 3 + 15) & ~15);
    const int seg_chunks = (a_hi - a_lo) >> 4;
    const size_t row_bytes = (size_t)p.W * 3;
    const uint8_t* fbase = p.frames + (size_t)b * p.H * row_bytes;

    // pad columns [3*P*P, ld) of the staged patch rows are emitted as zero
    if (p.layout == GVL_LAYOUT_BF16_PATCH) {
        const int padc = p.ld - 3 * p.patch * p.patch;
        const int npatch = p.TX / p.patch;
        for (int i = tid; i < npatch * padc; i += PRE_THREADS)
            reinterpret_cast<uint16_t*>(sOut)[(size_t)(i / padc) * p.ld + 3 * p.patch * p.patch + i % padc] = 0;
    }

    auto stage_group = [&](int g, int buf) {
        const int rbeg = r_lo + g * PRE_RG;
        const int nr = min(PRE_RG, r_hi - rbeg);
        uint8_t* dst = sRaw + (size_t)buf * PRE_RG * p.seg_stride;
        if (p.aligned16) {
            for (int i = tid; i < nr * seg_chunks; i += PRE_THREADS) {
                const int rr = i / seg_chunks, ch = i - rr * seg_chunks;
                pre_cp_async16(dst + rr * p.seg_stride + ch * 16, fbase + (size_t)(rbeg + rr) * row_bytes + a_lo + ch * 16);
            }
        } else {
            const int nbytes = min(a_hi, (int)row_bytes) - a_lo;
            for (int i = tid; i < nr * nbytes; i += PRE_THREADS) {
                const int rr = i / nbytes, o = i - rr * nbytes;
                dst[rr * p.seg_stride + o] = fbase[(size_t)(rbeg + rr) * row_bytes + a_lo + o];
            }
        }
    };

    // per-thread horizontal taps: this thread always produces column x0 + (tid % TXP)
    const int lanes_x = p.TX;               // threads are laid out as (row-in-group, x)
    const int rows_par = PRE_THREADS / lanes_x > 0 ? PRE_THREADS / lanes_x : 1;
    const int my_xx = tid % lanes_x;
    const int my_rr0 = tid / lanes_x;
    int hw[MAXT];
    int my_off = 0;
    const bool x_active = my_xx < ntx && my_rr0 < rows_par;
    {
        const int x = x0 + (x_active ? my_xx : 0);
        const int16_t* wp = p.h_w + (size_t)x * p.h_taps;
#pragma unroll
        for (int j = 0; j < MAXT; ++j) hw[j] = j < p.h_taps ? (int)wp[j] : 0;
        my_off = 3 * p.h_min[x] - a_lo;
    }

    const int ngroups = (nrows + PRE_RG - 1) / PRE_RG;
    stage_group(0, 0);
    asm volatile("cp.async.commit_group;" ::: "memory");
    const int h_round = 1 << (p.h_prec - 1);
    for (int g = 0; g < ngroups; ++g) {
        const int buf = g & 1;
        if (g + 1 < ngroups) stage_group(g + 1, buf ^ 1);
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        __syncthreads();
        const int nr = min(PRE_RG, nrows - g * PRE_RG);
        if (x_active) {
            for (int rr = my_rr0; rr < nr; rr += rows_par) {
                const uint8_t* src = sRaw + ((size_t)buf * PRE_RG + rr) * p.seg_stride + my_off;
                int a0 = h_round, a1 = h_round, a2 = h_round;
#pragma unroll
                for (int j = 0; j < MAXT; ++j) {
                    a0 += (int)src[3 * j + 0] * hw[j];
                    a1 += (int)src[3 * j + 1] * hw[j];
                    a2 += (int)src[3 * j + 2] * hw[j];
                }
                uint8_t* d = sH + (size_t)(g * PRE_RG + rr) * p.hs_stride + my_xx * 3;
                d[0] = (uint8_t)min(max(a0 >> p.h_prec, 0), 255);
                d[1] = (uint8_t)min(max(a1 >> p.h_prec, 0), 255);
                d[2] = (uint8_t)min(max(a2 >> p.h_prec, 0), 255);
            }
        }
        __syncthreads();
    }

    // ---- vertical pass + normalize, 4 consecutive (x,c) bytes per work item ----
    const int words = (ntx * 3 + 3) >> 2;
    const int v_round = 1 << (p.v_prec - 1);
    const int esize = p.layout == GVL_LAYOUT_U8_CHW ? 1 : (p.layout == GVL_LAYOUT_F32_CHW ? 4 : 2);
    const int PP = p.patch * p.patch;
    for (int item = tid; item < nty * words; item += PRE_THREADS) {
        const int yy = item / words, wi = item - yy * words;
        const int y = y0 + yy;
        const int rbeg = p.v_min[y] - r_lo;
        const int n = p.v_size[y];
        const int16_t* wv = p.v_w + (size_t)y * p.v_taps;
        int acc[4] = {v_round, v_round, v_round, v_round};
        for (int j = 0; j < n; ++j) {
            const uint32_t px = *reinterpret_cast<const uint32_t*>(sH + (size_t)(rbeg + j) * p.hs_stride + wi * 4);
            const int w = (int)wv[j];
            acc[0] += (int)(px & 0xFF) * w;
            acc[1] += (int)((px >> 8) & 0xFF) * w;
            acc[2] += (int)((px >> 16) & 0xFF) * w;
            acc[3] += (int)(px >> 24) * w;
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int f = wi * 4 + e;
            if (f >= ntx * 3) break;
            const int xx = f / 3, c = f - xx * 3;
            const int u = min(max(acc[e] >> p.v_prec, 0), 255);
            size_t o;  // element index inside the output tile
            if (p.layout == GVL_LAYOUT_BF16_PATCH) {
                const int x = x0 + xx;
                const int px_i = x / p.patch - x0 / p.patch, kx = x % p.patch, ky = y % p.patch;
                o = (size_t)px_i * p.ld + c * PP + ky * p.patch + kx;
            } else {
                o = ((size_t)c * p.TY + yy) * p.TX + xx;
            }
            if (p.layout == GVL_LAYOUT_U8_CHW) {
                sOut[o] = (uint8_t)u;
            } else {
                const float v = __fdiv_rn((float)u - p.sub[c], p.div[c]);
                if (p.layout == GVL_LAYOUT_F32_CHW)
                    reinterpret_cast<float*>(sOut)[o] = v;
                else
                    reinterpret_cast<__nv_bfloat16*>(sOut)[o] = __float2bfloat16_rn(v);
            }
        }
    }
    __syncthreads();

    // ---- copy-out: contiguous segments ----
    int nseg, seg_elems;
    if (p.layout == GVL_LAYOUT_BF16_PATCH) {
        nseg = ntx / p.patch;
        seg_elems = p.ld;
    } else {
        nseg = 3 * nty;
        seg_elems = ntx;
    }
    const int seg_bytes = seg_elems * esize;
    for (int s = 0; s < nseg; ++s) {
        uint8_t* gdst;
        const uint8_t* ssrc;
        if (p.layout == GVL_LAYOUT_BF16_PATCH) {
            const size_t prow = ((size_t)b * p.gh + y0 / p.patch) * p.gw + x0 / p.patch + s;
            gdst = reinterpret_cast<uint8_t*>(p.out) + prow * p.ld * 2;
            ssrc = sOut + (size_t)s * p.ld * 2;
        } else {
            const int c = s / nty, yy = s - c * nty;
            gdst = reinterpret_cast<uint8_t*>(p.out) +
                   ((((size_t)b * 3 + c) * p.out_h + y0 + yy) * p.out_w + x0) * esize;
            ssrc = sOut + (((size_t)c * p.TY + yy) * p.TX) * esize;
        }
        if (((reinterpret_cast<uintptr_t>(gdst) | (uintptr_t)seg_bytes | reinterpret_cast<uintptr_t>(ssrc)) & 15) == 0) {
            for (int i = tid; i < (seg_bytes >> 4); i += PRE_THREADS)
                reinterpret_cast<uint4*>(gdst)[i] = reinterpret_cast<const uint4*>(ssrc)[i];
        } else {
            for (int i = tid; i < seg_bytes; i += PRE_THREADS) gdst[i] = ssrc[i];
        }
    }
}

// pixel_values fp32 [B,3,H,W] -> bf16 im2col rows (the drop-in `get_image_features(pixel_values=...)` seam).
// One thread per 8 output elements (16-byte stores); reads are 4-byte, L1/L2 resident.
__global__ void __launch_bounds__(256)
patchify_f32_kernel(const float* __restrict__ pv, int B, int H, int W, int patch, int ld, int gh, int gw,
                    __nv_bfloat16* __restrict__ out) {
    const int chunks = ld >> 3;
    const size_t total = (size_t)B * gh * gw * chunks;
    const int PP = patch * patch, K = 3 * PP;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int ch = (int)(i % chunks);
        const size_t row = i / chunks;
        const int px = (int)(row % gw), py = (int)((row / gw) % gh), b = (int)(row / ((size_t)gw * gh));
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int k = ch * 8 + e;
            if (k < K) {
                const int c = k / PP, r = k - c * PP, ky = r / patch, kx = r - ky * patch;
                v[e] = pv[(((size_t)b * 3 + c) * H + py * patch + ky) * W + px * patch + kx];
            } else {
                v[e] = 0.f;
            }
        }
        uint4 o;
        o.x = pack_bf16x2(v[0], v[1]);
        o.y = pack_bf16x2(v[2], v[3]);
        o.z = pack_bf16x2(v[4], v[5]);
        o.w = pack_bf16x2(v[6], v[7]);
        reinterpret_cast<uint4*>(out + row * ld)[ch] = o;
    }
}

}  // namespace gvl

extern "C" int gvl_patchify_f32(const float* pixel_values, int B, int H, int W, int patch, int ld, void* out,
                                void* stream) {
    using namespace gvl;
    GVL_CHECK_ARG(pixel_values && out, "gvl_patchify_f32: null pointer");
    GVL_CHECK_ARG(B > 0 && patch >= 1 && H >= patch && W >= patch && ld >= 3 * patch * patch && ld % 8 == 0,
                  "gvl_patchify_f32: bad shape B=%d H=%d W=%d patch=%d ld=%d", B, H, W, patch, ld);
    GVL_CHECK_ARG((uintptr_t)out % 16 == 0, "gvl_patchify_f32: output must be 16-byte aligned");
    const int gh = H / patch, gw = W / patch;
    const size_t total = (size_t)B * gh * gw * (ld / 8);
    int grid = (int)std::min<size_t>((total + 255) / 256, (size_t)sm_count() * 16);
    ProfScope prof(GVL_K_PATCHIFY, (double)B * gh * gw * 3 * patch * patch * 6, reinterpret_cast<cudaStream_t>(stream));
    patchify_f32_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        pixel_values, B, H, W, patch, ld, gh, gw, reinterpret_cast<__nv_bfloat16*>(out));
    GVL_LAUNCH_CHECK("patchify_f32_kernel");
    return 0;
}

extern "C" int gvl_resize_taps(int in_size, int out_size, int resample, int max_taps, int32_t* h_xmin, int32_t* h_xsize,
                               int16_t* h_weights, int* h_precision, int* h_taps_used) {
    using namespace gvl;
    AxisTaps t;
    GVL_CHECK_ARG(compute_axis_taps(in_size, out_size, resample, t) == 0, "gvl_resize_taps: bad arguments");
    GVL_CHECK_ARG(max_taps >= t.taps, "gvl_resize_taps: max_taps %d < required %d", max_taps, t.taps);
    for (int i = 0; i < out_size; ++i) {
        h_xmin[i] = t.xmin[i];
        h_xsize[i] = t.xsize[i];
        for (int j = 0; j < max_taps; ++j) h_weights[(size_t)i * max_taps + j] = j < t.taps ? t.w[(size_t)i * t.taps + j] : 0;
    }
    *h_precision = t.precision;
    *h_taps_used = t.taps;
    return 0;
}

extern "C" int gvl_preprocess_u8(const uint8_t* frames, int B, int H, int W, int out_h, int out_w, int resample,
                                 const float* h_sub, const float* h_div, void* out, int layout, int patch, int ld,
                                 void* stream) {
    using namespace gvl;
    GVL_CHECK_ARG(frames && out && h_sub && h_div, "gvl_preprocess_u8: null pointer");
    GVL_CHECK_ARG(B > 0 && B <= 65535 && H > 0 && W > 0 && out_h > 0 && out_w > 0, "gvl_preprocess_u8: bad shape");
    GVL_CHECK_ARG(layout >= 0 && layout <= 3, "gvl_preprocess_u8: bad layout %d", layout);
    GVL_CHECK_ARG(H >= out_h && W >= out_w, "gvl_preprocess_u8: only downscaling is supported (%dx%d -> %dx%d)", H, W,
                  out_h, out_w);
    PreParams p;
    memset(&p, 0, sizeof(p));
    int TX = 128, TY = 16;
    p.eff_h = out_h;
    p.eff_w = out_w;
    p.gh = p.gw = 0;
    if (layout == GVL_LAYOUT_BF16_PATCH) {
        GVL_CHECK_ARG(patch >= 4 && patch <= 32 && ld >= 3 * patch * patch && ld % 8 == 0,
                      "gvl_preprocess_u8: bad patch/ld %d/%d", patch, ld);
        GVL_CHECK_ARG((uintptr_t)out % 16 == 0, "gvl_preprocess_u8: output must be 16-byte aligned");
        p.gh = out_h / patch;
        p.gw = out_w / patch;
        GVL_CHECK_ARG(p.gh > 0 && p.gw > 0, "gvl_preprocess_u8: output smaller than one patch");
        p.eff_h = p.gh * patch;
        p.eff_w = p.gw * patch;
        TY = patch;
        TX = patch * (128 / patch);
    }
    DevTables tb;
    int rc = get_tables(H, W, p.eff_h, p.eff_w, out_h, out_w, resample, TX, TY, tb);
    if (rc) return rc;
    GVL_CHECK_ARG(tb.h_taps <= PRE_MAXT_H, "gvl_preprocess_u8: %d horizontal taps exceed the kernel limit %d", tb.h_taps,
                  PRE_MAXT_H);
    p.frames = frames;
    p.B = B;
    p.H = H;
    p.W = W;
    p.out_h = out_h;
    p.out_w = out_w;
    p.h_min = tb.h_min;
    p.h_size = tb.h_size;
    p.v_min = tb.v_min;
    p.v_size = tb.v_size;
    p.h_w = tb.h_w;
    p.v_w = tb.v_w;
    p.h_taps = tb.h_taps;
    p.v_taps = tb.v_taps;
    p.h_prec = tb.h_prec;
    p.v_prec = tb.v_prec;
    p.TX = TX;
    p.TY = TY;
    // staged rows are over-read by up to 3*MAXT bytes past the last tap of the last column (zero weights)
    p.seg_stride = (tb.max_seg_bytes + 3 * PRE_MAXT_H + 16 + 15) & ~15;
    p.hs_stride = (TX * 3 + 3) & ~3;
    p.max_rows = tb.max_rows;
    for (int c = 0; c < 3; ++c) {
        p.sub[c] = h_sub[c];
        p.div[c] = h_div[c];
    }
    p.out = out;
    p.layout = layout;
    p.patch = patch;
    p.ld = ld;
    p.aligned16 = ((W * 3) % 16 == 0 && (uintptr_t)frames % 16 == 0) ? 1 : 0;
    const int esize = layout == GVL_LAYOUT_U8_CHW ? 1 : (layout == GVL_LAYOUT_F32_CHW ? 4 : 2);
    size_t out_tile_bytes =
        layout == GVL_LAYOUT_BF16_PATCH ? (size_t)(TX / patch) * ld * 2 : (size_t)3 * TY * TX * esize;
    size_t smem = (size_t)2 * PRE_RG * p.seg_stride + (((size_t)p.max_rows * p.hs_stride + 15) & ~(size_t)15) +
                  out_tile_bytes + 16;
    GVL_CHECK_ARG(smem <= 200 * 1024, "gvl_preprocess_u8: tile needs %zu bytes of shared memory (geometry too large)",
                  smem);
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    dim3 grid((p.eff_w + TX - 1) / TX, (p.eff_h + TY - 1) / TY, B);
    const double out_bytes = layout == GVL_LAYOUT_BF16_PATCH ? (double)p.gh * p.gw * 3 * patch * patch * 2
                                                              : (double)3 * out_h * out_w * esize;
    ProfScope prof(GVL_K_PREPROCESS, (double)B * ((double)H * W * 3 + out_bytes), s);
    auto launch = [&](auto kernel) -> int {
        GVL_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kernel<<<grid, PRE_THREADS, smem, s>>>(p);
        return 0;
    };
    rc = tb.h_taps <= 12 ? launch(preprocess_kernel<12>) : launch(preprocess_kernel<PRE_MAXT_H>);
    if (rc) return rc;
    GVL_LAUNCH_CHECK("preprocess_kernel");
    return 0;
}
