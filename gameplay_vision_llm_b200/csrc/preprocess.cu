// preprocess.cu — K1: fused uint8 frame resize (antialiased, integer two-pass) + normalize + patchify.
//
// Bit-exact restatement of what HF's SiglipImageProcessor (torchvision backend) does on CPU
// (ATen:native/cpu/UpSampleKernelAVXAntialias.h:304-437): a horizontal pass with int16 fixed-point
// weights into a uint8 intermediate, then a vertical pass, each `clip((2^(p-1) + sum px*w) >> p)`;
// then `(float(u8) - sub) / div` in fp32 (IEEE division) and, for the patch layout, a round-to-nearest
// bf16 cast scattered into im2col order for the stride-P patch-embedding GEMM.
//
// Two kernels.  `preprocess_planar_kernel` is the production path (v2): a CTA owns a tile of TY x TX
// output pixels; the source rows stream through registers in groups of 8 (three coalesced 128-bit loads
// per 16-pixel chunk, next group in flight while the current one is filtered), are de-interleaved with
// byte permutes into channel-planar shared memory, and both filter passes run on IDP.2A (two u8 x s16
// taps per instruction): every thread keeps its column's taps as zero-padded 16-bit weight pairs aligned
// to the 32-bit words it loads, so no per-byte extraction is ever needed.  The horizontal pass writes
// the uint8 intermediate TRANSPOSED (four consecutive source rows of one column per word), which makes
// the vertical taps contiguous bytes as well.  Finished bands are staged in their final element type
// and leave as full contiguous segments (whole 592-element patch rows / whole CHW row pieces).
// `preprocess_kernel` (v1, one byte per shared-memory load) remains as the fallback for geometries
// outside the planar kernel's limits and for A/B runs (GVL_PRE_LEGACY=1).
#include "preprocess_common.cuh"

#include <atomic>
#include <cmath>
#include <cstdlib>
#include <cstring>
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

int compute_axis_taps(int in_size, int out_size, int resample, AxisTaps& t) {
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
    // planar kernel
    float* lut = nullptr;   // [3][256] (u8 - sub) / div, built per call signature below
    int nch_max = 0;        // 16-pixel chunks per staged source row (max over x tiles)
    int max_hsize = 0, max_vsize = 0;  // largest tap count actually used per axis
};

std::mutex g_tab_mu;
typedef std::tuple<int, int, int, int, int, int, int, int, int, int, int, int> TabKey;
static std::map<TabKey, DevTables> g_tabs;

// eff_h x eff_w = region actually produced, starting at (cy0, cx0) of the resized image (crop window; the patch
// layout drops the remainder rows / columns)
static int get_tables(int H, int W, int eff_h, int eff_w, int out_h, int out_w, int resample, int TX, int TY,
                      DevTables& out, int cy0 = 0, int cx0 = 0) {
    int dev = 0;
    GVL_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_tab_mu);
    TabKey key = std::make_tuple(dev, H, W, out_h, out_w, resample, TX, TY, eff_h, eff_w, cy0, cx0);
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
    for (int x0 = cx0; x0 < cx0 + eff_w; x0 += TX) {
        const int x1 = std::min(x0 + TX, cx0 + eff_w);
        const int c_lo = th.xmin[x0], c_hi = th.xmin[x1 - 1] + th.xsize[x1 - 1];
        const int a_lo = (3 * c_lo) & ~15, a_hi = (3 * c_hi + 15) & ~15;
        d.max_seg_bytes = std::max(d.max_seg_bytes, a_hi - a_lo);
        d.nch_max = std::max(d.nch_max, ((c_hi + 15) >> 4) - (c_lo >> 4));
    }
    for (int i = cx0; i < cx0 + eff_w; ++i) d.max_hsize = std::max(d.max_hsize, (int)th.xsize[i]);
    for (int i = cy0; i < cy0 + eff_h; ++i) d.max_vsize = std::max(d.max_vsize, (int)tv.xsize[i]);
    for (int y0 = cy0; y0 < cy0 + eff_h; y0 += TY) {
        const int y1 = std::min(y0 + TY, cy0 + eff_h);
        d.max_rows = std::max(d.max_rows, tv.xmin[y1 - 1] + tv.xsize[y1 - 1] - tv.xmin[y0]);
    }
    int rc = upload(th.xmin, &d.h_min) || upload(th.xsize, &d.h_size) || upload(th.w, &d.h_w) ||
             upload(tv.xmin, &d.v_min) || upload(tv.xsize, &d.v_size) || upload(tv.w, &d.v_w);
    if (rc) return 2;
    g_tabs[key] = d;
    out = d;
    return 0;
}

// ---- planar-kernel tables: taps pre-packed as 16-bit weight pairs aligned to the 32-bit words a thread loads ----

static uint32_t host_weight_pair(const int16_t* w, int taps, int h, int o) {
    const int j0 = 2 * h - o, j1 = j0 + 1;
    const uint32_t lo = (j0 >= 0 && j0 < taps) ? (uint32_t)(uint16_t)w[j0] : 0u;
    const uint32_t hi = (j1 >= 0 && j1 < taps) ? (uint32_t)(uint16_t)w[j1] : 0u;
    return lo | (hi << 16);
}

struct PlanarTables {
    // hq: per output column, 4*nw + 4 words: [0, 2nw) weight pairs of the window starting at the word holding the
    //     first tap, [2nw, 4nw) zeros (so a warp can slide its window start), [4nw] = first | last << 16 (non-zero
    //     pair range), [4nw + 1] = source pixel of the window's first word (xmin & ~3).
    // vq: per output row, vq_stride words: [0, 2nwv) weight pairs relative to the row quads of the row's y tile,
    //     [2nwv] = first quad.
    uint32_t *hq = nullptr, *vq = nullptr;
    int vq_stride = 0;
};
static std::map<std::tuple<int, int, int, int, int, int, int, int, int, int>, PlanarTables> g_ptabs;

static int get_planar_tables(int H, int W, int out_h, int out_w, int resample, int TY, int cy0, int nw, int nwv,
                             PlanarTables& out) {
    int dev = 0;
    GVL_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_tab_mu);
    auto key = std::make_tuple(dev, H, W, out_h, out_w, resample, TY, cy0, nw, nwv);
    auto it = g_ptabs.find(key);
    if (it != g_ptabs.end()) {
        out = it->second;
        return 0;
    }
    AxisTaps th, tv;
    if (compute_axis_taps(W, out_w, resample, th) || compute_axis_taps(H, out_h, resample, tv)) return 1;
    PlanarTables t;
    const int hs = 4 * nw + 4;
    std::vector<uint32_t> hq((size_t)out_w * hs, 0u);
    for (int x = 0; x < out_w; ++x) {
        uint32_t* row = hq.data() + (size_t)x * hs;
        const int o = th.xmin[x] & 3;
        int first = 2 * nw, last = 0;
        for (int h = 0; h < 2 * nw; ++h) {
            row[h] = host_weight_pair(th.w.data() + (size_t)x * th.taps, th.taps, h, o);
            if (row[h]) {
                first = std::min(first, h);
                last = h + 1;
            }
        }
        row[4 * nw] = (uint32_t)first | ((uint32_t)last << 16);
        row[4 * nw + 1] = (uint32_t)(th.xmin[x] & ~3);
    }
    t.vq_stride = (2 * nwv + 1 + 3) & ~3;
    std::vector<uint32_t> vq((size_t)out_h * t.vq_stride, 0u);
    for (int y = cy0; y < out_h; ++y) {  // y tiles start at the crop origin
        uint32_t* row = vq.data() + (size_t)y * t.vq_stride;
        const int rel = tv.xmin[y] - tv.xmin[cy0 + ((y - cy0) / TY) * TY];
        for (int h = 0; h < 2 * nwv; ++h)
            row[h] = host_weight_pair(tv.w.data() + (size_t)y * tv.taps, tv.taps, h, rel & 3);
        row[2 * nwv] = (uint32_t)(rel >> 2);
    }
    if (upload(hq, &t.hq) || upload(vq, &t.vq)) return 2;
    g_ptabs[key] = t;
    out = t;
    return 0;
}

static std::atomic<int> g_pre_path{0};

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

// ---- v2: planar / IDP.2A kernel ---------------------------------------------------------------------

constexpr int PL_THREADS = 256;
constexpr int PL_RG = 8;  // source rows per staged group = two row quads of the transposed intermediate
constexpr int PL_PF = 3;  // L2 prefetch distance in groups
constexpr int PL_HALF = PL_THREADS / 2;

struct PlanarParams {
    const uint8_t* frames;
    int B, H, W;
    int out_h, out_w, eff_h, eff_w;  // out_* = dims of the written image (the crop window), eff_* = produced region
    int cy0, cx0;                    // origin of the produced region inside the resized image
    const int32_t *h_min, *h_size, *v_min, *v_size;
    const int16_t *h_w, *v_w;
    int h_taps, v_taps, h_prec, v_prec;
    int TX, TY, VB;   // tile and band (rows finished per staging round) sizes
    int pw;           // bytes per planar staged row (multiple of 16)
    int g_max;        // row quads of the intermediate
    int stage_bytes;  // max(planar staging, output band staging), multiple of 16
    const float* lut; // [3][256]
    const uint32_t *hq, *vq;  // packed weight-pair tables (PlanarTables)
    int vq_stride;
    void* out;
    int layout, patch, ld, gh, gw;
    int fast_rows;    // every stored row starts 16-byte aligned and holds whole 16-pixel chunks
    int src_x0, src_w; // stored column band of the frames (0, W for whole frames)
};

// 16 interleaved RGB pixels (12 words) -> 4 words per colour plane
__device__ __forceinline__ void deinterleave16(const uint32_t (&a)[12], uint4& r, uint4& g, uint4& b) {
    uint32_t rr[4], gg[4], bb[4];
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        const uint32_t w0 = a[3 * m], w1 = a[3 * m + 1], w2 = a[3 * m + 2];
        rr[m] = __byte_perm(__byte_perm(w0, w1, 0x0630), w2, 0x5210);  // bytes 0,3,6,9
        gg[m] = __byte_perm(__byte_perm(w0, w1, 0x0741), w2, 0x6210);  // bytes 1,4,7,10
        bb[m] = __byte_perm(__byte_perm(w0, w1, 0x0052), w2, 0x7410);  // bytes 2,5,8,11
    }
    r = make_uint4(rr[0], rr[1], rr[2], rr[3]);
    g = make_uint4(gg[0], gg[1], gg[2], gg[3]);
    b = make_uint4(bb[0], bb[1], bb[2], bb[3]);
}

// Horizontal filter of one row quad (4 staged rows x 3 planes) for this thread's column.  NWE = words per
// window actually loaded; SKIP_FIRST / SKIP_LAST drop the IDP of the first low / last high byte pair when it
// is zero-weighted for the whole warp (all three are warp-uniform, chosen once per warp at start-up).
template <int NWMAX, int NWE, bool SKIP_FIRST, bool SKIP_LAST>
__device__ __forceinline__ void h_pass_quad(const uint8_t* src_base, int pw, const uint32_t (&wq)[2 * NWMAX],
                                            int h_round, int h_prec, uint32_t* dst, int plane_stride) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        int u[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint32_t* src = reinterpret_cast<const uint32_t*>(src_base + (size_t)(c * PL_RG + i) * pw);
            int acc = h_round;
#pragma unroll
            for (int k = 0; k < NWE; ++k) {
                const uint32_t w = src[k];
                if (!(SKIP_FIRST && k == 0)) acc = dp2a_lo(wq[2 * k], w, acc);
                if (!(SKIP_LAST && k == NWE - 1)) acc = dp2a_hi(wq[2 * k + 1], w, acc);
            }
            u[i] = acc >> h_prec;
        }
        dst[c * plane_stride] = pack4_sat_u8(u[0], u[1], u[2], u[3]);
    }
}

// NW / NWV: 32-bit words a thread loads per horizontal / vertical window (o + taps <= 4 * words);
// LPT: 16-pixel chunks a thread fetches per staged group.
// FAST: every row is 16-byte aligned with W % 16 == 0 (128-bit loads only; the guarded byte loader is compiled out).
template <int NW, int NWV, int LPT, bool FAST>
__global__ void __launch_bounds__(PL_THREADS, FAST ? 3 : 2)
preprocess_planar_kernel(const PlanarParams p) {
    extern __shared__ __align__(16) uint8_t pl_smem[];
    // [stage: 3 planes x 8 rows x pw | aliased by the output band]
    // [sH: 3 planes x g_max quads x 2 column pairs x 32 x 2 words (+ NWV pad quads)] [sWV] [lut: 768 floats]
    uint8_t* sP = pl_smem;
    uint8_t* sOut = pl_smem;
    uint32_t* sH = reinterpret_cast<uint32_t*>(pl_smem + p.stage_bytes);
    uint32_t* sWV = sH + (3 * p.g_max + NWV) * 128;   // [TY][vq_stride]
    float* sLut = reinterpret_cast<float*>(sWV + p.TY * p.vq_stride);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.z;
    const int x0 = p.cx0 + blockIdx.x * p.TX, x1 = min(x0 + p.TX, p.cx0 + p.eff_w);
    const int y0 = p.cy0 + blockIdx.y * p.TY, y1 = min(y0 + p.TY, p.cy0 + p.eff_h);
    const int ntx = x1 - x0, nty = y1 - y0;
    const int r_lo = p.v_min[y0];
    const int r_hi = p.v_min[y1 - 1] + p.v_size[y1 - 1];
    const int nrows = r_hi - r_lo;
    const int c_lo = p.h_min[x0];
    const int c_hi = p.h_min[x1 - 1] + p.h_size[x1 - 1];
    const int chunk0 = c_lo >> 4;                     // first 16-pixel chunk of the staged rows
    const int nch = ((c_hi + 15) >> 4) - chunk0;      // chunks per staged row
    // the source tensor may hold only the column band [src_x0, src_x0 + src_w) of the W-wide frames
    const int chunk_base = chunk0 - (p.src_x0 >> 4);  // first staged chunk, counted from the stored row's start
    const int row_chunks = p.src_w >> 4;              // whole chunks in a stored row
    const size_t row_bytes = (size_t)p.src_w * 3;
    const uint8_t* fbase = p.frames + ((size_t)b * p.H + r_lo) * row_bytes;

    // ---- per-thread horizontal setup: column x0 + 4*lane + (warp & 3), row quad (warp >> 2) of each group.
    // The warp slides its window start past the leading words none of its columns need and picks the
    // h_pass_quad variant that skips the byte pairs that are zero-weighted for all of them.
    const int xm = warp & 3, slot = warp >> 2;
    const int xx = 4 * lane + xm;
    const bool x_active = xx < ntx;
    uint32_t wq[2 * NW];
    int my_off, variant;
    {
        const uint32_t* hrow = p.hq + (size_t)(x0 + (x_active ? xx : 0)) * (4 * NW + 4);
        const uint2 info = *reinterpret_cast<const uint2*>(hrow + 4 * NW);
        int first = x_active ? (int)(info.x & 0xffffu) : 2 * NW;
        int last = x_active ? (int)(info.x >> 16) : 0;
        first = __reduce_min_sync(0xffffffffu, first);
        last = __reduce_max_sync(0xffffffffu, last);
        if (last <= first) first = 0, last = 1;  // warp without columns
        const int skip = first >> 1;              // leading words nobody needs
        my_off = (int)info.y - (chunk0 << 4) + 4 * skip;
        const int hb = first - 2 * skip, he = last - 2 * skip;
        const int nwe = (he + 1) >> 1 <= NW - 1 ? NW - 1 : NW;
        variant = (nwe == NW ? 4 : 0) | (hb == 1 ? 2 : 0) | (he <= 2 * nwe - 1 ? 1 : 0);
#pragma unroll
        for (int k = 0; k < NW; ++k) {
            const uint2 w2 = *reinterpret_cast<const uint2*>(hrow + 2 * skip + 2 * k);
            wq[2 * k] = w2.x;
            wq[2 * k + 1] = w2.y;
        }
    }
    // ---- vertical tables of this tile's rows, LUT
    for (int i = tid; i < nty * p.vq_stride; i += PL_THREADS) sWV[i] = p.vq[(size_t)y0 * p.vq_stride + i];
    for (int i = tid; i < 768; i += PL_THREADS) sLut[i] = p.lut[i];

    // ---- staging: global -> registers (next group) while the current group is filtered.  A thread owns the same
    // LPT (row, chunk) slots of every group, so the address arithmetic is done once.
    // The two row-quad halves of the CTA (warps 0-3 / 4-7) stage and filter their own 4 rows of every group and
    // synchronise among themselves only (named barriers), so one half's loads overlap the other half's math.
    uint32_t pre[LPT][12];
    int src_off[LPT], dst_off[LPT], item_rr[LPT];
    const int htid = tid & (PL_HALF - 1);
#pragma unroll
    for (int k = 0; k < LPT; ++k) {
        const int i = htid + k * PL_HALF;
        const int rr = i / nch, ch = i - rr * nch;
        const bool ok = rr < 4 && (!FAST || chunk_base + ch < row_chunks);
        item_rr[k] = ok ? slot * 4 + rr : 0x40000000;  // never < rows_left
        src_off[k] = (slot * 4 + rr) * (int)row_bytes + (chunk_base + ch) * 48;
        dst_off[k] = (slot * 4 + rr) * p.pw + ch * 16;
    }
    auto fetch_group = [&](int g) {
        const int rows_left = min(nrows, p.H - r_lo) - g * PL_RG;  // staged rows of this group that exist
        const uint8_t* gbase = fbase + (size_t)g * PL_RG * row_bytes;
#pragma unroll
        for (int k = 0; k < LPT; ++k) {
            const bool in = item_rr[k] < rows_left;
            if (FAST) {
                uint4 v0 = make_uint4(0, 0, 0, 0), v1 = v0, v2 = v0;
                if (in) {
                    const uint4* src = reinterpret_cast<const uint4*>(gbase + src_off[k]);
                    v0 = __ldg(src);
                    v1 = __ldg(src + 1);
                    v2 = __ldg(src + 2);
                }
                pre[k][0] = v0.x; pre[k][1] = v0.y; pre[k][2] = v0.z; pre[k][3] = v0.w;
                pre[k][4] = v1.x; pre[k][5] = v1.y; pre[k][6] = v1.z; pre[k][7] = v1.w;
                pre[k][8] = v2.x; pre[k][9] = v2.y; pre[k][10] = v2.z; pre[k][11] = v2.w;
            } else {
                // ragged / unaligned rows: guarded byte loads (zero beyond the row end)
                const int rr = item_rr[k];
                const uint8_t* rowp = gbase + (size_t)(in ? rr : 0) * row_bytes;
                const int byte0 = src_off[k] - rr * (int)row_bytes;
#pragma unroll
                for (int w = 0; w < 12; ++w) {
                    uint32_t v = 0;
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int o = byte0 + 4 * w + e;
                        if (in && o < (int)row_bytes) v |= (uint32_t)rowp[o] << (8 * e);
                    }
                    pre[k][w] = v;
                }
            }
        }
    };
    auto store_group = [&]() {
#pragma unroll
        for (int k = 0; k < LPT; ++k) {
            if (htid + k * PL_HALF < 4 * nch) {
                uint4 r, g, bl;
                deinterleave16(pre[k], r, g, bl);
                uint8_t* d = sP + dst_off[k];
                *reinterpret_cast<uint4*>(d) = r;
                *reinterpret_cast<uint4*>(d + (size_t)PL_RG * p.pw) = g;
                *reinterpret_cast<uint4*>(d + (size_t)2 * PL_RG * p.pw) = bl;
            }
        }
    };

    const int ngroups = (nrows + PL_RG - 1) / PL_RG;
    const int h_round = 1 << (p.h_prec - 1);
    // DRAM latency is taken off the critical path with bulk L2 prefetches PL_PF groups ahead (one per source
    // row, no registers, no shared memory): the register loads of fetch_group then hit L2.
    auto prefetch_group = [&](int g) {
        if (FAST && g < ngroups && htid < 4) {
            const int row = g * PL_RG + slot * 4 + htid;
            if (row < min(nrows, p.H - r_lo)) {
                const int nb = min(nch, row_chunks - chunk_base) * 48;
                const uint8_t* src = fbase + (size_t)row * row_bytes + (size_t)chunk_base * 48;
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(nb) : "memory");
            }
        }
    };
#pragma unroll
    for (int g = 1; g <= PL_PF; ++g) prefetch_group(g);
    const uint8_t* h_src = sP + (size_t)slot * 4 * p.pw + my_off;
    uint32_t* h_dst = sH + (slot * 2 + (xm >> 1)) * 64 + lane * 2 + (xm & 1);
    const int plane_stride = p.g_max * 128;
    auto half_sync = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(1 + slot), "n"(PL_HALF) : "memory"); };
    __syncthreads();  // tables / LUT visible
    fetch_group(0);
    store_group();
    half_sync();
    for (int g = 0; g < ngroups; ++g) {
        prefetch_group(g + 1 + PL_PF);
        if (g + 1 < ngroups) fetch_group(g + 1);
        // horizontal pass: 4 rows x 3 planes of this thread's column -> one transposed word per plane
        if (x_active) {
            uint32_t* dst = h_dst + g * 256;
            switch (variant) {
                case 0: h_pass_quad<NW, NW - 1, false, false>(h_src, p.pw, wq, h_round, p.h_prec, dst, plane_stride); break;
                case 1: h_pass_quad<NW, NW - 1, false, true>(h_src, p.pw, wq, h_round, p.h_prec, dst, plane_stride); break;
                case 2: h_pass_quad<NW, NW - 1, true, false>(h_src, p.pw, wq, h_round, p.h_prec, dst, plane_stride); break;
                case 3: h_pass_quad<NW, NW - 1, true, true>(h_src, p.pw, wq, h_round, p.h_prec, dst, plane_stride); break;
                case 4: h_pass_quad<NW, NW, false, false>(h_src, p.pw, wq, h_round, p.h_prec, dst, plane_stride); break;
                case 5: h_pass_quad<NW, NW, false, true>(h_src, p.pw, wq, h_round, p.h_prec, dst, plane_stride); break;
                case 6: h_pass_quad<NW, NW, true, false>(h_src, p.pw, wq, h_round, p.h_prec, dst, plane_stride); break;
                default: h_pass_quad<NW, NW, true, true>(h_src, p.pw, wq, h_round, p.h_prec, dst, plane_stride); break;
            }
        }
        half_sync();  // this half is done reading its staged rows
        if (g + 1 < ngroups) store_group();
        half_sync();
    }
    __syncthreads();  // both halves' intermediate rows complete; the staging buffer becomes the output band

    // ---- vertical pass + normalise, one band (VB output rows) at a time through the staging buffer.
    // Work item = (column pair 4*lane + 2*xp + {0,1}, output row): one 64-bit load per row quad and plane
    // fetches both columns, all planes share the row's weight pairs; four warps per xp alternate rows.
    // Words beyond a row's window carry zero weights (the pad quads after sH keep the reads in bounds).
    const int v_round = 1 << (p.v_prec - 1);
    const int PP = p.patch * p.patch;
    const bool patch_layout = p.layout == GVL_LAYOUT_BF16_PATCH;
    const int xp = warp & 1, vslot = warp >> 1;
    const int xa = 4 * lane + 2 * xp;
    const bool a_on = xa < ntx, b_on = xa + 1 < ntx;
    const int oa = patch_layout ? (xa / p.patch) * p.ld + xa % p.patch : xa;
    const int ob = patch_layout ? ((xa + 1) / p.patch) * p.ld + (xa + 1) % p.patch : xa + 1;
    const bool pair_store = b_on && ob == oa + 1 && (oa & 1) == 0 && (p.patch & 1) == 0;
    const int c_stride = patch_layout ? PP : p.VB * p.TX;  // elements between planes inside the staged band
    const int y_stride = patch_layout ? p.patch : p.TX;
    const int esize = p.layout == GVL_LAYOUT_U8_CHW ? 1 : (p.layout == GVL_LAYOUT_F32_CHW ? 4 : 2);
    const uint2* v_src = reinterpret_cast<const uint2*>(sH) + xp * 32 + lane;
    constexpr int VS = (2 * NWV + 1 + 3) & ~3;
    for (int band0 = 0; band0 < nty; band0 += p.VB) {
        const int nb = min(p.VB, nty - band0);
        if (patch_layout) {
            const int padc = p.ld - 3 * PP, npatch = ntx / p.patch;
            for (int i = tid; i < npatch * padc; i += PL_THREADS)
                reinterpret_cast<uint16_t*>(sOut)[(size_t)(i / padc) * p.ld + 3 * PP + i % padc] = 0;
        }
        for (int yl = vslot; yl < nb; yl += 4) {
            uint32_t wv[VS];
#pragma unroll
            for (int h = 0; h < VS / 4; ++h) {
                const uint4 t = reinterpret_cast<const uint4*>(sWV + (band0 + yl) * VS)[h];
                wv[4 * h] = t.x; wv[4 * h + 1] = t.y; wv[4 * h + 2] = t.z; wv[4 * h + 3] = t.w;
            }
            const uint2* src = v_src + wv[2 * NWV] * 64;
            int ua[3], ub[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                int acc_a = v_round, acc_b = v_round;
#pragma unroll
                for (int k = 0; k < NWV; ++k) {
                    const uint2 w = src[(c * p.g_max + k) * 64];
                    acc_a = dp2a_lo(wv[2 * k], w.x, acc_a);
                    acc_b = dp2a_lo(wv[2 * k], w.y, acc_b);
                    acc_a = dp2a_hi(wv[2 * k + 1], w.x, acc_a);
                    acc_b = dp2a_hi(wv[2 * k + 1], w.y, acc_b);
                }
                ua[c] = min(max(acc_a >> p.v_prec, 0), 255);
                ub[c] = min(max(acc_b >> p.v_prec, 0), 255);
            }
            if (a_on) {
                const int o = oa + yl * y_stride, o2 = ob + yl * y_stride;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    if (p.layout == GVL_LAYOUT_U8_CHW) {
                        sOut[o + c * c_stride] = (uint8_t)ua[c];
                        if (b_on) sOut[o2 + c * c_stride] = (uint8_t)ub[c];
                    } else if (p.layout == GVL_LAYOUT_F32_CHW) {
                        reinterpret_cast<float*>(sOut)[o + c * c_stride] = sLut[c * 256 + ua[c]];
                        if (b_on) reinterpret_cast<float*>(sOut)[o2 + c * c_stride] = sLut[c * 256 + ub[c]];
                    } else if (pair_store) {
                        *reinterpret_cast<uint32_t*>(reinterpret_cast<uint16_t*>(sOut) + o + c * c_stride) =
                            pack_bf16x2(sLut[c * 256 + ua[c]], sLut[c * 256 + ub[c]]);
                    } else {
                        reinterpret_cast<__nv_bfloat16*>(sOut)[o + c * c_stride] = __float2bfloat16_rn(sLut[c * 256 + ua[c]]);
                        if (b_on)
                            reinterpret_cast<__nv_bfloat16*>(sOut)[o2 + c * c_stride] =
                                __float2bfloat16_rn(sLut[c * 256 + ub[c]]);
                    }
                }
            }
        }
        __syncthreads();
        // copy-out: contiguous segments
        if (patch_layout) {
            const size_t prow = ((size_t)b * p.gh + (y0 - p.cy0 + band0) / p.patch) * p.gw + (x0 - p.cx0) / p.patch;
            uint4* gdst = reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(p.out) + prow * p.ld * 2);
            const int n16 = (ntx / p.patch) * p.ld * 2 / 16;
            for (int i = tid; i < n16; i += PL_THREADS) gdst[i] = reinterpret_cast<const uint4*>(sOut)[i];
        } else {
            const int seg_bytes = ntx * esize;
            for (int s = warp; s < 3 * nb; s += PL_THREADS / 32) {
                const int c = s / nb, yl = s - c * nb;
                uint8_t* gdst = reinterpret_cast<uint8_t*>(p.out) +
                                ((((size_t)b * 3 + c) * p.out_h + (y0 - p.cy0) + band0 + yl) * p.out_w + (x0 - p.cx0)) * esize;
                const uint8_t* ssrc = sOut + (size_t)((c * p.VB + yl) * p.TX) * esize;
                if (((reinterpret_cast<uintptr_t>(gdst) | (uintptr_t)seg_bytes | reinterpret_cast<uintptr_t>(ssrc)) & 15) == 0) {
                    for (int i = lane; i < (seg_bytes >> 4); i += 32)
                        reinterpret_cast<uint4*>(gdst)[i] = reinterpret_cast<const uint4*>(ssrc)[i];
                } else {
                    for (int i = lane; i < seg_bytes; i += 32) gdst[i] = ssrc[i];
                }
            }
        }
        __syncthreads();
    }
}

// (u8 - sub[c]) / div[c] in fp32 (IEEE division on the host), cached per (device, sub, div)
static std::map<std::tuple<int, float, float, float, float, float, float>, float*> g_luts;
std::shared_mutex g_tab_rw;

void trim_table_caches_if_full() {
    {
        std::lock_guard<std::mutex> lk(g_tab_mu);
        if (g_tabs.size() < kTableCacheCap && g_ptabs.size() < kTableCacheCap && g_luts.size() < kTableCacheCap &&
            stream_table_cache_size() < kTableCacheCap)
            return;
    }
    std::unique_lock<std::shared_mutex> x(g_tab_rw);  // no launch is between its table fetch and its kernel
    std::lock_guard<std::mutex> lk(g_tab_mu);
    cudaDeviceSynchronize();
    for (auto& kv : g_tabs) {
        DevTables& e = kv.second;
        cudaFree(e.h_min); cudaFree(e.h_size); cudaFree(e.v_min); cudaFree(e.v_size); cudaFree(e.h_w); cudaFree(e.v_w);
    }
    g_tabs.clear();
    for (auto& kv : g_ptabs) {
        cudaFree(kv.second.hq);
        cudaFree(kv.second.vq);
    }
    g_ptabs.clear();
    for (auto& kv : g_luts) cudaFree(kv.second);
    g_luts.clear();
    stream_table_cache_clear();
}

int get_lut(const float* sub, const float* div, float** out, float* h_copy) {
    int dev = 0;
    GVL_CUDA(cudaGetDevice(&dev));
    std::vector<float> h(768);
    for (int c = 0; c < 3; ++c)
        for (int u = 0; u < 256; ++u) {
            volatile float num = (float)u - sub[c];
            volatile float q = num / div[c];
            h[c * 256 + u] = q;
        }
    if (h_copy) memcpy(h_copy, h.data(), 768 * sizeof(float));
    std::lock_guard<std::mutex> lk(g_tab_mu);
    auto key = std::make_tuple(dev, sub[0], sub[1], sub[2], div[0], div[1], div[2]);
    auto it = g_luts.find(key);
    if (it != g_luts.end()) {
        *out = it->second;
        return 0;
    }
    float* d = nullptr;
    if (upload(h, &d)) return 2;
    g_luts[key] = d;
    *out = d;
    return 0;
}

// returns -1 when the geometry is outside the planar kernel's limits (caller falls back to v1)
// (cy0, cx0, crop_h, crop_w): window of the resized out_h x out_w image that is written (CHW layouts); the whole
// image when crop_h == 0.
static int launch_planar(const uint8_t* frames, int B, int H, int W, int out_h, int out_w, int resample,
                         const float* h_sub, const float* h_div, void* out, int layout, int patch, int ld,
                         cudaStream_t s, int cy0 = 0, int cx0 = 0, int crop_h = 0, int crop_w = 0, int src_x0 = 0,
                         int src_w = 0) {
    PlanarParams p;
    memset(&p, 0, sizeof(p));
    if (crop_h == 0) crop_h = out_h, crop_w = out_w;
    if (src_w == 0) src_w = W;
    p.eff_h = crop_h;
    p.eff_w = crop_w;
    p.cy0 = cy0;
    p.cx0 = cx0;
    int TX = 128, VB = 8, ty_unit = 8, ty_mult = 4;
    const int esize = layout == GVL_LAYOUT_U8_CHW ? 1 : (layout == GVL_LAYOUT_F32_CHW ? 4 : 2);
    if (layout == GVL_LAYOUT_BF16_PATCH) {
        p.gh = out_h / patch;
        p.gw = out_w / patch;
        p.eff_h = p.gh * patch;
        p.eff_w = p.gw * patch;
        TX = patch * (128 / patch);
        VB = patch;
        ty_unit = patch;
        ty_mult = 3;
    }
    DevTables tb;
    int nw = 0, nwv = 0, lpt = 0;
    size_t smem = 0;
    int TY = 0, pw = 0, g_max = 0, stage_bytes = 0;
    for (; ty_mult >= 1; --ty_mult) {
        TY = ty_unit * ty_mult;
        int rc = get_tables(H, W, p.eff_h, p.eff_w, out_h, out_w, resample, TX, TY, tb, cy0, cx0);
        if (rc) return rc;
        nw = tb.max_hsize + 3 <= 16 ? 4 : (tb.max_hsize + 3 <= 24 ? 6 : (tb.max_hsize + 3 <= 32 ? 8 : 0));
        nwv = tb.max_vsize + 3 <= 12 ? 3 : (tb.max_vsize + 3 <= 16 ? 4 : (tb.max_vsize + 3 <= 32 ? 8 : 0));
        lpt = 4 * tb.nch_max <= 2 * PL_HALF ? 2 : (4 * tb.nch_max <= 3 * PL_HALF ? 3 : 0);
        if (!nw || !nwv || !lpt) return -1;
        const bool fast_rows = src_w % 16 == 0 && src_x0 % 16 == 0 && (uintptr_t)frames % 16 == 0;
        if (!fast_rows || (!(nw <= 4 && nwv <= 3 && lpt <= 2) && !(nw <= 6 && nwv <= 4 && lpt <= 2))) nw = 8, nwv = 8, lpt = 3;
        else if (!(nw <= 4 && nwv <= 3)) nw = 6, nwv = 4;
        pw = tb.nch_max * 16 + 32;
        g_max = 2 * ((tb.max_rows + PL_RG - 1) / PL_RG);
        const int band_bytes = layout == GVL_LAYOUT_BF16_PATCH ? (TX / patch) * ld * 2 : 3 * VB * TX * esize;
        stage_bytes = (std::max(3 * PL_RG * pw, band_bytes) + 15) & ~15;
        const int vq_stride = (2 * nwv + 1 + 3) & ~3;
        smem = (size_t)stage_bytes + (size_t)(3 * g_max + nwv) * 128 * 4 + (size_t)TY * vq_stride * 4 + 768 * 4;
        if (smem <= 74 * 1024) break;  // three CTAs per SM
    }
    if (smem > 200 * 1024) return -1;
    float* lut = nullptr;
    int rc = get_lut(h_sub, h_div, &lut);
    if (rc) return rc;
    PlanarTables pt;
    rc = get_planar_tables(H, W, out_h, out_w, resample, TY, cy0, nw, nwv, pt);
    if (rc) return rc;
    p.hq = pt.hq;
    p.vq = pt.vq;
    p.vq_stride = pt.vq_stride;
    p.frames = frames;
    p.B = B;
    p.H = H;
    p.W = W;
    p.out_h = crop_h;
    p.out_w = crop_w;
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
    p.VB = VB;
    p.pw = pw;
    p.g_max = g_max;
    p.stage_bytes = stage_bytes;
    p.lut = lut;
    p.out = out;
    p.layout = layout;
    p.patch = patch > 0 ? patch : 1;
    p.ld = ld;
    p.fast_rows = (src_w % 16 == 0 && src_x0 % 16 == 0 && (uintptr_t)frames % 16 == 0) ? 1 : 0;
    p.src_x0 = src_x0;
    p.src_w = src_w;
    dim3 grid((p.eff_w + TX - 1) / TX, (p.eff_h + TY - 1) / TY, B);
    const double out_bytes = layout == GVL_LAYOUT_BF16_PATCH ? (double)p.gh * p.gw * 3 * patch * patch * 2
                                                              : (double)3 * crop_h * crop_w * esize;
    ProfScope prof(GVL_K_PREPROCESS, (double)B * ((double)H * W * 3 + out_bytes), s);
    auto launch = [&](auto kernel) -> int {
        GVL_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kernel<<<grid, PL_THREADS, smem, s>>>(p);
        return 0;
    };
    if (p.fast_rows) {
        if (nw == 4) rc = launch(preprocess_planar_kernel<4, 3, 2, true>);
        else if (nw == 6) rc = launch(preprocess_planar_kernel<6, 4, 2, true>);
        else rc = launch(preprocess_planar_kernel<8, 8, 3, true>);
    } else {
        rc = launch(preprocess_planar_kernel<8, 8, 3, false>);
    }
    if (rc) return rc;
    GVL_LAUNCH_CHECK("preprocess_planar_kernel");
    return 0;
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

// bf16 pixel_values [clips*frames, 3, H, W] (frame-major inside a clip) -> bf16 tubelet im2col rows for the
// Conv3d(kernel = stride = (tubelet, patch, patch)) patch embedding: row = ((clip*(frames/tubelet) + t/tubelet)*gh +
// py)*gw + px, col = c*tubelet*p*p + (t % tubelet)*p*p + ky*p + kx.  One thread moves 8 consecutive kx (16 bytes).
namespace gvl {
__global__ void __launch_bounds__(256)
patchify_tubelet_kernel(const __nv_bfloat16* __restrict__ pv, int nframes_total, int frames, int H, int W, int patch,
                        int tubelet, __nv_bfloat16* __restrict__ out) {
    const int gh = H / patch, gw = W / patch, PP = patch * patch, K = 3 * tubelet * PP;
    const int kchunks = patch >> 3;  // 16-byte chunks per patch row
    const size_t total = (size_t)nframes_total * 3 * gh * patch * gw * kchunks;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        size_t r = i;
        const int kc = (int)(r % kchunks); r /= kchunks;
        const int px = (int)(r % gw); r /= gw;
        const int ky = (int)(r % patch); r /= patch;
        const int py = (int)(r % gh); r /= gh;
        const int c = (int)(r % 3); r /= 3;
        const int f = (int)r;  // global frame index
        const int clip = f / frames, t = f - clip * frames;
        const uint4 v = *reinterpret_cast<const uint4*>(pv + (((size_t)f * 3 + c) * H + py * patch + ky) * W + px * patch + kc * 8);
        const size_t row = (((size_t)clip * (frames / tubelet) + t / tubelet) * gh + py) * gw + px;
        *reinterpret_cast<uint4*>(out + row * K + (size_t)c * tubelet * PP + (t % tubelet) * PP + ky * patch + kc * 8) = v;
    }
}
}  // namespace gvl

extern "C" int gvl_patchify_tubelet_bf16(const void* pixel_values, int clips, int frames, int H, int W, int patch,
                                         int tubelet, void* out, void* stream) {
    using namespace gvl;
    GVL_CHECK_ARG(pixel_values && out, "gvl_patchify_tubelet_bf16: null pointer");
    GVL_CHECK_ARG(clips > 0 && frames > 0 && tubelet > 0 && frames % tubelet == 0 && patch % 8 == 0 && H % patch == 0 &&
                      W % patch == 0 && W % 8 == 0,
                  "gvl_patchify_tubelet_bf16: bad shape clips=%d frames=%d H=%d W=%d patch=%d tubelet=%d", clips, frames, H,
                  W, patch, tubelet);
    GVL_CHECK_ARG((uintptr_t)pixel_values % 16 == 0 && (uintptr_t)out % 16 == 0,
                  "gvl_patchify_tubelet_bf16: pointers must be 16-byte aligned");
    const size_t total = (size_t)clips * frames * 3 * H * (W / 8);
    const int grid = (int)std::min<size_t>((total + 255) / 256, (size_t)sm_count() * 16);
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    ProfScope prof(GVL_K_PATCHIFY, (double)clips * frames * 3 * H * W * 4, s);
    patchify_tubelet_kernel<<<grid, 256, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(pixel_values), clips * frames,
                                                 frames, H, W, patch, tubelet, reinterpret_cast<__nv_bfloat16*>(out));
    GVL_LAUNCH_CHECK("patchify_tubelet_kernel");
    return 0;
}

extern "C" int gvl_preprocess_u8_crop(const uint8_t* frames, int B, int H, int W, int out_h, int out_w, int crop_y0,
                                      int crop_x0, int crop_h, int crop_w, int resample, const float* h_sub,
                                      const float* h_div, void* out, int layout, void* stream) {
    using namespace gvl;
    trim_table_caches_if_full();
    std::shared_lock<std::shared_mutex> tab_guard(g_tab_rw);
    GVL_CHECK_ARG(frames && out && h_sub && h_div, "gvl_preprocess_u8_crop: null pointer");
    GVL_CHECK_ARG(B > 0 && B <= 65535 && H > 0 && W > 0 && out_h > 0 && out_w > 0, "gvl_preprocess_u8_crop: bad shape");
    GVL_CHECK_ARG(layout >= GVL_LAYOUT_U8_CHW && layout <= GVL_LAYOUT_BF16_CHW,
                  "gvl_preprocess_u8_crop: only the CHW layouts take a crop window (layout %d)", layout);
    GVL_CHECK_ARG(H >= out_h && W >= out_w, "gvl_preprocess_u8_crop: only downscaling is supported (%dx%d -> %dx%d)", H,
                  W, out_h, out_w);
    GVL_CHECK_ARG(crop_y0 >= 0 && crop_x0 >= 0 && crop_h > 0 && crop_w > 0 && crop_y0 + crop_h <= out_h &&
                      crop_x0 + crop_w <= out_w,
                  "gvl_preprocess_u8_crop: crop window [%d,%d)+%dx%d outside the %dx%d resized image", crop_y0, crop_x0,
                  crop_h, crop_w, out_h, out_w);
    const int rc = launch_planar(frames, B, H, W, out_h, out_w, resample, h_sub, h_div, out, layout, 0, 0,
                                 reinterpret_cast<cudaStream_t>(stream), crop_y0, crop_x0, crop_h, crop_w);
    if (rc == -1) {
        set_error("gvl_preprocess_u8_crop: resize geometry %dx%d -> %dx%d is outside the kernel's tap window", H, W, out_h,
                  out_w);
        return 1;
    }
    return rc;
}

extern "C" int gvl_preprocess_u8_crop_band(const uint8_t* frames, int B, int H, int W, int band_x0, int band_w, int out_h,
                                           int out_w, int crop_y0, int crop_x0, int crop_h, int crop_w, int resample,
                                           const float* h_sub, const float* h_div, void* out, int layout, void* stream) {
    using namespace gvl;
    trim_table_caches_if_full();
    std::shared_lock<std::shared_mutex> tab_guard(g_tab_rw);
    GVL_CHECK_ARG(frames && out && h_sub && h_div, "gvl_preprocess_u8_crop_band: null pointer");
    GVL_CHECK_ARG(B > 0 && B <= 65535 && H > 0 && W > 0 && out_h > 0 && out_w > 0, "gvl_preprocess_u8_crop_band: bad shape");
    GVL_CHECK_ARG(layout >= GVL_LAYOUT_U8_CHW && layout <= GVL_LAYOUT_BF16_CHW,
                  "gvl_preprocess_u8_crop_band: only the CHW layouts take a crop window (layout %d)", layout);
    GVL_CHECK_ARG(H >= out_h && W >= out_w, "gvl_preprocess_u8_crop_band: only downscaling is supported (%dx%d -> %dx%d)",
                  H, W, out_h, out_w);
    GVL_CHECK_ARG(crop_y0 >= 0 && crop_x0 >= 0 && crop_h > 0 && crop_w > 0 && crop_y0 + crop_h <= out_h &&
                      crop_x0 + crop_w <= out_w,
                  "gvl_preprocess_u8_crop_band: crop window [%d,%d)+%dx%d outside the %dx%d resized image", crop_y0,
                  crop_x0, crop_h, crop_w, out_h, out_w);
    GVL_CHECK_ARG(band_x0 >= 0 && band_w > 0 && band_x0 + band_w <= W && band_x0 % 16 == 0,
                  "gvl_preprocess_u8_crop_band: band [%d, %d) must start on a multiple of 16 inside the %d-wide frame",
                  band_x0, band_x0 + band_w, W);
    {
        AxisTaps th;
        GVL_CHECK_ARG(compute_axis_taps(W, out_w, resample, th) == 0, "gvl_preprocess_u8_crop_band: bad resize geometry");
        const int need_lo = th.xmin[crop_x0];
        const int need_hi = th.xmin[crop_x0 + crop_w - 1] + th.xsize[crop_x0 + crop_w - 1];
        GVL_CHECK_ARG(band_x0 <= need_lo && need_hi <= band_x0 + band_w,
                      "gvl_preprocess_u8_crop_band: the crop reads source columns [%d, %d), the band holds [%d, %d)", need_lo,
                      need_hi, band_x0, band_x0 + band_w);
    }
    const int rc = launch_planar(frames, B, H, W, out_h, out_w, resample, h_sub, h_div, out, layout, 0, 0,
                                 reinterpret_cast<cudaStream_t>(stream), crop_y0, crop_x0, crop_h, crop_w, band_x0, band_w);
    if (rc == -1) {
        set_error("gvl_preprocess_u8_crop_band: resize geometry %dx%d -> %dx%d is outside the kernel's tap window", H, W,
                  out_h, out_w);
        return 1;
    }
    return rc;
}

extern "C" int gvl_copy_band_h2d(void* dst, const void* src_host, long long rows, long long src_row_bytes,
                                 long long band_offset_bytes, long long band_bytes, void* stream) {
    using namespace gvl;
    GVL_CHECK_ARG(dst && src_host && rows > 0 && band_bytes > 0 && band_offset_bytes >= 0 &&
                      band_offset_bytes + band_bytes <= src_row_bytes,
                  "gvl_copy_band_h2d: bad arguments");
    GVL_CUDA(cudaMemcpy2DAsync(dst, (size_t)band_bytes, reinterpret_cast<const uint8_t*>(src_host) + band_offset_bytes,
                               (size_t)src_row_bytes, (size_t)band_bytes, (size_t)rows, cudaMemcpyHostToDevice,
                               reinterpret_cast<cudaStream_t>(stream)));
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

extern "C" int gvl_preprocess_path(int path) {
    using namespace gvl;
    GVL_CHECK_ARG(path >= GVL_PRE_PATH_AUTO && path <= GVL_PRE_PATH_V1, "gvl_preprocess_path: bad path %d", path);
    g_pre_path.store(path, std::memory_order_relaxed);
    return 0;
}

extern "C" int gvl_preprocess_u8(const uint8_t* frames, int B, int H, int W, int out_h, int out_w, int resample,
                                 const float* h_sub, const float* h_div, void* out, int layout, int patch, int ld,
                                 void* stream) {
    using namespace gvl;
    trim_table_caches_if_full();
    std::shared_lock<std::shared_mutex> tab_guard(g_tab_rw);
    GVL_CHECK_ARG(frames && out && h_sub && h_div, "gvl_preprocess_u8: null pointer");
    GVL_CHECK_ARG(B > 0 && B <= 65535 && H > 0 && W > 0 && out_h > 0 && out_w > 0, "gvl_preprocess_u8: bad shape");
    GVL_CHECK_ARG(layout >= 0 && layout <= 3, "gvl_preprocess_u8: bad layout %d", layout);
    GVL_CHECK_ARG(H >= out_h && W >= out_w, "gvl_preprocess_u8: only downscaling is supported (%dx%d -> %dx%d)", H, W,
                  out_h, out_w);
    if (layout == GVL_LAYOUT_BF16_PATCH) {
        GVL_CHECK_ARG(patch >= 4 && patch <= 32 && ld >= 3 * patch * patch && ld % 8 == 0,
                      "gvl_preprocess_u8: bad patch/ld %d/%d", patch, ld);
        GVL_CHECK_ARG((uintptr_t)out % 16 == 0, "gvl_preprocess_u8: output must be 16-byte aligned");
        GVL_CHECK_ARG(out_h >= patch && out_w >= patch, "gvl_preprocess_u8: output smaller than one patch");
    }
    {
        const int path = g_pre_path.load(std::memory_order_relaxed);
        cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
        if (path == GVL_PRE_PATH_AUTO && layout == GVL_LAYOUT_BF16_PATCH) {
            const int src = launch_stream5(frames, B, H, W, out_h, out_w, resample, h_sub, h_div, out, patch, ld, st);
            if (src != -1) return src;
        }
        if (path != GVL_PRE_PATH_V1) {
            const int prc = launch_planar(frames, B, H, W, out_h, out_w, resample, h_sub, h_div, out, layout, patch, ld, st);
            if (prc != -1) return prc;
        }
    }
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
