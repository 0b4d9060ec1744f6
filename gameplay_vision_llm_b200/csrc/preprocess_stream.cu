// preprocess_stream.cu — K1s: streaming form of the fused uint8 resize + normalize + patchify kernel for frames
// whose width is exactly five times the output width (1920 -> 384: the 1080p -> SigLIP2-so400m geometry of the
// headline workload), GVL_LAYOUT_BF16_PATCH only.  Same arithmetic as preprocess.cu (ATen's integer two-pass
// antialiased resize, horizontal then vertical, int16 weights, uint8 intermediate; fp32 normalisation; bf16 im2col)
// and bit-identical output; what changes is how the work is laid out.
//
//   * A CTA owns one strip of 128 output columns (9 patches of 14 + 2 slack columns) and a run of consecutive patch
//     rows of one frame, and walks DOWN the frame: every source row of the run is filtered horizontally exactly
//     once (the planar kernel re-filters the ~4.6 rows two tiles share: 12 % at 14-row tiles).
//   * Source rows arrive as raw interleaved RGB through 1-D bulk async copies (cp.async.bulk, mbarrier completion):
//     each warp owns a double-buffered slot of two rows and re-arms it itself as soon as the slot's words are in
//     registers, i.e. before the arithmetic (two task times of lead).  An additional cp.async.bulk.prefetch.L2
//     stream was measured and made it slower (87 us without, 91 / 94 / 96 / 104 us at distance 1 / 2 / 3 / 6).
//   * Horizontal pass: with an exact 5:1 scale the window of column x starts at pixel 5x - 2, so a lane that filters
//     four adjacent columns (4g .. 4g+3) of a row needs the 28 planar bytes 20g-4 .. 20g+23 — 21 interleaved words at
//     a lane stride of 15 words (conflict-free) — de-interleaves them once (42 PRMT) and feeds every colour plane's
//     7 words to compile-time-placed IDP.2A pairs: 5 per output, no window search, no per-column branches.  The
//     weight pairs come from a per-lane table (edge columns renormalised as ATen does) — the kernel falls back to
//     the planar one when the host finds a column whose taps do not fit this placement.
//   * The intermediate is a ring of row PAIRS: word = {row 2p, row 2p+1} x {column 2c, 2c+1}, so the vertical pass
//     is IDP.2A as well (4 per output: windows of <= 8 rows), one 64-bit load per row pair for four columns.
//   * Normalisation is fmaf(u8, 1/div, -sub/div) when the host has verified that it rounds to the same bf16 as the
//     IEEE division for all 3 x 256 inputs (true for SigLIP's 0.5 / 0.5), else a shared-memory LUT of the divisions.
//   * A finished patch row of the strip (9 x 592 bf16, contiguous in the im2col matrix) leaves with ONE bulk store.
#include "preprocess_common.cuh"

#include <cmath>
#include <cstring>
#include <map>
#include <tuple>

namespace gvl {

constexpr int PS_NW = 5;                 // warps per CTA (all of them filter)
constexpr int PS_THREADS = PS_NW * 32;
constexpr int PS_S = 5;                  // horizontal scale
constexpr int PS_COLS = 128;             // output columns a strip computes (4 per lane)
constexpr int PS_PITCH = 1968;           // bytes per staged source row: 32 lanes x 60 + 24 + alignment slack, 16 B multiple
constexpr int PS_SH_PAIRS = 24;          // ring of intermediate row pairs (48 rows)
constexpr int PS_VPAIRS = 4;             // row pairs per vertical window
constexpr int PS_MAX_STRIPS = 8;
// first 16-bit pair (of the 14 a lane holds per plane) of column 4g + j: its window starts at planar byte 2 + 5j
__host__ __device__ constexpr int ps_hp(int j) { return (2 + PS_S * j) >> 1; }

struct PsStrip {
    int g_start;   // first byte of a source row that is copied (16 B aligned)
    int dst_off;   // where it lands inside the staged row
    int len;       // bytes copied per row
    int delta;     // staged-row offset of the strip's planar byte (5 xs - 4), 4 B aligned
    int dx;        // first wanted column minus first computed column
    int n_patch;   // patches of the strip
    int p0;        // first patch
    int pad;
};

struct PsParams {
    const uint8_t* frames;
    int B, H, W;
    int gh, gw, patch, ld;
    int n_strips, run_len;
    int h_prec, v_prec;
    const uint32_t* hq;  // [n_strips][32][20] weight pairs
    const uint32_t* vq;  // [gh * patch][8]: 4 weight pairs, 4 ring positions (16 bit each), pad
    const int2* unit;    // [gh] first / last row pair of a patch row's vertical windows
    float na[3], nb[3];  // value = fmaf(u8, na, nb)
    const float* lut;    // [3][256], used when the fmaf form is not exact
    void* out;
    PsStrip strip[PS_MAX_STRIPS];
};

__device__ __forceinline__ void ps_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void ps_bulk_s2g(void* dst, const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes)
                 : "memory");
}

template <bool ARITH>
__global__ void __launch_bounds__(PS_THREADS, 3)
preprocess_stream5_kernel(const __grid_constant__ PsParams p) {
    extern __shared__ __align__(128) uint8_t ps_smem[];
    // [raw: NW warps x 2 slots x 2 rows x PITCH] [sH: 3 planes x SH_PAIRS x 64 words] [band: n_patch x ld bf16]
    // [vertical tables: 32 x 8 words] [lut: 768 floats (LUT form only)] [full barriers: NW x 2]
    uint8_t* raw = ps_smem;
    uint32_t* sH = reinterpret_cast<uint32_t*>(raw + PS_NW * 4 * PS_PITCH);
    uint16_t* band = reinterpret_cast<uint16_t*>(sH + 3 * PS_SH_PAIRS * 64);
    const int band_elems = ((PS_COLS / p.patch) * p.ld + 7) & ~7;
    uint32_t* sVQ = reinterpret_cast<uint32_t*>(band + band_elems);  // [patch][8] vertical tables of the current patch row
    float* sLut = reinterpret_cast<float*>(sVQ + 32 * 8);
    uint64_t* full = reinterpret_cast<uint64_t*>(sLut + (ARITH ? 0 : 768));

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // (tells the compiler it is warp-uniform)
    const int strip = blockIdx.x % p.n_strips, run = blockIdx.x / p.n_strips, b = blockIdx.y;
    const PsStrip& st = p.strip[strip];
    const int u0 = run * p.run_len, u1 = min(u0 + p.run_len, p.gh);
    const int p_begin = __ldg(&p.unit[u0]).x, p_end = __ldg(&p.unit[u1 - 1]).y;
    const int row_last = min(2 * p_end + 1, p.H - 1);
    const size_t row_bytes = (size_t)p.W * 3;
    const uint8_t* fbase = p.frames + (size_t)b * p.H * row_bytes + st.g_start;

    if (tid == 0) {
        for (int i = 0; i < PS_NW * 2; ++i) mbar_init(&full[i], 1);
        fence_barrier_init();
    }
    for (int i = tid; i < band_elems / 2; i += PS_THREADS) reinterpret_cast<uint32_t*>(band)[i] = 0;  // pad columns stay 0
    if (!ARITH)
        for (int i = tid; i < 768; i += PS_THREADS) sLut[i] = p.lut[i];
    __syncthreads();

    // ---- this warp's row pairs: p_begin + warp + k * NW, k = 0 .. n_my-1; pair k lives in slot k & 1.  Lane 0 re-arms
    // a slot as soon as the warp has pulled its words into registers.
    const int n_my = p_end - p_begin - warp >= 0 ? (p_end - p_begin - warp) / PS_NW + 1 : 0;
    // the frame's last pair has one row when H is odd
    const int k_single = (2 * p_end + 1 > p.H - 1 && (p_end - p_begin - warp) % PS_NW == 0) ? n_my - 1 : -1;
    const size_t pair_stride = (size_t)2 * PS_NW * row_bytes;
    const uint8_t* my_src = fbase + (size_t)2 * (p_begin + warp) * row_bytes;
    uint8_t* my_dst = raw + (size_t)(warp * 4) * PS_PITCH + st.dst_off;
    const uint32_t len = (uint32_t)st.len;
    auto issue = [&](int k) {  // lane 0 only, k < n_my
        uint64_t* bar = &full[warp * 2 + (k & 1)];
        uint8_t* dst = my_dst + (k & 1) * (2 * PS_PITCH);
        const uint8_t* src = my_src + (size_t)k * pair_stride;
        const bool two = k != k_single;
        mbar_arrive_expect_tx(bar, two ? 2 * len : len);
        ps_bulk_g2s(dst, src, len, bar);
        if (two) ps_bulk_g2s(dst + PS_PITCH, src + row_bytes, len, bar);
    };
    if (lane == 0) {
        if (0 < n_my) issue(0);
        if (1 < n_my) issue(1);
    }

    // ---- per-lane constants
    uint32_t wq[20];
    {
        const uint4* src = reinterpret_cast<const uint4*>(p.hq + ((size_t)strip * 32 + lane) * 20);
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            const uint4 t = __ldg(src + i);
            wq[4 * i] = t.x; wq[4 * i + 1] = t.y; wq[4 * i + 2] = t.z; wq[4 * i + 3] = t.w;
        }
    }
    const int h_round = 1 << (p.h_prec - 1), v_round = 1 << (p.v_prec - 1);
    const uint32_t* my_raw = reinterpret_cast<const uint32_t*>(raw + (size_t)(warp * 4) * PS_PITCH + st.delta + 60 * lane);
    // vertical pass: this lane's four columns 4*lane .. 4*lane+3 of the strip's computed columns; the wanted ones
    // start at dx.  Two bf16 pairs per (row, plane): element offsets inside the band, -1 = not wanted.
    const int ncols = st.n_patch * p.patch;
    int idx_a = -1, idx_b = -1;
    {
        const int ca = 4 * lane - st.dx, cb = ca + 2;
        if (ca >= 0 && ca + 1 < ncols) idx_a = (ca / p.patch) * p.ld + ca % p.patch;
        if (cb >= 0 && cb + 1 < ncols) idx_b = (cb / p.patch) * p.ld + cb % p.patch;
    }
    const int PP = p.patch * p.patch;
    const float na0 = p.na[0], na1 = p.na[1], na2 = p.na[2], nb0 = p.nb[0], nb1 = p.nb[1], nb2 = p.nb[2];

    int k = 0;
    int pp = p_begin + warp;          // this warp's next row pair
    int h_slot = pp % PS_SH_PAIRS;    // and its place in the intermediate ring
    for (int u = u0; u < u1; ++u) {
        const int pair_hi = __ldg(&p.unit[u]).y;
        // vertical tables of this patch row (read after the barrier below)
        if (tid < 2 * p.patch)
            reinterpret_cast<uint4*>(sVQ)[tid] = __ldg(reinterpret_cast<const uint4*>(p.vq + (size_t)u * p.patch * 8) + tid);
        // ---- horizontal pass of the row pairs this patch row still needs
        for (; pp <= pair_hi; ++k, pp += PS_NW) {
            mbar_wait(&full[warp * 2 + (k & 1)], (uint32_t)(k >> 1) & 1u);
            const uint32_t* src = my_raw + (k & 1) * (2 * PS_PITCH / 4);
            // 19 interleaved words per row (the lane's bytes 4 .. 79); once both rows are in registers the slot is
            // re-armed with the warp's pair after next, so that copy has two task times to land
            uint32_t a[2][21];
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int i = 1; i < 20; ++i) a[r][i] = src[r * (PS_PITCH / 4) + i];
            __syncwarp();
            if (lane == 0 && k + 2 < n_my) issue(k + 2);
            int acc[2][3][4];
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                // -> 7 words per plane; of the first / last planar word only the upper / lower half is ever multiplied
                uint32_t pl[3][7];
                pl[0][0] = __byte_perm(a[r][1], a[r][2], 0x5200);  // planar bytes 2, 3 (interleaved 6, 9) in place
                pl[1][0] = __byte_perm(a[r][1], a[r][2], 0x6300);  // 7, 10
                pl[2][0] = __byte_perm(a[r][1], a[r][2], 0x7400);  // 8, 11
#pragma unroll
                for (int m = 1; m < 6; ++m) {
                    const uint32_t w0 = a[r][3 * m], w1 = a[r][3 * m + 1], w2 = a[r][3 * m + 2];
                    pl[0][m] = __byte_perm(__byte_perm(w0, w1, 0x0630), w2, 0x5210);  // bytes 0,3,6,9
                    pl[1][m] = __byte_perm(__byte_perm(w0, w1, 0x0741), w2, 0x6210);  // bytes 1,4,7,10
                    pl[2][m] = __byte_perm(__byte_perm(w0, w1, 0x0052), w2, 0x7410);  // bytes 2,5,8,11
                }
                pl[0][6] = __byte_perm(a[r][18], a[r][19], 0x0030);  // planar bytes 24, 25 (interleaved 72, 75)
                pl[1][6] = __byte_perm(a[r][18], a[r][19], 0x0041);  // 73, 76
                pl[2][6] = __byte_perm(a[r][18], a[r][19], 0x0052);  // 74, 77
#pragma unroll
                for (int c = 0; c < 3; ++c)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        int s = h_round;
#pragma unroll
                        for (int i = 0; i < 5; ++i) {
                            const int h = ps_hp(j) + i;
                            s = (h & 1) ? dp2a_hi(wq[5 * j + i], pl[c][h >> 1], s) : dp2a_lo(wq[5 * j + i], pl[c][h >> 1], s);
                        }
                        acc[r][c][j] = s >> p.h_prec;
                    }
            }
            uint32_t* dst = sH + h_slot * 64 + 2 * lane;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                uint2 w;
                w.x = pack4_sat_u8(acc[0][c][0], acc[1][c][0], acc[0][c][1], acc[1][c][1]);
                w.y = pack4_sat_u8(acc[0][c][2], acc[1][c][2], acc[0][c][3], acc[1][c][3]);
                *reinterpret_cast<uint2*>(dst + c * (PS_SH_PAIRS * 64)) = w;
            }
            h_slot += PS_NW;
            if (h_slot >= PS_SH_PAIRS) h_slot -= PS_SH_PAIRS;
        }
        if (tid == 0) tma_store_wait_read<0>();  // the previous patch row has left the band buffer
        __syncthreads();                         // intermediate rows and tables of this patch row complete
        // ---- vertical pass + normalise: task = output row of the patch row (all three planes)
        for (int yy = warp; yy < p.patch; yy += PS_NW) {
            const uint4 q0 = reinterpret_cast<const uint4*>(sVQ)[2 * yy], q1 = reinterpret_cast<const uint4*>(sVQ)[2 * yy + 1];
            const uint32_t wv[4] = {q0.x, q0.y, q0.z, q0.w};
            const int so[4] = {(int)(q1.x & 0xffffu), (int)(q1.x >> 16), (int)(q1.y & 0xffffu), (int)(q1.y >> 16)};
            const uint2* src = reinterpret_cast<const uint2*>(sH) + lane;
            uint16_t* brow = band + yy * p.patch;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                int s0 = v_round, s1 = v_round, s2 = v_round, s3 = v_round;
#pragma unroll
                for (int i = 0; i < PS_VPAIRS; ++i) {
                    const uint2 w = src[c * (PS_SH_PAIRS * 32) + so[i]];
                    s0 = dp2a_lo(wv[i], w.x, s0);
                    s1 = dp2a_hi(wv[i], w.x, s1);
                    s2 = dp2a_lo(wv[i], w.y, s2);
                    s3 = dp2a_hi(wv[i], w.y, s3);
                }
                const uint32_t u4 = pack4_sat_u8(s0 >> p.v_prec, s1 >> p.v_prec, s2 >> p.v_prec, s3 >> p.v_prec);
                float f0, f1, f2, f3;
                if (ARITH) {
                    const float na = c == 0 ? na0 : (c == 1 ? na1 : na2), nb = c == 0 ? nb0 : (c == 1 ? nb1 : nb2);
                    f0 = fmaf((float)(u4 & 0xffu), na, nb);
                    f1 = fmaf((float)((u4 >> 8) & 0xffu), na, nb);
                    f2 = fmaf((float)((u4 >> 16) & 0xffu), na, nb);
                    f3 = fmaf((float)(u4 >> 24), na, nb);
                } else {
                    const float* l = sLut + c * 256;
                    f0 = l[u4 & 0xffu];
                    f1 = l[(u4 >> 8) & 0xffu];
                    f2 = l[(u4 >> 16) & 0xffu];
                    f3 = l[u4 >> 24];
                }
                if (idx_a >= 0) *reinterpret_cast<uint32_t*>(brow + idx_a + c * PP) = pack_bf16x2(f0, f1);
                if (idx_b >= 0) *reinterpret_cast<uint32_t*>(brow + idx_b + c * PP) = pack_bf16x2(f2, f3);
            }
        }
        fence_proxy_async_smem();
        __syncthreads();
        if (tid == 0) {
            uint8_t* gdst = reinterpret_cast<uint8_t*>(p.out) + (((size_t)b * p.gh + u) * p.gw + st.p0) * p.ld * 2;
            ps_bulk_s2g(gdst, band, (uint32_t)(st.n_patch * p.ld * 2));
            tma_store_commit();
        }
    }
    if (tid == 0) tma_store_wait_all();
}

// ---- host: tables of the 5:1 placement, cached per geometry ----

struct PsTables {
    uint32_t *hq = nullptr, *vq = nullptr;
    int2* unit = nullptr;
    int n_strips = 0, h_prec = 0, v_prec = 0, gh = 0, gw = 0;
    PsStrip strip[PS_MAX_STRIPS];
    bool ok = false;
};
static std::map<std::tuple<int, int, int, int, int, int, int, int>, PsTables> g_ps_tabs;

size_t stream_table_cache_size() { return g_ps_tabs.size(); }
void stream_table_cache_clear() {
    for (auto& kv : g_ps_tabs) {
        cudaFree(kv.second.hq);
        cudaFree(kv.second.vq);
        cudaFree(kv.second.unit);
    }
    g_ps_tabs.clear();
}

static uint16_t ps_bf16_rn(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    u += 0x7fffu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}

static int ps_get_tables(int H, int W, int out_h, int out_w, int resample, int patch, int ld, PsTables& out) {
    int dev = 0;
    GVL_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_tab_mu);
    auto key = std::make_tuple(dev, H, W, out_h, out_w, resample, patch, ld);
    auto it = g_ps_tabs.find(key);
    if (it != g_ps_tabs.end()) {
        out = it->second;
        return 0;
    }
    PsTables t;
    auto done = [&]() {
        g_ps_tabs[key] = t;
        out = t;
        return 0;
    };
    AxisTaps th, tv;
    if (compute_axis_taps(W, out_w, resample, th) || compute_axis_taps(H, out_h, resample, tv)) return done();
    t.gh = out_h / patch;
    t.gw = out_w / patch;
    t.h_prec = th.precision;
    t.v_prec = tv.precision;
    const int nps = PS_COLS / patch;  // patches per strip
    t.n_strips = (t.gw + nps - 1) / nps;
    if (t.n_strips > PS_MAX_STRIPS || t.h_prec < 1 || t.v_prec < 1) return done();
    const int row_bytes = W * 3;
    std::vector<uint32_t> hq((size_t)t.n_strips * 32 * 20, 0u);
    for (int s = 0; s < t.n_strips; ++s) {
        PsStrip& st = t.strip[s];
        const int x0 = s * nps * patch, xs = x0 & ~3;
        st.p0 = s * nps;
        st.n_patch = std::min(nps, t.gw - st.p0);
        st.dx = x0 - xs;
        if (st.dx + st.n_patch * patch > PS_COLS || (st.dx & 1) || (patch & 1)) return done();
        const int base_px = PS_S * xs - 4;            // planar pixel of the strip's byte 0 (may be -4)
        const int seg0 = 3 * base_px;                 // interleaved byte where lane 0's 84 bytes start
        const int a0 = seg0 >= 0 ? (seg0 & ~15) : -16;
        st.delta = seg0 - a0;
        st.g_start = std::max(a0, 0);
        st.dst_off = st.g_start - a0;
        int end = (seg0 + 31 * 60 + 84 + 15) & ~15;
        end = std::min(end, row_bytes);
        st.len = end - st.g_start;
        st.pad = 0;
        if (st.len <= 0 || st.dst_off + st.len > PS_PITCH || (st.delta & 3) || st.delta + 31 * 60 + 84 > PS_PITCH) return done();
        for (int g = 0; g < 32; ++g)
            for (int j = 0; j < 4; ++j) {
                const int x = xs + 4 * g + j;
                uint32_t* wrow = hq.data() + ((size_t)s * 32 + g) * 20 + 5 * j;
                if (x >= out_w) continue;  // slack column beyond the image: zero weights
                const int first_px = base_px + 20 * g + 2 * ps_hp(j);  // pixel of the column's first pair
                const int16_t* w = th.w.data() + (size_t)x * th.taps;
                for (int tp = 0; tp < th.xsize[x]; ++tp) {
                    if (w[tp] == 0) continue;
                    const int rel = th.xmin[x] + tp - first_px;
                    if (rel < 0 || rel >= 10) return done();  // taps outside the fixed placement
                    // the bytes past the copied segment are never read with a non-zero weight
                    if (3 * (th.xmin[x] + tp) + 2 >= st.g_start + st.len) return done();
                    wrow[rel >> 1] |= (uint32_t)(uint16_t)w[tp] << (16 * (rel & 1));
                }
            }
    }
    const int rows = t.gh * patch;
    std::vector<uint32_t> vq((size_t)rows * 8, 0u);
    std::vector<int2> unit(t.gh);
    for (int y = 0; y < rows; ++y) {
        const int rel = tv.xmin[y] & 1;
        if (tv.xsize[y] + rel > 2 * PS_VPAIRS) return done();
        uint32_t* row = vq.data() + (size_t)y * 8;
        const int16_t* w = tv.w.data() + (size_t)y * tv.taps;
        for (int tp = 0; tp < tv.xsize[y]; ++tp) row[(tp + rel) >> 1] |= (uint32_t)(uint16_t)w[tp] << (16 * ((tp + rel) & 1));
        for (int i = 0; i < PS_VPAIRS; ++i)  // ring positions of the window's row pairs, as uint2 indices
            row[4 + (i >> 1)] |= (uint32_t)((((tv.xmin[y] >> 1) + i) % PS_SH_PAIRS) * 32) << (16 * (i & 1));
    }
    for (int u = 0; u < t.gh; ++u) {
        const int ya = u * patch, yb = ya + patch - 1;
        unit[u].x = tv.xmin[ya] >> 1;
        unit[u].y = (tv.xmin[yb] + std::max(tv.xsize[yb], 1) - 1) >> 1;
        for (int y = ya; y <= yb; ++y) {
            unit[u].x = std::min(unit[u].x, tv.xmin[y] >> 1);
            unit[u].y = std::max(unit[u].y, (tv.xmin[y] + std::max(tv.xsize[y], 1) - 1) >> 1);
        }
        if (unit[u].y - unit[u].x + 1 > PS_SH_PAIRS || 2 * unit[u].y + 1 >= H + 1) return done();
        if (u > 0 && (unit[u].x < unit[u - 1].x || unit[u].y < unit[u - 1].y)) return done();
    }
    if (upload(hq, &t.hq) || upload(vq, &t.vq) || upload(unit, &t.unit)) return 2;
    t.ok = true;
    return done();
}

int launch_stream5(const uint8_t* frames, int B, int H, int W, int out_h, int out_w, int resample, const float* h_sub,
                   const float* h_div, void* out, int patch, int ld, cudaStream_t s) {
    if (W != PS_S * out_w || resample != GVL_RESAMPLE_BILINEAR || H < out_h) return -1;
    if ((uintptr_t)frames % 16 != 0 || (W * 3) % 16 != 0 || patch > 32 || patch * 3 > 255) return -1;
    PsTables t;
    int rc = ps_get_tables(H, W, out_h, out_w, resample, patch, ld, t);
    if (rc) return rc;
    if (!t.ok) return -1;
    PsParams p;
    memset(&p, 0, sizeof(p));
    float h_lut[768];
    float* lut = nullptr;
    rc = get_lut(h_sub, h_div, &lut, h_lut);
    if (rc) return rc;
    bool arith = true;
    for (int c = 0; c < 3; ++c) {
        volatile float na = 1.0f / h_div[c];
        volatile float nb = -h_sub[c] / h_div[c];
        p.na[c] = na;
        p.nb[c] = nb;
        for (int u = 0; u < 256 && arith; ++u)
            if (ps_bf16_rn(fmaf((float)u, p.na[c], p.nb[c])) != ps_bf16_rn(h_lut[c * 256 + u])) arith = false;
    }
    p.frames = frames;
    p.B = B;
    p.H = H;
    p.W = W;
    p.gh = t.gh;
    p.gw = t.gw;
    p.patch = patch;
    p.ld = ld;
    p.n_strips = t.n_strips;
    p.h_prec = t.h_prec;
    p.v_prec = t.v_prec;
    p.hq = t.hq;
    p.vq = t.vq;
    p.unit = t.unit;
    p.lut = lut;
    p.out = out;
    memcpy(p.strip, t.strip, sizeof(p.strip));
    // patch rows per CTA: long runs filter fewer rows twice, short runs fill the machine when the batch is small
    const int slots = sm_count() * 3;
    int run_len = 1;
    for (int cand : {3, 2}) {
        const long ctas = (long)t.n_strips * ((t.gh + cand - 1) / cand) * B;
        if (ctas >= 3L * slots) {
            run_len = cand;
            break;
        }
    }
    p.run_len = run_len;
    const int n_runs = (t.gh + run_len - 1) / run_len;
    const int band_elems = ((PS_COLS / patch) * ld + 7) & ~7;
    const size_t smem = (size_t)PS_NW * 4 * PS_PITCH + (size_t)3 * PS_SH_PAIRS * 64 * 4 + (size_t)band_elems * 2 +
                        32 * 8 * 4 + (arith ? 0 : 768 * 4) + PS_NW * 2 * 8;
    if (smem > 200 * 1024) return -1;
    dim3 grid(t.n_strips * n_runs, B, 1);
    ProfScope prof(GVL_K_PREPROCESS, (double)B * ((double)H * W * 3 + (double)t.gh * t.gw * 3 * patch * patch * 2), s);
    if (arith) {
        GVL_CUDA(cudaFuncSetAttribute(preprocess_stream5_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        preprocess_stream5_kernel<true><<<grid, PS_THREADS, smem, s>>>(p);
    } else {
        GVL_CUDA(cudaFuncSetAttribute(preprocess_stream5_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        preprocess_stream5_kernel<false><<<grid, PS_THREADS, smem, s>>>(p);
    }
    GVL_LAUNCH_CHECK("preprocess_stream5_kernel");
    return 0;
}

}  // namespace gvl
