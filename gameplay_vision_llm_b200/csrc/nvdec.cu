// nvdec.cu — hardware video decode for the frame ingest (replaces the decode half of
// `extract_frames`, scripts/extract_features.py:230-264, which decodes every frame of the file to a host PIL image).
//
// The B200 has NVDEC engines; their user-mode driver (libnvcuvid.so.1) ships with the GPU driver, but this image has no
// Video Codec SDK headers, so the handful of entry points and structs used here are declared below from the public
// nvcuvid.h / cuviddec.h interface (the ABI is frozen across driver generations: every struct carries reserved tails
// that must be zero) and bound with dlopen at first use.  Nothing links against the library: on a box without it
// gvl_nvdec_available() returns 0 and the caller raises.
//
// Flow: the host demuxes the container (Python: frame_ingest.mp4_samples) and feeds elementary-stream bytes;
// cuvidParseVideoData splits them into pictures and calls back: sequence -> create the decoder at the display size;
// decode -> cuvidDecodePicture (asynchronous, on the engine); display (display order) -> every `interval`-th frame is
// mapped (NV12 in device memory) and converted to packed RGB straight into the caller's batch buffer — the frames the
// sampling rule skips are decoded (inter prediction needs them) but never converted, copied or sent to the host.
#include "common.cuh"
#include "cuvid_abi.h"

#include <dlfcn.h>

#include <cstdlib>

#include <mutex>
#include <new>

namespace gvl {
namespace cuvid {

struct Api {
    void* lib = nullptr;
    CUresult (*GetDecoderCaps)(CUVIDDECODECAPS*) = nullptr;
    CUresult (*CreateVideoParser)(CUvideoparser*, CUVIDPARSERPARAMS*) = nullptr;
    CUresult (*ParseVideoData)(CUvideoparser, CUVIDSOURCEDATAPACKET*) = nullptr;
    CUresult (*DestroyVideoParser)(CUvideoparser) = nullptr;
    CUresult (*CreateDecoder)(CUvideodecoder*, CUVIDDECODECREATEINFO*) = nullptr;
    CUresult (*DestroyDecoder)(CUvideodecoder) = nullptr;
    CUresult (*DecodePicture)(CUvideodecoder, void*) = nullptr;
    CUresult (*MapVideoFrame64)(CUvideodecoder, int, unsigned long long*, unsigned int*, CUVIDPROCPARAMS*) = nullptr;
    CUresult (*UnmapVideoFrame64)(CUvideodecoder, unsigned long long) = nullptr;
    bool ok = false;
};

static Api& api() {
    static Api a;
    static std::once_flag once;
    std::call_once(once, [] {
        // GVL_NVCUVID_LIB: another library with the cuvid entry points (tests/mock_nvcuvid: a software test double)
        const char* override_lib = getenv("GVL_NVCUVID_LIB");
        for (const char* name : {override_lib, "libnvcuvid.so.1", "libnvcuvid.so"}) {
            if (!name || !name[0]) continue;
            a.lib = dlopen(name, RTLD_NOW | RTLD_LOCAL);
            if (a.lib || name == override_lib) break;  // a named override that fails to load is an error, not a fallback
        }
        if (!a.lib) return;
        bool all = true;
        auto sym = [&](const char* n) {
            void* p = dlsym(a.lib, n);
            if (!p) all = false;
            return p;
        };
        a.GetDecoderCaps = reinterpret_cast<decltype(a.GetDecoderCaps)>(sym("cuvidGetDecoderCaps"));
        a.CreateVideoParser = reinterpret_cast<decltype(a.CreateVideoParser)>(sym("cuvidCreateVideoParser"));
        a.ParseVideoData = reinterpret_cast<decltype(a.ParseVideoData)>(sym("cuvidParseVideoData"));
        a.DestroyVideoParser = reinterpret_cast<decltype(a.DestroyVideoParser)>(sym("cuvidDestroyVideoParser"));
        a.CreateDecoder = reinterpret_cast<decltype(a.CreateDecoder)>(sym("cuvidCreateDecoder"));
        a.DestroyDecoder = reinterpret_cast<decltype(a.DestroyDecoder)>(sym("cuvidDestroyDecoder"));
        a.DecodePicture = reinterpret_cast<decltype(a.DecodePicture)>(sym("cuvidDecodePicture"));
        a.MapVideoFrame64 = reinterpret_cast<decltype(a.MapVideoFrame64)>(sym("cuvidMapVideoFrame64"));
        a.UnmapVideoFrame64 = reinterpret_cast<decltype(a.UnmapVideoFrame64)>(sym("cuvidUnmapVideoFrame64"));
        a.ok = all;
    });
    return a;
}

}  // namespace cuvid

// NV12 (pitch-linear luma plane + interleaved CbCr plane at half resolution) -> packed RGB [H, W, 3].
// One thread per horizontal pixel pair (they share a chroma sample; chroma is replicated, not interpolated — what
// swscale's default yuv420p -> rgb24 path does).  m = {ky, krv, kgu, kgv, kbu, y_offset}.
struct YuvMatrix {
    float ky, krv, kgu, kgv, kbu, yoff;
};

__global__ void __launch_bounds__(256)
nv12_to_rgb_kernel(const uint8_t* __restrict__ luma, const uint8_t* __restrict__ chroma, int pitch, int H, int W,
                   YuvMatrix m, uint8_t* __restrict__ rgb) {
    const int xp = blockIdx.x * blockDim.x + threadIdx.x;  // pixel pair
    const int y = blockIdx.y;
    if (2 * xp >= W) return;
    const uint8_t* lrow = luma + (size_t)y * pitch + 2 * xp;
    const uint8_t* crow = chroma + (size_t)(y >> 1) * pitch + 2 * xp;
    const float cb = (float)crow[0] - 128.f, cr = (float)crow[1] - 128.f;
    const float rr = m.krv * cr, gg = m.kgu * cb + m.kgv * cr, bb = m.kbu * cb;
    uint8_t* o = rgb + ((size_t)y * W + 2 * xp) * 3;
    const int n = (2 * xp + 1 < W) ? 2 : 1;
    for (int i = 0; i < n; ++i) {
        const float yy = m.ky * ((float)lrow[i] - m.yoff);
        o[3 * i + 0] = (uint8_t)__float2int_rn(fminf(fmaxf(yy + rr, 0.f), 255.f));
        o[3 * i + 1] = (uint8_t)__float2int_rn(fminf(fmaxf(yy + gg, 0.f), 255.f));
        o[3 * i + 2] = (uint8_t)__float2int_rn(fminf(fmaxf(yy + bb, 0.f), 255.f));
    }
}

static YuvMatrix yuv_matrix(int matrix_coefficients, int full_range) {
    // Kr / Kb of BT.709 when the stream says so (matrix_coefficients == 1), BT.601 otherwise (5, 6, unspecified):
    // ffmpeg's scale filter (decord's conversion) resolves "auto" the same way
    const double kr = matrix_coefficients == 1 ? 0.2126 : 0.299, kb = matrix_coefficients == 1 ? 0.0722 : 0.114;
    const double kg = 1.0 - kr - kb;
    const double ys = full_range ? 1.0 : 255.0 / 219.0, cs = full_range ? 1.0 : 255.0 / 224.0;
    YuvMatrix m;
    m.ky = (float)ys;
    m.krv = (float)(2.0 * (1.0 - kr) * cs);
    m.kbu = (float)(2.0 * (1.0 - kb) * cs);
    m.kgu = (float)(-2.0 * kb * (1.0 - kb) / kg * cs);
    m.kgv = (float)(-2.0 * kr * (1.0 - kr) / kg * cs);
    m.yoff = full_range ? 0.f : 16.f;
    return m;
}

}  // namespace gvl

struct gvl_nvdec {
    gvl::cuvid::CUvideoparser parser = nullptr;
    gvl::cuvid::CUvideodecoder decoder = nullptr;
    int codec = 0;
    gvl::cuvid::CUVIDEOFORMAT fmt = {};
    bool have_fmt = false;
    int disp_w = 0, disp_h = 0;
    // output of the current feed
    uint8_t* out = nullptr;
    int cap = 0, out_h = 0, out_w = 0, written = 0;
    long long first = 0, interval = 1, displayed = 0;
    cudaStream_t stream = nullptr;
    int matrix_override = -1, range_override = -1;
    char err[256] = {0};
    bool failed = false;
    void fail(const char* fmt_, ...) {
        if (failed) return;
        failed = true;
        va_list ap;
        va_start(ap, fmt_);
        vsnprintf(err, sizeof(err), fmt_, ap);
        va_end(ap);
    }
};

namespace gvl {

static int on_sequence(void* user, cuvid::CUVIDEOFORMAT* f) {
    gvl_nvdec* d = static_cast<gvl_nvdec*>(user);
    auto& a = cuvid::api();
    const int w = f->display_area.right - f->display_area.left, h = f->display_area.bottom - f->display_area.top;
    int surfaces = f->min_num_decode_surfaces > 0 ? f->min_num_decode_surfaces + 2 : 10;
    if (d->decoder) {
        if (d->have_fmt && f->coded_width == d->fmt.coded_width && f->coded_height == d->fmt.coded_height && w == d->disp_w &&
            h == d->disp_h && f->chroma_format == d->fmt.chroma_format && f->bit_depth_luma_minus8 == d->fmt.bit_depth_luma_minus8)
            return surfaces;  // the same sequence header again (every IDR of a camera / capture file repeats it)
        a.DestroyDecoder(d->decoder);
        d->decoder = nullptr;
    }
    if (f->chroma_format != cuvid::CHROMA_420 || f->bit_depth_luma_minus8 != 0) {
        d->fail("unsupported video format: chroma_format %d, %d-bit (8-bit 4:2:0 only)", f->chroma_format,
                8 + f->bit_depth_luma_minus8);
        return 0;
    }
    cuvid::CUVIDDECODECAPS caps = {};
    caps.eCodecType = f->codec;
    caps.eChromaFormat = f->chroma_format;
    caps.nBitDepthMinus8 = f->bit_depth_luma_minus8;
    CUresult rc = a.GetDecoderCaps(&caps);
    if (rc != CUDA_SUCCESS || !caps.bIsSupported) {
        d->fail("this GPU's NVDEC does not decode codec %d (cuvidGetDecoderCaps rc %d, supported %d)", f->codec, (int)rc,
                (int)caps.bIsSupported);
        return 0;
    }
    if (f->coded_width > caps.nMaxWidth || f->coded_height > caps.nMaxHeight) {
        d->fail("%ux%u exceeds NVDEC's %ux%u", f->coded_width, f->coded_height, caps.nMaxWidth, caps.nMaxHeight);
        return 0;
    }
    cuvid::CUVIDDECODECREATEINFO ci = {};
    ci.ulWidth = f->coded_width;
    ci.ulHeight = f->coded_height;
    ci.ulNumDecodeSurfaces = surfaces;
    ci.CodecType = f->codec;
    ci.ChromaFormat = f->chroma_format;
    ci.ulCreationFlags = cuvid::CREATE_PREFER_CUVID;
    ci.bitDepthMinus8 = 0;
    ci.ulMaxWidth = f->coded_width;
    ci.ulMaxHeight = f->coded_height;
    ci.display_area.left = (short)f->display_area.left;
    ci.display_area.top = (short)f->display_area.top;
    ci.display_area.right = (short)f->display_area.right;
    ci.display_area.bottom = (short)f->display_area.bottom;
    ci.OutputFormat = cuvid::SURFACE_NV12;
    ci.DeinterlaceMode = f->progressive_sequence ? cuvid::DEINTERLACE_WEAVE : cuvid::DEINTERLACE_ADAPTIVE;
    ci.ulTargetWidth = w;
    ci.ulTargetHeight = h;
    ci.ulNumOutputSurfaces = 2;
    rc = a.CreateDecoder(&d->decoder, &ci);
    if (rc != CUDA_SUCCESS) {
        d->decoder = nullptr;
        d->fail("cuvidCreateDecoder failed (CUresult %d) for %ux%u codec %d", (int)rc, f->coded_width, f->coded_height, f->codec);
        return 0;
    }
    d->fmt = *f;
    d->have_fmt = true;
    d->disp_w = w;
    d->disp_h = h;
    return surfaces;
}

static int on_decode(void* user, void* pic) {
    gvl_nvdec* d = static_cast<gvl_nvdec*>(user);
    if (!d->decoder) {
        d->fail("picture before any sequence header");
        return 0;
    }
    const CUresult rc = cuvid::api().DecodePicture(d->decoder, pic);
    if (rc != CUDA_SUCCESS) {
        d->fail("cuvidDecodePicture failed (CUresult %d)", (int)rc);
        return 0;
    }
    return 1;
}

static int on_display(void* user, cuvid::CUVIDPARSERDISPINFO* info) {
    gvl_nvdec* d = static_cast<gvl_nvdec*>(user);
    if (!info) return 1;  // end-of-stream notification
    const long long idx = d->displayed++;
    if (idx < d->first || (idx - d->first) % d->interval) return 1;  // decoded, never converted
    if (!d->out || d->written >= d->cap) {
        d->fail("output buffer full: %d frames kept in one feed (feed smaller chunks or pass a larger buffer)", d->cap);
        return 0;
    }
    if (d->out_w != d->disp_w || d->out_h != d->disp_h) {
        d->fail("output buffer is %dx%d, the video is %dx%d", d->out_w, d->out_h, d->disp_w, d->disp_h);
        return 0;
    }
    auto& a = cuvid::api();
    cuvid::CUVIDPROCPARAMS pp = {};
    pp.progressive_frame = info->progressive_frame;
    pp.second_field = info->repeat_first_field + 1;
    pp.top_field_first = info->top_field_first;
    pp.unpaired_field = info->repeat_first_field < 0;
    pp.output_stream = reinterpret_cast<CUstream>(d->stream);
    unsigned long long dptr = 0;
    unsigned int pitch = 0;
    CUresult rc = a.MapVideoFrame64(d->decoder, info->picture_index, &dptr, &pitch, &pp);
    if (rc != CUDA_SUCCESS) {
        d->fail("cuvidMapVideoFrame failed (CUresult %d)", (int)rc);
        return 0;
    }
    const uint8_t* luma = reinterpret_cast<const uint8_t*>(dptr);
    const uint8_t* chroma = luma + (size_t)pitch * ((d->disp_h + 1) & ~1);
    const int mat = d->matrix_override >= 0 ? d->matrix_override : d->fmt.video_signal_description.matrix_coefficients;
    const int full = d->range_override >= 0 ? d->range_override : d->fmt.video_signal_description.video_full_range_flag;
    uint8_t* dst = d->out + (size_t)d->written * d->out_h * d->out_w * 3;
    const dim3 grid(((d->disp_w + 1) / 2 + 255) / 256, d->disp_h);
    nv12_to_rgb_kernel<<<grid, 256, 0, d->stream>>>(luma, chroma, (int)pitch, d->disp_h, d->disp_w, yuv_matrix(mat, full), dst);
    cudaError_t le = cudaGetLastError();
    // the surface goes back to the decoder at unmap: the conversion must have read it by then
    cudaError_t se = cudaStreamSynchronize(d->stream);
    a.UnmapVideoFrame64(d->decoder, dptr);
    if (le != cudaSuccess || se != cudaSuccess) {
        d->fail("NV12 -> RGB conversion failed: %s", cudaGetErrorString(le != cudaSuccess ? le : se));
        return 0;
    }
    count_launch();
    d->written++;
    return 1;
}

}  // namespace gvl

extern "C" int gvl_nvdec_available(void) { return gvl::cuvid::api().ok ? 1 : 0; }

extern "C" int gvl_nvdec_caps(int codec, int* supported, int* max_w, int* max_h, int* n_engines) {
    using namespace gvl;
    auto& a = cuvid::api();
    GVL_CHECK_ARG(a.ok, "gvl_nvdec_caps: libnvcuvid.so.1 is not available on this machine");
    GVL_CUDA(cudaFree(0));  // the primary context must be current for the driver-level library
    cuvid::CUVIDDECODECAPS caps = {};
    caps.eCodecType = codec;
    caps.eChromaFormat = cuvid::CHROMA_420;
    const CUresult rc = a.GetDecoderCaps(&caps);
    GVL_CHECK_ARG(rc == CUDA_SUCCESS, "cuvidGetDecoderCaps failed (CUresult %d)", (int)rc);
    if (supported) *supported = caps.bIsSupported;
    if (max_w) *max_w = (int)caps.nMaxWidth;
    if (max_h) *max_h = (int)caps.nMaxHeight;
    if (n_engines) *n_engines = caps.nNumNVDECs;
    return 0;
}

extern "C" int gvl_nvdec_open(int codec, int max_display_delay, gvl_nvdec** out) {
    using namespace gvl;
    GVL_CHECK_ARG(out, "gvl_nvdec_open: null pointer");
    auto& a = cuvid::api();
    GVL_CHECK_ARG(a.ok, "gvl_nvdec_open: libnvcuvid.so.1 is not available on this machine (no hardware decode)");
    GVL_CUDA(cudaFree(0));
    gvl_nvdec* d = new (std::nothrow) gvl_nvdec();
    GVL_CHECK_ARG(d, "gvl_nvdec_open: out of memory");
    d->codec = codec;
    cuvid::CUVIDPARSERPARAMS pp = {};
    pp.CodecType = codec;
    pp.ulMaxNumDecodeSurfaces = 1;  // the sequence callback's return value sets the real count
    pp.ulClockRate = 0;
    pp.ulErrorThreshold = 0;
    pp.ulMaxDisplayDelay = max_display_delay < 0 ? 0 : (unsigned)max_display_delay;
    pp.pUserData = d;
    pp.pfnSequenceCallback = on_sequence;
    pp.pfnDecodePicture = on_decode;
    pp.pfnDisplayPicture = on_display;
    const CUresult rc = a.CreateVideoParser(&d->parser, &pp);
    if (rc != CUDA_SUCCESS) {
        delete d;
        set_error("cuvidCreateVideoParser failed (CUresult %d) for codec %d", (int)rc, codec);
        return 2;
    }
    *out = d;
    return 0;
}

extern "C" int gvl_nvdec_sampling(gvl_nvdec* d, long long first_frame, long long interval, int matrix_override,
                                  int full_range_override) {
    using namespace gvl;
    GVL_CHECK_ARG(d && first_frame >= 0 && interval >= 1, "gvl_nvdec_sampling: bad arguments");
    d->first = first_frame;
    d->interval = interval;
    d->matrix_override = matrix_override;
    d->range_override = full_range_override;
    return 0;
}

extern "C" int gvl_nvdec_feed(gvl_nvdec* d, const uint8_t* data, size_t bytes, int end_of_stream, uint8_t* out_rgb, int cap,
                              int H, int W, int* frames_written, long long* frames_displayed, void* stream) {
    using namespace gvl;
    GVL_CHECK_ARG(d && d->parser, "gvl_nvdec_feed: closed decoder");
    GVL_CHECK_ARG(!d->failed, "gvl_nvdec_feed: decoder already failed: %s", d->err);
    GVL_CHECK_ARG(data || bytes == 0, "gvl_nvdec_feed: null data");
    GVL_CHECK_ARG(cap >= 0 && (out_rgb || cap == 0), "gvl_nvdec_feed: bad output buffer");
    d->out = out_rgb;
    d->cap = cap;
    d->out_h = H;
    d->out_w = W;
    d->written = 0;
    d->stream = reinterpret_cast<cudaStream_t>(stream);
    cuvid::CUVIDSOURCEDATAPACKET pkt = {};
    pkt.payload = data;
    pkt.payload_size = bytes;
    pkt.flags = end_of_stream ? cuvid::PKT_ENDOFSTREAM : 0;
    const CUresult rc = cuvid::api().ParseVideoData(d->parser, &pkt);
    d->out = nullptr;
    if (frames_written) *frames_written = d->written;
    if (frames_displayed) *frames_displayed = d->displayed;
    GVL_CHECK_ARG(!d->failed, "gvl_nvdec_feed: %s", d->err);
    GVL_CHECK_ARG(rc == CUDA_SUCCESS, "cuvidParseVideoData failed (CUresult %d)", (int)rc);
    return 0;
}

extern "C" int gvl_nvdec_info(gvl_nvdec* d, int32_t* info8) {
    using namespace gvl;
    GVL_CHECK_ARG(d && info8, "gvl_nvdec_info: null pointer");
    GVL_CHECK_ARG(d->have_fmt, "gvl_nvdec_info: no sequence header seen yet");
    info8[0] = (int32_t)d->fmt.coded_width;
    info8[1] = (int32_t)d->fmt.coded_height;
    info8[2] = d->disp_w;
    info8[3] = d->disp_h;
    info8[4] = (int32_t)d->fmt.frame_rate.numerator;
    info8[5] = (int32_t)d->fmt.frame_rate.denominator;
    info8[6] = d->fmt.video_signal_description.matrix_coefficients;
    info8[7] = d->fmt.video_signal_description.video_full_range_flag;
    return 0;
}

extern "C" int gvl_nvdec_close(gvl_nvdec* d) {
    if (!d) return 0;
    auto& a = gvl::cuvid::api();
    if (d->parser) a.DestroyVideoParser(d->parser);
    if (d->decoder) a.DestroyDecoder(d->decoder);
    delete d;
    return 0;
}
