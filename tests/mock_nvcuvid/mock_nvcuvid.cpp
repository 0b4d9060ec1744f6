// mock_nvcuvid.cpp — TEST DOUBLE for libnvcuvid (test infrastructure only; never loaded by the product unless a test sets
// GVL_NVCUVID_LIB).  It implements the nine cuvid entry points csrc/nvdec.cu binds, in software, for the known-answer
// H.264 streams gameplay_vision_llm_b200/synth_video.py writes (baseline profile, every coded macroblock I_PCM, P pictures
// made of skipped macroblocks): an Annex-B splitter, an SPS parser, a PCM "decoder" that copies the stored samples, NV12
// surfaces in device memory, and the parser's callback protocol (sequence -> decode -> display in display order with a
// display delay, flush at end of stream).
//
// What it proves: the binding's callback flow, surface mapping, sampling rule, NV12 -> RGB kernel, batch assembly and
// error paths run end to end on a GPU.  What it cannot prove: that the struct layouts in csrc/cuvid_abi.h match the
// real driver's — both sides of this test include the same header.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstring>
#include <deque>
#include <vector>

#include "cuvid_abi.h"

using namespace gvl::cuvid;

namespace {

struct BitReader {
    const uint8_t* p;
    size_t n, pos = 0;  // pos in bits
    BitReader(const uint8_t* d, size_t len) : p(d), n(len) {}
    uint32_t u(int bits) {
        uint32_t v = 0;
        for (int i = 0; i < bits; ++i, ++pos) v = (v << 1) | ((pos >> 3) < n ? (p[pos >> 3] >> (7 - (pos & 7))) & 1u : 0u);
        return v;
    }
    uint32_t ue() {
        int zeros = 0;
        while (u(1) == 0 && zeros < 32) ++zeros;
        return zeros ? ((1u << zeros) - 1 + u(zeros)) : 0;
    }
    int se() {
        const uint32_t k = ue();
        return (k & 1) ? (int)((k + 1) / 2) : -(int)(k / 2);
    }
    void align() { pos = (pos + 7) & ~(size_t)7; }
};

std::vector<uint8_t> unescape(const uint8_t* d, size_t n) {  // remove emulation prevention bytes
    std::vector<uint8_t> o;
    o.reserve(n);
    int zeros = 0;
    for (size_t i = 0; i < n; ++i) {
        if (zeros >= 2 && d[i] == 3) {
            zeros = 0;
            continue;
        }
        o.push_back(d[i]);
        zeros = d[i] == 0 ? zeros + 1 : 0;
    }
    return o;
}

struct Picture {  // what the parser hands to pfnDecodePicture (opaque to the binding, like CUVIDPICPARAMS)
    int CurrPicIdx;
    int nal_type;
    std::vector<uint8_t> rbsp;
};

struct Decoder {
    int coded_w, coded_h, out_w, out_h, crop_l, crop_t;
    size_t pitch;
    std::vector<uint8_t*> surfaces;
    std::vector<uint8_t> ref;  // last decoded picture: Y plane then Cb then Cr, coded size
};

struct Parser {
    CUVIDPARSERPARAMS pp;
    std::vector<uint8_t> buf;
    bool have_sps = false;
    CUVIDEOFORMAT fmt;
    int surfaces = 1, next_surface = 0;
    std::deque<int> display_queue;
    long long decoded = 0;
};

bool parse_sps(const std::vector<uint8_t>& r, CUVIDEOFORMAT& f) {
    BitReader b(r.data() + 1, r.size() - 1);
    const uint32_t profile = b.u(8);
    b.u(8);
    b.u(8);
    b.ue();
    if (profile != 66) return false;
    b.ue();                      // log2_max_frame_num_minus4
    if (b.ue() != 2) return false;  // pic_order_cnt_type
    b.ue();
    b.u(1);
    const uint32_t mbw = b.ue() + 1, mbh = b.ue() + 1;
    if (!b.u(1)) return false;   // frame_mbs_only
    b.u(1);
    uint32_t cl = 0, cr = 0, ct = 0, cb = 0;
    if (b.u(1)) {
        cl = b.ue(); cr = b.ue(); ct = b.ue(); cb = b.ue();
    }
    std::memset(&f, 0, sizeof(f));
    f.codec = 4;
    f.frame_rate.numerator = 30;
    f.frame_rate.denominator = 1;
    f.progressive_sequence = 1;
    f.min_num_decode_surfaces = 4;
    f.coded_width = mbw * 16;
    f.coded_height = mbh * 16;
    f.display_area.left = (int)(2 * cl);
    f.display_area.top = (int)(2 * ct);
    f.display_area.right = (int)(mbw * 16 - 2 * cr);
    f.display_area.bottom = (int)(mbh * 16 - 2 * cb);
    f.chroma_format = CHROMA_420;
    f.video_signal_description.matrix_coefficients = 2;  // unspecified
    if (b.u(1)) {                // vui
        if (b.u(1)) { if (b.u(8) == 255) { b.u(16); b.u(16); } }
        if (b.u(1)) b.u(1);
        if (b.u(1)) {
            f.video_signal_description.video_format = b.u(3);
            f.video_signal_description.video_full_range_flag = b.u(1);
            if (b.u(1)) {
                f.video_signal_description.color_primaries = (unsigned char)b.u(8);
                f.video_signal_description.transfer_characteristics = (unsigned char)b.u(8);
                f.video_signal_description.matrix_coefficients = (unsigned char)b.u(8);
            }
        }
        if (b.u(1)) { b.ue(); b.ue(); }
        if (b.u(1)) {
            const uint32_t tick = b.u(32), scale = b.u(32);
            b.u(1);
            if (tick) {
                f.frame_rate.numerator = scale;
                f.frame_rate.denominator = 2 * tick;
            }
        }
    }
    return true;
}

void emit_display(Parser* ps, bool flush) {
    while (!ps->display_queue.empty() && (flush || ps->display_queue.size() > ps->pp.ulMaxDisplayDelay)) {
        CUVIDPARSERDISPINFO di = {};
        di.picture_index = ps->display_queue.front();
        di.progressive_frame = 1;
        di.top_field_first = 1;
        ps->display_queue.pop_front();
        if (ps->pp.pfnDisplayPicture && !ps->pp.pfnDisplayPicture(ps->pp.pUserData, &di)) return;
    }
}

bool handle_nal(Parser* ps, const uint8_t* d, size_t n) {
    if (n == 0) return true;
    const int type = d[0] & 31;
    if (type == 7) {
        CUVIDEOFORMAT f;
        if (!parse_sps(unescape(d, n), f)) return false;
        ps->fmt = f;
        ps->have_sps = true;
        const int r = ps->pp.pfnSequenceCallback ? ps->pp.pfnSequenceCallback(ps->pp.pUserData, &ps->fmt) : 1;
        if (r <= 0) return false;
        if (r > 1) ps->surfaces = r;
        return true;
    }
    if (type != 5 && type != 1) return true;  // PPS, SEI, ...: nothing to do for these streams
    if (!ps->have_sps) return false;
    Picture pic;
    pic.CurrPicIdx = ps->next_surface;
    ps->next_surface = (ps->next_surface + 1) % ps->surfaces;
    pic.nal_type = type;
    pic.rbsp = unescape(d, n);
    if (ps->pp.pfnDecodePicture && !ps->pp.pfnDecodePicture(ps->pp.pUserData, &pic)) return false;
    ps->decoded++;
    ps->display_queue.push_back(pic.CurrPicIdx);
    emit_display(ps, false);
    return true;
}

}  // namespace

extern "C" {

CUresult cuvidGetDecoderCaps(CUVIDDECODECAPS* c) {
    c->bIsSupported = (c->eCodecType == 4 && c->eChromaFormat == CHROMA_420 && c->nBitDepthMinus8 == 0) ? 1 : 0;
    c->nNumNVDECs = 1;
    c->nOutputFormatMask = 1;
    c->nMaxWidth = 8192;
    c->nMaxHeight = 8192;
    c->nMaxMBCount = 262144;
    c->nMinWidth = 48;
    c->nMinHeight = 16;
    return CUDA_SUCCESS;
}

CUresult cuvidCreateVideoParser(CUvideoparser* out, CUVIDPARSERPARAMS* p) {
    if (p->CodecType != 4) return CUDA_ERROR_NOT_SUPPORTED;
    Parser* ps = new Parser();
    ps->pp = *p;
    *out = ps;
    return CUDA_SUCCESS;
}

CUresult cuvidDestroyVideoParser(CUvideoparser h) {
    delete static_cast<Parser*>(h);
    return CUDA_SUCCESS;
}

CUresult cuvidParseVideoData(CUvideoparser h, CUVIDSOURCEDATAPACKET* pkt) {
    Parser* ps = static_cast<Parser*>(h);
    if (pkt->payload && pkt->payload_size) ps->buf.insert(ps->buf.end(), pkt->payload, pkt->payload + pkt->payload_size);
    const bool eos = (pkt->flags & PKT_ENDOFSTREAM) != 0;
    // NAL units are delimited by start codes; the last one in the buffer is complete only at end of stream
    const std::vector<uint8_t>& b = ps->buf;
    std::vector<size_t> starts;  // offset of the first byte after each 00 00 01
    for (size_t i = 0; i + 3 <= b.size(); ++i)
        if (b[i] == 0 && b[i + 1] == 0 && b[i + 2] == 1) {
            starts.push_back(i + 3);
            i += 2;
        }
    size_t consumed = 0;
    for (size_t k = 0; k < starts.size(); ++k) {
        const bool last = k + 1 == starts.size();
        if (last && !eos) break;
        size_t end = last ? b.size() : starts[k + 1] - 3;
        while (end > starts[k] && b[end - 1] == 0) --end;  // trailing zero_byte of a 4-byte start code
        if (!handle_nal(ps, b.data() + starts[k], end - starts[k])) return CUDA_ERROR_UNKNOWN;
        consumed = last ? b.size() : starts[k + 1] - 3;
    }
    ps->buf.erase(ps->buf.begin(), ps->buf.begin() + (long)consumed);
    if (eos) {
        emit_display(ps, true);
        ps->buf.clear();
    }
    return CUDA_SUCCESS;
}

CUresult cuvidCreateDecoder(CUvideodecoder* out, CUVIDDECODECREATEINFO* ci) {
    if (ci->CodecType != 4 || ci->ChromaFormat != CHROMA_420 || ci->OutputFormat != SURFACE_NV12) return CUDA_ERROR_NOT_SUPPORTED;
    Decoder* d = new Decoder();
    d->coded_w = (int)ci->ulWidth;
    d->coded_h = (int)ci->ulHeight;
    d->out_w = (int)ci->ulTargetWidth;
    d->out_h = (int)ci->ulTargetHeight;
    d->crop_l = ci->display_area.left;
    d->crop_t = ci->display_area.top;
    d->pitch = ((size_t)d->out_w + 255) & ~(size_t)255;
    const size_t rows = (size_t)((d->out_h + 1) & ~1) * 3 / 2;
    for (unsigned long i = 0; i < ci->ulNumDecodeSurfaces; ++i) {
        uint8_t* s = nullptr;
        if (cudaMalloc(&s, d->pitch * rows) != cudaSuccess) return CUDA_ERROR_OUT_OF_MEMORY;
        d->surfaces.push_back(s);
    }
    d->ref.assign((size_t)d->coded_w * d->coded_h * 3 / 2, 0);
    *out = d;
    return CUDA_SUCCESS;
}

CUresult cuvidDestroyDecoder(CUvideodecoder h) {
    Decoder* d = static_cast<Decoder*>(h);
    for (uint8_t* s : d->surfaces) cudaFree(s);
    delete d;
    return CUDA_SUCCESS;
}

CUresult cuvidDecodePicture(CUvideodecoder h, void* pic_) {
    Decoder* d = static_cast<Decoder*>(h);
    Picture* pic = static_cast<Picture*>(pic_);
    if (pic->CurrPicIdx < 0 || pic->CurrPicIdx >= (int)d->surfaces.size()) return CUDA_ERROR_INVALID_VALUE;
    const int W = d->coded_w, H = d->coded_h, mbw = W / 16, mbh = H / 16;
    uint8_t* Y = d->ref.data();
    uint8_t* Cb = Y + (size_t)W * H;
    uint8_t* Cr = Cb + (size_t)W * H / 4;
    if (pic->nal_type == 5) {  // IDR picture, every macroblock I_PCM
        BitReader b(pic->rbsp.data() + 1, pic->rbsp.size() - 1);
        b.ue();                       // first_mb_in_slice
        b.ue();                       // slice_type
        b.ue();                       // pic_parameter_set_id
        b.u(4);                       // frame_num
        b.ue();                       // idr_pic_id
        b.u(2);                       // dec_ref_pic_marking
        b.se();                       // slice_qp_delta
        b.ue();                       // disable_deblocking_filter_idc
        for (int mb = 0; mb < mbw * mbh; ++mb) {
            if (b.ue() != 25) return CUDA_ERROR_NOT_SUPPORTED;  // anything but I_PCM is beyond this test double
            b.align();
            const uint8_t* s = pic->rbsp.data() + 1 + (b.pos >> 3);
            if ((b.pos >> 3) + 384 > pic->rbsp.size() - 1) return CUDA_ERROR_UNKNOWN;
            const int my = mb / mbw, mx = mb % mbw;
            for (int r = 0; r < 16; ++r) std::memcpy(Y + (size_t)(my * 16 + r) * W + mx * 16, s + r * 16, 16);
            for (int r = 0; r < 8; ++r) std::memcpy(Cb + (size_t)(my * 8 + r) * (W / 2) + mx * 8, s + 256 + r * 8, 8);
            for (int r = 0; r < 8; ++r) std::memcpy(Cr + (size_t)(my * 8 + r) * (W / 2) + mx * 8, s + 320 + r * 8, 8);
            b.pos += 384 * 8;
        }
    }  // else: a P picture of skipped macroblocks = the reference picture unchanged
    // cropped NV12 into the surface
    std::vector<uint8_t> nv((size_t)d->pitch * ((d->out_h + 1) & ~1) * 3 / 2, 0);
    for (int y = 0; y < d->out_h; ++y) std::memcpy(nv.data() + (size_t)y * d->pitch, Y + (size_t)(y + d->crop_t) * W + d->crop_l, d->out_w);
    uint8_t* c = nv.data() + d->pitch * (size_t)((d->out_h + 1) & ~1);
    for (int y = 0; y < (d->out_h + 1) / 2; ++y)
        for (int x = 0; x < (d->out_w + 1) / 2; ++x) {
            const size_t src = (size_t)(y + d->crop_t / 2) * (W / 2) + x + d->crop_l / 2;
            c[(size_t)y * d->pitch + 2 * x] = Cb[src];
            c[(size_t)y * d->pitch + 2 * x + 1] = Cr[src];
        }
    return cudaMemcpy(d->surfaces[pic->CurrPicIdx], nv.data(), nv.size(), cudaMemcpyHostToDevice) == cudaSuccess
               ? CUDA_SUCCESS : CUDA_ERROR_UNKNOWN;
}

CUresult cuvidMapVideoFrame64(CUvideodecoder h, int idx, unsigned long long* dptr, unsigned int* pitch, CUVIDPROCPARAMS*) {
    Decoder* d = static_cast<Decoder*>(h);
    if (idx < 0 || idx >= (int)d->surfaces.size()) return CUDA_ERROR_INVALID_VALUE;
    *dptr = reinterpret_cast<unsigned long long>(d->surfaces[idx]);
    *pitch = (unsigned int)d->pitch;
    return CUDA_SUCCESS;
}

CUresult cuvidUnmapVideoFrame64(CUvideodecoder, unsigned long long) { return CUDA_SUCCESS; }

}  // extern "C"
