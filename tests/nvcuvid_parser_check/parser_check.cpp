// parser_check.cpp — test infrastructure: drives the REAL libnvcuvid bitstream parser (host code: it works even where
// the decode engine is not reachable) through the struct layouts of csrc/cuvid_abi.h and prints what its callbacks
// receive.  A wrong CUVIDPARSERPARAMS / CUVIDSOURCEDATAPACKET layout means no callbacks or a crash; a wrong
// CUVIDEOFORMAT / CUVIDPARSERDISPINFO layout means wrong numbers.  tests/test_nvdec_gpu.py feeds it a known stream.
//   usage: parser_check <annexb-file> [library]
#include <dlfcn.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "cuvid_abi.h"

using namespace gvl::cuvid;

struct State {
    int sequences = 0, decodes = 0, displays = 0, last_idx = -1, max_idx = -1;
};

static int on_seq(void* u, CUVIDEOFORMAT* f) {
    State* s = static_cast<State*>(u);
    if (s->sequences++ == 0)
        printf("sequence codec=%d coded=%ux%u display=%d,%d,%d,%d fps=%u/%u progressive=%d chroma=%d bitdepth=%d "
               "min_surfaces=%d matrix=%d full_range=%d\n",
               f->codec, f->coded_width, f->coded_height, f->display_area.left, f->display_area.top, f->display_area.right,
               f->display_area.bottom, f->frame_rate.numerator, f->frame_rate.denominator, (int)f->progressive_sequence,
               f->chroma_format, 8 + f->bit_depth_luma_minus8, (int)f->min_num_decode_surfaces,
               (int)f->video_signal_description.matrix_coefficients, (int)f->video_signal_description.video_full_range_flag);
    return f->min_num_decode_surfaces > 0 ? f->min_num_decode_surfaces + 2 : 8;
}
static int on_dec(void* u, void*) {
    static_cast<State*>(u)->decodes++;
    return 1;  // no engine here: pretend the picture was submitted
}
static int on_disp(void* u, CUVIDPARSERDISPINFO* d) {
    State* s = static_cast<State*>(u);
    if (!d) return 1;
    s->displays++;
    s->last_idx = d->picture_index;
    if (d->picture_index > s->max_idx) s->max_idx = d->picture_index;
    if (s->displays <= 3)
        printf("display #%d picture_index=%d progressive=%d timestamp=%lld\n", s->displays, d->picture_index,
               d->progressive_frame, d->timestamp);
    return 1;
}

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    void* lib = dlopen(argc > 2 ? argv[2] : "libnvcuvid.so.1", RTLD_NOW);
    if (!lib) {
        printf("nolib %s\n", dlerror());
        return 3;
    }
    auto create = reinterpret_cast<CUresult (*)(CUvideoparser*, CUVIDPARSERPARAMS*)>(dlsym(lib, "cuvidCreateVideoParser"));
    auto parse = reinterpret_cast<CUresult (*)(CUvideoparser, CUVIDSOURCEDATAPACKET*)>(dlsym(lib, "cuvidParseVideoData"));
    auto destroy = reinterpret_cast<CUresult (*)(CUvideoparser)>(dlsym(lib, "cuvidDestroyVideoParser"));
    if (!create || !parse || !destroy) return 4;
    FILE* f = fopen(argv[1], "rb");
    if (!f) return 5;
    std::vector<unsigned char> data;
    unsigned char buf[65536];
    size_t n;
    while ((n = fread(buf, 1, sizeof(buf), f)) > 0) data.insert(data.end(), buf, buf + n);
    fclose(f);
    State st;
    CUVIDPARSERPARAMS pp = {};
    pp.CodecType = 4;
    pp.ulMaxNumDecodeSurfaces = 1;
    pp.ulMaxDisplayDelay = 0;
    pp.pUserData = &st;
    pp.pfnSequenceCallback = on_seq;
    pp.pfnDecodePicture = on_dec;
    pp.pfnDisplayPicture = on_disp;
    CUvideoparser parser = nullptr;
    CUresult rc = create(&parser, &pp);
    printf("create rc=%d\n", (int)rc);
    if (rc != CUDA_SUCCESS) return 6;
    for (size_t off = 0; off < data.size(); off += 4096) {  // arbitrary chunking, like nvdec_ingest's raw mode
        CUVIDSOURCEDATAPACKET pkt = {};
        pkt.payload = data.data() + off;
        pkt.payload_size = data.size() - off < 4096 ? data.size() - off : 4096;
        rc = parse(parser, &pkt);
        if (rc != CUDA_SUCCESS) {
            printf("parse rc=%d at offset %zu\n", (int)rc, off);
            return 7;
        }
    }
    CUVIDSOURCEDATAPACKET eos = {};
    eos.flags = PKT_ENDOFSTREAM;
    rc = parse(parser, &eos);
    printf("eos rc=%d\n", (int)rc);
    printf("totals sequences=%d decodes=%d displays=%d max_picture_index=%d\n", st.sequences, st.decodes, st.displays, st.max_idx);
    destroy(parser);
    return 0;
}
