// cuvid_abi.h — the subset of the NVDEC user-mode interface (cuviddec.h / nvcuvid.h of the Video Codec SDK) that
// nvdec.cu binds with dlopen, declared here because this image ships the driver library's name but not the SDK headers
// (Linux x86-64 layout; every struct carries the SDK's reserved tail, which must be zero).  Also included by the test
// double tests/mock_nvcuvid/mock_nvcuvid.cpp, which implements these entry points in software for known-answer streams.
#pragma once

#include <cuda.h>

namespace gvl {
namespace cuvid {

// ---- the subset of cuviddec.h / nvcuvid.h used here (Linux x86-64 layout) ---------------------------------------
typedef void* CUvideodecoder;
typedef void* CUvideoparser;
typedef long long CUvideotimestamp;

struct CUVIDDECODECAPS {
    int eCodecType;               // IN
    int eChromaFormat;            // IN
    unsigned int nBitDepthMinus8; // IN
    unsigned int reserved1[3];
    unsigned char bIsSupported;   // OUT
    unsigned char nNumNVDECs;
    unsigned short nOutputFormatMask;
    unsigned int nMaxWidth, nMaxHeight, nMaxMBCount;
    unsigned short nMinWidth, nMinHeight;
    unsigned char bIsHistogramSupported, nCounterBitDepth;
    unsigned short nMaxHistogramBins;
    unsigned int reserved3[10];
    unsigned int tail_pad[16];  // not in the SDK: slack in case a newer driver writes a longer struct
};

struct CUVIDDECODECREATEINFO {
    unsigned long ulWidth, ulHeight, ulNumDecodeSurfaces;
    int CodecType, ChromaFormat;
    unsigned long ulCreationFlags, bitDepthMinus8, ulIntraDecodeOnly, ulMaxWidth, ulMaxHeight, Reserved1;
    struct { short left, top, right, bottom; } display_area;
    int OutputFormat, DeinterlaceMode;
    unsigned long ulTargetWidth, ulTargetHeight, ulNumOutputSurfaces;
    void* vidLock;
    struct { short left, top, right, bottom; } target_rect;
    unsigned long enableHistogram;
    unsigned long Reserved2[4];
    unsigned long tail_pad[8];
};

struct CUVIDPROCPARAMS {
    int progressive_frame, second_field, top_field_first, unpaired_field;
    unsigned int reserved_flags, reserved_zero;
    unsigned long long raw_input_dptr;
    unsigned int raw_input_pitch, raw_input_format;
    unsigned long long raw_output_dptr;
    unsigned int raw_output_pitch, Reserved1;
    CUstream output_stream;
    unsigned int Reserved[46];
    unsigned long long* histogram_dptr;
    void* Reserved2[1];
    unsigned long tail_pad[8];
};

struct CUVIDEOFORMAT {
    int codec;
    struct { unsigned int numerator, denominator; } frame_rate;
    unsigned char progressive_sequence, bit_depth_luma_minus8, bit_depth_chroma_minus8, min_num_decode_surfaces;
    unsigned int coded_width, coded_height;
    struct { int left, top, right, bottom; } display_area;
    int chroma_format;
    unsigned int bitrate;
    struct { int x, y; } display_aspect_ratio;
    struct {
        unsigned char video_format : 3;
        unsigned char video_full_range_flag : 1;
        unsigned char reserved_zero_bits : 4;
        unsigned char color_primaries, transfer_characteristics, matrix_coefficients;
    } video_signal_description;
    unsigned int seqhdr_data_length;
};

struct CUVIDSOURCEDATAPACKET {
    unsigned long flags, payload_size;
    const unsigned char* payload;
    CUvideotimestamp timestamp;
};

struct CUVIDPARSERDISPINFO {
    int picture_index, progressive_frame, top_field_first, repeat_first_field;
    CUvideotimestamp timestamp;
};

typedef int (*PFNVIDSEQUENCECALLBACK)(void*, CUVIDEOFORMAT*);
typedef int (*PFNVIDDECODECALLBACK)(void*, void* /* CUVIDPICPARAMS*: passed through untouched */);
typedef int (*PFNVIDDISPLAYCALLBACK)(void*, CUVIDPARSERDISPINFO*);

struct CUVIDPARSERPARAMS {
    int CodecType;
    unsigned int ulMaxNumDecodeSurfaces, ulClockRate, ulErrorThreshold, ulMaxDisplayDelay;
    unsigned int bAnnexb : 1;
    unsigned int uReserved : 31;
    unsigned int uReserved1[4];
    void* pUserData;
    PFNVIDSEQUENCECALLBACK pfnSequenceCallback;
    PFNVIDDECODECALLBACK pfnDecodePicture;
    PFNVIDDISPLAYCALLBACK pfnDisplayPicture;
    void* pfnGetOperatingPoint;  // AV1 only
    void* pfnGetSEIMsg;
    void* pvReserved2[5];
    void* pExtVideoInfo;
    void* tail_pad[8];
};

enum { PKT_ENDOFSTREAM = 0x01, PKT_TIMESTAMP = 0x02 };
enum { SURFACE_NV12 = 0, CHROMA_420 = 1, DEINTERLACE_WEAVE = 0, DEINTERLACE_ADAPTIVE = 2, CREATE_PREFER_CUVID = 4 };

}  // namespace cuvid
}  // namespace gvl
