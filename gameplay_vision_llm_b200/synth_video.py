"""Synthetic video files with exactly known content, for the ingest tests and benches (there is no encoder, no
network and no sample video in this image).

`write_h264_pcm` emits a valid H.264 (ITU-T H.264 §7.3) Annex-B elementary stream — or the same stream in an MP4
container — whose macroblocks are all I_PCM: the samples are stored uncompressed, so ANY conforming decoder (NVDEC,
ffmpeg) must reproduce the input Y'CbCr planes bit for bit.  That makes the decoder side of the ingest a known-answer
test: frame i of the file is `planes[i]`, exactly.  Optional P frames made of skipped macroblocks (a copy of the previous
picture) exercise inter prediction / reference handling without changing the known answer.

Baseline profile, CAVLC, one slice per picture, POC type 2 (output order = decode order), deblocking disabled.
"""
from __future__ import annotations

import re
import struct

import numpy as np


class _Bits:
    def __init__(self):
        self.out = bytearray()
        self.cur = 0
        self.n = 0

    def u(self, bits: int, value: int) -> None:
        for i in range(bits - 1, -1, -1):
            self.cur = (self.cur << 1) | ((value >> i) & 1)
            self.n += 1
            if self.n == 8:
                self.out.append(self.cur)
                self.cur, self.n = 0, 0

    def ue(self, v: int) -> None:
        v += 1
        nbits = v.bit_length()
        self.u(nbits - 1, 0)
        self.u(nbits, v)

    def se(self, v: int) -> None:
        self.ue(2 * v - 1 if v > 0 else -2 * v)

    def align_zero(self) -> None:
        while self.n:
            self.u(1, 0)

    def bytes_aligned(self, data: bytes) -> None:
        assert self.n == 0
        self.out += data

    def trailing(self) -> None:
        self.u(1, 1)
        self.align_zero()


_EPB = re.compile(rb"\x00\x00(?=[\x00-\x03])")


def _nal(ref_idc: int, nal_type: int, rbsp: bytes) -> bytes:
    """NAL unit payload (no start code): header byte + RBSP with emulation prevention bytes (§7.4.1)."""
    return bytes([(ref_idc << 5) | nal_type]) + _EPB.sub(b"\x00\x00\x03", rbsp)


def _sps(width: int, height: int, fps_num: int, fps_den: int, full_range: bool, matrix: int) -> bytes:
    mbw, mbh = (width + 15) // 16, (height + 15) // 16
    b = _Bits()
    b.u(8, 66)        # profile_idc: baseline
    b.u(8, 0xC0)      # constraint_set0/1
    b.u(8, 42 if mbw * mbh > 1620 else 31)  # level_idc
    b.ue(0)           # seq_parameter_set_id
    b.ue(0)           # log2_max_frame_num_minus4
    b.ue(2)           # pic_order_cnt_type
    b.ue(1)           # max_num_ref_frames
    b.u(1, 0)         # gaps_in_frame_num_value_allowed_flag
    b.ue(mbw - 1)
    b.ue(mbh - 1)
    b.u(1, 1)         # frame_mbs_only_flag
    b.u(1, 1)         # direct_8x8_inference_flag
    crop_r, crop_b = mbw * 16 - width, mbh * 16 - height
    if crop_r or crop_b:
        assert crop_r % 2 == 0 and crop_b % 2 == 0, "4:2:0 cropping is in units of two samples"
        b.u(1, 1)
        b.ue(0); b.ue(crop_r // 2); b.ue(0); b.ue(crop_b // 2)  # noqa: E702
    else:
        b.u(1, 0)
    b.u(1, 1)         # vui_parameters_present_flag
    b.u(1, 0)         # aspect_ratio_info_present_flag
    b.u(1, 0)         # overscan_info_present_flag
    b.u(1, 1)         # video_signal_type_present_flag
    b.u(3, 5)         # video_format: unspecified
    b.u(1, 1 if full_range else 0)
    b.u(1, 1)         # colour_description_present_flag
    b.u(8, matrix); b.u(8, matrix); b.u(8, matrix)  # primaries / transfer / matrix (6 = BT.601, 1 = BT.709)  # noqa: E702
    b.u(1, 0)         # chroma_loc_info_present_flag
    b.u(1, 1)         # timing_info_present_flag
    b.u(32, fps_den)  # num_units_in_tick
    b.u(32, 2 * fps_num)  # time_scale (field rate)
    b.u(1, 1)         # fixed_frame_rate_flag
    b.u(1, 0); b.u(1, 0)  # nal / vcl hrd  # noqa: E702
    b.u(1, 0)         # pic_struct_present_flag
    b.u(1, 1)         # bitstream_restriction_flag
    b.u(1, 1)         # motion_vectors_over_pic_boundaries_flag
    b.ue(0); b.ue(0)  # max_bytes_per_pic_denom, max_bits_per_mb_denom: unlimited (PCM)  # noqa: E702
    b.ue(16); b.ue(16)  # log2_max_mv_length_*  # noqa: E702
    b.ue(0)           # max_num_reorder_frames
    b.ue(1)           # max_dec_frame_buffering
    b.trailing()
    return _nal(3, 7, bytes(b.out))


def _pps() -> bytes:
    b = _Bits()
    b.ue(0); b.ue(0)  # pps id, sps id  # noqa: E702
    b.u(1, 0)         # entropy_coding_mode_flag: CAVLC
    b.u(1, 0)         # bottom_field_pic_order_in_frame_present_flag
    b.ue(0)           # num_slice_groups_minus1
    b.ue(0); b.ue(0)  # num_ref_idx_l0/l1_default_active_minus1  # noqa: E702
    b.u(1, 0); b.u(2, 0)  # weighted_pred_flag, weighted_bipred_idc  # noqa: E702
    b.se(0); b.se(0); b.se(0)  # pic_init_qp/qs, chroma_qp_index_offset  # noqa: E702
    b.u(1, 1)         # deblocking_filter_control_present_flag
    b.u(1, 0); b.u(1, 0)  # constrained_intra_pred_flag, redundant_pic_cnt_present_flag  # noqa: E702
    b.trailing()
    return _nal(3, 8, bytes(b.out))


def _idr_slice(y: np.ndarray, cb: np.ndarray, cr: np.ndarray, idr_pic_id: int) -> bytes:
    """One IDR picture, every macroblock I_PCM.  y [16*mbh, 16*mbw], cb / cr [8*mbh, 8*mbw] uint8."""
    mbh, mbw = y.shape[0] // 16, y.shape[1] // 16
    b = _Bits()
    b.ue(0)           # first_mb_in_slice
    b.ue(7)           # slice_type: I (all slices of the picture)
    b.ue(0)           # pic_parameter_set_id
    b.u(4, 0)         # frame_num
    b.ue(idr_pic_id)
    b.u(1, 0); b.u(1, 0)  # no_output_of_prior_pics_flag, long_term_reference_flag  # noqa: E702
    b.se(0)           # slice_qp_delta
    b.ue(1)           # disable_deblocking_filter_idc
    # macroblock layer: mb_type ue(25) = I_PCM, pcm_alignment_zero_bits, 256 + 64 + 64 raw samples.  After the first
    # macroblock every one starts byte aligned, so its prefix is the two bytes of ue(25) padded with zero bits.
    ymb = y.reshape(mbh, 16, mbw, 16).transpose(0, 2, 1, 3).reshape(mbh * mbw, 256)
    cbmb = cb.reshape(mbh, 8, mbw, 8).transpose(0, 2, 1, 3).reshape(mbh * mbw, 64)
    crmb = cr.reshape(mbh, 8, mbw, 8).transpose(0, 2, 1, 3).reshape(mbh * mbw, 64)
    b.ue(25)
    b.align_zero()
    rest = np.empty((mbh * mbw, 2 + 384), np.uint8)
    rest[:, 0], rest[:, 1] = 0x0D, 0x00          # ue(25) = 0000 1101 0 + 7 alignment zero bits
    rest[:, 2:258], rest[:, 258:322], rest[:, 322:386] = ymb, cbmb, crmb
    flat = rest.reshape(-1)[2:]                   # the first macroblock's prefix is already in the bit writer
    b.bytes_aligned(flat.tobytes())
    b.trailing()
    return _nal(3, 5, bytes(b.out))


def _p_skip_slice(n_mbs: int, frame_num: int) -> bytes:
    """A P picture whose macroblocks are all skipped: a copy of the previous picture."""
    b = _Bits()
    b.ue(0)           # first_mb_in_slice
    b.ue(5)           # slice_type: P
    b.ue(0)           # pic_parameter_set_id
    b.u(4, frame_num & 15)
    b.u(1, 0)         # num_ref_idx_active_override_flag
    b.u(1, 0)         # ref_pic_list_modification_flag_l0
    b.u(1, 0)         # adaptive_ref_pic_marking_mode_flag
    b.se(0)           # slice_qp_delta
    b.ue(1)           # disable_deblocking_filter_idc
    b.ue(n_mbs)       # mb_skip_run
    b.trailing()
    return _nal(2, 1, bytes(b.out))


def rgb_to_ycbcr420(frames: np.ndarray, full_range: bool = False, bt709: bool = False):
    """uint8 RGB [N,H,W,3] -> (Y [N,H,W], Cb [N,H/2,W/2], Cr) uint8, 2x2 box-averaged chroma.  Only used to make
    plausible content; the known answer of a decode test is the planes, not the RGB."""
    f = frames.astype(np.float64)
    kr, kb = (0.2126, 0.0722) if bt709 else (0.299, 0.114)
    yy = kr * f[..., 0] + (1 - kr - kb) * f[..., 1] + kb * f[..., 2]
    cb = (f[..., 2] - yy) / (2 * (1 - kb))
    cr = (f[..., 0] - yy) / (2 * (1 - kr))
    if full_range:
        y8, cb8, cr8 = yy, cb + 128, cr + 128
    else:
        y8, cb8, cr8 = 16 + yy * 219 / 255, 128 + cb * 224 / 255, 128 + cr * 224 / 255
    N, H, W = yy.shape
    box = lambda p: p.reshape(N, H // 2, 2, W // 2, 2).mean(axis=(2, 4))  # noqa: E731
    q = lambda p: np.clip(np.rint(p), 0, 255).astype(np.uint8)  # noqa: E731
    return q(y8), q(box(cb8)), q(box(cr8))


def h264_pcm_access_units(y: np.ndarray, cb: np.ndarray, cr: np.ndarray, fps=(30, 1), skip_every: int = 0,
                          full_range: bool = False, matrix: int = 6):
    """(sps, pps, [access unit = list of NAL payloads]) for planes y [N,H,W], cb / cr [N,H/2,W/2] (H, W even).
    skip_every = k > 0: after every coded picture, k - 1 P pictures of skipped macroblocks follow (so the stream has
    N * k frames and frame j shows planes[j // k])."""
    N, H, W = y.shape
    mbw, mbh = (W + 15) // 16, (H + 15) // 16
    sps, pps = _sps(W, H, fps[0], fps[1], full_range, matrix), _pps()
    aus = []
    for i in range(N):
        yp = np.zeros((mbh * 16, mbw * 16), np.uint8)
        cbp = np.full((mbh * 8, mbw * 8), 128, np.uint8)
        crp = np.full((mbh * 8, mbw * 8), 128, np.uint8)
        yp[:H, :W], cbp[:H // 2, :W // 2], crp[:H // 2, :W // 2] = y[i], cb[i], cr[i]
        aus.append([sps, pps, _idr_slice(yp, cbp, crp, i & 1)])
        for j in range(1, max(1, skip_every)):
            aus.append([_p_skip_slice(mbw * mbh, j)])
    return sps, pps, aus


def write_h264_annexb(path: str, y, cb, cr, **kw) -> int:
    """Annex-B elementary stream (start-code prefixed NAL units).  Returns the number of frames."""
    _, _, aus = h264_pcm_access_units(y, cb, cr, **kw)
    with open(path, "wb") as f:
        for au in aus:
            for n in au:
                f.write(b"\x00\x00\x00\x01" + n)
    return len(aus)


def _box(kind: bytes, *payload: bytes) -> bytes:
    body = b"".join(payload)
    return struct.pack(">I4s", 8 + len(body), kind) + body


def _full(kind: bytes, version_flags: int, *payload: bytes) -> bytes:
    return _box(kind, struct.pack(">I", version_flags), *payload)


def write_h264_mp4(path: str, y, cb, cr, fps=(30, 1), **kw) -> int:
    """The same stream in a minimal ISO-BMFF (.mp4) file: ftyp, moov (one avc1 video track, avcC with the SPS / PPS,
    4-byte NAL length prefixes, one chunk), mdat.  What OpenCV / decord / ffmpeg open like any camera or capture file."""
    N, H, W = y.shape
    sps, pps, aus = h264_pcm_access_units(y, cb, cr, fps=fps, **kw)
    samples = []
    for au in aus:
        vcl = [n for n in au if (n[0] & 31) in (1, 5)]
        samples.append(b"".join(struct.pack(">I", len(n)) + n for n in vcl))
    n = len(samples)
    timescale, delta = fps[0] * 1000, fps[1] * 1000
    sync = [i + 1 for i, au in enumerate(aus) if any((x[0] & 31) == 5 for x in au)]
    avcc = _box(b"avcC", bytes([1, sps[1], sps[2], sps[3], 0xFF, 0xE1]), struct.pack(">H", len(sps)), sps, bytes([1]),
                struct.pack(">H", len(pps)), pps)
    avc1 = _box(b"avc1", b"\x00" * 6, struct.pack(">H", 1), b"\x00" * 16, struct.pack(">HH", W, H),
                struct.pack(">II", 0x00480000, 0x00480000), b"\x00" * 4, struct.pack(">H", 1), b"\x00" * 32,
                struct.pack(">Hh", 0x18, -1), avcc)
    matrix = struct.pack(">9i", 0x10000, 0, 0, 0, 0x10000, 0, 0, 0, 0x40000000)
    ftyp = _box(b"ftyp", b"isom", struct.pack(">I", 0x200), b"isomiso2avc1mp41")

    def moov(chunk_offset: int) -> bytes:
        stbl = _box(b"stbl", _full(b"stsd", 0, struct.pack(">I", 1), avc1),
                    _full(b"stts", 0, struct.pack(">III", 1, n, delta)),
                    _full(b"stss", 0, struct.pack(">I", len(sync)), b"".join(struct.pack(">I", s) for s in sync)),
                    _full(b"stsc", 0, struct.pack(">IIII", 1, 1, n, 1)),
                    _full(b"stsz", 0, struct.pack(">II", 0, n), b"".join(struct.pack(">I", len(s)) for s in samples)),
                    _full(b"co64", 0, struct.pack(">IQ", 1, chunk_offset)))
        dinf = _box(b"dinf", _full(b"dref", 0, struct.pack(">I", 1), _full(b"url ", 1)))
        minf = _box(b"minf", _full(b"vmhd", 1, b"\x00" * 8), dinf, stbl)
        mdhd = _full(b"mdhd", 0, struct.pack(">IIIIHH", 0, 0, timescale, n * delta, 0x55C4, 0))
        hdlr = _full(b"hdlr", 0, struct.pack(">I4s", 0, b"vide"), b"\x00" * 12, b"VideoHandler\x00")
        tkhd = _full(b"tkhd", 3, struct.pack(">IIIII", 0, 0, 1, 0, n * delta), b"\x00" * 8, struct.pack(">hhhH", 0, 0, 0, 0),
                     matrix, struct.pack(">II", W << 16, H << 16))
        trak = _box(b"trak", tkhd, _box(b"mdia", mdhd, hdlr, minf))
        mvhd = _full(b"mvhd", 0, struct.pack(">IIII", 0, 0, timescale, n * delta), struct.pack(">IH", 0x10000, 0x100),
                     b"\x00" * 10, matrix, b"\x00" * 24, struct.pack(">I", 2))
        return _box(b"moov", mvhd, trak)

    head = len(ftyp) + len(moov(0)) + 16  # mdat with a 64-bit size
    body = b"".join(samples)
    with open(path, "wb") as f:
        f.write(ftyp)
        f.write(moov(head))
        f.write(struct.pack(">I4sQ", 1, b"mdat", 16 + len(body)))
        f.write(body)
    return n
