"""Independent MP3 decoder for the tests: FFmpeg's `mp3float` from the libavcodec that ships inside the image's
opencv wheel, driven through ctypes (no headers: the few struct offsets used are checked at run time — every decoded
frame must report 1152 samples).  Replaces AVFoundation's AVAudioFile of the reference's decode tests (TST:653-660)."""
import ctypes as C
import glob
import os

import numpy as np

_LIBS = None


def _load():
    global _LIBS
    if _LIBS is not None:
        return _LIBS
    import importlib.util
    spec = importlib.util.find_spec("cv2")
    if spec is None:
        raise RuntimeError("opencv wheel (source of libavcodec) not installed")
    base = os.path.join(os.path.dirname(os.path.dirname(spec.origin)), "opencv_python_headless.libs")
    order = ["libdrm", "libcrypto", "libssl", "libpng16", "libaom", "libvpx", "libavutil", "libswresample", "libavcodec"]
    libs = {}
    for name in order:
        hits = sorted(glob.glob(os.path.join(base, name + "-*.so*")))
        if not hits:
            raise RuntimeError("missing " + name)
        libs[name] = C.CDLL(hits[0], mode=C.RTLD_GLOBAL)
    av, au = libs["libavcodec"], libs["libavutil"]
    av.avcodec_find_decoder_by_name.restype = C.c_void_p; av.avcodec_find_decoder_by_name.argtypes = [C.c_char_p]
    av.avcodec_alloc_context3.restype = C.c_void_p; av.avcodec_alloc_context3.argtypes = [C.c_void_p]
    av.avcodec_open2.restype = C.c_int; av.avcodec_open2.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    av.avcodec_free_context.argtypes = [C.POINTER(C.c_void_p)]
    av.av_packet_alloc.restype = C.c_void_p
    av.av_packet_free.argtypes = [C.POINTER(C.c_void_p)]
    av.avcodec_send_packet.restype = C.c_int; av.avcodec_send_packet.argtypes = [C.c_void_p, C.c_void_p]
    av.avcodec_receive_frame.restype = C.c_int; av.avcodec_receive_frame.argtypes = [C.c_void_p, C.c_void_p]
    au.av_frame_alloc.restype = C.c_void_p
    au.av_frame_free.argtypes = [C.POINTER(C.c_void_p)]
    au.av_frame_unref.argtypes = [C.c_void_p]
    au.av_log_set_level.argtypes = [C.c_int]
    au.av_log_set_level(-8)                       # AV_LOG_QUIET
    _LIBS = (av, au)
    return _LIBS


def split_frames(data):
    """Frame boundaries from the MPEG-1 Layer III headers alone."""
    br = [0, 32, 40, 48, 56, 64, 80, 96, 112, 128, 160, 192, 224, 256, 320, 0]
    sr = [44100, 48000, 32000, 0]
    out, i = [], 0
    while i + 4 <= len(data):
        h = int.from_bytes(data[i:i + 4], "big")
        assert (h >> 21) == 0x7FF, "lost sync at byte %d" % i
        size = 144 * br[(h >> 12) & 15] * 1000 // sr[(h >> 10) & 3] + ((h >> 9) & 1)
        out.append(data[i:i + size]); i += size
    return out


def decode(data):
    """Decode a raw MP3 byte stream; returns (pcm float32 [channels, samples], frames decoded, frames rejected)."""
    av, au = _load()
    codec = av.avcodec_find_decoder_by_name(b"mp3float")
    assert codec, "mp3float decoder not in this libavcodec"
    ctx = av.avcodec_alloc_context3(codec)
    assert av.avcodec_open2(ctx, codec, None) == 0
    pkt, frame = av.av_packet_alloc(), au.av_frame_alloc()
    pad = 64                                       # AV_INPUT_BUFFER_PADDING_SIZE
    chans, ok, bad = None, 0, 0
    for fr in split_frames(data):
        buf = C.create_string_buffer(bytes(fr) + b"\0" * pad, len(fr) + pad)
        C.c_void_p.from_address(pkt + 24).value = C.addressof(buf)      # AVPacket.data
        C.c_int.from_address(pkt + 32).value = len(fr)                   # AVPacket.size
        if av.avcodec_send_packet(ctx, pkt) != 0:
            bad += 1
            continue
        while av.avcodec_receive_frame(ctx, frame) == 0:
            nb = C.c_int.from_address(frame + 112).value                # AVFrame.nb_samples
            assert nb == 1152, "unexpected AVFrame layout (nb_samples = %d)" % nb
            planes = []
            for c in range(2):
                p = C.c_void_p.from_address(frame + 8 * c).value         # AVFrame.data[c], planar float
                if p:
                    planes.append(np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), (nb,)).copy())
            if chans is None:
                chans = [[] for _ in planes]
            for c, pl in enumerate(planes[:len(chans)]):
                chans[c].append(pl)
            ok += 1
            au.av_frame_unref(frame)
    fp, pp, cp = C.c_void_p(frame), C.c_void_p(pkt), C.c_void_p(ctx)
    au.av_frame_free(C.byref(fp)); av.av_packet_free(C.byref(pp)); av.avcodec_free_context(C.byref(cp))
    pcm = np.stack([np.concatenate(c) for c in chans]) if chans else np.zeros((0, 0), np.float32)
    return pcm, ok, bad


def shift_global_gain(stream, delta):
    """A copy of an MPEG-1 Layer III stream with `delta` added to every global_gain field (the decoded signal scales by
    2^(delta / 4); CRC words are left as they are — the decoder does not verify them).  Used because this libavcodec build keeps
    its samples on the 16-bit scale and its l3_unscale() overflows to 0 for escape-coded values (|ix| >= 15) at the gains a
    full-scale signal needs: decoding the same bits 15 dB or more lower and scaling back shows the encoder's own accuracy."""
    import mp3parse
    m = bytearray(stream)
    for f in mp3parse.parse_frames(stream):
        ch = f["ch"]
        base = (f["pos"] + 4 + (0 if f["protection"] else 2)) * 8 + 9 + (5 if ch == 1 else 3) + 4 * ch
        for k in range(2 * ch):
            bp = base + 59 * k + 21
            g = 0
            for i in range(8):
                g = g << 1 | (m[(bp + i) >> 3] >> (7 - ((bp + i) & 7)) & 1)
            g = min(max(g + delta, 0), 255)
            for i in range(8):
                q = bp + i
                m[q >> 3] = (m[q >> 3] & ~(0x80 >> (q & 7))) | ((0x80 >> (q & 7)) if (g >> (7 - i)) & 1 else 0)
    return bytes(m)


def decode_unit_scale(stream, delta=-60):
    """decode() on the [-1, 1] scale of the encoder's input, through shift_global_gain (see there)."""
    pcm, ok, bad = decode(shift_global_gain(stream, delta))
    return pcm.astype(np.float64) * (2.0 ** (-delta / 4.0) / 32768.0), ok, bad
