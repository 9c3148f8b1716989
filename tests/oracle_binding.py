"""ctypes binding of the CPU oracle (oracle/libmp3oracle.so).  TEST INFRASTRUCTURE ONLY: imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, never by the product package."""
import ctypes as C
import os
import subprocess
import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.path.join(_ROOT, "oracle", "libmp3oracle.so")


class Options(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ("sample_rate", "bitrate_kbps", "vbr", "mode", "quality", "crc_protected", "original", "copyright")]


class ID3(C.Structure):
    _fields_ = [("title", C.c_char_p), ("artist", C.c_char_p), ("album", C.c_char_p), ("genre", C.c_char_p),
                ("comment", C.c_char_p), ("track", C.c_int32), ("track_total", C.c_int32), ("year", C.c_int32),
                ("album_art", C.c_void_p), ("album_art_len", C.c_size_t), ("album_art_mime", C.c_char_p)]


GC_TRACE = np.dtype([("spectrum", "<f4", 576), ("subband", "<f4", 576), ("thresholds", "<f4", 576), ("ix", "<i4", 576),
                     ("energy", "<f4"), ("sub_energy", "<f4", 3), ("block_type", "<i4"), ("mixed", "<i4"),
                     ("window_switching", "<i4"), ("subblock_gain", "<i4", 3), ("g0", "<i4"), ("gain_out", "<i4"),
                     ("gain_used", "<i4"), ("iterations", "<i4"), ("bits", "<i4"), ("max_bits", "<i4"),
                     ("big_values", "<i4"), ("region0", "<i4"), ("region1", "<i4"), ("preflag", "<i4"),
                     ("frame", "<i4"), ("gr", "<i4"), ("ch", "<i4")])
FRAME_TRACE = np.dtype([("frame_energy", "<f4"), ("ms", "<i4"), ("bitrate_kbps", "<i4"), ("bitrate_index", "<i4"),
                        ("padding", "<i4"), ("frame_size", "<i4"), ("main_data_size", "<i4"), ("main_data_begin", "<i4"),
                        ("reservoir_bits", "<i4"), ("huff_bytes", "<i4"), ("is_final", "<i4")])

_lib = None
_native = None


def build():
    subprocess.check_call(["make", "-s", "-C", os.path.join(_ROOT, "oracle")])


def _declare(L):
    L.orc_create.restype = C.c_void_p; L.orc_create.argtypes = [C.POINTER(Options)]
    L.orc_destroy.argtypes = [C.c_void_p]
    L.orc_clone.restype = C.c_void_p; L.orc_clone.argtypes = [C.c_void_p]
    for f in (L.orc_encode,):
        f.restype = C.POINTER(C.c_uint8); f.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
    for f in (L.orc_flush, L.orc_xing_header):
        f.restype = C.POINTER(C.c_uint8); f.argtypes = [C.c_void_p, C.POINTER(C.c_size_t)]
    L.orc_frame_count.restype = C.c_uint32; L.orc_frame_count.argtypes = [C.c_void_p]
    L.orc_byte_count.restype = C.c_uint32; L.orc_byte_count.argtypes = [C.c_void_p]
    L.orc_id3_build.restype = C.c_void_p; L.orc_id3_build.argtypes = [C.POINTER(ID3), C.POINTER(C.c_size_t)]
    L.orc_free.argtypes = [C.c_void_p]
    L.orc_trace_enable.argtypes = [C.c_void_p, C.c_int]
    L.orc_trace_gc_count.restype = C.c_size_t; L.orc_trace_gc_count.argtypes = [C.c_void_p]
    L.orc_trace_gc.restype = C.c_void_p; L.orc_trace_gc.argtypes = [C.c_void_p]
    L.orc_trace_frame_count.restype = C.c_size_t; L.orc_trace_frame_count.argtypes = [C.c_void_p]
    L.orc_trace_frames.restype = C.c_void_p; L.orc_trace_frames.argtypes = [C.c_void_p]
    L.orc_trace_clear.argtypes = [C.c_void_p]
    for name, n, t in (("window", 512, C.c_float), ("analysis", 2048, C.c_float), ("mdct_long", 648, C.c_float),
                       ("mdct_short", 72, C.c_float), ("win_long", 36, C.c_float), ("win_short", 12, C.c_float),
                       ("len15", 256, C.c_uint8), ("code15", 256, C.c_uint8)):
        getattr(L, "orc_table_" + name).restype = C.POINTER(t * n)
    L.orc_inv_step.restype = C.c_float; L.orc_inv_step.argtypes = [C.c_int]
    L.orc_pow34.restype = C.c_float; L.orc_pow34.argtypes = [C.c_float]
    L.orc_bitrate_index.restype = C.c_int; L.orc_bitrate_index.argtypes = [C.c_int, C.c_int]
    L.orc_bitrate_value.restype = C.c_int; L.orc_bitrate_value.argtypes = [C.c_int]
    L.orc_encode_streams.restype = C.c_size_t
    L.orc_encode_streams.argtypes = [C.POINTER(Options), C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.c_size_t,
                                     C.c_int, C.POINTER(C.c_uint64)]
    L.orc_compare_streams.restype = C.c_size_t
    L.orc_compare_streams.argtypes = [C.POINTER(Options), C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.c_size_t, C.c_size_t,
                                      C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.POINTER(C.c_int64)]
    L.orc_synth_fill.restype = None
    L.orc_synth_fill.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_uint64]
    return L


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_SO):
        build()
    _lib = _declare(C.CDLL(_SO))
    return _lib


def native_lib():
    """The oracle rebuilt -O3 -march=native ON THIS HOST (oracle/Makefile `native`): the timed CPU baseline of bench.py.
    Same source, same results (no fast-math, no contraction); the default build stays the checker."""
    global _native
    if _native is None:
        subprocess.check_call(["make", "-s", "-B", "-C", os.path.join(_ROOT, "oracle"), "native"])
        _native = _declare(C.CDLL(os.path.join(_ROOT, "oracle", "libmp3oracle_native.so")))
    return _native


MODES = {"mono": 0, "stereo": 1, "jointStereo": 2}


def make_options(sample_rate=44100, bitrate_kbps=128, vbr=False, mode="stereo", quality=5, crc_protected=False,
                 original=True, copyright=False):
    return Options(sample_rate, bitrate_kbps, int(vbr), MODES[mode] if isinstance(mode, str) else int(mode), quality,
                   int(crc_protected), int(original), int(copyright))


def table(name):
    return np.array(getattr(lib(), "orc_table_" + name)().contents)


class Session:
    """Mirror of the reference's EncoderSession (MP3Encoder.swift:237-350) over the oracle."""

    def __init__(self, trace=False, **opts):
        self.L = lib()
        self.opts = make_options(**opts)
        self.h = self.L.orc_create(C.byref(self.opts))
        if trace:
            self.L.orc_trace_enable(self.h, 1)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_destroy(self.h); self.h = None

    def clone(self):
        o = object.__new__(Session); o.L = self.L; o.opts = self.opts; o.h = self.L.orc_clone(self.h); return o

    def encode(self, samples):
        a = np.ascontiguousarray(samples, dtype=np.float32)
        n = C.c_size_t()
        p = self.L.orc_encode(self.h, a.ctypes.data_as(C.c_void_p), a.size, C.byref(n))
        return C.string_at(p, n.value) if n.value else b""

    def flush(self):
        n = C.c_size_t(); p = self.L.orc_flush(self.h, C.byref(n))
        return C.string_at(p, n.value) if n.value else b""

    def xing_header(self):
        n = C.c_size_t(); p = self.L.orc_xing_header(self.h, C.byref(n))
        return C.string_at(p, n.value)

    @property
    def frame_count(self):
        return self.L.orc_frame_count(self.h)

    @property
    def byte_count(self):
        return self.L.orc_byte_count(self.h)

    def gc_trace(self):
        n = self.L.orc_trace_gc_count(self.h)
        if not n:
            return np.zeros(0, GC_TRACE)
        buf = C.string_at(self.L.orc_trace_gc(self.h), n * GC_TRACE.itemsize)
        return np.frombuffer(buf, GC_TRACE).copy()

    def frame_trace(self):
        n = self.L.orc_trace_frame_count(self.h)
        if not n:
            return np.zeros(0, FRAME_TRACE)
        buf = C.string_at(self.L.orc_trace_frames(self.h), n * FRAME_TRACE.itemsize)
        return np.frombuffer(buf, FRAME_TRACE).copy()


def encode_all(pcm, trace=False, **opts):
    """encode(samples:) + flush() of one stream; returns (bytes, session)."""
    s = Session(trace=trace, **opts)
    out = s.encode(pcm) + s.flush()
    return out, s


def id3_build(title=None, artist=None, album=None, genre=None, comment=None, track=None, track_total=None, year=None,
              album_art=None, album_art_mime="image/jpeg"):
    e = lambda v: None if v is None else v.encode("utf-8")
    art = None if album_art is None else (C.c_uint8 * len(album_art)).from_buffer_copy(album_art)
    t = ID3(e(title), e(artist), e(album), e(genre), e(comment), -1 if track is None else track,
            -1 if track_total is None else track_total, -1 if year is None else year,
            C.cast(art, C.c_void_p) if art is not None else None, 0 if album_art is None else len(album_art),
            e(album_art_mime))
    n = C.c_size_t(); p = lib().orc_id3_build(C.byref(t), C.byref(n))
    if not p:
        return b""
    out = C.string_at(p, n.value); lib().orc_free(p)
    return out


def encode_streams(pcms, n_threads, native=False, **opts):
    """Multi-threaded CPU baseline: returns (total_bytes, digest).  native=True: the -O3 -march=native build."""
    o = make_options(**opts)
    arrs = [np.ascontiguousarray(p, dtype=np.float32) for p in pcms]
    ptrs = (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])
    lens = (C.c_size_t * len(arrs))(*[a.size for a in arrs])
    dig = C.c_uint64()
    L = native_lib() if native else lib()
    n = L.orc_encode_streams(C.byref(o), ptrs, lens, len(arrs), n_threads, C.byref(dig))
    return n, dig.value


def compare_streams_raw(pcm_ptrs, n_floats, expect_ptrs, expect_lens, n_threads, chunk_floats=0, **opts):
    """orc_compare_streams on raw addresses (ints): every stream is encoded by the oracle and compared with the bytes at
    expect_ptrs[i].  Returns the list of (stream, first differing byte offset) — empty when everything is identical."""
    o = make_options(**opts)
    n = len(pcm_ptrs)
    pp = (C.c_void_p * n)(*pcm_ptrs); nn = (C.c_size_t * n)(*n_floats)
    ep = (C.c_void_p * n)(*expect_ptrs); en = (C.c_size_t * n)(*expect_lens)
    fd = (C.c_int64 * n)()
    bad = lib().orc_compare_streams(C.byref(o), pp, nn, n, chunk_floats, n_threads, ep, en, fd)
    out = [(i, int(fd[i])) for i in range(n) if fd[i] >= 0]
    assert len(out) == bad
    return out


def compare_streams(pcms, outputs, n_threads=0, chunk_floats=0, **opts):
    """The same for numpy PCM arrays and bytes objects."""
    arrs = [np.ascontiguousarray(p, dtype=np.float32).reshape(-1) for p in pcms]
    bufs = [np.frombuffer(o, dtype=np.uint8) if len(o) else np.zeros(1, np.uint8) for o in outputs]
    return compare_streams_raw([a.ctypes.data for a in arrs], [a.size for a in arrs], [b.ctypes.data for b in bufs],
                               [len(o) for o in outputs], n_threads or (os.cpu_count() or 1), chunk_floats, **opts)


def synth_fill(n_per_channel, channels=2, sample_rate=44100, f_left=440.0, f_right=554.37, amp=0.5, noise=0.05, seed=1234):
    """CPU twin of mp3b_synth_fill: interleaved float32, bit-identical to the device generator."""
    out = np.empty(n_per_channel * channels, dtype=np.float32)
    lib().orc_synth_fill(out.ctypes.data, n_per_channel, channels, sample_rate, f_left, f_right, amp, noise, seed)
    return out
