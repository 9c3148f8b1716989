"""ctypes mirror of include/mp3b200.h plus the reference-shaped host classes.

Reference interface mirrored (Sources/SwiftMP3/MP3Encoder.swift = SRC): ID3Tag SRC:8-54, MP3EncoderOptions SRC:57-116,
MP3Encoder SRC:132-230, EncoderSession SRC:237-350 (encode(samples:) 297, flush() 318, generateXingHeader() 367,
generateID3Tag() 355, encodedFrameCount / encodedByteCount 261-264).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libmp3b200.so")


ERR_BUFFER_TOO_SMALL = -4      # mp3b_status, include/mp3b200.h


class MP3BError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("mp3b200 error %d: %s" % (code, message))
        self.code = code


class _Options(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ("sample_rate", "bitrate_kbps", "vbr", "mode", "quality", "crc_protected", "original", "copyright")]


class _ID3(C.Structure):
    _fields_ = [("title", C.c_char_p), ("artist", C.c_char_p), ("album", C.c_char_p), ("genre", C.c_char_p),
                ("comment", C.c_char_p), ("track", C.c_int32), ("track_total", C.c_int32), ("year", C.c_int32),
                ("album_art", C.c_char_p), ("album_art_len", C.c_size_t), ("album_art_mime", C.c_char_p)]


GC_RECORD = np.dtype([("part23_length", "<i4"), ("big_values", "<i4"), ("global_gain", "<i4"), ("gain_used", "<i4"),
                      ("block_type", "<i4"), ("subblock_gain", "<i4", 3), ("region0", "<i4"), ("region1", "<i4"),
                      ("preflag", "<i4"), ("g0", "<i4"), ("max_bits", "<i4"), ("iterations", "<i4"), ("energy", "<f4"),
                      ("table_select", "<i4", 3), ("count1table_select", "<i4"), ("scalefac_compress", "<i4"), ("part2_length", "<i4")])
FRAME_RECORD = np.dtype([("bitrate_index", "<i4"), ("padding", "<i4"), ("frame_size", "<i4"), ("main_data_size", "<i4"),
                         ("main_data_begin", "<i4"), ("reservoir_bits", "<i4"), ("huff_bytes", "<i4"), ("ms", "<i4"),
                         ("is_final", "<i4"), ("frame_energy", "<f4")])
STAGES = ("h2d", "prepass", "filterbank", "granule", "scan", "pack", "frames", "d2h", "total")   # MP3B_STAGE_* order

_lib = None


def library_path():
    return _SO


def build_library():
    """Compile csrc/ into libmp3b200.so for sm_100a (nvcc cross-compiles without a GPU)."""
    subprocess.check_call(["make", "-s", "-C", os.path.join(_HERE, "csrc")])


def lib():
    """Load libmp3b200.so.  Raises (loudly) when the CUDA extension has not been built — there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_SO):
        raise ImportError("libmp3b200.so is missing: run `make -C swift-mp3_b200/csrc` (or __graft_entry__.build()); "
                          "this package has no CPU fallback")
    L = C.CDLL(_SO)
    vp, sz, i32, u8p = C.c_void_p, C.c_size_t, C.c_int, C.POINTER(C.c_uint8)
    szp = C.POINTER(C.c_size_t)
    sig = {
        "mp3b_version": (i32, []), "mp3b_last_error": (C.c_char_p, []),
        "mp3b_options_default": (None, [C.POINTER(_Options)]), "mp3b_device_count": (i32, [C.POINTER(i32)]),
        "mp3b_session_create": (i32, [C.POINTER(_Options), i32, C.POINTER(vp)]), "mp3b_session_destroy": (None, [vp]),
        "mp3b_session_encode": (i32, [vp, vp, sz, vp, sz, szp]), "mp3b_session_flush": (i32, [vp, vp, sz, szp]),
        "mp3b_session_take_output": (i32, [vp, vp, sz, szp]), "mp3b_session_output_bound": (sz, [vp, sz]),
        "mp3b_session_xing_header": (i32, [vp, vp, sz, szp]), "mp3b_xing_frame_size": (i32, [C.POINTER(_Options)]),
        "mp3b_session_frame_count": (C.c_uint32, [vp]), "mp3b_session_byte_count": (C.c_uint32, [vp]),
        "mp3b_id3_build": (i32, [C.POINTER(_ID3), vp, sz, szp]),
        "mp3b_batch_create": (i32, [C.POINTER(_Options), i32, i32, C.POINTER(vp)]),
        "mp3b_batch_create_ex": (i32, [C.POINTER(_Options), i32, i32, i32, C.POINTER(vp)]),
        "mp3b_batch_frames_per_pass": (i32, [vp]),
        "mp3b_batch_create_multi": (i32, [C.POINTER(_Options), i32, C.POINTER(i32), i32, i32, C.POINTER(vp)]),
        "mp3b_batch_set_iso_mode": (i32, [vp, i32]), "mp3b_batch_iso_mode": (i32, [vp]), "mp3b_session_set_iso_mode": (i32, [vp, i32]),
        "mp3b_batch_set_matrixing": (i32, [vp, i32]), "mp3b_batch_matrixing": (i32, [vp]),
        "mp3b_batch_device_count": (i32, [vp]), "mp3b_batch_stream_device": (i32, [vp, i32]),
        "mp3b_batch_destroy": (None, [vp]), "mp3b_batch_stream_count": (i32, [vp]),
        "mp3b_batch_encode": (i32, [vp, C.POINTER(vp), szp, i32, vp]),
        "mp3b_batch_encode_device": (i32, [vp, C.POINTER(vp), szp, i32, i32]),
        "mp3b_batch_encode_strided": (i32, [vp, vp, sz, szp, i32, vp]),
        "mp3b_batch_encode_i16": (i32, [vp, C.POINTER(vp), szp, i32, vp]),
        "mp3b_batch_output": (i32, [vp, i32, C.POINTER(vp), szp]),
        "mp3b_batch_output_device": (i32, [vp, i32, C.POINTER(vp), szp]),
        "mp3b_batch_output_total": (sz, [vp]),
        "mp3b_batch_xing_header": (i32, [vp, i32, vp, sz, szp]),
        "mp3b_batch_frame_count": (C.c_uint32, [vp, i32]), "mp3b_batch_byte_count": (C.c_uint32, [vp, i32]),
        "mp3b_host_alloc": (i32, [sz, C.POINTER(vp)]), "mp3b_host_free": (None, [vp]),
        "mp3b_device_alloc": (i32, [i32, sz, C.POINTER(vp)]), "mp3b_device_free": (None, [i32, vp]),
        "mp3b_device_copy": (i32, [i32, vp, vp, sz, i32]), "mp3b_device_sync": (i32, [i32]),
        "mp3b_batch_stage_ms": (i32, [vp, C.POINTER(C.c_float), i32]), "mp3b_batch_launch_count": (i32, [vp]),
        "mp3b_batch_pass_count": (i32, [vp]), "mp3b_batch_stream": (vp, [vp]), "mp3b_batch_reset": (i32, [vp]),
        "mp3b_batch_set_trace": (i32, [vp, i32]), "mp3b_batch_trace_frames": (i32, [vp, i32]),
        "mp3b_batch_trace_frame_records": (i32, [vp, i32, vp, i32]),
        "mp3b_batch_trace_gc_records": (i32, [vp, i32, vp, i32]),
        "mp3b_batch_trace_gc_array": (i32, [vp, i32, i32, vp, i32]),
        "mp3b_table": (i32, [i32, vp, sz]),
        "mp3b_synth_fill": (i32, [i32, vp, sz, i32, i32, C.c_float, C.c_float, C.c_float, C.c_float, C.c_uint64]),
        "mp3b_selftest": (i32, [i32, C.POINTER(C.c_uint64)]),
        "mp3b_batch_reset_stream": (i32, [vp, i32]),
        "mp3b_session_clone": (i32, [vp, C.POINTER(vp)]), "mp3b_batch_clone": (i32, [vp, C.POINTER(vp)]),
        "mp3b_pool_create": (i32, [C.POINTER(_Options), i32, i32, i32, C.POINTER(vp)]), "mp3b_pool_destroy": (None, [vp]),
        "mp3b_pool_open": (i32, [vp, C.POINTER(i32)]),
        "mp3b_pool_encode": (i32, [vp, i32, vp, sz, vp, sz, C.POINTER(sz)]),
        "mp3b_pool_flush": (i32, [vp, i32, vp, sz, C.POINTER(sz)]),
        "mp3b_pool_take_output": (i32, [vp, i32, vp, sz, C.POINTER(sz)]),
        "mp3b_pool_stats": (i32, [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
        "mp3b_pool_last_error": (C.c_char_p, []),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args
    _lib = L
    return L


EXPORTS = None  # filled lazily by tests from include/mp3b200.h


def _check(rc):
    if rc < 0:
        raise MP3BError(rc, lib().mp3b_last_error().decode("utf-8", "replace"))
    return rc


def device_count():
    n = C.c_int(0)
    _check(lib().mp3b_device_count(C.byref(n)))
    return n.value


_TABLES = {"window": (0, "<f4"), "analysis": (1, "<f4"), "mdct_long": (2, "<f4"), "mdct_short": (3, "<f4"),
           "win_long": (4, "<f4"), "win_short": (5, "<f4"), "inv_step": (6, "<f4"), "len15": (7, "u1"),
           "code15": (8, "u1"), "gain_thr": (9, "<f8"), "alias_cs": (10, "<f4"), "alias_ca": (11, "<f4"),
           "sfb_cum": (12, "<i4"), "len31s": (13, "u1"), "tab31": (14, "<u2")}


def table(name):
    """Product constant table as a numpy array (host copy of what the kernels use)."""
    which, dt = _TABLES[name]
    n = _check(lib().mp3b_table(which, None, 1 << 20))
    out = np.zeros(n, dtype=dt)
    _check(lib().mp3b_table(which, out.ctypes.data, out.nbytes))
    return out


class Mode:
    """MP3EncoderOptions.Mode, SRC:59-63."""
    mono, stereo, jointStereo = 0, 1, 2


class ID3Tag:
    """ID3Tag, SRC:8-54."""

    def __init__(self, title=None, artist=None, album=None, genre=None, year=None, track=None, trackTotal=None,
                 comment=None, albumArt=None, albumArtMIME="image/jpeg"):
        self.title, self.artist, self.album, self.genre, self.year = title, artist, album, genre, year
        self.track, self.trackTotal, self.comment, self.albumArt, self.albumArtMIME = track, trackTotal, comment, albumArt, albumArtMIME

    def build(self):
        """ID3TagWriter.build, SRC:1040-1075."""
        enc = lambda v: None if v is None else v.encode("utf-8")
        t = _ID3(enc(self.title), enc(self.artist), enc(self.album), enc(self.genre), enc(self.comment),
                 -1 if self.track is None else self.track, -1 if self.trackTotal is None else self.trackTotal,
                 -1 if self.year is None else self.year, self.albumArt, len(self.albumArt) if self.albumArt else 0,
                 enc(self.albumArtMIME))
        n = C.c_size_t(0)
        rc = lib().mp3b_id3_build(C.byref(t), None, 0, C.byref(n))
        if rc == 0 and n.value == 0:
            return b""
        buf = C.create_string_buffer(n.value)
        _check(lib().mp3b_id3_build(C.byref(t), buf, n.value, C.byref(n)))
        return buf.raw[:n.value]


class MP3EncoderOptions:
    """MP3EncoderOptions, SRC:57-116 (same names, same defaults, quality clamped to 0...9)."""

    def __init__(self, sampleRate=44100, bitrateKbps=128, vbr=False, mode=Mode.stereo, quality=5, crcProtected=False,
                 original=True, copyright=False, id3Tag=None):
        self.sampleRate, self.bitrateKbps, self.vbr, self.mode = sampleRate, bitrateKbps, vbr, mode
        self.quality = min(max(quality, 0), 9)
        self.crcProtected, self.original, self.copyright, self.id3Tag = crcProtected, original, copyright, id3Tag

    def _c(self):
        return _Options(self.sampleRate, self.bitrateKbps, int(self.vbr), int(self.mode), self.quality,
                        int(self.crcProtected), int(self.original), int(self.copyright))

    @property
    def channels(self):
        return 1 if self.mode == Mode.mono else 2


def _as_f32(samples):
    a = np.ascontiguousarray(samples, dtype=np.float32)
    return a.reshape(-1)


class EncoderSession:
    """EncoderSession, SRC:237-350: one stream on one device (a batch of one)."""

    def __init__(self, options, device=0):
        self.options = options
        self._h = C.c_void_p()
        o = options._c()
        _check(lib().mp3b_session_create(C.byref(o), device, C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            lib().mp3b_session_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def _call(self, fn, *head):
        cap = int(lib().mp3b_session_output_bound(self._h, head[1] if head else 0))
        buf = (C.c_uint8 * max(cap, 1))()
        n = C.c_size_t(0)
        _check(fn(self._h, *head, buf, cap, C.byref(n)))
        return bytes(memoryview(buf)[:n.value])

    def set_iso_mode(self, on=True):
        _check(lib().mp3b_session_set_iso_mode(self._h, int(on)))

    def clone(self):
        """`var copy = session`: an independent snapshot (the reference's EncoderSession is a value type)."""
        c = object.__new__(EncoderSession)
        c.options = self.options
        c._h = C.c_void_p()
        _check(lib().mp3b_session_clone(self._h, C.byref(c._h)))
        return c

    def encode(self, samples):
        """encode(samples:), SRC:297-310: interleaved float32 in [-1, 1]; returns 0...k whole frames."""
        a = _as_f32(samples)
        return self._call(lib().mp3b_session_encode, a.ctypes.data, a.size)

    def flush(self):
        """flush(), SRC:318-350."""
        return self._call(lib().mp3b_session_flush)

    def generateXingHeader(self):
        buf = (C.c_uint8 * 2048)()
        n = C.c_size_t(0)
        _check(lib().mp3b_session_xing_header(self._h, buf, 2048, C.byref(n)))
        return bytes(memoryview(buf)[:n.value])

    def generateID3Tag(self):
        return self.options.id3Tag.build() if self.options.id3Tag is not None else b""

    @property
    def encodedFrameCount(self):
        return int(lib().mp3b_session_frame_count(self._h))

    @property
    def encodedByteCount(self):
        return int(lib().mp3b_session_byte_count(self._h))


class EncoderBatch:
    """N independent EncoderSessions with the same options advancing together (the batch plane): on one device, or — devices =
    a list of CUDA ordinals — partitioned by stream over several (mp3b_batch_create_multi; one host thread per device)."""

    def __init__(self, options, n_streams, device=0, frames_per_pass=0, devices=None):
        self.options, self.n_streams, self.device = options, n_streams, device
        self._h = C.c_void_p()
        o = options._c()
        if devices is not None:
            self.device = devices[0]
            arr = (C.c_int * len(devices))(*devices)
            _check(lib().mp3b_batch_create_multi(C.byref(o), n_streams, arr, len(devices), frames_per_pass, C.byref(self._h)))
        else:
            _check(lib().mp3b_batch_create_ex(C.byref(o), n_streams, device, frames_per_pass, C.byref(self._h)))

    @property
    def device_count(self):
        return lib().mp3b_batch_device_count(self._h)

    def set_iso_mode(self, on=True):
        """Opt-in ISO mode (include/mp3b200.h): ISO quantizer, table selection, count1, real main_data_begin."""
        _check(lib().mp3b_batch_set_iso_mode(self._h, int(on)))

    @property
    def iso_mode(self):
        return lib().mp3b_batch_iso_mode(self._h)

    @property
    def matrixing(self):
        return lib().mp3b_batch_matrixing(self._h)

    def set_matrixing(self, mode):
        """0 = FP32 FMA (default, bit-exact with the oracle), 1 = 3xTF32 on the tensor cores (include/mp3b200.h)."""
        _check(lib().mp3b_batch_set_matrixing(self._h, int(mode)))

    def stream_device(self, stream):
        return _check(lib().mp3b_batch_stream_device(self._h, stream))

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            lib().mp3b_batch_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    @property
    def frames_per_pass(self):
        return lib().mp3b_batch_frames_per_pass(self._h)

    def encode(self, chunks, flush=False, flush_mask=None, fetch=True):
        """One encode(samples:) per stream (chunks[i] may be None / empty), optionally followed by flush()."""
        arrs = [None if c is None else _as_f32(c) for c in chunks]
        ptrs = (C.c_void_p * self.n_streams)(*[None if a is None or a.size == 0 else a.ctypes.data for a in arrs])
        ns = (C.c_size_t * self.n_streams)(*[0 if a is None else a.size for a in arrs])
        mask = None
        if flush_mask is not None:
            mask = np.ascontiguousarray(flush_mask, dtype=np.uint8).ctypes.data
        _check(lib().mp3b_batch_encode(self._h, ptrs, ns, int(flush), mask))
        return self.outputs() if fetch else None

    def encode_ptrs(self, host_ptrs, n_floats, flush=False):
        """Raw variant: host_ptrs / n_floats are ctypes arrays (c_void_p / c_size_t) prepared by the caller."""
        _check(lib().mp3b_batch_encode(self._h, host_ptrs, n_floats, int(flush), None))

    def encode_strided(self, base_ptr, pitch_floats, n_floats, flush=False, flush_mask=None):
        """mp3b_batch_encode_strided: stream i = base_ptr + i * pitch_floats * 4 (one host arena, one strided upload);
        n_floats is a ctypes c_size_t array."""
        mask = None if flush_mask is None else np.ascontiguousarray(flush_mask, dtype=np.uint8).ctypes.data
        _check(lib().mp3b_batch_encode_strided(self._h, base_ptr, pitch_floats, n_floats, int(flush), mask))

    def encode_i16(self, chunks, flush=False):
        """Extension: one interleaved int16 array per stream; equals encode([c.astype(float32) / 32768 for c in chunks])."""
        arrs = [np.ascontiguousarray(c, dtype=np.int16).reshape(-1) for c in chunks]
        ptrs = (C.c_void_p * self.n_streams)(*[a.ctypes.data if a.size else None for a in arrs])
        ns = (C.c_size_t * self.n_streams)(*[a.size for a in arrs])
        _check(lib().mp3b_batch_encode_i16(self._h, ptrs, ns, int(flush), None))
        return self.outputs()

    def encode_i16_ptrs(self, host_ptrs, n_samples, flush=False):
        _check(lib().mp3b_batch_encode_i16(self._h, host_ptrs, n_samples, int(flush), None))

    def encode_device(self, dev_ptrs, n_floats, flush=False, download=False):
        _check(lib().mp3b_batch_encode_device(self._h, dev_ptrs, n_floats, int(flush), int(download)))

    def output(self, stream):
        p, n = C.c_void_p(), C.c_size_t(0)
        _check(lib().mp3b_batch_output(self._h, stream, C.byref(p), C.byref(n)))
        return C.string_at(p.value, n.value) if n.value else b""

    def outputs(self):
        return [self.output(i) for i in range(self.n_streams)]

    def output_device(self, stream):
        p, n = C.c_void_p(), C.c_size_t(0)
        _check(lib().mp3b_batch_output_device(self._h, stream, C.byref(p), C.byref(n)))
        return p.value, n.value

    @property
    def output_total(self):
        return int(lib().mp3b_batch_output_total(self._h))

    def frame_count(self, stream):
        return int(lib().mp3b_batch_frame_count(self._h, stream))

    def byte_count(self, stream):
        return int(lib().mp3b_batch_byte_count(self._h, stream))

    def xing_header(self, stream):
        buf = (C.c_uint8 * 2048)()
        n = C.c_size_t(0)
        _check(lib().mp3b_batch_xing_header(self._h, stream, buf, 2048, C.byref(n)))
        return bytes(memoryview(buf)[:n.value])

    def stage_ms(self):
        ms = (C.c_float * len(STAGES))()
        _check(lib().mp3b_batch_stage_ms(self._h, ms, len(STAGES)))
        return dict(zip(STAGES, [float(v) for v in ms]))

    @property
    def launch_count(self):
        return lib().mp3b_batch_launch_count(self._h)

    @property
    def pass_count(self):
        return lib().mp3b_batch_pass_count(self._h)

    @property
    def cuda_stream(self):
        """cudaStream_t (as an int) all work of this batch is issued on."""
        return lib().mp3b_batch_stream(self._h)

    def reset(self):
        _check(lib().mp3b_batch_reset(self._h))

    def clone(self):
        """Snapshot of all sessions of the batch (mp3b_batch_clone)."""
        c = object.__new__(EncoderBatch)
        c.options, c.n_streams, c.device = self.options, self.n_streams, self.device
        c._h = C.c_void_p()
        _check(lib().mp3b_batch_clone(self._h, C.byref(c._h)))
        return c

    # ---- traces (tests) ----
    def set_trace(self, spectrum=False, ix=False, thresholds=False, records=True):
        flags = (1 if spectrum else 0) | (2 if ix else 0) | (4 if thresholds else 0) | (8 if records else 0)
        _check(lib().mp3b_batch_set_trace(self._h, flags))

    def trace_frames(self, stream):
        n = lib().mp3b_batch_trace_frames(self._h, stream)
        out = np.zeros(n, dtype=FRAME_RECORD)
        _check(lib().mp3b_batch_trace_frame_records(self._h, stream, out.ctypes.data, n))
        return out

    def trace_gc(self, stream):
        n = lib().mp3b_batch_trace_frames(self._h, stream) * 2 * self.options.channels
        out = np.zeros(n, dtype=GC_RECORD)
        _check(lib().mp3b_batch_trace_gc_records(self._h, stream, out.ctypes.data, n))
        return out

    def trace_array(self, stream, kind):
        k = {"spectrum": 0, "ix": 1, "thresholds": 2, "psy": 3, "scalefactors": 4}[kind]
        n = lib().mp3b_batch_trace_frames(self._h, stream) * 2 * self.options.channels
        out = np.zeros((n, 576 if k < 3 else 24), dtype="<i4" if k in (1, 4) else "<f4")
        _check(lib().mp3b_batch_trace_gc_array(self._h, stream, k, out.ctypes.data, n))
        return out


class SessionPool:
    """mp3b_pool: n_sessions EncoderSessions that may be driven from concurrent threads; the calls that arrive together
    run as one step on the GPU (BASELINE config 5)."""

    class Session:
        def __init__(self, pool, slot, bound):
            self._pool, self._slot, self._bound = pool, slot, bound

        def _call(self, fn, *head):
            cap = self._bound(head[1] if head else 0)
            buf = (C.c_uint8 * cap)()
            n = C.c_size_t(0)
            rc = fn(self._pool._h, self._slot, *head, buf, cap, C.byref(n))
            if rc == ERR_BUFFER_TOO_SMALL:                       # the bytes are kept by the pool: fetch them with a buffer that fits
                cap = n.value
                buf = (C.c_uint8 * cap)()
                rc = lib().mp3b_pool_take_output(self._pool._h, self._slot, buf, cap, C.byref(n))
            if rc != 0:
                raise MP3BError(rc, (lib().mp3b_pool_last_error() or b"").decode())
            return bytes(memoryview(buf)[:n.value])

        def encode(self, samples):
            a = _as_f32(samples)
            return self._call(lib().mp3b_pool_encode, a.ctypes.data, a.size)

        def flush(self):
            return self._call(lib().mp3b_pool_flush)

    def __init__(self, options, n_sessions, device=0, max_wait_us=200):
        self.options = options
        self._h = C.c_void_p()
        o = options._c()
        _check(lib().mp3b_pool_create(C.byref(o), n_sessions, device, max_wait_us, C.byref(self._h)))
        self._frame_bytes = 1441 + 8

    def newSession(self):
        slot = C.c_int(-1)
        _check(lib().mp3b_pool_open(self._h, C.byref(slot)))
        fsc = 1152 * self.options.channels
        return SessionPool.Session(self, slot.value, lambda n: (n // fsc + 3) * self._frame_bytes)

    def stats(self):
        a, b = C.c_uint64(0), C.c_uint64(0)
        _check(lib().mp3b_pool_stats(self._h, C.byref(a), C.byref(b)))
        return {"steps": a.value, "requests": b.value}

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            lib().mp3b_pool_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close


class MP3Encoder:
    """MP3Encoder, SRC:132-230: immutable holder of the options; sessions do the work."""

    def __init__(self, options=None):
        self.options = options if options is not None else MP3EncoderOptions()

    def newSession(self, device=0):
        """newSession(), SRC:143-145."""
        return EncoderSession(self.options, device)

    def newBatch(self, n_streams, device=0, frames_per_pass=0, devices=None):
        return EncoderBatch(self.options, n_streams, device, frames_per_pass, devices)

    def encode(self, chunks, device=0):
        """Counterpart of encode(_:) -> AsyncThrowingStream (SRC:151-179): yields every non-empty chunk of frames."""
        s = self.newSession(device)
        try:
            for c in chunks:
                out = s.encode(c)
                if out:
                    yield out
            out = s.flush()
            if out:
                yield out
        finally:
            s.close()

    def encode_to(self, chunks, path, device=0):
        """Counterpart of encode(_:to:) (SRC:189-230): ID3 tag, Xing placeholder, frames, then the real Xing frame."""
        s = self.newSession(device)
        try:
            id3 = s.generateID3Tag()
            o = self.options._c()
            with open(path, "wb") as f:
                f.write(id3)
                f.write(bytes(_check(lib().mp3b_xing_frame_size(C.byref(o)))))   # SRC:198-205: from the snapped bitrate
                for c in chunks:
                    f.write(s.encode(c))
                f.write(s.flush())
                f.seek(len(id3))
                f.write(s.generateXingHeader())                                    # SRC:227-229
        finally:
            s.close()
