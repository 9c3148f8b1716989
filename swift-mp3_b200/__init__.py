"""B200-native MP3 encode path behind the API of mierau/swift-mp3 (MP3Encoder / MP3EncoderOptions /
EncoderSession.encode(samples:) / flush()).

This Python layer is a thin ctypes mirror of the C ABI in include/mp3b200.h (libmp3b200.so, built by
swift-mp3_b200/csrc/Makefile).  All compute runs in the CUDA kernels of csrc/kernels.cu; there is no CPU fallback:
importing works without a GPU (so the symbols can be checked), creating an encoder session does not.

The package directory name contains a hyphen, so import it with
    importlib.import_module("swift-mp3_b200")
"""
from .binding import (  # noqa: F401
    ID3Tag, MP3Encoder, MP3EncoderOptions, EncoderSession, EncoderBatch, SessionPool, Mode, MP3BError, lib, library_path,
    build_library, table, device_count, GC_RECORD, FRAME_RECORD, STAGES,
)
