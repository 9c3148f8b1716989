"""CPU tests of the drop-in boundary: libmp3b200.so loads, exports every symbol include/mp3b200.h declares, the
host-only entry points work without a GPU, product tables equal the oracle's, and there is no CPU fallback."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mp3b200.h")


def declared_symbols():
    text = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    return sorted(set(re.findall(r"\b(mp3b_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(mp3):
    L = mp3.lib()
    names = declared_symbols()
    assert len(names) >= 40
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    out = subprocess.run(["nm", "-D", "--defined-only", mp3.library_path()], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (mp3b_\w+)", out))
    assert set(names) <= exported
    assert not [n for n in exported if n not in names], "exported but undeclared"


def test_header_cites_reference_lines():
    text = open(HEADER).read()
    assert text.count("SRC:") >= 12 and 'extern "C"' in text and "torch" not in text.replace("No torch", "")


def test_version_and_defaults(mp3):
    L = mp3.lib()
    assert L.mp3b_version() >= 100
    from importlib import import_module
    b = import_module("swift-mp3_b200.binding")
    o = b._Options()
    L.mp3b_options_default(C.byref(o))
    assert (o.sample_rate, o.bitrate_kbps, o.vbr, o.mode, o.quality, o.crc_protected, o.original, o.copyright) == \
        (44100, 128, 0, 1, 5, 0, 1, 0)                         # TST:69-76 optionsDefaults
    d = mp3.MP3EncoderOptions()
    assert (d.sampleRate, d.bitrateKbps, d.vbr, d.mode, d.quality) == (44100, 128, False, mp3.Mode.stereo, 5)
    assert mp3.MP3EncoderOptions(quality=99).quality == 9 and mp3.MP3EncoderOptions(quality=-3).quality == 0   # SRC:110


def test_product_tables_equal_oracle(mp3, orc):
    assert np.array_equal(mp3.table("window"), orc.table("window"))
    assert np.array_equal(mp3.table("analysis"), orc.table("analysis").reshape(-1))
    assert np.array_equal(mp3.table("mdct_long"), orc.table("mdct_long").reshape(-1))
    assert np.array_equal(mp3.table("mdct_short"), orc.table("mdct_short").reshape(-1))
    assert np.array_equal(mp3.table("win_long"), orc.table("win_long"))
    assert np.array_equal(mp3.table("win_short"), orc.table("win_short"))
    assert np.array_equal(mp3.table("len15"), orc.table("len15")) and np.array_equal(mp3.table("code15"), orc.table("code15"))
    # the quantizer-indexed views the kernels use: u = min(floor(2 t), 30) per value, q = (u + 1) >> 1 (kernels.cu quant30)
    len15, code15 = orc.table("len15").astype(int), orc.table("code15").astype(int)
    len31, tab31 = mp3.table("len31s").reshape(31, 32), mp3.table("tab31").reshape(31, 32)
    for ux in range(31):
        for uy in range(31):
            qx, qy = (ux + 1) >> 1, (uy + 1) >> 1
            assert len31[ux, uy] == len15[qx * 16 + qy] + (qx != 0) + (qy != 0)
            assert tab31[ux, uy] == code15[qx * 16 + qy] | len15[qx * 16 + qy] << 8
    inv = np.array([orc.lib().orc_inv_step(g) for g in range(256)], np.float32)
    assert np.array_equal(mp3.table("inv_step"), inv)
    thr = mp3.table("gain_thr")
    assert thr[210] == 1.0 and np.all(np.diff(thr) > 0)
    # the threshold table must reproduce 210 + Int(4*log2(r)) (SRC:1004) on a dense sweep of float32 ratios
    r = np.exp(np.random.default_rng(1).uniform(np.log(1e-12), np.log(50.0), 200000)).astype(np.float32).astype(np.float64)
    want = np.clip(210 + np.trunc(4.0 * np.log2(r)).astype(np.int64), 0, 255)
    idx = np.searchsorted(thr, r, side="right") - 1          # largest i with thr[i] <= r
    got = np.where(idx < 0, 0, np.where((r < 1.0) & (thr[np.clip(idx, 0, 255)] != r), idx + 1, idx))
    assert np.array_equal(np.clip(got, 0, 255), want)


def test_id3_matches_oracle(mp3, orc):                        # ID3TagWriter SRC:1040-1075, TST:189-302
    art = bytes(range(64))
    t = mp3.ID3Tag(title="Title", artist="Artist", album="Album", genre="Genre", year=2024, track=3, trackTotal=12,
                   comment="hello", albumArt=art, albumArtMIME="image/png")
    ref = orc.id3_build(title="Title", artist="Artist", album="Album", genre="Genre", year=2024, track=3, track_total=12,
                        comment="hello", album_art=art, album_art_mime="image/png")
    assert t.build() == ref and ref[:5] == b"ID3\x03\x00"
    assert mp3.ID3Tag().build() == b"" == orc.id3_build()     # id3EmptyFields TST:290-302
    assert mp3.ID3Tag(track=7).build() == orc.id3_build(track=7)


def test_bad_arguments(mp3):
    L = mp3.lib()
    h = C.c_void_p()
    assert L.mp3b_batch_create(None, 1, 0, C.byref(h)) == -1
    assert b"null" in L.mp3b_last_error()
    n = C.c_size_t()
    assert L.mp3b_session_encode(None, None, 0, None, 0, C.byref(n)) == -1
    assert L.mp3b_batch_stream_count(None) == 0 and L.mp3b_batch_frame_count(None, 0) == 0


def test_no_cpu_fallback(mp3):
    """Without a usable sm_100 device, session creation must fail loudly (never encode on the CPU)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(mp3.MP3BError) as e:
        mp3.MP3Encoder().newSession()
    assert e.value.code == -2


def test_product_does_not_import_oracle():
    for root, _, files in os.walk(os.path.join(ROOT, "swift-mp3_b200")):
        for f in files:
            if f.endswith((".py", ".cc", ".cu", ".h")) and f != "tables_gen.h":
                text = open(os.path.join(root, f)).read()
                assert "oracle_binding" not in text and "libmp3oracle" not in text and "dlopen" not in text, f
                assert not re.search(r'#include\s+"[^"]*oracle', text) and not re.search(r"^\s*(import|from)\s+\S*oracle", text, re.M), f


def test_generated_tables_are_current(tmp_path):
    """swift-mp3_b200/csrc/tables_gen.h is what tools/gen_device_tables.py writes today (no hand edits, no drift)."""
    import subprocess, sys
    out = tmp_path / "tables_gen.h"
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_device_tables.py"), str(out)], check=True, capture_output=True, timeout=120)
    assert out.read_text() == open(os.path.join(ROOT, "swift-mp3_b200", "csrc", "tables_gen.h")).read()


def test_xing_frame_size_uses_the_snapped_bitrate(mp3):
    """SRC:198-200: bitrateValue(bitrateIndex(125, 44100)) = 128 -> 144 * 128000 / 44100 = 417; 64 kbps at 22.05 kHz goes through the
    MPEG-2 index table but the MPEG-1 value table: index 8 -> 112 kbps -> 731.  Host-only entry point."""
    from importlib import import_module
    b = import_module("swift-mp3_b200.binding")
    L = mp3.lib()
    for kbps, sr, want in ((125, 44100, 417), (128, 44100, 417), (130, 44100, 417), (320, 48000, 960), (32, 48000, 96), (64, 22050, 731)):
        o = b._Options(sr, kbps, 0, 1, 5, 0, 1, 0)
        assert L.mp3b_xing_frame_size(C.byref(o)) == want, (kbps, sr)
    assert L.mp3b_xing_frame_size(None) < 0


def test_stream_limit_is_rejected_at_creation(mp3):
    """More streams than a grid dimension holds fail cleanly at creation (bad argument), not at the first launch."""
    from importlib import import_module
    b = import_module("swift-mp3_b200.binding")
    L = mp3.lib()
    o = b._Options(44100, 128, 0, 1, 5, 0, 1, 0)
    h = C.c_void_p()
    assert L.mp3b_batch_create(C.byref(o), 65536, 0, C.byref(h)) == -1 and b"65535" in L.mp3b_last_error()


def test_synth_twin_is_pinned(orc):
    """The CPU generator of the bench inputs (twin of the device kernel): statistics of the C1 recipe and a pinned digest."""
    import hashlib
    x = orc.synth_fill(44100, 2, 44100, 440.0, 554.37, 0.5, 0.05, 1234)
    t = np.arange(44100) / 44100.0
    nl, nr = x[0::2] - 0.5 * np.sin(2 * np.pi * 440.0 * t), x[1::2] - 0.5 * np.sin(2 * np.pi * 554.37 * t + 0.3)
    assert abs(nl.std() - 0.05) < 1e-3 and abs(nr.std() - 0.05) < 1e-3 and abs(np.corrcoef(nl, nr)[0, 1]) < 0.02
    assert abs(nl.mean()) < 1e-3 and np.abs(x).max() <= 1.0
    assert hashlib.sha256(x.tobytes()).hexdigest() == open(os.path.join(ROOT, "tests", "golden", "synth_c1_sha256.txt")).read().strip()
