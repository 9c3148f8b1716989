"""CPU tests: the oracle (oracle/mp3_oracle.c) against the reference's own structural known-answer tests
(Tests/SwiftMP3Tests/SwiftMP3Tests.swift = TST; every test cites the TST line it restates), the literals of the
reference source where it is available, and the committed golden fixtures."""
import hashlib
import json
import os
import re

import numpy as np
import pytest

import mp3parse
import signals

REF = "/root/reference/Sources/SwiftMP3/MP3Encoder.swift"
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "oracle_golden.json")


def frames_of(data):
    return mp3parse.parse_frames(data)


def test_encode_silence_sync_word(orc):                       # TST:7-23
    s = orc.Session()
    out = s.encode(np.zeros(1152 * 2 * 4, np.float32)) + s.flush()
    assert len(out) > 0 and out[0] == 0xFF and out[1] & 0xE0 == 0xE0


def test_encode_mono(orc):                                    # TST:25-36
    s = orc.Session(mode="mono")
    out = s.encode(np.zeros(1152 * 3, np.float32)) + s.flush()
    assert out[0] == 0xFF and (out[3] >> 6) & 3 == 3


def test_one_frame_delay_and_flush(orc):                      # TST:38-50, 408-455
    s = orc.Session()
    frame = signals.sine440(1)
    assert s.encode(frame) == b""                             # first full frame returns empty Data
    second = s.encode(frame)
    assert len(frames_of(second)) == 1
    s2 = orc.Session()
    s2.encode(frame)
    out = s2.flush()                                          # flush emits the buffered frame with empty PCM
    assert len(frames_of(out)) == 1
    assert s2.flush() == b""                                  # double flush returns empty
    s3 = orc.Session()
    out = s3.encode(frame[:1000]) + s3.flush()                # partial frame is zero padded
    assert len(frames_of(out)) == 1


def test_counters_and_frame_sizes(orc):                       # TST:457-477
    s = orc.Session()
    out = s.encode(signals.sine440(10)) + s.flush()
    assert s.frame_count == 10 and s.byte_count == len(out)
    assert 417 <= len(out) / 10 <= 418
    assert {f["size"] for f in frames_of(out)} <= {417, 418}


def test_padding_ratio_matches_theory(orc):                   # TST:363-406, 801-845
    s = orc.Session()
    out = s.encode(np.zeros(1152 * 2 * 1000, np.float32)) + s.flush()
    fr = frames_of(out)
    assert len(fr) == 1000
    ratio = sum(f["padding"] for f in fr) / 1000.0
    assert 0.93 < ratio < 0.98 and abs(ratio - 42300 / 44100) < 0.002
    assert {f["size"] for f in fr} == {417, 418}


def test_contiguous_frames(orc):                              # TST:560-624
    s = orc.Session()
    out = s.encode(signals.sine440(20)) + s.flush()
    assert len(frames_of(out)) == 20                          # parse_frames asserts zero trailing bytes


def test_bit_reservoir_main_data_begin(orc):                  # TST:304-361, 479-529
    s = orc.Session()
    out = s.encode(signals.sine440(12, amp=0.1)) + s.flush()
    fr = frames_of(out)
    assert any(f["mdb"] > 0 for f in fr[1:])
    assert fr[0]["mdb"] == 0
    s = orc.Session()
    out = s.encode(signals.sine440(5, amp=0.1)[:-700]) + s.flush()
    assert frames_of(out)[-1]["mdb"] == 0                     # final (flush) frame has main_data_begin == 0


def test_mono_reservoir(orc):                                 # TST:531-558
    s = orc.Session(mode="mono")
    out = s.encode(signals.sine440(8, channels=1, amp=0.1)) + s.flush()
    fr = frames_of(out)
    assert len(fr) == 8 and all(f["mode"] == 3 for f in fr)


def test_deterministic(orc):                                  # TST:775-799
    pcm = signals.sine_noise(0.5, seed=3)
    assert orc.encode_all(pcm)[0] == orc.encode_all(pcm)[0]


def test_xing_header(orc):                                    # TST:52-67, 129-169
    s = orc.Session()
    s.encode(signals.sine440(4)); s.flush()
    x = s.xing_header()
    assert x[0] == 0xFF and x[1] & 0xE0 == 0xE0 and x[36:40] == b"Info" and len(x) == 417
    v = orc.Session(vbr=True); v.encode(signals.sine440(4)); v.flush()
    assert v.xing_header()[36:40] == b"Xing"


def test_id3(orc):                                            # TST:189-302
    t = orc.id3_build(title="T", artist="A", album="B", album_art=b"\x89PNG", album_art_mime="image/png")
    assert t[:5] == b"ID3\x03\x00" and b"TIT2" in t and b"TPE1" in t and b"TALB" in t and b"APIC" in t
    size = (t[6] << 21) | (t[7] << 14) | (t[8] << 7) | t[9]
    assert size == len(t) - 10
    assert orc.id3_build() == b""


def test_silence_known_answer(orc):                           # gain 170, nothing coded
    out, s = orc.encode_all(np.zeros(1152 * 2 * 6, np.float32), trace=True)
    g = s.gc_trace()
    assert (g["gain_out"] == 170).all() and (g["big_values"] == 0).all() and (g["bits"] == 0).all()
    assert all(f["mdb"] == 0 for f in frames_of(out))


@pytest.mark.parametrize("cfg", [dict(sample_rate=44100, bitrate_kbps=128), dict(sample_rate=44100, bitrate_kbps=128, mode="mono"),
                                 dict(sample_rate=48000, bitrate_kbps=192), dict(sample_rate=32000, bitrate_kbps=64),
                                 dict(sample_rate=44100, bitrate_kbps=128, mode="jointStereo")])
def test_bitstream_round_trip(orc, cfg):
    """The configurations of TST:727-755.  Instead of AVFoundation: the bytes alone must decode (frame walk, FIFO
    replay, table-15 Huffman) to exactly the ix the quantizer produced, and side info must equal the trace."""
    ch = 1 if cfg.get("mode") == "mono" else 2
    pcm = signals.sine_noise(0.4, sr=cfg["sample_rate"], channels=ch, seed=21)
    out, s = orc.encode_all(pcm, trace=True, **cfg)
    frames, ix = mp3parse.decode_stream(out, orc.table("len15"), orc.table("code15"))
    g = s.gc_trace()
    assert np.array_equal(ix, g["ix"])
    flat = [q for f in frames for q in f["gc"]]
    assert [q["global_gain"] for q in flat] == list(g["gain_out"]) and [q["big_values"] for q in flat] == list(g["big_values"])
    assert all(q["table_select"][0] == 15 and q["scalefac_compress"] == 0 for q in flat)


def test_pow34_matches_libm(orc):
    """[OD3]: (float)(sqrt(d) * sqrt(sqrt(d))) equals the float64 pow rounded to float32."""
    rng = np.random.default_rng(0)
    a = np.abs(rng.standard_normal(200000)).astype(np.float32) * np.float32(0.3) + np.float32(1e-10)
    got = np.array([orc.lib().orc_pow34(float(v)) for v in a[:20000]], np.float32)
    want = (a[:20000].astype(np.float64) ** 0.75).astype(np.float32)
    assert np.array_equal(got, want)


@pytest.mark.skipif(not os.path.exists(REF), reason="reference source not on this box")
def test_tables_equal_reference_literals(orc):
    L = open(REF).read().split("\n")
    win = [np.float32(v) for v in re.findall(r"-?\d+\.\d+", "\n".join(L[1209:1353]))]
    assert len(win) == 512 and np.array_equal(np.array(win, np.float32), orc.table("window"))
    lens = [int(v) for v in re.findall(r"-?\d+", "\n".join(L[2457:2473]))]
    codes = [int(v) for v in re.findall(r"-?\d+", "\n".join(L[2476:2492]))]
    assert lens == list(orc.table("len15")) and codes == list(orc.table("code15"))


def test_golden_fixtures(orc):
    """Regression pins generated by tools/make_golden.py from the oracle (the reference cannot run here)."""
    gold = json.load(open(GOLDEN))
    for case in gold["cases"]:
        pcm = getattr(signals, case["signal"])(**case["signal_args"])
        out, s = orc.encode_all(pcm, trace=True, **case["options"])
        g = s.gc_trace()
        assert hashlib.sha256(out).hexdigest() == case["sha256"], case["name"]
        assert len(out) == case["bytes"] and s.frame_count == case["frames"]
        assert hashlib.sha256(g["ix"].tobytes()).hexdigest() == case["ix_sha256"], case["name"]
        assert hashlib.sha256(g["spectrum"].tobytes()).hexdigest() == case["spectrum_sha256"], case["name"]
