"""An independent restatement of the spectral front end, in float64 numpy and in matrix form, against the oracle's traces.

The oracle (oracle/mp3_oracle.c) walks the reference's loops one sample at a time in float32; this file states the same
mathematics the other way round — whole-signal matrix products in float64, written from the formulas of
Sources/SwiftMP3/MP3Encoder.swift (SRC) and not from the oracle's code — so that a transcription slip in either shows up
as a disagreement.  The two differ only by float32 rounding, which is what north_star's tiers allow:
  tier 1  subband samples and MDCT spectra within 1e-5 of the granule's peak;
  tier 2  the quantizer's ix identical on (nearly) every line at the gain the oracle used.
Block types and gains are taken from the oracle's trace (the transient detector and the gain loop have their own
known-answer tests); everything between the PCM and ix is recomputed here."""
import numpy as np
import pytest

import signals

# ISO 11172-3 table B.9 as printed in SRC:1568-1575
CS = np.array([0.857492926, 0.881741997, 0.949628649, 0.983314592, 0.995517816, 0.999160558, 0.999899195, 0.999993155])
CA = np.array([-0.514495755, -0.471731969, -0.313377454, -0.181913200, -0.094574193, -0.040965583, -0.014198569, -0.003699975])


def _subbands(x, window):
    """PolyphaseFilterbank.analyze SRC:1367-1411 for a whole channel: returns S[step, subband]."""
    n_steps = len(x) // 32
    xp = np.concatenate([np.zeros(480), x.astype(np.float64)])               # the 512-buffer starts as zeros (SRC:275)
    idx = 32 * np.arange(n_steps)[:, None] + (511 - np.arange(512))[None, :]  # reversed buffer: newest sample first (SRC:1386-1387)
    z = xp[idx] * window[None, :]                                             # SRC:1389
    y = z.reshape(n_steps, 8, 64).sum(axis=1)                                 # Y[j] = sum_i Z[j + 64 i], SRC:1392-1399
    k, n = np.arange(32)[:, None], np.arange(64)[None, :]
    m = np.cos(np.pi / 64.0 * (2 * k + 1) * (n - 16))                         # SRC:1197-1206
    return y @ m.T


def _mdct(sub, block_types):
    """MDCT.apply SRC:1512-1565 (+ mdctLong 1619-1636, mdctShort 1639-1662, alias reduction 1581-1616) for a whole channel."""
    n_gr = sub.shape[0] // 18
    s = sub.reshape(n_gr, 18, 32).transpose(0, 2, 1).copy()                   # [granule, subband, t]
    s[:, 1::2, 1::2] *= -1.0                                                  # frequency inversion, SRC:1520-1524
    prev = np.concatenate([np.zeros((1, 32, 18)), s[:-1]])                    # overlap = the previous granule's (flipped) samples
    comb = np.concatenate([prev, s], axis=2)                                  # [granule, subband, 36]
    kk, mm = np.arange(36)[None, :], np.arange(18)[:, None]
    long_m = np.cos(np.pi / 72.0 * (2 * kk + 1 + 18) * (2 * mm + 1))          # SRC:1422-1433
    long_w = np.sin(np.pi / 36.0 * (np.arange(36) + 0.5))                     # SRC:1450-1457
    k2, m2 = np.arange(12)[None, :], np.arange(6)[:, None]
    short_m = np.cos(np.pi / 24.0 * (2 * k2 + 1 + 6) * (2 * m2 + 1))          # SRC:1436-1447
    short_w = np.sin(np.pi / 12.0 * (np.arange(12) + 0.5))                    # SRC:1460-1467
    out_long = (comb * long_w) @ long_m.T / 9.0                               # [granule, subband, 18]
    out_short = np.zeros_like(out_long)
    for w in range(3):                                                        # SRC:1645-1659: window w at offset 6 w + 6, out[w + 3 m]
        seg = comb[:, :, 6 * w + 6: 6 * w + 18] * short_w
        out_short[:, :, w::3] = seg @ short_m.T / 3.0
    bt = np.asarray(block_types)
    use_long = (bt == 0)[:, None] | ((bt == 1)[:, None] & (np.arange(32) < 2)[None, :])   # SRC:1542-1553 (mixed = raw value 1)
    spec = np.where(use_long[:, :, None], out_long, out_short).reshape(n_gr, 576)
    for g in np.nonzero(bt == 0)[0]:                                          # alias reduction, long blocks only (SRC:1560-1562)
        x = spec[g]
        sb = np.arange(31)[:, None]
        iu, il = sb * 18 + 17 - np.arange(8)[None, :], (sb + 1) * 18 + np.arange(8)[None, :]
        upper, lower = x[iu].copy(), x[il].copy()
        x[iu] = lower * CA + upper * CS
        x[il] = lower * CS - upper * CA
    return spec


def _quantize(spec, gains):
    """quantizeWithGain SRC:797-825 at the gain the oracle's loop ended on."""
    step = np.maximum(2.0 ** ((np.asarray(gains, dtype=np.float64) - 210.0) / 4.0), 1e-4)[:, None]
    scaled = np.maximum(np.abs(spec), 1e-10) ** 0.75 / step
    q = np.minimum(np.floor(scaled + 0.5), 15).astype(np.int64)               # .rounded(): ties away from zero (scaled >= 0)
    return np.where(spec < 0, -q, q)


@pytest.mark.parametrize("name", ["sine_noise_stereo", "white_mono_48k", "castanets_stereo"])
def test_front_end_against_float64_restatement(orc, name):
    if name == "sine_noise_stereo":
        pcm, opts, ch = signals.sine_noise(1.0), dict(mode="stereo"), 2
    elif name == "white_mono_48k":
        pcm, opts, ch = signals.white(1.0), dict(mode="mono", sample_rate=48000, bitrate_kbps=320), 1
    else:
        pcm, opts, ch = signals.castanets(2.0), dict(mode="stereo", bitrate_kbps=128), 2
    n_frames = len(pcm) // (1152 * ch)
    pcm = pcm[: n_frames * 1152 * ch]                                         # whole frames: no flush padding to model
    _, rs = orc.encode_all(pcm, trace=True, **opts)
    gt = rs.gc_trace()
    assert len(gt) == n_frames * 2 * ch
    window = orc.table("window").astype(np.float64)
    worst_sub = worst_spec = 0.0
    lines = same = 0
    for c in range(ch):
        x = pcm[c::ch]
        tr = gt[c::ch]                                                        # gc order is (granule, channel)
        sub = _subbands(x, window)
        ref_sub = tr["subband"].reshape(-1, 32, 18).transpose(0, 2, 1).reshape(-1, 32)   # oracle layout [sb][t] per granule
        peak = np.abs(ref_sub).reshape(-1, 18 * 32).max(axis=1).repeat(18)[:, None]
        worst_sub = max(worst_sub, float((np.abs(sub - ref_sub) / np.maximum(peak, 1e-12)).max()))
        spec = _mdct(sub, tr["block_type"])
        gpeak = np.abs(tr["spectrum"]).max(axis=1)[:, None]
        worst_spec = max(worst_spec, float((np.abs(spec - tr["spectrum"]) / np.maximum(gpeak, 1e-12)).max()))
        ix = _quantize(spec, tr["gain_used"])
        lines += ix.size
        same += int((ix == tr["ix"]).sum())
        if name == "castanets_stereo":
            assert (tr["block_type"] == 2).any() and (tr["block_type"] == 1).any()      # short and mixed blocks are exercised
    assert worst_sub < 2e-6, worst_sub                                        # tier 1 asks for 1e-5; measured 4e-7 (float32 rounding)
    assert worst_spec < 2e-6, worst_spec
    assert same / lines > 0.9999, (same, lines)                               # tier 2, per line; measured: every line identical


SFB_LONG = {44100: [4, 4, 4, 4, 4, 4, 6, 6, 8, 8, 10, 12, 16, 20, 24, 28, 34, 42, 50, 54, 76],          # SRC:1814
            48000: [4, 4, 4, 4, 4, 4, 6, 6, 6, 8, 10, 12, 16, 18, 22, 28, 34, 40, 46, 54, 54],          # SRC:1817
            32000: [4, 4, 4, 4, 4, 4, 6, 6, 8, 10, 12, 16, 20, 24, 30, 38, 46, 56, 68, 84, 102]}        # SRC:1820


@pytest.mark.parametrize("name", ["sine_noise_stereo", "white_mono_48k", "castanets_stereo", "sine_32k"])
def test_decisions_and_side_fields_against_restatement(orc, name):
    """The per-granule decisions and side-info fields, restated from SRC in numpy on the oracle's own spectrum / ix:
    block type + subblock_gain (SRC:1944-1968), initial gain (989-1006), big_values (692-700), the table-15 bit count
    (828-853) = part2_3_length, preflag (2042-2066), region counts (856-887), and the gain loop's outcome (734-794): the
    count fits the budget unless the loop ran out of iterations, and the gain before it did not fit."""
    sr = 44100
    if name == "sine_noise_stereo":
        pcm, opts, ch = signals.sine_noise(1.0), dict(mode="stereo"), 2
    elif name == "white_mono_48k":
        pcm, opts, ch, sr = signals.white(1.0), dict(mode="mono", sample_rate=48000, bitrate_kbps=320), 1, 48000
    elif name == "castanets_stereo":
        pcm, opts, ch = signals.castanets(2.0), dict(mode="stereo", bitrate_kbps=128), 2
    else:
        pcm, opts, ch, sr = signals.sine_noise(1.0, sr=32000), dict(mode="stereo", sample_rate=32000, bitrate_kbps=64), 2, 32000
    n_frames = len(pcm) // (1152 * ch)
    pcm = pcm[: n_frames * 1152 * ch]
    _, rs = orc.encode_all(pcm, trace=True, **opts)
    gt = rs.gc_trace()
    len15 = orc.table("len15").astype(int).reshape(16, 16)
    bounds = np.cumsum(SFB_LONG[sr])
    bt_same = 0
    for i, t in enumerate(gt):
        c, g = i % ch, i // ch
        x = pcm[c::ch][576 * g: 576 * (g + 1)].astype(np.float64)
        e3 = (x.reshape(3, 192) ** 2).sum(axis=1) / 192.0
        mx, mn = e3.max(), e3.min()
        ratio = mx / max(mn, 1e-4)
        bt = (1 if int(np.argmax(e3)) == 0 else 2) if ratio > 6.0 else 0
        sbg = [int((1.0 - min(max(e / max(mx, 1e-4), 0.0), 1.0)) * 7.0) for e in e3]
        bt_same += int(bt == t["block_type"] and sbg == list(t["subblock_gain"]))
        spec = t["spectrum"].astype(np.float64)
        peak = np.abs(spec).max()
        g0 = 210 if peak <= 0 else min(max(210 + int(4.0 * np.log2(peak ** 0.75 / 15.0)), 0), 255)
        assert abs(g0 - t["g0"]) <= 1, (i, g0, t["g0"])                        # float32 powf against float64 at a truncation edge
        ix = t["ix"]
        nz = np.nonzero(ix)[0]
        last = int(nz[-1]) + 1 if nz.size else 0
        bv = min(((last + 1) & ~1) // 2, 288)
        assert bv == t["big_values"]
        a = np.minimum(np.abs(ix[: 2 * bv]), 15).reshape(-1, 2)
        bits = int(len15[a[:, 0], a[:, 1]].sum() + np.count_nonzero(a))
        assert bits == t["bits"], (i, bits, t["bits"])
        assert int((spec[432:] ** 2).sum() > 1.5 * (spec[:432] ** 2).sum()) == t["preflag"] or \
            abs((spec[432:] ** 2).sum() - 1.5 * (spec[:432] ** 2).sum()) < 1e-6 * (spec ** 2).sum()
        r0 = 0
        for k in range(15):
            if bounds[k] <= 2 * bv:
                r0 = k
            else:
                break
        r1 = 0
        for k in range(r0 + 1, min(r0 + 8, 21)):
            if bounds[k] <= 2 * bv:
                r1 = k - r0 - 1
            else:
                break
        assert (min(r0, 15), min(r1, 7)) == (t["region0"], t["region1"])
        # the loop's outcome: it stopped because the count fits, or it ran out of iterations / hit gain 255
        assert t["bits"] <= t["max_bits"] or t["iterations"] == 20 or t["gain_out"] >= 255, (i, t["bits"], t["max_bits"])
        assert t["gain_used"] <= t["gain_out"] <= 255
    assert bt_same >= 0.995 * len(gt), (bt_same, len(gt))                      # float64 sums against float32 at the ratio / truncation edges


@pytest.mark.parametrize("cfg", [dict(mode="stereo"), dict(mode="mono", bitrate_kbps=64), dict(mode="stereo", crc_protected=True, bitrate_kbps=320, sample_rate=48000),
                                 dict(mode="jointStereo", vbr=True, quality=2), dict(mode="stereo", sample_rate=32000, bitrate_kbps=96)])
def test_frame_chain_against_restatement(orc, cfg):
    """encodeFrame's serial chain (SRC:475-568: padding accumulator, frame / slot sizes, reservoir snapshot, per-granule
    budget, stream FIFO, updateReservoir) replayed in Python from the oracle's own per-frame bitrate and per-granule bit
    counts; every frame field and every granule's max_bits must come out the same."""
    ch = 1 if cfg["mode"] == "mono" else 2
    sr = cfg.get("sample_rate", 44100)
    pcm = signals.castanets(2.0, sr=sr) if cfg.get("vbr") else signals.sine_noise(1.3, sr=sr)
    if ch == 1:
        pcm = np.ascontiguousarray(pcm[0::2])
    pcm = pcm[: len(pcm) - 333 * ch]                                          # ragged: the last frame is flush()'s padded one
    _, rs = orc.encode_all(pcm, trace=True, **cfg)
    ft, gt = rs.frame_trace(), rs.gc_trace()
    side, crc = (17 if ch == 1 else 32), (2 if cfg.get("crc_protected") else 0)
    pad_rem = avail = backlog = 0
    prev_slot = None
    for f, t in enumerate(ft):
        num = 144 * int(t["bitrate_kbps"]) * 1000
        pad_rem += num % sr                                                   # shouldPad SRC:456-463
        padding = 0
        if pad_rem >= sr:
            pad_rem -= sr; padding = 1
        frame_size = num // sr + padding
        mds = frame_size - 4 - crc - side
        final = f == len(ft) - 1                                              # the ragged tail: encodeFrame(isFinal: true) from flush()
        mdb = 0 if final else min(backlog, 511)                               # SRC:499, 2099-2101
        res_bits = 0 if final else avail * 8                                  # SRC:500
        bpg = (mds * 8 + res_bits * 9 // 10) // (2 * ch)                      # SRC:647-650
        gcs = gt[f * 2 * ch: (f + 1) * 2 * ch]
        huff = (int(gcs["bits"].sum()) + 7) // 8                              # one byte pad per frame, SRC:729
        assert (padding, frame_size, mds, mdb, res_bits, huff, int(final)) == \
            (t["padding"], t["frame_size"], t["main_data_size"], t["main_data_begin"], t["reservoir_bits"], t["huff_bytes"], t["is_final"]), f
        assert (gcs["max_bits"] == bpg).all(), f
        backlog += huff                                                       # appendHuffmanData SRC:511
        if prev_slot is not None:
            backlog -= min(prev_slot, backlog)                                # fillSlot of the buffered frame SRC:548-556, 2110-2121
        prev_slot = mds
        avail = min(max(avail + mds - huff, 0), 511)                          # updateReservoir SRC:2125-2128
    assert rs.frame_count == len(ft)


@pytest.mark.parametrize("cfg", [dict(mode="jointStereo", vbr=True, quality=2), dict(mode="mono", vbr=True, quality=7, bitrate_kbps=96),
                                 dict(mode="stereo", vbr=True, quality=0, bitrate_kbps=256, sample_rate=48000)])
def test_vbr_bitrate_choice_against_restatement(orc, cfg):
    """VBRState.chooseBitrate (SRC:1177-1189) + VBRState.update (1144-1153) + MP3Tables.bitrateIndex (2509-2523) replayed in
    float32 numpy from the oracle's own frame / granule energies: the bitrate of every frame must come out the same."""
    f32 = np.float32
    ch = 1 if cfg["mode"] == "mono" else 2
    sr, base, q = cfg.get("sample_rate", 44100), cfg.get("bitrate_kbps", 128), cfg["quality"]
    pcm = signals.castanets(3.0, sr=sr)
    if ch == 1:
        pcm = np.ascontiguousarray(pcm[0::2])
    _, rs = orc.encode_all(pcm, trace=True, **cfg)
    ft, gt = rs.frame_trace(), rs.gc_trace()
    table = [0, 32, 40, 48, 56, 64, 80, 96, 112, 128, 160, 192, 224, 256, 320, 0]             # SRC:2511
    hist, seen = [], set()
    for f, t in enumerate(ft):
        e = f32(t["frame_energy"])
        if hist:
            s = f32(0.0)
            for h in hist:
                s = f32(s + h)
            average = f32(s / f32(len(hist)))
        else:
            average = e
        ratio = min(max(f32(e / max(average, f32(0.0001))), f32(0.5)), f32(2.0))
        qf = f32(f32(9 - q) / f32(9.0))
        max_adj = int(f32(f32(32.0) + f32(f32(32.0) * qf)))
        adj = int(f32(f32(ratio - f32(1.0)) * f32(max_adj)))
        lo, hi = max(32, base - 64 + q * 8), min(320, base + 64 - q * 4)
        want = max(lo, min(base + adj, hi))
        idx = table.index(want) if want in table else min(range(16), key=lambda i: abs(table[i] - want))   # first minimum, like min(by:)
        assert (idx, table[idx]) == (t["bitrate_index"], t["bitrate_kbps"]), (f, want)
        seen.add(int(t["bitrate_kbps"]))
        for g in gt[f * 2 * ch: (f + 1) * 2 * ch]:                            # VBRState.update: ten most recent granule energies
            hist.append(f32(g["energy"]))
        hist = hist[-10:]
    assert len(seen) >= 2                                                     # the bitrate really moves on this signal
