"""Test-only float64 numpy restatement of the engine's ISO-mode psychoacoustic model (swift-mp3_b200/csrc/iso_psy.cuh, tables:
engine.cc build_psy_tab), written from the model's definition, not from the kernel: whole-array FFTs (numpy.fft) instead of the
warp's radix-4 passes, matrix products instead of per-lane loops.  There is no reference behaviour for this stage (the reference's
thresholds are a band mean nobody reads, SRC:1983-2013, 737), so the definition is the engine's own and this file is what pins it.
"""
import numpy as np

SFB_LONG = {0: [4, 8, 12, 16, 20, 24, 30, 36, 44, 52, 62, 74, 90, 110, 134, 162, 196, 238, 288, 342, 418],      # 44.1 kHz
            1: [4, 8, 12, 16, 20, 24, 30, 36, 42, 50, 60, 72, 88, 106, 128, 156, 190, 230, 276, 330, 384],      # 48 kHz
            2: [4, 8, 12, 16, 20, 24, 30, 36, 44, 54, 66, 82, 102, 126, 156, 194, 240, 296, 364, 448, 550]}     # 32 kHz
MAXP = 80


def bark(f):
    return 13.0 * np.arctan(0.00076 * f) + 3.5 * np.arctan((f / 7500.0) ** 2)


def tables(sample_rate, sfb_index):
    fs = float(sample_rate)
    lo, parts = 0, []
    while lo < 512 and len(parts) < MAXP:
        hi = lo + 1
        while hi < 512 and bark(hi * fs / 1024.0) - bark(lo * fs / 1024.0) < 1.0 / 3.0:
            hi += 1
        if len(parts) == MAXP - 1:
            hi = 512
        parts.append((lo, hi - lo)); lo = hi
    P = len(parts)
    plo = np.array([p[0] for p in parts]); pn = np.array([p[1] for p in parts])
    bval = bark((plo + 0.5 * (pn - 1)) * fs / 1024.0)
    s3 = np.zeros((P, P))                                  # [target][source]
    for i in range(P):
        for j in range(P):
            tx = (3.0 if j >= i else 1.5) * (bval[i] - bval[j])
            x = 0.0
            if 0.5 <= tx <= 2.5:
                u = tx - 0.5; x = 8.0 * (u * u - 2.0 * u)
            tx += 0.474
            ty = 15.811389 + 7.5 * tx - 17.5 * np.sqrt(1.0 + tx * tx)
            s3[i, j] = 0.0 if ty <= -60.0 else 10.0 ** ((x + ty) / 10.0)
    rnorm = 1.0 / s3.sum(axis=1)
    minval = np.clip(24.5 - 2.0 * bval, 4.5, 24.5)
    k = np.arange(512)
    f = np.maximum(k * fs / 1024.0, 20.0) / 1000.0
    ath = np.minimum(3.64 * f ** -0.8 - 6.5 * np.exp(-0.6 * (f - 3.3) ** 2) + 1e-3 * f ** 4, 80.0)
    qline = (32768.0 * 256.0) ** 2 * 10.0 ** ((ath - 96.0) / 10.0)
    qthr = np.array([qline[a:a + n].sum() for a, n in parts])
    line_part = np.repeat(np.arange(P), pn)
    sfb_line = [0] + [(c * 8 + 4) // 9 for c in SFB_LONG[sfb_index]] + [512]
    return dict(P=P, plo=plo, pn=pn, s3=s3, rnorm=rnorm, minval=minval, qthr=qthr, line_part=line_part, sfb_line=sfb_line)


def hann(n):
    return 0.5 * (1.0 - np.cos(2.0 * np.pi * (np.arange(n) + 0.5) / n))


def analyse(x1024, T):
    """x1024: the 1024 samples [576 g - 768, 576 g + 256) of the coded channel, full scale = 1.0.  -> (ratio[22], pe, mean tonality)."""
    v = np.asarray(x1024, dtype=np.float64) * 32768.0
    w256 = hann(256)
    S = [np.fft.fft(v[192 + 192 * w: 448 + 192 * w] * w256)[:129] for w in range(3)]
    r = [np.abs(a) for a in S]
    u0 = np.where(r[0] > 0, S[0] / np.maximum(r[0], 1e-300), 1.0)
    u1 = np.where(r[1] > 0, S[1] / np.maximum(r[1], 1e-300), 1.0)
    pred = (2.0 * r[1] - r[0]) * (u1 * u1 * np.conj(u0))
    den = r[2] + np.abs(2.0 * r[1] - r[0])
    cw_s = np.where(den > 0, np.abs(S[2] - pred) / np.maximum(den, 1e-300), 0.0)
    X = np.fft.fft(v * hann(1024))[:512]
    e = np.abs(X) ** 2
    k = np.arange(512)
    cw = np.where(k < 206, cw_s[np.minimum((k + 2) >> 2, 128)], 0.4)
    P = T["P"]
    eb = np.array([e[a:a + n].sum() for a, n in zip(T["plo"], T["pn"])])
    cb = np.array([(e * cw)[a:a + n].sum() for a, n in zip(T["plo"], T["pn"])])
    ecb = T["s3"] @ eb; ctb = T["s3"] @ cb
    cbb = np.where(ecb > 0, ctb / np.maximum(ecb, 1e-300), 0.0)
    tb = np.where(cbb > 0, np.clip(-0.299 - 0.43 * np.log(np.maximum(cbb, 1e-300)), 0.0, 1.0), 1.0)
    snr = np.maximum(T["minval"], 29.0 * tb + 6.0 * (1.0 - tb))
    nb = ecb * T["rnorm"] * 10.0 ** (-0.1 * snr)
    thr = np.maximum(T["qthr"], nb)
    pe = float(np.sum(T["pn"] * np.log((eb + 1.0) / (thr + 1.0))))
    thr_line = (thr / T["pn"])[T["line_part"]]
    sl = T["sfb_line"]
    ratio = np.array([thr_line[sl[b]:sl[b + 1]].sum() / max(e[sl[b]:sl[b + 1]].sum(), 1e-20) for b in range(22)])
    return ratio, max(pe, 0.0), float(tb.mean())
