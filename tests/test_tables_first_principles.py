"""The constant tables against FIRST PRINCIPLES — independent of the reference's literals and of the generator that wrote both
copies of `iso_tables.inc` (oracle/ and csrc/ share one generated file: a slip there would be common-mode).

  * the cosine / sine tables are their closed forms (SRC:1402-1408 analysis matrix, 1619-1662 MDCT, 1470-1503 windows);
  * the 512-tap analysis window (ISO 11172-3 Table C.1) has no closed form, but together with the analysis matrix it must BE a
    32-band pseudo-QMF bank: a constant input comes out of subband 0 with unit gain and nowhere else, a sinusoid at the centre of
    subband k comes out of subband k with unit gain (rms 1 / sqrt 2) and of its neighbours below -100 dB; its magnitudes are
    symmetric about tap 256, its peak is the standard's 0.035780907;
  * the alias-reduction butterflies are cs = 1 / sqrt(1 + c^2), ca = c / sqrt(1 + c^2) of the eight ISO coefficients c;
  * the long scalefactor-band widths are ISO 11172-3 Table B.8's band edges for 44.1 / 48 / 32 kHz.
"""
import os, re
import numpy as np
import oracle_binding as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _inc(name):
    txt = open(os.path.join(ROOT, "oracle", "iso_tables.inc")).read()
    m = re.search(name + r"\[\d+\] = \{(.*?)\};", txt, re.S)
    return np.array([float(v.rstrip("f")) for v in re.findall(r"-?\d+\.?\d*(?:e-?\d+)?f?", m.group(1))])


def test_both_copies_of_the_generated_tables_are_one_file():
    a = open(os.path.join(ROOT, "oracle", "iso_tables.inc")).read()
    b = open(os.path.join(ROOT, "swift-mp3_b200", "csrc", "iso_tables.inc")).read()
    assert a == b


def test_trigonometric_tables_are_their_closed_forms():
    tol = 6e-8                                                                     # half an ulp of a float32 near 1
    m, k = np.meshgrid(np.arange(18), np.arange(36), indexing="ij")
    assert np.abs(orc.table("mdct_long").reshape(18, 36) - np.cos(np.pi / 72 * (2 * k + 1 + 18) * (2 * m + 1))).max() < tol
    m, k = np.meshgrid(np.arange(6), np.arange(12), indexing="ij")
    assert np.abs(orc.table("mdct_short").reshape(6, 12) - np.cos(np.pi / 24 * (2 * k + 1 + 6) * (2 * m + 1))).max() < tol
    assert np.abs(orc.table("win_long") - np.sin(np.pi / 36 * (np.arange(36) + 0.5))).max() < tol
    assert np.abs(orc.table("win_short") - np.sin(np.pi / 12 * (np.arange(12) + 0.5))).max() < tol
    kk, nn = np.meshgrid(np.arange(32), np.arange(64), indexing="ij")
    assert np.abs(orc.table("analysis").reshape(32, 64) - np.cos((2 * kk + 1) * (nn - 16) * np.pi / 64)).max() < tol


def _analyze(x512):
    """One filterbank step on a 512-sample window vector in the reference's order (SRC:1386-1408): Y[n] = sum_j C[n + 64 j] X[n + 64 j],
    S[k] = sum_n M[k][n] Y[n]."""
    w = orc.table("window").astype(np.float64)
    M = orc.table("analysis").astype(np.float64).reshape(32, 64)
    return M @ (w * x512).reshape(8, 64).sum(0)


def test_window_and_matrix_are_a_32_band_filterbank():
    w = orc.table("window").astype(np.float64)
    i = np.arange(1, 256)
    assert np.array_equal(np.abs(w[i]), np.abs(w[512 - i])) and w[0] == 0.0          # magnitudes symmetric about tap 256
    assert abs(w[256] - 0.035780907) < 1e-9 and np.argmax(np.abs(w)) == 256           # ISO Table C.1's centre tap
    dc = _analyze(np.ones(512))
    assert abs(dc[0] - 1.0) < 1e-4 and np.abs(dc[1:]).max() < 1e-5                    # DC: subband 0, unit gain, -100 dB elsewhere
    n = np.arange(512)
    for k in range(32):
        om = (k + 0.5) * np.pi / 32                                                   # centre of subband k
        out = np.array([_analyze(np.cos(om * n + ph)) for ph in np.linspace(0, np.pi, 9)[:-1]])
        rms = np.sqrt((out ** 2).mean(0))
        assert abs(rms[k] - np.sqrt(0.5)) < 1e-4
        assert np.delete(rms, k).max() < 1e-5


def test_alias_butterflies_and_scalefactor_bands_are_the_standards():
    c = np.array([-0.6, -0.535, -0.33, -0.185, -0.095, -0.041, -0.0142, -0.0037])    # ISO 11172-3 Table B.9
    assert np.abs(_inc("ISO_ALIAS_CS") - 1 / np.sqrt(1 + c * c)).max() < 1e-8
    assert np.abs(_inc("ISO_ALIAS_CA") - c / np.sqrt(1 + c * c)).max() < 1e-8
    edges = {44100: [0, 4, 8, 12, 16, 20, 24, 30, 36, 44, 52, 62, 74, 90, 110, 134, 162, 196, 238, 288, 342, 418, 576],
             48000: [0, 4, 8, 12, 16, 20, 24, 30, 36, 42, 50, 60, 72, 88, 106, 128, 156, 190, 230, 276, 330, 384, 576],
             32000: [0, 4, 8, 12, 16, 20, 24, 30, 36, 44, 54, 66, 82, 102, 126, 156, 194, 240, 296, 364, 448, 550, 576]}   # Table B.8, long blocks
    for sr, e in edges.items():
        assert list(_inc("ISO_SFB_LONG_%d" % sr).astype(int)) == list(np.diff(e)[:21])


def test_huffman_table_15_is_a_complete_prefix_code():
    """Kraft equality: the 256 code lengths of ISO table 15 fill the code space exactly, and the codes are prefix-free."""
    ln, code = orc.table("len15").astype(int), orc.table("code15").astype(int)
    assert abs(sum(2.0 ** -int(l) for l in ln) - 1.0) < 1e-12
    words = sorted(format(int(c), "b").zfill(int(l)) for c, l in zip(code, ln))
    assert len(set(words)) == 256 and all(not b.startswith(a) for a, b in zip(words, words[1:]))
