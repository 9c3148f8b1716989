"""Test-only numpy restatement of the engine's ISO-mode bit counting (swift-mp3_b200/csrc/iso_mode.cuh: iso_evaluate, iso_candidates),
written from the description of the rules, with the Huffman tables of tests/iso_huffman.json: rzero / count1 / big_values
partition, three regions on scalefactor-band boundaries (two for window-switched granules), per region the cheapest table of
the group its largest value falls in, count1 table A or B.  It is the oracle for that stage: the GPU's part2_3_length,
table_select, count1table_select, big_values and regions must equal it, and the gain the engine settles on must be the smallest
whose count fits the budget."""
import numpy as np

import isoparse

SFB = isoparse.SFB_LONG


def _candidates(m, linbits):
    if m == 0: return [0]
    if m == 1: return [1]
    if m == 2: return [2, 3]
    if m == 3: return [5, 6]
    if m <= 5: return [7, 8, 9]
    if m <= 7: return [10, 11, 12]
    if m <= 15: return [13, 15]
    need = m - 15
    a = sum(1 for t in range(16, 24) if (1 << linbits[t]) <= need)
    b = sum(1 for t in range(24, 32) if (1 << linbits[t]) <= need)
    return [16 + min(a, 7), 24 + min(b, 7)]


def count(ix_abs, sr_index, ws=False):
    """ix_abs: 576 non-negative ints in bitstream order -> dict(bits, big_values, count1, table_select, count1table, region0, region1)."""
    luts, linbits, quad, raw = isoparse.tables()
    tabs = raw["tables"]
    v = np.asarray(ix_abs, dtype=np.int64)
    px, py = v[0::2], v[1::2]
    nz = np.nonzero((px | py) != 0)[0]; top = int(nz[-1]) + 1 if len(nz) else 0
    bg = np.nonzero((px > 1) | (py > 1))[0]; big = int(bg[-1]) + 1 if len(bg) else 0
    c1 = (top - big) >> 1
    bv = top - 2 * c1
    cum = SFB[sr_index][1:22]                                  # band ends 0..20
    if ws:
        r0 = r1 = 0; a1, a2 = 36, 576
    else:
        nb = sum(1 for e in cum if e <= 2 * bv)
        k0 = min(max((nb + 1) // 3, 1), 16); k1 = min(max((nb + 1) // 3, 1), 8)
        r0, r1 = k0 - 1, k1 - 1
        a1 = cum[k0 - 1]; a2 = cum[k0 + k1 - 1] if k0 + k1 - 1 < 21 else 576
    bits, sel = 0, []
    for lo, hi in ((0, a1), (a1, a2), (a2, 576)):
        p = np.arange(288)
        inr = (p < bv) & (2 * p >= lo) & (2 * p < hi)
        x, y = px[inr], py[inr]
        m = int(max(x.max(initial=0), y.max(initial=0)))
        best, bt = None, 0
        for t in _candidates(m, linbits):
            if t == 0:
                tot = 0
            else:
                src = str(t if t < 16 else 16 if t < 24 else 24)
                L = np.array(tabs[src]["len"])
                tot = int(L[np.minimum(x, 15), np.minimum(y, 15)].sum()) + linbits.get(t, 0) * int((x >= 15).sum() + (y >= 15).sum())
            if best is None or tot < best:
                best, bt = tot, t
        bits += best; sel.append(bt)
    bits += int((px[:bv] != 0).sum() + (py[:bv] != 0).sum())
    q = v[2 * bv:2 * bv + 4 * c1].reshape(-1, 4)
    idx = (q[:, 0] << 3) | (q[:, 1] << 2) | (q[:, 2] << 1) | q[:, 3] if len(q) else np.zeros(0, np.int64)
    sg = q.sum(axis=1) if len(q) else np.zeros(0, np.int64)
    ca = int(np.array(raw["quad"]["len"][0])[idx].sum() + sg.sum()) if len(q) else 0
    cb = int(4 * len(q) + sg.sum())
    return dict(bits=bits + min(ca, cb), big_values=bv, count1=c1, table_select=sel, count1table=int(cb < ca), region0=r0, region1=r1)
