#!/usr/bin/env python3
"""Order-sensitivity study: what "parity unpinned" is worth in numbers, and the control the tensor-core decision needs.

The oracle fixes choices Apple's closed Accelerate / libm do not document ([OD1]..[OD3] in oracle/mp3_oracle.c: summation
order of vDSP_dotpr / vDSP_sve / vDSP_svesq, fused or unfused multiply-add, the last bit of powf).  The CUDA path is bit-exact
against the oracle; against the REAL reference the residual is whatever those choices move.  This tool measures it: the
oracle's front end is re-run with each alternative (oracle/variants.inc, built into libmp3oracle_var.so) on BASELINE configs
1-3 and, per granule-channel, the variant's spectrum is quantized at the gain the baseline used (SRC:797-825).  A
granule-channel "changes" when its block type, its initial gain (SRC:989-1006) or any of its 576 quantized values differ;
because the spectra do not depend on the bit reservoir, every granule-channel is an independent trial.  The whole-stream
bytes are compared too (one changed value moves the reservoir and with it every later frame of the stream).
Variants 6-8 put the 32x64 matrixing (SRC:1402-1408) on a split-TF32 tensor-core model; variant 12 is their control
(the same matrixing in plain FP32, other summation order).

usage: tools/order_sensitivity.py [scale, default 1.0 = about 1.1e6 granule-channels per variant] > profiles/r02_order_sensitivity.txt
"""
import ctypes as C
import os
import subprocess
import sys
import time
from concurrent.futures import ProcessPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_binding as orc   # noqa: E402
import signals                 # noqa: E402

VARIANTS = [(1, "dot products in 4 lanes (128-bit SIMD order), fused"), (2, "dot products in 8 lanes (256-bit SIMD order), fused"),
            (3, "dot products unfused (mul, then add), ascending"), (4, "|x|^0.75 one ulp up"), (5, "|x|^0.75 one ulp down"),
            (9, "energies in 4 lanes"), (10, "energies in 8 lanes"), (11, "window sums as a pairwise tree"),
            (12, "CONTROL: matrixing only in 8 lanes, plain FP32"), (6, "matrixing 3xTF32 (2-term split, 3 products)"),
            (7, "matrixing exact 3-term TF32 split (6 products), RN accumulate"), (8, "matrixing exact 3-term TF32 split (6 products), RZ accumulate")]


def var_lib():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "variants"])
    L = orc._declare(C.CDLL(os.path.join(ROOT, "oracle", "libmp3oracle_var.so")))
    L.orc_set_variant.argtypes = [C.c_int]
    L.orc_requantize.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    return L


def pink_fast(seconds, sr, seed, peak=0.5):
    from scipy.signal import lfilter
    w = np.random.default_rng(seed).standard_normal(int(round(seconds * sr)))
    out = lfilter([0.0990460], [1, -0.99765], w) + lfilter([0.2965164], [1, -0.96300], w) + lfilter([1.0526913], [1, -0.57000], w) + 0.1848 * w
    return (out * (peak / np.max(np.abs(out)))).astype(np.float32)


def make_stream(kind, i):
    if kind == "c1":        # C1 / C4 recipe, 30 s stereo CBR 128
        fl = 110.0 * 2.0 ** ((i % 48) / 12.0)
        return orc.synth_fill(30 * 44100, 2, 44100, fl, fl * 1.26, 0.5, 0.05, 1000 + i), dict(sample_rate=44100, bitrate_kbps=128, mode="stereo")
    if kind == "c2white":   # 60 s 48 kHz mono CBR 320, U(-0.5, 0.5)
        return signals.white(60.0, seed=2000 + i), dict(sample_rate=48000, bitrate_kbps=320, mode="mono")
    if kind == "c2pink":
        return pink_fast(60.0, 48000, 3000 + i), dict(sample_rate=48000, bitrate_kbps=320, mode="mono")
    return signals.castanets(30.0, seed=4000 + i, period=0.25), dict(sample_rate=44100, bitrate_kbps=128, mode="jointStereo", vbr=True, quality=2)


def session_trace(L, pcm, opts):
    o = orc.make_options(**opts)
    h = L.orc_create(C.byref(o))
    L.orc_trace_enable(h, 1)
    n = C.c_size_t()
    a = np.ascontiguousarray(pcm, np.float32)
    p = L.orc_encode(h, a.ctypes.data_as(C.c_void_p), a.size, C.byref(n)); out = C.string_at(p, n.value)
    p = L.orc_flush(h, C.byref(n)); out += C.string_at(p, n.value)
    k = L.orc_trace_gc_count(h)
    tr = np.frombuffer(C.string_at(L.orc_trace_gc(h), k * orc.GC_TRACE.itemsize), orc.GC_TRACE).copy()
    L.orc_destroy(h)
    return out, tr


def work(job):
    kind, i = job
    L = var_lib()
    pcm, opts = make_stream(kind, i)
    L.orc_set_variant(0)
    base_bytes, base = session_trace(L, pcm, opts)
    res = {}
    ix = np.zeros(576, np.int32)
    for v, _ in VARIANTS:
        L.orc_set_variant(v)
        vb, tr = session_trace(L, pcm, opts)
        changed = bt = g0 = lines = 0
        max_rel = 0.0
        for b, t in zip(base, tr):
            bad = False
            if b["block_type"] != t["block_type"]:
                bt += 1; bad = True
            elif b["g0"] != t["g0"]:
                g0 += 1; bad = True
            else:
                spec = np.ascontiguousarray(t["spectrum"])
                L.orc_requantize(spec.ctypes.data, int(b["gain_used"]), ix.ctypes.data)
                d = int(np.count_nonzero(ix != b["ix"]))
                if d:
                    lines += d; bad = True
            changed += bad
            pk = float(np.max(np.abs(b["spectrum"])))
            if pk > 0 and b["block_type"] == t["block_type"]:
                max_rel = max(max_rel, float(np.max(np.abs(t["spectrum"].astype(np.float64) - b["spectrum"]))) / pk)
        res[v] = (len(base), changed, bt, g0, lines, int(vb != base_bytes), max_rel)
    L.orc_set_variant(0)
    return kind, res


def main():
    scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
    plan = [("c1", max(1, round(100 * scale))), ("c2white", max(1, round(40 * scale))), ("c2pink", max(1, round(40 * scale))), ("c3", max(1, round(50 * scale)))]
    jobs = [(k, i) for k, n in plan for i in range(n)]
    t0 = time.time()
    agg = {}
    with ProcessPoolExecutor(max_workers=os.cpu_count()) as ex:
        for kind, res in ex.map(work, jobs, chunksize=1):
            for v, r in res.items():
                a = agg.setdefault((kind, v), [0, 0, 0, 0, 0, 0, 0.0, 0])
                for j in range(6):
                    a[j] += r[j]
                a[6] = max(a[6], r[6]); a[7] += 1
    print("# tools/order_sensitivity.py %g   (%d streams, %.0f s wall, %d processes)" % (scale, len(jobs), time.time() - t0, os.cpu_count()))
    print("# per variant and signal: granule-channels (gc) tested | gc changed (%) = block type + initial gain + ix | quantized values changed |")
    print("# streams whose BYTES change | max |spectrum difference| / granule peak (tier 1 allows 1e-5).  Tier 2 allows 0.01 % of gc.")
    names = {"c1": "C1/C4 sine+noise stereo 128k", "c2white": "C2 white mono 48k 320k", "c2pink": "C2 pink mono 48k 320k", "c3": "C3 castanets joint VBR q2"}
    for v, label in VARIANTS:
        print("\nvariant %2d: %s" % (v, label))
        tot = [0, 0, 0, 0]
        for kind, _ in plan:
            a = agg[(kind, v)]
            print("  %-30s gc %8d  changed %6d (%.4f %%: block type %d, initial gain %d, ix %d)  values %7d  streams %3d / %3d  max rel %.2e" %
                  (names[kind], a[0], a[1], 100.0 * a[1] / a[0], a[2], a[3], a[1] - a[2] - a[3], a[4], a[5], a[7], a[6]))
            tot[0] += a[0]; tot[1] += a[1]; tot[2] += a[5]; tot[3] += a[7]
        print("  %-30s gc %8d  changed %6d (%.4f %%)  streams with different bytes %d / %d" % ("ALL", tot[0], tot[1], 100.0 * tot[1] / tot[0], tot[2], tot[3]))


if __name__ == "__main__":
    main()
