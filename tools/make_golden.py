#!/usr/bin/env python3
"""Generate tests/golden/oracle_golden.json: SHA-256 pins of the oracle's output bytes, quantized ix and MDCT spectra
for small seeded inputs.  The reference (Swift + Accelerate) cannot run on Linux, so these pin the ORACLE (regression
guard for it and a second, size-independent target for the GPU tests), not the reference: parity stays "unpinned"."""
import hashlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_binding as orc, signals
CASES = [
    ("c1_sine_noise_stereo_128", "sine_noise", dict(seconds=1.0), dict()),
    ("c2_white_mono_48k_320", "white", dict(seconds=1.0), dict(sample_rate=48000, bitrate_kbps=320, mode="mono")),
    ("c3_castanets_joint_vbr_q2", "castanets", dict(seconds=2.0), dict(mode="jointStereo", vbr=True, quality=2)),
    ("sine440_12_frames_reservoir", "sine440", dict(frames=12, amp=0.1), dict()),
    ("crc_32k_64", "sine_noise", dict(seconds=0.5, sr=32000, seed=9), dict(sample_rate=32000, bitrate_kbps=64, crc_protected=True)),
]
out = {"generator": "tools/make_golden.py", "source": "oracle/mp3_oracle.c", "cases": []}
for name, sig, sargs, opts in CASES:
    pcm = getattr(signals, sig)(**sargs)
    data, s = orc.encode_all(pcm, trace=True, **opts)
    g = s.gc_trace()
    out["cases"].append(dict(name=name, signal=sig, signal_args=sargs, options=opts, bytes=len(data), frames=int(s.frame_count),
                             sha256=hashlib.sha256(data).hexdigest(), ix_sha256=hashlib.sha256(g["ix"].tobytes()).hexdigest(),
                             spectrum_sha256=hashlib.sha256(g["spectrum"].tobytes()).hexdigest(),
                             block_types=[int((g["block_type"] == k).sum()) for k in range(3)],
                             mean_iterations=float(g["iterations"].mean())))
os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "oracle_golden.json"), "w"), indent=1)
print(json.dumps([(c["name"], c["frames"], c["block_types"], c["mean_iterations"]) for c in out["cases"]], indent=1))
