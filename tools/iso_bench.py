#!/usr/bin/env python3
"""Device-plane throughput of the opt-in ISO mode (levels 1, 2 and 3) beside the reference-compatible default, on a C4-shaped batch
(stereo 44.1 kHz CBR 128, synthetic sine + noise streams).  usage: tools/iso_bench.py [streams] [seconds] [steps]  -> one JSON line"""
import ctypes as C, importlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
mp3 = importlib.import_module("swift-mp3_b200")
sharding = importlib.import_module("swift-mp3_b200.sharding")
L = mp3.lib()
S = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
secs = float(sys.argv[2]) if len(sys.argv) > 2 else 10.0
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
n_per = int(round(secs * 44100)) // 2 * 2
pcm = torch.empty((S, n_per * 2), dtype=torch.float32, device="cuda")
for i in range(S):
    fl, fr, seed = sharding.stream_params(i)
    assert L.mp3b_synth_fill(0, pcm[i].data_ptr(), n_per, 2, 44100, fl, fr, 0.5, 0.05, seed) == 0
ptrs = (C.c_void_p * S)(*[pcm[i].data_ptr() for i in range(S)])
ns = (C.c_size_t * S)(*([n_per * 2] * S))
out = {"streams": S, "seconds": n_per / 44100.0, "steps": steps}
for level in (0, 1, 2, 3):
    b = mp3.EncoderBatch(mp3.MP3EncoderOptions(mode=mp3.Mode.stereo), S, 0)
    if level:
        b.set_iso_mode(level)
    tot = {}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ext = torch.cuda.ExternalStream(b.cuda_stream, device=torch.device("cuda", 0))
    for k in range(steps + 2):
        if k == 2:
            e0.record(ext)
        b.reset(); b.encode_device(ptrs, ns, flush=True, download=False)
        if k >= 2:
            for n, v in b.stage_ms().items():
                tot[n] = tot.get(n, 0.0) + v / steps
    e1.record(ext); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    out["level%d" % level] = {"x_realtime": S * n_per / 44100.0 / (ms / 1000.0), "ms_per_step": ms, "stage_ms": {k: round(v, 3) for k, v in tot.items()},
                              "bytes": int(b.output_total)}
    b.close()
print(json.dumps(out))
