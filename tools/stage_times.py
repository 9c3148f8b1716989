#!/usr/bin/env python3
"""Per-stage device time of one device-plane step (512 streams x 30 s by default), no parity check: for timing
experiments on kernel variants (MP3B_FB_DEBUG etc.).   usage: tools/stage_times.py [streams] [seconds] [steps] [channels]"""
import ctypes as C, importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
mp3 = importlib.import_module("swift-mp3_b200")
sharding = importlib.import_module("swift-mp3_b200.sharding")
L = mp3.lib()
S = int(sys.argv[1]) if len(sys.argv) > 1 else 512
secs = float(sys.argv[2]) if len(sys.argv) > 2 else 30.0
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
CHN = int(sys.argv[4]) if len(sys.argv) > 4 else 2
n_per = int(round(secs * 44100))
pcm = torch.empty((S, n_per * CHN), dtype=torch.float32, device="cuda")
for i in range(S):
    fl, fr, seed = sharding.stream_params(i)
    assert L.mp3b_synth_fill(0, pcm[i].data_ptr(), n_per, CHN, 44100, fl, fr, 0.5, 0.05, seed) == 0
ptrs = (C.c_void_p * S)(*[pcm[i].data_ptr() for i in range(S)])
ns = (C.c_size_t * S)(*([n_per * CHN] * S))
b = mp3.EncoderBatch(mp3.MP3EncoderOptions(mode=mp3.Mode.stereo if CHN == 2 else mp3.Mode.mono), S, 0)
if os.environ.get("MP3B_MATRIXING"):
    b.set_matrixing(int(os.environ["MP3B_MATRIXING"]))
tot = {}
for k in range(steps + 2):
    b.reset(); b.encode_device(ptrs, ns, flush=True, download=False)
    if k >= 2:
        for n, v in b.stage_ms().items(): tot[n] = tot.get(n, 0.0) + v / steps
print(os.environ.get("MP3B_FB_DEBUG", "0"), {k: round(v, 3) for k, v in tot.items()})
