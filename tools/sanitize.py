#!/usr/bin/env python3
"""Small end-to-end runs of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck):
default mode (stereo CBR, joint-stereo VBR with short blocks, multi-pass), ISO mode levels 1 and 2, tensor-core matrixing."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
mp3 = importlib.import_module("swift-mp3_b200")
import signals as sg
cases = [("cbr", sg.sine_noise(0.6, seed=1), dict(), {}),
         ("vbr-joint", sg.castanets(0.6), dict(mode=mp3.Mode.jointStereo, vbr=True, quality=2), {}),
         ("mono-48k", sg.white(0.4), dict(sampleRate=48000, bitrateKbps=320, mode=mp3.Mode.mono), {}),
         ("iso1", sg.sine_noise(0.4, seed=2), dict(), dict(iso=1)),
         ("iso2", sg.sine_noise(0.4, seed=3), dict(), dict(iso=2)),
         ("iso2-joint", sg.castanets(0.4), dict(mode=mp3.Mode.jointStereo, vbr=True, quality=2), dict(iso=2)),
         ("tc", sg.sine_noise(0.6, seed=4), dict(), dict(tc=1)),
         ("tc-mono", sg.white(0.4), dict(sampleRate=48000, bitrateKbps=320, mode=mp3.Mode.mono), dict(tc=1))]
for name, pcm, o, x in cases:
    for fpp in (0, 5):
        b = mp3.EncoderBatch(mp3.MP3EncoderOptions(**o), 3, 0, fpp)
        if x.get("iso"): b.set_iso_mode(x["iso"])
        if x.get("tc"): b.set_matrixing(1)
        outs = b.encode([pcm, pcm[: len(pcm) // 3], pcm[7:]], flush=True)
        print(name, fpp, [len(v) for v in outs])
        b.close()
print("done")
