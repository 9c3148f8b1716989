"""Debug helper (GPU box): encode the ISO-mode test cases at levels 1 and 2 and leave bytes + traces in gpurun_out/ for offline analysis."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
mp3 = importlib.import_module("swift-mp3_b200")
import signals as sg
cases = [("c1", sg.sine_noise(3.0), dict(sampleRate=44100, bitrateKbps=128, mode=mp3.Mode.stereo)),
         ("c2", sg.white(2.0), dict(sampleRate=48000, bitrateKbps=320, mode=mp3.Mode.mono))]
for name, pcm, o in cases:
    for level in (1, 2):
        opts = mp3.MP3EncoderOptions(**o)
        b = mp3.EncoderBatch(opts, 1, 0, 0)
        b.set_iso_mode(level)
        b.set_trace(spectrum=True, ix=True)
        out = b.encode([pcm], flush=True)[0]
        np.savez_compressed(os.path.join(ROOT, "gpurun_out", "iso_%s_l%d.npz" % (name, level)), mp3=np.frombuffer(out, np.uint8), pcm=pcm,
                            spec=b.trace_array(0, "spectrum"), ix=b.trace_array(0, "ix"), gc=b.trace_gc(0), fr=b.trace_frames(0))
        b.close()
print("ok")
