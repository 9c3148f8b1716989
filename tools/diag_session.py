import importlib, sys, time, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
mp3 = importlib.import_module("swift-mp3_b200")
import signals
L = mp3.lib()
for name, pcm, mo in (("c2", signals.white(60.0), mp3.MP3EncoderOptions(sampleRate=48000, bitrateKbps=320, mode=mp3.Mode.mono)),
                      ("c3", signals.castanets(30.0), mp3.MP3EncoderOptions(sampleRate=44100, bitrateKbps=128, mode=mp3.Mode.jointStereo, vbr=True, quality=2))):
    for rep in range(4):
        b = mp3.EncoderBatch(mo, 1, 0)
        t0 = time.perf_counter()
        out = b.encode([pcm], flush=True)
        t1 = time.perf_counter()
        st = b.stage_ms()
        s = mp3.EncoderSession(mo, 0)
        t2 = time.perf_counter(); o1 = s.encode(pcm); t3 = time.perf_counter(); o2 = s.flush(); t4 = time.perf_counter()
        print(name, rep, "batch.encode %.2f ms" % (1e3 * (t1 - t0)), {k: round(v, 2) for k, v in st.items()}, "passes", b.pass_count,
              "| session encode %.2f flush %.2f ms" % (1e3 * (t3 - t2), 1e3 * (t4 - t3)), len(o1), len(o2))
        b.close(); s.close()
