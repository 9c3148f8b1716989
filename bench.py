#!/usr/bin/env python3
"""bench.py — headline benchmark of the B200-native MP3 encode path.

Metric (BASELINE.json): encoded audio seconds per second (x realtime), whole job over all GPUs.
Workload (BASELINE config 4, strong scaling): the batch of 4096 independent 30 s 44.1 kHz stereo CBR 128 kbps streams,
sharded by stream over the N GPUs (4096 / N streams each; at N = 1 the whole batch, 43 GB of PCM, sits on one B200).  One
"step" = every stream of the shard through fresh EncoderSessions: encode(samples:) of the whole stream + flush().
`--streams K` switches to K streams per GPU (weak scaling; reduced runs for profiling).

  value : PCM already resident in HBM, MP3 frames left in HBM (device plane of the C ABI), timed with CUDA events on
          the engine's own stream, max over ranks.
  e2e   : the same step through the reference-facing call with HOST buffers (pinned): H2D of the PCM and D2H of the
          MP3 bytes inside the timed region, wall clock bracketed by synchronize + barrier, max over ranks.
  --impl reference : the CPU restatement of the reference (oracle/, the Swift original cannot be built here) on all
          host cores, on a bounded sample of the same workload (the same synthetic streams, from the generator's CPU twin).
  parity: EVERY stream of every rank's shard is compared byte for byte with the oracle (untimed; config.parity reads
          "4096/4096"), the e2e leg's bytes with the device leg's.
  other_configs: BASELINE configs 1, 2, 3 (single streams: session plane with host buffers, device plane, the oracle on one
          core) and 5 (1024 sessions x 1152-sample chunks: p50 / p99 per call) ride in the default line; `--workload cN`
          prints one of them as the line of its own.
"""
import argparse
import ctypes as C
import importlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

SR, CH, KBPS = 44100, 2, 128
TOTAL_STREAMS = 4096                               # BASELINE config 4
# Algorithmic bytes / flops per granule-channel (gc = 576 samples of one channel), DESIGN.md section 4
ALGO_BYTES = {"prepass": 2304 + 8,                 # PCM in, decisions out
              "filterbank": 2304 + 2304,           # PCM in, 18 x 32 subband samples out
              "granule": 2304 + 2304 + 84,         # subband samples in, sign * |x|^0.75 out, curve tables out
              "scan": 84 + 45}                     # curve tables in, records out ("pack" / "frames" depend on the output size)
FLOP_FILTERBANK = 18 * (512 + 448 + 2 * 2048)      # window products + sums + 32 x 64 matrixing, direct form as executed
FLOP_GRANULE = 32 * (36 + 2 * 648 + 18 * 3) + 8 * 31 * 6 + 576 * 4   # long-block MDCT + scaling, alias butterflies, energies
FP32_MEASURED_FRACTION = 0.905                     # tools/microbench/mb.cu, kernel C (FFMA2 from registers only), three runs
KERNEL_OF = {"prepass": "k_prepass", "filterbank": "k_filterbank (polyphase analysis: windowing + 32x64 matrixing)",
             "granule": "k_granule (MDCT + alias reduction + |x|^0.75 + bits-vs-gain curve)", "scan": "k_scan",
             "pack": "k_pack", "frames": "k_frames"}


sharding = importlib.import_module("swift-mp3_b200.sharding")
stream_params = sharding.stream_params      # BASELINE C4 recipe: seed 1000+i, f_L = 110 * 2^((i mod 48)/12), f_R = 1.26 f_L


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons, one row every 50 ms for the whole run (started before the warm-up: nvidia-smi
    needs about a second to come up); window(t0, t1) summarises the rows that arrived inside a timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.proc = index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))
        except Exception:
            pass

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        self.join(timeout=2)

    def window(self, t0, t1):
        rows = [r for t, r in self.rows if t0 <= t <= t1]
        note = "sampled inside the timed region"
        if not rows and self.rows:                 # region shorter than the sampling period: nearest sample
            rows = [min(self.rows, key=lambda tr: abs(tr[0] - 0.5 * (t0 + t1)))[1]]
            note = "timed region shorter than the 50 ms sampling period: nearest sample"
        sm = sorted(int(r[0]) for r in rows if r and r[0].isdigit())
        mx = max([int(r[1]) for r in rows if len(r) > 1 and r[1].isdigit()] or [0])
        pw = [float(r[2]) for r in rows if len(r) > 2 and r[2].replace(".", "", 1).isdigit()]
        reasons = set()
        for r in rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm), "power_w": max(pw) if pw else None, "note": note}


def c4_stream(i, seconds):
    """Stream i of BASELINE config 4 from the CPU twin of the device generator (bit-identical to mp3b_synth_fill)."""
    import oracle_binding as orc
    fl, fr, seed = stream_params(i)
    return orc.synth_fill(int(round(seconds * SR)), CH, SR, fl, fr, 0.5, 0.05, seed)


def cpu_reference(seconds_per_stream, min_wall=10.0, max_wall=40.0):
    """The CPU restatement of the reference, built -O3 -march=native on this host (oracle/Makefile `native`), on every host
    core and on one core: a bounded sample of the C4 workload (streams 0, 1, 2, ... of the same recipe the GPU arm encodes)."""
    import oracle_binding as orc
    cores = os.cpu_count() or 1
    n = max(2 * cores, 8)
    pcms = [c4_stream(i, seconds_per_stream) for i in range(n)]
    opts = dict(sample_rate=SR, bitrate_kbps=KBPS, mode="stereo")
    _, d_native = orc.encode_streams(pcms[:cores], cores, native=True, **opts)            # warm-up
    _, d_check = orc.encode_streams(pcms[:cores], cores, **opts)
    assert d_native == d_check, "the -march=native build of the oracle disagrees with the checker build"
    t1 = time.perf_counter(); k1 = 0
    while time.perf_counter() - t1 < 3.0:                                                  # one session on one core
        orc.encode_streams(pcms[k1 % n:k1 % n + 1], 1, native=True, **opts); k1 += 1
    one_core = k1 * seconds_per_stream / (time.perf_counter() - t1)
    done, t0 = 0, time.perf_counter()
    while True:
        orc.encode_streams(pcms, cores, native=True, **opts)
        done += n
        el = time.perf_counter() - t0
        if el >= min_wall or el >= max_wall:
            break
    audio = done * seconds_per_stream
    return {"value": audio / el, "unit": "x realtime (audio s / s)", "cores": cores, "kind": "port", "one_core": one_core,
            "build": "gcc -O3 -march=native -ffp-contract=off (digest-equal to the -O2 x86-64-v3 checker build)",
            "sample": "%d streams x %.0f s of the C4 recipe (44.1 kHz stereo CBR128; the GPU arm's streams 0..%d from the generator's "
                      "CPU twin), one oracle session per stream, %d threads, %.1f s wall; one_core = one session at a time on one thread"
                      % (done, seconds_per_stream, n - 1, cores, el)}, el, audio


def check_all_streams(mp3, L, b, pcm, S, local):
    """Untimed parity of a whole shard: every stream's bytes (the batch's downloaded output) against an oracle session fed the
    same PCM (downloaded from the device in slices).  Returns (streams checked, list of (stream, first differing byte))."""
    import oracle_binding as orc
    bad = []
    cores = os.cpu_count() or 1
    for lo in range(0, S, 128):
        hi = min(S, lo + 128)
        host = pcm[lo:hi].cpu().numpy()
        ep, en = [], []
        for i in range(lo, hi):
            p_, n_ = C.c_void_p(), C.c_size_t(0)
            assert L.mp3b_batch_output(b._h, i, C.byref(p_), C.byref(n_)) == 0, L.mp3b_last_error()
            ep.append(p_.value or host.ctypes.data); en.append(n_.value)
        r = orc.compare_streams_raw([host[i].ctypes.data for i in range(hi - lo)], [host.shape[1]] * (hi - lo), ep, en, cores,
                                    sample_rate=SR, bitrate_kbps=KBPS, mode="stereo")
        bad += [(lo + i, off) for i, off in r]
    return S, bad


def stream_digests(L, b, S):
    import hashlib
    out = []
    for i in range(S):
        p_, n_ = C.c_void_p(), C.c_size_t(0)
        assert L.mp3b_batch_output(b._h, i, C.byref(p_), C.byref(n_)) == 0
        out.append(hashlib.blake2b(C.string_at(p_.value, n_.value) if n_.value else b"", digest_size=16).digest())
    return out


def single_stream_config(name, mp3, L, local, repeats=5):
    """BASELINE configs 1-3: ONE stream ("replicas only": no sharding).  Session plane with host buffers (the reference-facing
    call: encode(samples:) of the whole stream + flush()), device plane (PCM in HBM), and the oracle on one core."""
    import numpy as np
    import torch
    import oracle_binding as orc
    import signals
    if name == "c1":
        pcm, o, label = orc.synth_fill(10 * 44100, 2, 44100, 440.0, 554.37, 0.5, 0.05, 1234), dict(sample_rate=44100, bitrate_kbps=128, mode="stereo"), \
            "C1: 44.1 kHz stereo CBR 128, 10 s sine + noise, single stream"
    elif name == "c2":
        pcm, o, label = signals.white(60.0), dict(sample_rate=48000, bitrate_kbps=320, mode="mono"), "C2: 48 kHz mono CBR 320, 60 s white noise, single stream"
    elif name == "c2pink":
        pcm, o, label = signals.pink(60.0), dict(sample_rate=48000, bitrate_kbps=320, mode="mono"), "C2: 48 kHz mono CBR 320, 60 s pink noise, single stream"
    else:
        pcm, o, label = signals.castanets(30.0), dict(sample_rate=44100, bitrate_kbps=128, mode="jointStereo", vbr=True, quality=2), \
            "C3: 44.1 kHz joint-stereo VBR q2, 30 s castanet-like transients, single stream"
    ch = 1 if o["mode"] == "mono" else 2
    seconds = pcm.size / ch / o["sample_rate"]
    mo = mp3.MP3EncoderOptions(sampleRate=o["sample_rate"], bitrateKbps=o["bitrate_kbps"], vbr=o.get("vbr", False),
                               mode={"mono": 0, "stereo": 1, "jointStereo": 2}[o["mode"]], quality=o.get("quality", 5))
    ref, rs = orc.encode_all(pcm, **o)
    # session plane, host buffers
    ts = []
    for _ in range(repeats + 2):
        s = mp3.EncoderSession(mo, local)
        t0 = time.perf_counter()
        out = s.encode(pcm) + s.flush()
        ts.append(time.perf_counter() - t0)
        s.close()
        assert out == ref, name + ": session output differs from the oracle"
    t_sess = sorted(ts[2:])[len(ts[2:]) // 2]
    # device plane
    dev = torch.from_numpy(pcm).cuda()
    b = mp3.EncoderBatch(mo, 1, local)
    ptrs, ns = (C.c_void_p * 1)(dev.data_ptr()), (C.c_size_t * 1)(pcm.size)
    td = []
    for _ in range(repeats + 2):
        b.reset()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        b.encode_device(ptrs, ns, flush=True, download=True)
        td.append(time.perf_counter() - t0)
    assert b.output(0) == ref, name + ": device-plane output differs from the oracle"
    t_dev = sorted(td[2:])[len(td[2:]) // 2]
    launches, passes = b.launch_count, b.pass_count
    b.close()
    # the oracle on one core (native build)
    t0 = time.perf_counter(); k = 0
    while time.perf_counter() - t0 < 2.0:
        orc.encode_streams([pcm], 1, native=True, **o); k += 1
    t_cpu = (time.perf_counter() - t0) / k
    gc = rs.frame_count * 2 * ch
    return {"workload": label, "audio_seconds": seconds, "frames": int(rs.frame_count), "granule_channels": int(gc),
            "value": seconds / t_dev, "unit": "x realtime", "ms_device_plane": 1e3 * t_dev,
            "e2e": {"value": seconds / t_sess, "unit": "x realtime", "ms": 1e3 * t_sess, "h2d_bytes": int(pcm.nbytes), "d2h_bytes": len(ref),
                    "call": "mp3b_session_encode(whole stream) + mp3b_session_flush, host buffers"},
            "cpu_one_core": {"value": seconds / t_cpu, "unit": "x realtime", "ms": 1e3 * t_cpu, "kind": "port", "cores": 1},
            "gpu_launches": int(launches), "passes": int(passes), "scaling": "replicas only (one stream)",
            "parity": "byte-identical to the oracle (session plane and device plane)"}


def bench_c5(steps, mp3, L, local):
    """BASELINE config 5: 1024 concurrent sessions (C4 streams 0...1023) fed 1152-sample stereo chunks through
    encode(samples:) — one mp3b_batch_encode_strided call per chunk with a pinned host arena in, MP3 bytes out; p50 / p99
    latency of a call (= of every frame in it).  EVERY session's bytes are compared with an oracle session fed the same chunks."""
    import numpy as np
    import torch
    import oracle_binding as orc
    S, steps = 1024, max(steps, 200)
    chunk = 1152 * CH
    opts = mp3.MP3EncoderOptions(sampleRate=SR, bitrateKbps=KBPS, mode=mp3.Mode.stereo)
    b = mp3.EncoderBatch(opts, S, local, 8)
    n_chunks = 64                                                   # distinct chunks per stream, cycled
    pcm = torch.empty((S, n_chunks * chunk), dtype=torch.float32, device="cuda")
    for i in range(S):
        fl, fr, seed = stream_params(i)
        assert L.mp3b_synth_fill(local, pcm[i].data_ptr(), n_chunks * 1152, CH, SR, fl, fr, 0.5, 0.05, seed) == 0
    host = pcm.cpu().numpy()
    hp = C.c_void_p()
    assert L.mp3b_host_alloc(S * chunk * 4, C.byref(hp)) == 0, L.mp3b_last_error()
    arena = np.ctypeslib.as_array(C.cast(hp, C.POINTER(C.c_float)), shape=(S, chunk))
    ns = (C.c_size_t * S)(*([chunk] * S))
    lat, got = [], [bytearray() for _ in range(S)]
    total = steps + 20
    for k in range(total):
        arena[:] = host[:, (k % n_chunks) * chunk:(k % n_chunks + 1) * chunk]     # the caller's side: chunks arrive in its own buffers
        t0 = time.perf_counter()
        b.encode_strided(hp.value, chunk, ns, flush=False)
        lat.append(time.perf_counter() - t0)
        for i in range(S):
            got[i] += b.output(i)
    stage, launches = b.stage_ms(), b.launch_count
    lat = np.array(lat[20:]) * 1e3
    # parity: all sessions; the oracle is fed the same chunk sequence (the 64 chunks cycled)
    reps = -(-total // n_chunks)
    seq = [np.ascontiguousarray(np.tile(host[i], reps)[: total * chunk]) for i in range(S)]
    refs = [bytes(g) for g in got]
    cores = os.cpu_count() or 1
    # no flush was issued: compare against encode() only by appending the oracle's flush-free prefix -> use sessions directly
    bad = 0
    ep = [np.frombuffer(r, np.uint8) if r else np.zeros(1, np.uint8) for r in refs]
    diff = orc.compare_streams_raw([x.ctypes.data for x in seq], [x.size for x in seq], [e.ctypes.data for e in ep], [len(r) for r in refs],
                                   cores, chunk_floats=chunk, sample_rate=SR, bitrate_kbps=KBPS, mode="stereo")
    # the oracle driver flushes at the end (one more frame than the un-flushed batch emitted): a session is identical when the
    # first difference is exactly at the end of the batch's bytes
    bad = [(i, off) for i, off in diff if off != len(refs[i])]
    assert not bad, "streaming output differs from the oracle: %r" % bad[:4]
    assert len(diff) == S, "the oracle's flush frame is missing for some sessions"
    p50, p99 = float(np.percentile(lat, 50)), float(np.percentile(lat, 99))
    res = {"workload": "C5: 1024 sessions x 1152-sample stereo chunks, host arena in (mp3b_batch_encode_strided), bytes out",
           "metric": "per-frame latency, 1024 concurrent sessions (config 5)", "value": p50, "unit": "ms (p50)",
           "p99_ms": p99, "mean_ms": float(lat.mean()), "steps": steps, "higher_is_better": False,
           "frame_period_ms": 1152 / SR * 1e3, "realtime_factor_p50": S * 1152 / SR * 1e3 / p50,
           "stage_ms_last_call": stage, "gpu_launches_per_call": launches,
           "parity": "%d/%d sessions byte-identical to the oracle over %d chunks" % (S, S, total)}
    L.mp3b_host_free(hp)
    b.close()
    return res


def bench_multi(a, mp3, L, metric):
    """BASELINE config 4 through the multi-device batch plane of the C ABI (mp3b_batch_create_multi) from ONE process — what a
    Swift host would call: 4096 streams partitioned by stream over `--gpus` devices, one host thread per device inside the
    library.  value = device plane (each stream's PCM on its own device), e2e = pinned host buffers in, bytes out; wall clock
    around the call with every device synchronised.  All streams are checked against the oracle."""
    import numpy as np
    import torch
    import oracle_binding as orc
    G = a.gpus
    S = a.streams if a.streams > 0 else TOTAL_STREAMS
    n_per = int(round(a.seconds * SR)); n_floats = n_per * CH
    opts = mp3.MP3EncoderOptions(sampleRate=SR, bitrateKbps=KBPS, mode=mp3.Mode.stereo)
    b = mp3.EncoderBatch(opts, S, devices=list(range(G)))
    dev_of = [b.stream_device(i) for i in range(S)]
    pcm = [torch.empty((dev_of.count(d), n_floats), dtype=torch.float32, device="cuda:%d" % d) for d in range(G)]
    row, seen = [], [0] * G
    for i in range(S):
        d = dev_of[i]; t = pcm[d][seen[d]]; seen[d] += 1
        fl, fr, seed = stream_params(i)
        assert L.mp3b_synth_fill(d, t.data_ptr(), n_per, CH, SR, fl, fr, 0.5, 0.05, seed) == 0, L.mp3b_last_error()
        row.append(t)
    sync = lambda: [L.mp3b_device_sync(d) for d in range(G)]
    sync()
    dptrs = (C.c_void_p * S)(*[t.data_ptr() for t in row]); ns = (C.c_size_t * S)(*([n_floats] * S))
    # parity of every stream (device plane, downloaded)
    b.encode_device(dptrs, ns, flush=True, download=True)
    bad, cores = [], os.cpu_count() or 1
    for lo in range(0, S, 128):
        hi = min(S, lo + 128)
        host = [row[i].cpu().numpy() for i in range(lo, hi)]
        outs = [b.output(i) for i in range(lo, hi)]
        bad += [(lo + i, off) for i, off in orc.compare_streams(host, outs, cores, sample_rate=SR, bitrate_kbps=KBPS, mode="stereo")]
    assert not bad, "multi-device batch differs from the oracle: %r" % bad[:4]
    digests = stream_digests(L, b, S)

    def timed(fn, steps, warm):
        for _ in range(warm):
            b.reset(); fn()
        sync(); t0 = time.perf_counter()
        for _ in range(steps):
            b.reset(); fn()
        sync()
        return (time.perf_counter() - t0) / steps
    t_dev = timed(lambda: b.encode_device(dptrs, ns, flush=True, download=False), a.steps, a.warmup)
    stage_dev, launches = b.stage_ms(), b.launch_count
    hp = C.c_void_p()
    assert L.mp3b_host_alloc(S * n_floats * 4, C.byref(hp)) == 0, L.mp3b_last_error()
    for i in range(S):
        assert L.mp3b_device_copy(dev_of[i], hp.value + i * n_floats * 4, row[i].data_ptr(), n_floats * 4, 1) == 0
    hptrs = (C.c_void_p * S)(*[hp.value + i * n_floats * 4 for i in range(S)])
    e2e_steps = a.e2e_steps or min(a.steps, 5)
    t_e2e = timed(lambda: b.encode_ptrs(hptrs, ns, flush=True), e2e_steps, 1)
    assert stream_digests(L, b, S) == digests, "e2e output differs from the oracle-checked device-plane output"
    audio = S * a.seconds
    emit({"metric": metric, "value": audio / t_dev, "unit": "x realtime", "n_gpus": G, "steps": a.steps, "warmup": a.warmup,
          "ms_per_step": 1e3 * t_dev, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
          "config": {"workload": "C4: %d streams x %.0f s partitioned by stream over %d GPU(s) by mp3b_batch_create_multi, ONE process "
                                 "(one host thread per device inside the library)" % (S, a.seconds, G),
                     "parity": "%d/%d streams byte-identical to the oracle" % (S, S), "timing": "wall clock around the C-ABI call, all devices synchronised"},
          "e2e": {"value": audio / t_e2e, "unit": "x realtime", "ms_per_step": 1e3 * t_e2e, "steps": e2e_steps,
                  "h2d_bytes_per_step": S * n_floats * 4, "d2h_bytes_per_step": int(b.output_total),
                  "h2d_GBps_aggregate": S * n_floats * 4 / t_e2e / 1e9},
          "gpu_launches": int(launches) * a.steps, "stage_ms_device_plane_slowest_device": stage_dev})
    L.mp3b_host_free(hp)
    b.close()
    return 0


_REAL_STDOUT = None


def quiet_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version line from C), so file
    descriptor 1 points at stderr while the bench runs and emit() puts it back for the one line."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    if _REAL_STDOUT is not None:
        os.dup2(_REAL_STDOUT, 1)
    print(json.dumps(line))
    sys.stdout.flush()
    if _REAL_STDOUT is not None:
        os.dup2(2, 1)                  # whatever a library still prints while shutting down goes to stderr again


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native")
    ap.add_argument("--streams", type=int, default=0,
                    help="streams per GPU; 0 (default) = BASELINE config 4: 4096 streams in total, sharded over the GPUs (strong scaling)")
    ap.add_argument("--seconds", type=float, default=30.0)
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = min(steps, 5)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-tc", action="store_true", help="skip the tensor-core matrixing leg (reported beside value)")
    ap.add_argument("--i16", action="store_true", help="also time the 16-bit-input extension end to end (reported as e2e_i16, not the headline)")
    ap.add_argument("--workload", default="c4", choices=["c4", "c1", "c2", "c3", "c5"],
                    help="c4 (default, the headline): batch of 30 s streams; c1 / c2 / c3: the single-stream configs; "
                         "c5: streaming latency, 1024 sessions x 1152-sample chunks")
    ap.add_argument("--single-process", action="store_true", help="drive all --gpus devices from this one process through "
                    "mp3b_batch_create_multi (no torchrun): the multi-device plane of the C ABI")
    ap.add_argument("--no-others", action="store_true", help="skip the other_configs legs (C1, C2, C3, C5) of the default line")
    ap.add_argument("--parity", default="all", choices=["all", "spot"], help="all (default): every stream of the shard against "
                    "the oracle; spot: two streams (profiling runs)")
    a = ap.parse_args()
    quiet_stdout()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    metric = "encoded audio sec/sec (x realtime)"
    strong = a.streams <= 0
    if strong:
        a.streams = TOTAL_STREAMS // world
    workload = ("C4: batch of %d independent %.0f s 44.1 kHz stereo CBR 128 kbps streams sharded by stream over %d GPU(s), %d per "
                "GPU; inputs %.1f GB per GPU > L2" % (a.streams * world, a.seconds, world, a.streams, a.streams * a.seconds * SR * CH * 4 / 1e9))

    if a.impl == "reference":
        if rank != 0:
            return 0
        res, el, audio = cpu_reference(a.seconds, min_wall=max(5.0, 2.0 * a.steps), max_wall=120.0)
        line = {"impl": "reference", "metric": metric, "value": res["value"], "unit": "x realtime", "n_gpus": a.gpus,
                "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1000.0 * el / max(a.steps, 1), "higher_is_better": True,
                "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload, "reference": "C restatement of SwiftMP3 (oracle/, built -O3 -march=native on this host); "
                           "the Swift + Accelerate original cannot be built on Linux", "one_core_x_realtime": res["one_core"]},
                "cpu_baseline": res,
                "e2e": {"value": res["value"], "unit": "x realtime", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        emit(line)
        return 0

    import numpy as np
    import torch
    mp3 = importlib.import_module("swift-mp3_b200")
    L = mp3.lib()
    if a.single_process:
        assert world == 1, "--single-process is not run under torchrun"
        return bench_multi(a, mp3, L, metric)
    torch.cuda.set_device(local)
    numa = sharding.bind_near_gpu(local)           # before any pinned allocation
    sampler = ClockSampler(local); sampler.start()
    dist = None
    if world > 1:
        # NCCL's log (rank / topology lines the driver checks) goes to stderr: file descriptor 1 points there while the bench
        # runs (quiet_stdout), so stdout stays the one JSON line
        os.environ.setdefault("NCCL_DEBUG", os.environ.get("MP3B_NCCL_DEBUG", "INFO"))
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        return sharding.max_over_ranks(v, dist, "cuda")

    def sum_over_ranks(v):
        return sharding.sum_over_ranks(v, dist, "cuda")

    if a.workload != "c4":                                          # single-GPU configs: rank 0 runs them, "replicas only"
        if rank == 0:
            if a.workload == "c5":
                res = bench_c5(a.steps, mp3, L, local)
            else:
                res = single_stream_config(a.workload, mp3, L, local)
                if a.workload == "c2":
                    res["pink"] = single_stream_config("c2pink", mp3, L, local)
            line = dict(res)
            line.setdefault("metric", metric)
            line.update({"n_gpus": 1, "warmup": a.warmup, "dtype": "f32", "data": "synthetic", "vs_baseline": None,
                         "config": {"workload": res["workload"]}})
            line.setdefault("steps", a.steps); line.setdefault("higher_is_better", True)
            emit(line)
        sampler.stop()
        if dist is not None:
            dist.barrier(); dist.destroy_process_group()
        return 0
    S = a.streams                                                  # streams of this rank
    shard_lo, shard_hi = sharding.shard_range(S * world, rank, world)
    assert shard_hi - shard_lo == S
    n_per = int(round(a.seconds * SR))
    n_floats = n_per * CH
    pcm = torch.empty((S, n_floats), dtype=torch.float32, device="cuda")
    for i in range(S):
        fl, fr, seed = stream_params(shard_lo + i)
        rc = L.mp3b_synth_fill(local, pcm[i].data_ptr(), n_per, CH, SR, fl, fr, 0.5, 0.05, seed)
        assert rc == 0, L.mp3b_last_error()
    torch.cuda.synchronize()
    dptrs = (C.c_void_p * S)(*[pcm[i].data_ptr() for i in range(S)])
    ns = (C.c_size_t * S)(*([n_floats] * S))
    opts = mp3.MP3EncoderOptions(sampleRate=SR, bitrateKbps=KBPS, mode=mp3.Mode.stereo)
    b = mp3.EncoderBatch(opts, S, local)
    frames_per_pass = b.frames_per_pass
    ext = torch.cuda.ExternalStream(b.cuda_stream, device=torch.device("cuda", local))
    audio_per_step = S * a.seconds

    # ---- parity (untimed): EVERY stream of this rank's shard against the CPU oracle, on every rank; the inputs are the CPU
    # twin's bit for bit, so the reference arm encodes the same streams
    import oracle_binding as orc
    b.reset()
    b.encode_device(dptrs, ns, flush=True, download=True)
    t_par = time.perf_counter()
    if a.parity == "all":
        checked, bad = check_all_streams(mp3, L, b, pcm, S, local)
    else:
        checked, bad = 0, []
        for i in (0, S - 1):
            ref, _ = orc.encode_all(pcm[i].cpu().numpy(), sample_rate=SR, bitrate_kbps=KBPS, mode="stereo")
            checked += 1
            if b.output(i) != ref:
                bad.append((i, -1))
    t_par = time.perf_counter() - t_par
    for i in (0, S // 2, S - 1):
        twin = c4_stream(shard_lo + i, a.seconds)
        assert np.array_equal(pcm[i].cpu().numpy().view("<u4"), twin.view("<u4")), "device PCM differs from the CPU twin (stream %d)" % (shard_lo + i)
    assert not bad, "rank %d: %d stream(s) differ from the oracle, first %r" % (rank, len(bad), bad[:4])
    device_digests = stream_digests(L, b, S)
    checked_total = int(sum_over_ranks(checked))
    parity = "%d/%d streams byte-identical to the oracle (every rank checks its own shard; %.0f s on rank 0); inputs bit-identical " \
             "to the generator's CPU twin" % (checked_total, S * world, t_par)

    # ---- value: device plane, K steps
    def step_device():
        b.reset()
        b.encode_device(dptrs, ns, flush=True, download=False)

    for _ in range(a.warmup):
        step_device()
    barrier()
    t_dev0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ext)
    stages = {k: 0.0 for k in mp3.STAGES}
    launches = passes = 0
    for _ in range(a.steps):
        step_device()
        for k, v in b.stage_ms().items():
            stages[k] += v
        launches += b.launch_count; passes += b.pass_count
    e1.record(ext)
    barrier()
    t_dev1 = time.perf_counter()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    ms_per_step = ms_total / a.steps
    value = world * audio_per_step / (ms_per_step / 1000.0)
    out_bytes = b.output_total

    # ---- roofline: per-stage table from the engine's CUDA events (recorded on its own stream around every kernel), the
    # dominant kernel = the stage with the largest share of the step
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak, which = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json)") if "hbm_gbs" in peaks else (6650.0, "fallback")
    frames = (n_per + 1151) // 1152
    gc_per_step = S * frames * 2 * CH
    gc_per_launch = gc_per_step * a.steps / max(passes, 1)
    out_per_gc = out_bytes / max(gc_per_step, 1)
    algo = dict(ALGO_BYTES, pack=2304 + out_per_gc, frames=2 * out_per_gc + 30)
    fp32_peak = 148 * 128 * 2 * (peaks.get("sm_max_mhz", 1965.0) / 1000.0) / 1000.0   # nominal TFLOP/s, non-tensor FP32
    flops = {"filterbank": FLOP_FILTERBANK, "granule": FLOP_GRANULE}
    ncu, ncu_file = {}, None
    for name in ("r02_kernel_traffic.json", "r01_kernel_traffic.json"):       # the newest committed ncu --set full summary
        try:
            ncu = json.load(open(os.path.join(ROOT, "profiles", name)))
            ncu_file = name
            break
        except Exception:
            pass
    stage_table = {}
    fused_prepass = stages["prepass"] / max(passes, 1) < 0.02     # CBR without joint stereo: k_prepass is not launched at all
    if fused_prepass:
        algo["granule"] += 2304                                   # k_granule then reads the granule's PCM itself (block-type decision)
    sm_clock_hz = peaks.get("sm_max_mhz", 1965.0) * 1e6
    issue_peak = 148 * 4 * sm_clock_hz / 1e9                      # G warp-instructions / s: one per scheduler per cycle
    for k in algo:
        if k == "prepass" and fused_prepass:
            stage_table[k] = {"ms_per_step": 0.0, "share": 0.0, "skipped": "CBR without joint stereo: the block types are decided inside k_granule, "
                                                                            "k_prepass (energies for VBR / mid-side / the trace plane) is not launched"}
            continue
        sec = stages[k] / a.steps / 1000.0
        row = {"ms_per_step": 1000.0 * sec, "share": stages[k] / max(stages["total"], 1e-9), "algo_bytes_per_gc": algo[k],
               "achieved_GBps": gc_per_step * algo[k] / sec / 1e9, "hbm_frac": gc_per_step * algo[k] / sec / 1e9 / hbm_peak}
        if k in flops:
            row["fp32_TFLOPs"] = gc_per_step * flops[k] / sec / 1e12
            row["fp32_frac_of_nominal"] = row["fp32_TFLOPs"] / fp32_peak
        if k in ncu:
            row["ncu"] = ncu[k]
            if ncu[k].get("inst_executed_per_gc"):
                row["issue_Ginst_per_s"] = gc_per_step * ncu[k]["inst_executed_per_gc"] / sec / 1e9
                row["issue_frac"] = row["issue_Ginst_per_s"] / issue_peak
        stage_table[k] = row
    dom = max(algo, key=lambda k: stages[k])
    dom_ms_per_launch = stages[dom] / max(passes, 1)
    dom_s = dom_ms_per_launch / 1000.0
    hbm_achieved = gc_per_launch * algo[dom] / dom_s / 1e9
    traffic = ncu[dom]["dram_bytes_per_gc"] * gc_per_launch if dom in ncu and "dram_bytes_per_gc" in ncu[dom] else None
    # The bound is the limiter ncu shows for the kernel, not an assumption: k_filterbank is bound by the FP32 FMA pipe (direct-form
    # matrixing kept operation for operation), k_granule by instruction issue (MDCT immediates + FP64 |x|^0.75 + integer bit
    # counting), the small kernels by latency.  The HBM view (algorithmic bytes / time vs the measured copy bandwidth) rides along.
    hbm_view = {"achieved": hbm_achieved, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_achieved / hbm_peak, "traffic": traffic,
                "peak_source": which, "algo_bytes_per_gc": algo[dom]}
    if dom in flops and (dom == "filterbank" or not (dom in ncu and ncu[dom].get("inst_executed_per_gc"))):
        ach = gc_per_launch * flops[dom] / dom_s / 1e12
        roofline = {"bound": "fp32", "achieved": ach, "peak": fp32_peak, "unit": "TFLOP/s", "frac": ach / fp32_peak,
                    "peak_source": "nominal non-tensor FP32: 148 SM x 128 lanes x 2 x sm_max_mhz (MEASURED_PEAKS.json holds no FP32 figure)",
                    "peak_measured": FP32_MEASURED_FRACTION * fp32_peak, "frac_of_measured": ach / (FP32_MEASURED_FRACTION * fp32_peak),
                    "peak_measured_source": "tools/microbench/mb.cu kernel C on a B200 of this pool (profiles/r02_microbench.txt): a register-only "
                                            "FFMA2 stream reaches 90.5-91.0 % of the nominal issue rate",
                    "flop_per_gc": flops[dom]}
    elif dom in ncu and ncu[dom].get("inst_executed_per_gc"):
        ach = gc_per_launch * ncu[dom]["inst_executed_per_gc"] / dom_s / 1e9
        roofline = {"bound": "issue", "achieved": ach, "peak": issue_peak, "unit": "Ginst/s (warp instructions)", "frac": ach / issue_peak,
                    "peak_source": "148 SM x 4 schedulers x 1 warp instruction per cycle x sm_max_mhz",
                    "inst_per_gc": ncu[dom]["inst_executed_per_gc"], "inst_source": "smsp__inst_executed.sum, profiles/" + ncu_file}
    else:
        roofline = dict(hbm_view, bound="hbm")
    roofline.update({"kernel": KERNEL_OF[dom], "traffic": traffic, "hbm": hbm_view, "ms_per_launch": dom_ms_per_launch,
                     "gc_per_launch": gc_per_launch, "share_of_step": stages[dom] / max(stages["total"], 1e-9),
                     "limiter": ("FP32 FMA pipe and shared-memory wavefronts, not HBM: the reference's direct-form 32x64 matrixing is kept "
                                 "operation for operation (bit-exact parity); see DESIGN.md section 4") if dom == "filterbank" else
                                "instruction issue (MDCT with immediate coefficients, FP64 |x|^0.75, integer bit counting); see DESIGN.md section 4",
                     "stages": stage_table, "stage_ms_per_step": {k: v / a.steps for k, v in stages.items()}})

    # ---- beside `value`, never instead of it: the same device-plane step with the filterbank's matrixing on the tensor cores
    # (tcgen05, three-term TF32 split; opt-in because its subband samples differ from the FP32 path in the last bits)
    tc = None
    if not a.no_tc:
        b.set_matrixing(1)
        b.reset(); b.encode_device(dptrs, ns, flush=True, download=True)
        tc_digests = stream_digests(L, b, S)
        for _ in range(2):
            step_device()
        barrier()
        t0e, t1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0e.record(ext)
        tc_stages = {k: 0.0 for k in mp3.STAGES}
        tc_steps = max(2, min(a.steps, 5))
        for _ in range(tc_steps):
            step_device()
            for k, v in b.stage_ms().items():
                tc_stages[k] += v
        t1e.record(ext)
        barrier()
        tc_ms = max_over_ranks(t0e.elapsed_time(t1e)) / tc_steps
        same = int(sum_over_ranks(sum(1 for x, y in zip(tc_digests, device_digests) if x == y)))
        tc = {"value": world * audio_per_step / (tc_ms / 1000.0), "unit": "x realtime", "ms_per_step": tc_ms, "steps": tc_steps,
              "stage_ms_per_step": {k: v / tc_steps for k, v in tc_stages.items()},
              "filterbank_tensor_TFLOPs": gc_per_step * 6 * 73728 / (tc_stages["filterbank"] / tc_steps / 1000.0) / 1e12,
              "streams_byte_identical_to_fp32_path": "%d/%d" % (same, S * world),
              "note": "mp3b_batch_set_matrixing(b, 1): k_filterbank_tc (tcgen05.mma kind::tf32, 6 products of a 3-term split, TMEM accumulators); "
                      "tier 1 (1e-5) holds with a 10x margin, see tests/test_gpu_parity.py::test_tensor_core_matrixing"}
        b.set_matrixing(0)

    # ---- e2e: host plane (pinned PCM in, MP3 bytes out), wall clock
    e2e_steps = a.e2e_steps or min(a.steps, 5)
    hp = C.c_void_p()
    assert L.mp3b_host_alloc(S * n_floats * 4, C.byref(hp)) == 0, L.mp3b_last_error()
    assert L.mp3b_device_copy(local, hp, pcm.data_ptr(), S * n_floats * 4, 1) == 0
    hptrs = (C.c_void_p * S)(*[hp.value + i * n_floats * 4 for i in range(S)])

    def step_host():
        b.reset()
        b.encode_ptrs(hptrs, ns, flush=True)

    for _ in range(max(1, min(a.warmup, 2))):
        step_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_host()
    barrier()
    t_e2e1 = time.perf_counter()
    e2e_s = max_over_ranks(t_e2e1 - t0) / e2e_steps
    sampler.stop()
    clocks = sampler.window(t_dev0, t_dev1)
    clocks["e2e_region"] = sampler.window(t0, t_e2e1)
    e2e_stage = b.stage_ms()
    e2e = {"value": world * audio_per_step / e2e_s, "unit": "x realtime", "h2d_bytes_per_step": S * n_floats * 4,
           "d2h_bytes_per_step": int(b.output_total), "ms_per_step": 1000.0 * e2e_s, "steps": e2e_steps,
           "stage_ms_last_step": e2e_stage}
    assert stream_digests(L, b, S) == device_digests, "e2e output differs from the (oracle-checked) device-plane output"
    e2e["parity"] = "all %d streams of every rank equal to the device-plane bytes checked against the oracle" % S
    L.mp3b_host_free(hp)

    # ---- optional: the same end-to-end step from 16-bit PCM (mp3b_batch_encode_i16: half the PCIe bytes); an extension of
    # the reference API, so it is reported beside e2e, never instead of it
    e2e_i16 = None
    if a.i16:
        pcm16 = (pcm * 32767.0).round().to(torch.int16)
        hp16 = C.c_void_p()
        assert L.mp3b_host_alloc(S * n_floats * 2, C.byref(hp16)) == 0, L.mp3b_last_error()
        assert L.mp3b_device_copy(local, hp16, pcm16.data_ptr(), S * n_floats * 2, 1) == 0
        hptrs16 = (C.c_void_p * S)(*[hp16.value + i * n_floats * 2 for i in range(S)])

        def step_host16():
            b.reset()
            b.encode_i16_ptrs(hptrs16, ns, flush=True)

        step_host16()
        barrier()
        t16 = time.perf_counter()
        for _ in range(e2e_steps):
            step_host16()
        barrier()
        s16 = max_over_ranks(time.perf_counter() - t16) / e2e_steps
        e2e_i16 = {"value": world * audio_per_step / s16, "unit": "x realtime", "h2d_bytes_per_step": S * n_floats * 2,
                   "d2h_bytes_per_step": int(b.output_total), "ms_per_step": 1000.0 * s16, "steps": e2e_steps,
                   "note": "extension: int16 PCM in, widened on the device to Float(s) / 32768"}
        if rank == 0:
            f = (pcm16[1].cpu().numpy().astype(np.float32) / np.float32(32768.0))
            ref, _ = orc.encode_all(f, sample_rate=SR, bitrate_kbps=KBPS, mode="stereo")
            assert b.output(1) == ref, "int16 e2e output differs from the oracle"
        L.mp3b_host_free(hp16)
        del pcm16

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu:
        cpu, _, _ = cpu_reference(a.seconds, min_wall=10.0, max_wall=30.0)
    others = None
    if rank == 0 and world == 1 and not a.no_others:
        b.close(); del pcm                                          # the big batch's memory is not needed any more
        torch.cuda.empty_cache()
        others = {"c1": single_stream_config("c1", mp3, L, local), "c2": single_stream_config("c2", mp3, L, local),
                  "c3": single_stream_config("c3", mp3, L, local), "c5": bench_c5(200, mp3, L, local)}
        others["c2"]["pink"] = single_stream_config("c2pink", mp3, L, local, repeats=3)
    total_launches = int(sum_over_ranks(launches))
    if rank == 0:
        line = {"metric": metric, "value": value, "unit": "x realtime", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": {"workload": workload, "streams_per_gpu": S, "seconds_per_stream": a.seconds,
                                                "frames_per_pass": frames_per_pass, "l2": "inputs larger than L2",
                                                "parity": parity, "output_bytes_per_step_per_gpu": int(out_bytes), "host_numa": numa},
                "clocks": clocks, "e2e": e2e, "gpu_launches": total_launches, "roofline": roofline}
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if others is not None:
            line["other_configs"] = others
        if e2e_i16 is not None:
            line["e2e_i16"] = e2e_i16
        if tc is not None:
            line["tensor_core_matrixing"] = tc
        emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
