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
          host cores, on a bounded sample of the same workload.
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


def cpu_reference(seconds_per_stream, min_wall=10.0, max_wall=40.0):
    """The CPU restatement of the reference on every host core: bounded sample of the C4 workload."""
    import numpy as np
    import oracle_binding as orc
    import signals
    cores = os.cpu_count() or 1
    n = max(2 * cores, 8)
    pcms = [signals.sine_noise(seconds_per_stream, sr=SR, f_left=stream_params(i)[0], f_right=stream_params(i)[1],
                               seed=stream_params(i)[2]) for i in range(min(n, 16))]
    pcms = [pcms[i % len(pcms)] for i in range(n)]
    opts = dict(sample_rate=SR, bitrate_kbps=KBPS, mode="stereo")
    orc.encode_streams(pcms[:cores], cores, **opts)            # warm-up
    done, t0 = 0, time.perf_counter()
    while True:
        orc.encode_streams(pcms, cores, **opts)
        done += n
        el = time.perf_counter() - t0
        if el >= min_wall or el >= max_wall:
            break
    audio = done * seconds_per_stream
    return {"value": audio / el, "unit": "x realtime (audio s / s)", "cores": cores, "kind": "port",
            "sample": "%d streams x %.0f s of the C4 recipe (44.1 kHz stereo CBR128), one oracle session per stream, "
                      "%d threads, %.1f s wall" % (done, seconds_per_stream, cores, el)}, el, audio


def bench_c5(a, mp3, L, local, rank, world):
    """BASELINE config 5: 1024 concurrent sessions fed 1152-sample stereo chunks through encode(samples:) — one batch call
    per chunk with pinned host buffers in, MP3 bytes out; p50 / p99 latency of a call (= of every frame in it)."""
    import numpy as np
    import torch
    S, steps = 1024, max(a.steps, 200)
    chunk = 1152 * CH
    opts = mp3.MP3EncoderOptions(sampleRate=SR, bitrateKbps=KBPS, mode=mp3.Mode.stereo)
    b = mp3.EncoderBatch(opts, S, local, 8)
    n_chunks = 64                                                   # distinct chunks per stream, cycled
    hp = C.c_void_p()
    assert L.mp3b_host_alloc(S * n_chunks * chunk * 4, C.byref(hp)) == 0, L.mp3b_last_error()
    pcm = torch.empty((S, n_chunks * chunk), dtype=torch.float32, device="cuda")
    for i in range(S):
        fl, fr, seed = stream_params(i)
        assert L.mp3b_synth_fill(local, pcm[i].data_ptr(), n_chunks * 1152, CH, SR, fl, fr, 0.5, 0.05, seed) == 0
    assert L.mp3b_device_copy(local, hp, pcm.data_ptr(), S * n_chunks * chunk * 4, 1) == 0
    ns = (C.c_size_t * S)(*([chunk] * S))
    lat = []
    for k in range(steps + 20):
        ptrs = (C.c_void_p * S)(*[hp.value + (i * n_chunks + k % n_chunks) * chunk * 4 for i in range(S)])
        t0 = time.perf_counter()
        b.encode_ptrs(ptrs, ns, flush=False)
        lat.append(time.perf_counter() - t0)
    lat = np.array(lat[20:]) * 1e3
    if rank == 0:
        import oracle_binding as orc
        host = pcm[3].cpu().numpy()
        rs = orc.Session(sample_rate=SR, bitrate_kbps=KBPS, mode="stereo")
        ref = b"".join(rs.encode(host[(k % n_chunks) * chunk:(k % n_chunks + 1) * chunk]) for k in range(steps + 20))
        assert b.byte_count(3) == len(ref) and b.output(3) == ref[-len(b.output(3)):], "streaming output differs from the oracle"
        p50, p99 = float(np.percentile(lat, 50)), float(np.percentile(lat, 99))
        emit({"metric": "per-frame latency, 1024 concurrent sessions (config 5)", "value": p50, "unit": "ms (p50)",
                          "p99_ms": p99, "mean_ms": float(lat.mean()), "n_gpus": 1, "steps": steps, "higher_is_better": False,
                          "frame_period_ms": 1152 / SR * 1e3, "realtime_factor_p50": S * 1152 / SR * 1e3 / p50,
                          "data": "synthetic", "config": {"workload": "C5: 1024 sessions x 1152-sample stereo chunks, host buffers in, bytes out"},
                          "stage_ms_last_call": b.stage_ms(), "gpu_launches_per_call": b.launch_count,
                          "parity": "session 3 byte-identical to the oracle over %d chunks" % (steps + 20)})
    L.mp3b_host_free(hp)
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
    ap.add_argument("--i16", action="store_true", help="also time the 16-bit-input extension end to end (reported as e2e_i16, not the headline)")
    ap.add_argument("--workload", default="c4", choices=["c4", "c5"],
                    help="c4 (default, the headline): batch of 30 s streams; c5: streaming latency, 1024 sessions x 1152-sample chunks")
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
                "config": {"workload": workload, "reference": "C restatement of SwiftMP3 (oracle/); the Swift + Accelerate "
                           "original cannot be built on Linux"},
                "cpu_baseline": res,
                "e2e": {"value": res["value"], "unit": "x realtime", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        emit(line)
        return 0

    import numpy as np
    import torch
    mp3 = importlib.import_module("swift-mp3_b200")
    L = mp3.lib()
    torch.cuda.set_device(local)
    numa = sharding.bind_near_gpu(local)           # before any pinned allocation
    sampler = ClockSampler(local); sampler.start()
    dist = None
    if world > 1:
        os.environ["NCCL_DEBUG"] = os.environ.get("MP3B_NCCL_DEBUG", "WARN")   # keep stdout to the one JSON line
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

    if a.workload == "c5":
        rc = bench_c5(a, mp3, L, local, rank, world)
        sampler.stop()
        return rc
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
    ext = torch.cuda.ExternalStream(b.cuda_stream, device=torch.device("cuda", local))
    audio_per_step = S * a.seconds

    # ---- parity spot check (untimed): two streams of this shard against the CPU oracle
    parity = None
    if rank == 0:
        import oracle_binding as orc
        b.reset()
        b.encode_device(dptrs, ns, flush=True, download=True)
        checked = 0
        for i in (0, S - 1):
            ref, _ = orc.encode_all(pcm[i].cpu().numpy(), sample_rate=SR, bitrate_kbps=KBPS, mode="stereo")
            assert b.output(i) == ref, "stream %d differs from the oracle" % i
            checked += 1
        parity = "%d streams bit-identical to the oracle" % checked

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
    ncu = {}
    try:
        ncu = json.load(open(os.path.join(ROOT, "profiles", "r01_kernel_traffic.json")))
    except Exception:
        pass
    stage_table = {}
    fused_prepass = stages["prepass"] / max(passes, 1) < 0.02     # CBR without joint stereo: k_prepass is not launched at all
    if fused_prepass:
        algo["granule"] += 2304                                   # k_granule then reads the granule's PCM itself (block-type decision)
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
        stage_table[k] = row
    dom = max(algo, key=lambda k: stages[k])
    dom_ms_per_launch = stages[dom] / max(passes, 1)
    achieved = gc_per_launch * algo[dom] / (dom_ms_per_launch / 1000.0) / 1e9
    traffic = ncu[dom]["dram_bytes_per_gc"] * gc_per_launch if dom in ncu and "dram_bytes_per_gc" in ncu[dom] else None
    roofline = {"kernel": KERNEL_OF[dom], "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                "frac": achieved / hbm_peak, "traffic": traffic, "peak_source": which, "ms_per_launch": dom_ms_per_launch,
                "gc_per_launch": gc_per_launch, "share_of_step": stages[dom] / max(stages["total"], 1e-9),
                "limiter": ("FP32 FMA pipe and shared-memory wavefronts, not HBM: the reference's direct-form 32x64 matrixing is kept "
                            "operation for operation (bit-exact parity); see fp32 and DESIGN.md section 4") if dom == "filterbank" else
                           "instruction issue (MDCT with immediate coefficients, FP64 |x|^0.75, integer bit counting); see DESIGN.md section 4",
                "stages": stage_table, "stage_ms_per_step": {k: v / a.steps for k, v in stages.items()}}
    if dom in flops:
        ach = gc_per_launch * flops[dom] / (dom_ms_per_launch / 1000.0) / 1e12
        roofline["fp32"] = {"achieved": ach, "peak": fp32_peak, "unit": "TFLOP/s", "frac": ach / fp32_peak,
                            "peak_source": "nominal 148 SM x 128 lanes x 2 x sm_max_mhz (no measured FP32 peak in MEASURED_PEAKS.json)",
                            "peak_measured": FP32_MEASURED_FRACTION * fp32_peak, "frac_of_measured": ach / (FP32_MEASURED_FRACTION * fp32_peak),
                            "peak_measured_source": "tools/microbench/mb.cu kernel C on a B200 of this pool: register-only FFMA2 stream "
                                                    "reaches 90.5-91.0 % of the nominal issue rate"}

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
    if rank == 0:
        import oracle_binding as orc
        ref, _ = orc.encode_all(pcm[1].cpu().numpy(), sample_rate=SR, bitrate_kbps=KBPS, mode="stereo")
        assert b.output(1) == ref, "e2e output differs from the oracle"
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
            import oracle_binding as orc
            f = (pcm16[1].cpu().numpy().astype(np.float32) / np.float32(32768.0))
            ref, _ = orc.encode_all(f, sample_rate=SR, bitrate_kbps=KBPS, mode="stereo")
            assert b.output(1) == ref, "int16 e2e output differs from the oracle"
        L.mp3b_host_free(hp16)
        del pcm16

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu:
        cpu, _, _ = cpu_reference(a.seconds, min_wall=10.0, max_wall=30.0)
    total_launches = int(sum_over_ranks(launches))
    if rank == 0:
        line = {"metric": metric, "value": value, "unit": "x realtime", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": {"workload": workload, "streams_per_gpu": S, "seconds_per_stream": a.seconds,
                                                "frames_per_pass": b.frames_per_pass, "l2": "inputs larger than L2",
                                                "parity": parity, "output_bytes_per_step_per_gpu": int(out_bytes), "host_numa": numa},
                "clocks": clocks, "e2e": e2e, "gpu_launches": total_launches, "roofline": roofline}
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if e2e_i16 is not None:
            line["e2e_i16"] = e2e_i16
        emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
