#!/usr/bin/env python3
"""Per-kernel DRAM traffic, executed instructions and pipe utilisation from an `ncu --set full` report of ONE device-plane
pass at the bench shape (4096 streams x 64 frames = 1048576 granule-channels per launch) -> profiles/<round>_kernel_traffic.json,
which bench.py reads for `roofline.traffic`, the issue-rate roofline and the per-stage `ncu` rows.
usage: tools/ncu_traffic.py gpurun_out/prof.ncu-rep [gc_per_launch] [round tag, default r02] [description of the run]"""
import csv, io, json, os, subprocess, sys
rep = sys.argv[1]; gc = int(sys.argv[2]) if len(sys.argv) > 2 else 1048576
tag = sys.argv[3] if len(sys.argv) > 3 else "r02"
what = sys.argv[4] if len(sys.argv) > 4 else "one device-plane pass at the bench shape (4096 streams x 64 frames)"
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}
stage_of = {"k_prepass": "prepass", "k_filterbank": "filterbank", "k_granule": "granule", "k_scan": "scan", "k_pack": "pack", "k_frames": "frames"}
def num(r, k):
    try: return float(r[col[k]].replace(",", ""))
    except Exception: return None
out = {}
for r in rows[2:]:
    name = r[col["Kernel Name"]].split("(")[0].replace("void ", "").split("<")[0].strip()      # "void k_granule<0>(...)" -> k_granule
    if name not in stage_of or stage_of[name] in out: continue
    unit = rows[1][col["dram__bytes_read.sum"]]
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(unit, 1.0)
    rd, wr = num(r, "dram__bytes_read.sum") * scale, num(r, "dram__bytes_write.sum") * scale
    out[stage_of[name]] = {
        "kernel": name, "gc_per_launch": gc, "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_gc": (rd + wr) / gc,
        "duration_us_under_ncu": num(r, "gpu__time_duration.sum") * {"s": 1e6, "ms": 1e3, "us": 1.0, "ns": 1e-3}.get(rows[1][col["gpu__time_duration.sum"]], 1.0),
        "inst_executed": num(r, "smsp__inst_executed.sum"), "inst_executed_per_gc": (num(r, "smsp__inst_executed.sum") or 0) / gc,
        "dram_throughput_pct": num(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        "issue_active_pct": num(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "pipe_fma_cycles_active_pct": num(r, "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
        "pipe_fp64_cycles_active_pct": num(r, "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
        "lsu_shared_wavefronts_pct": num(r, "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
        "warps_active_pct": num(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
        "registers_per_thread": num(r, "launch__registers_per_thread"),
        "source": "ncu --set full --clock-control none, " + what + ", " + os.path.basename(rep)}
dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", tag + "_kernel_traffic.json")
json.dump(out, open(dst, "w"), indent=1)
print("wrote", dst, sorted(out))
