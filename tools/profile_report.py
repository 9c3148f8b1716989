#!/usr/bin/env python3
"""Turn the outputs of tools/gpu_round.sh (gpurun_out/) into the committed evidence under profiles/ (TAG = argv[1], default r02):
  r01_launches.csv            ncu --metrics gpu__time_duration.sum launch list of `bench.py --steps 1 --warmup 3` (first 400 launches)
  r01_launch_shares.txt       per kernel and grid: launches, mean duration, share of the device-plane step
  r01_ncu_full_summary.txt    headline metrics + per-phase stall / opcode mix of the hot kernels (ncu --set full)
  r01_kernel_traffic.json     DRAM bytes and pipe utilisation per kernel (read by bench.py)
  r01_filterbank_lds.txt      shared-memory wavefronts per LDS / STS instruction of k_filterbank
  r01_bench.json, r01_bench_reference.json, r01_config5.json   the bench lines of the same session"""
import collections, csv, io, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TAG = sys.argv[1] if len(sys.argv) > 1 else "r02"
GC = sys.argv[2] if len(sys.argv) > 2 else "1048576"
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
shutil.copy(os.path.join(G, "launches.csv"), os.path.join(P, TAG + "_launches.csv"))
rows = [r for r in csv.reader(open(os.path.join(G, "launches.csv"))) if len(r) > 10 and r[0].isdigit()]
by = collections.OrderedDict()
for r in rows:
    key = (r[4].split("(")[0].replace("void ", "").split("<")[0].strip(), r[8])
    by.setdefault(key, []).append(float(r[-1]) / 1e3)
# device-plane passes = for each kernel the grid with the largest mean duration
best = {}
for (k, g), v in by.items():
    m = sum(v) / len(v)
    if k not in best or m > best[k][1]:
        best[k] = (g, m, len(v))
tot = sum(m for _, m, _ in best.values())
with open(os.path.join(P, TAG + "_launch_shares.txt"), "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none -c 400, `python bench.py --steps 1 --warmup 3 --e2e-steps 1 --no-cpu`\n")
    f.write("# (serialised, cold-cache launches: compare SHARES with bench.py's roofline.stages[*].share, not absolutes)\n")
    f.write("# device-plane pass (4096 streams x 64 frames) = per kernel the grid with the largest mean duration; k_gather / k_offsets belong\n# to the (untimed) parity download, k_scan / k_carry have one grid for every pass size:\n")
    for k, (g, m, n) in sorted(best.items(), key=lambda kv: -kv[1][1]):
        f.write("%-14s grid %-18s launches %3d  mean %9.1f us  share of pass %5.1f %%\n" % (k, g, n, m, 100 * m / tot))
    f.write("\n# every (kernel, grid) group of the list:\n")
    for (k, g), v in by.items():
        f.write("%-14s grid %-18s launches %3d  mean %9.1f us\n" % (k, g, len(v), sum(v) / len(v)))
rep = os.path.join(G, "prof.ncu-rep")
out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep], capture_output=True, text=True).stdout
seen, keep = set(), []
for block in out.split("=== source page:"):
    head = block.split("\n", 1)[0]
    if head in seen: continue
    seen.add(head); keep.append(block)
open(os.path.join(P, TAG + "_ncu_full_summary.txt"), "w").write("=== source page:".join(keep))
subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_traffic.py"), rep, GC, TAG], check=True)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(src)))
hi = [i for i, r in enumerate(rr) if r and r[0] == "Kernel Name"]
with open(os.path.join(P, TAG + "_filterbank_lds.txt"), "w") as f:
    f.write("# k_filterbank, ncu --set full source page: shared-memory wavefronts per executed LDS / STS / LDGSTS instruction\n")
    for a, b in zip(hi, hi[1:] + [len(rr)]):
        if "k_filterbank" not in rr[a][1]: continue
        hdr = rr[a + 1]; ci = {h: i for i, h in enumerate(hdr)}
        for r in rr[a + 2:b]:
            if len(r) != len(hdr): continue
            s = r[ci["Source"]]
            if not any(t in s for t in ("LDS", "STS", "LDGSTS")): continue
            ex = int(r[ci["Instructions Executed"]] or 0); wf = int(r[ci["L1 Wavefronts Shared"]] or 0)
            if ex: f.write("%-58s executed %9d  wavefronts %9d  per instruction %.2f\n" % (s.strip()[:58], ex, wf, wf / ex))
        break
for a, b in (("bench.json", TAG + "_bench.json"), ("bench_ref.json", TAG + "_bench_reference.json")):
    line = [l for l in open(os.path.join(G, a)) if l.startswith("{")][-1]
    json.dump(json.loads(line), open(os.path.join(P, b), "w"), indent=1)
bench_line = json.loads([l for l in open(os.path.join(G, "bench.json")) if l.startswith("{")][-1])
c5 = {"batch_call": bench_line.get("other_configs", {}).get("c5")}       # the C5 leg of the default bench line
if os.path.exists(os.path.join(G, "c5_pool.json")):
    c5["session_pool_1024_threads"] = json.loads([l for l in open(os.path.join(G, "c5_pool.json")) if l.startswith("{")][-1])
json.dump(c5, open(os.path.join(P, TAG + "_config5.json"), "w"), indent=1)
print("profiles/ updated:", sorted(os.listdir(P)))
