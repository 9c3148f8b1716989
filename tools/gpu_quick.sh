#!/bin/bash
# Quick GPU check: parity tests + short bench; prints value, stage times, e2e.   usage: gpu_quick.sh <tag>
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/pytest_$1.txt
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_$1.json 2> gpurun_out/bench_$1.err; echo rc=$?
tail -3 gpurun_out/bench_$1.err
python - <<PY
import json
d = json.load(open("gpurun_out/bench_$1.json"))
print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"])
print({k: round(v, 3) for k, v in d["roofline"]["stage_ms_per_step"].items()}); print("dominant", d["roofline"]["kernel"][:14], "hbm frac", round(d["roofline"]["frac"], 3), "fp32", d["roofline"].get("fp32", {}).get("frac"))
print(open("gpurun_out/pytest_$1.txt").read().strip().splitlines()[-1])
PY
