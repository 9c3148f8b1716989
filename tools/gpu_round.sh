#!/bin/bash
# One GPU session: parity tests, the default bench, the reference arm, config 5, then the ncu launch list of the bench
# command (fewer steps) and a full-set capture of the hot kernels on a reduced bench.  Outputs land in gpurun_out/;
# tools/profile_report.py turns them into the files under profiles/.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo bench rc=$?
tail -c 1200 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo ref rc=$?
python bench.py --workload c5 --steps 300 > gpurun_out/c5.json 2> gpurun_out/c5.err; echo c5 rc=$?
[ -x tools/pool_latency ] && ./tools/pool_latency 1024 200 300 > gpurun_out/c5_pool.json 2> gpurun_out/c5_pool.err
KERN="regex:k_(prepass|bitrate|filterbank|granule|scan|pack|frames|carry|offsets|gather)"
FULL="python bench.py --steps 1 --warmup 3 --e2e-steps 1 --no-cpu"
$FULL > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KERN" -c 400 --csv --log-file gpurun_out/launches.csv $FULL > gpurun_out/ncu1.log 2>&1
echo ncu1 rc=$?
SMALL="python bench.py --streams 128 --seconds 10 --steps 1 --warmup 1 --no-cpu"
$SMALL > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k "regex:k_(prepass|filterbank|granule|scan|pack|frames)" -s 7 -c 6 -o gpurun_out/prof $SMALL > gpurun_out/ncu2.log 2>&1
echo ncu2 rc=$?
ls -la gpurun_out | tail -25
