#!/bin/bash
# One GPU session: parity tests, the default bench, the reference arm, then the ncu launch list of the bench command (fewer
# steps) and a full-set capture of the hot kernels on ONE device-plane pass at the bench shape (4096 streams x 64 frames).
# Outputs land in gpurun_out/; tools/profile_report.py turns them into the files under profiles/.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv; nproc; free -g | head -2
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo bench rc=$?
tail -c 1500 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo ref rc=$?
[ -x tools/pool_latency ] && ./tools/pool_latency 1024 200 300 > gpurun_out/c5_pool.json 2> gpurun_out/c5_pool.err
[ -x tools/microbench/mb ] && ./tools/microbench/mb > gpurun_out/microbench.txt 2>&1
if [ "$1" != "noprof" ]; then
KERN="regex:k_(prepass|bitrate|filterbank|granule|scan|pack|frames|carry|offsets|gather)"
FULL="python bench.py --steps 1 --warmup 3 --e2e-steps 1 --no-cpu --no-others --parity spot"
$FULL > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KERN" -c 400 --csv --log-file gpurun_out/launches.csv $FULL > gpurun_out/ncu1.log 2>&1
echo ncu1 rc=$?
# one pass at the bench shape: 4096 streams x 64 frames (1.67 s of audio each) = 1 048 576 granule-channels per launch
ONE="python bench.py --streams 4096 --seconds 1.67 --steps 1 --warmup 1 --e2e-steps 1 --no-cpu --no-others --parity spot"
$ONE > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k "regex:k_(filterbank|granule|scan|pack|frames|carry)" -s 12 -c 6 -o gpurun_out/prof $ONE > gpurun_out/ncu2.log 2>&1
echo ncu2 rc=$?
fi
ls -la gpurun_out | tail -12
