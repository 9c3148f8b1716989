#!/bin/bash
# One GPU session: parity tests, bench, then ncu launch list + full capture of the hot kernels on a reduced bench.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo bench rc=$?
tail -c 2500 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
SMALL="python bench.py --streams 128 --seconds 10 --steps 1 --warmup 1 --no-cpu"
KERN="regex:k_(prepass|bitrate|filterbank|granule|scan|pack|frames|carry|offsets|gather)"
$SMALL > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KERN" -c 60 --csv --log-file gpurun_out/launches.csv $SMALL > gpurun_out/ncu1.log 2>&1
echo ncu1 rc=$?
$SMALL > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k "regex:k_(prepass|filterbank|granule|scan|pack|frames)" -s 7 -c 6 -o gpurun_out/prof $SMALL > gpurun_out/ncu2.log 2>&1
echo ncu2 rc=$?
ls -la gpurun_out
