#!/bin/bash
# One GPU session: parity tests, a full-set ncu capture of the hot kernels on ONE device-plane pass at the bench shape (4096
# streams x 64 frames), the default bench (computed from that capture), the reference arm, then the ncu launch list of the bench
# command (fewer steps) and the opt-in kernels.  Outputs land in gpurun_out/; tools/profile_report.py turns them into the files
# under profiles/.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv; nproc; free -g | head -2
python -m pytest tests -m gpu -x -q -s 2>&1 | tee gpurun_out/pytest_gpu_full.txt | tail -15
grep -h "^\.*ISO mode:\|^\.*ISO level\|^\.*tensor-core matrixing" gpurun_out/pytest_gpu_full.txt | sed 's/^\.*//' > gpurun_out/quality_report.txt
# First the full-set capture of ONE pass at the bench shape — 4096 streams x 64 frames (1.67 s of audio each) = 1 048 576
# granule-channels per launch — and its per-kernel summary: bench.py reads profiles/r02_kernel_traffic.json for roofline.traffic and
# the instruction counts of the issue roofline, so the bench line below is computed from THIS build's capture.
if [ "$1" != "noprof" ]; then
ONE="python bench.py --streams 4096 --seconds 1.67 --steps 1 --warmup 1 --e2e-steps 1 --no-cpu --no-others --no-tc --parity spot"
$ONE > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k "regex:k_(filterbank|granule|scan|pack|frames|carry)" -s 14 -c 7 -o gpurun_out/prof $ONE > gpurun_out/ncu2.log 2>&1
echo ncu2 rc=$?
python tools/ncu_traffic.py gpurun_out/prof.ncu-rep 1048576 r02 && cp profiles/r02_kernel_traffic.json gpurun_out/kernel_traffic.json
fi
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo bench rc=$?
tail -c 1500 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo ref rc=$?
[ -x tools/pool_latency ] && ./tools/pool_latency 1024 200 300 > gpurun_out/c5_pool.json 2> gpurun_out/c5_pool.err
[ -x tools/microbench/mb ] && ./tools/microbench/mb > gpurun_out/microbench.txt 2>&1
if [ "$1" != "noprof" ]; then
KERN="regex:k_(prepass|bitrate|filterbank|granule|scan|pack|frames|carry|offsets|gather)"
TCK="regex:k_filterbank_tc"
FULL="python bench.py --steps 1 --warmup 3 --e2e-steps 1 --no-cpu --no-others --no-tc --parity spot"
$FULL > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KERN" -c 400 --csv --log-file gpurun_out/launches.csv $FULL > gpurun_out/ncu1.log 2>&1
echo ncu1 rc=$?
# the opt-in kernels: tensor-core filterbank (one launch at the bench shape), ISO mode psychoacoustic model and outer loop
MP3B_MATRIXING=1 python tools/stage_times.py 4096 1.671837 1 > gpurun_out/plain3.log 2>&1 && \
MP3B_MATRIXING=1 ncu --set full --clock-control none --import-source on -k "$TCK" -s 2 -c 1 -o gpurun_out/prof_tc python tools/stage_times.py 4096 1.671837 1 > gpurun_out/ncu3.log 2>&1
echo ncu3 rc=$?
python tools/iso_bench.py 1024 10 3 > gpurun_out/iso_bench.json 2> gpurun_out/iso_bench.err; echo iso rc=$?
# (sections only, no source import: gpurun copies back at most 64 MiB and the two reports above take 38)
ncu --section SpeedOfLight --section LaunchStats --section Occupancy --section WarpStateStats --clock-control none -k "regex:k_(psy|outer|iso_blocktype)" -s 0 -c 2 -o gpurun_out/prof_iso python tools/iso_bench.py 256 10 1 > gpurun_out/ncu4.log 2>&1
echo ncu4 rc=$?
fi
ls -la gpurun_out | tail -12
