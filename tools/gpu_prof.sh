#!/bin/bash
# ncu --set full on selected kernels of a reduced bench run.  usage: gpu_prof.sh <kernel-regex> <skip> <count> <outname>
set -x
SMALL="python bench.py --streams 128 --seconds 10 --steps 1 --warmup 1 --no-cpu"
$SMALL > gpurun_out/plain_$4.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k "regex:$1" -s $2 -c $3 -o gpurun_out/$4 $SMALL > gpurun_out/ncu_$4.log 2>&1
echo ncu rc=$?
