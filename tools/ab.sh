#!/bin/bash
# Same-box A/B of two builds: swift-mp3_b200/libmp3b200.so (new) against libmp3b200_prev.so (a build of HEAD), alternating, so that
# box-to-box clock differences (power caps: 2-3 %) cancel.   usage (under gpurun): bash tools/ab.sh [streams] [seconds]
S=${1:-4096}; T=${2:-30}
cp swift-mp3_b200/libmp3b200.so /tmp/new.so
for r in 1 2; do
  cp /tmp/new.so swift-mp3_b200/libmp3b200.so; echo -n "new  "; python tools/stage_times.py $S $T 3 2>&1 | tail -1
  cp swift-mp3_b200/libmp3b200_prev.so swift-mp3_b200/libmp3b200.so; echo -n "prev "; python tools/stage_times.py $S $T 3 2>&1 | tail -1
done
cp /tmp/new.so swift-mp3_b200/libmp3b200.so
