#!/bin/bash
# phase-elimination timing of the tensor-core filterbank (MP3B_FB_DEBUG bits: 1 windowing off, 2 MMA off, 4 PCM loads off, 8 stores off, 16 epilogue off)
# 1.671837 s = 73728 samples = 64 frames: rows of
# the PCM matrix stay 16-byte aligned (the bulk-copy path); 1.67 s: every other row is not
for v in 2; do for d in ${TC_PHASES:-0 2 18 19 23}; do echo -n "variant $v "; MP3B_TC_VARIANT=$v MP3B_MATRIXING=1 MP3B_FB_DEBUG=$d python tools/stage_times.py 4096 1.671837 3 2>&1 | tail -1; done; done
echo -n "unaligned "; MP3B_MATRIXING=1 python tools/stage_times.py 4096 1.67 3 2>&1 | tail -1
echo -n "fp32 "; MP3B_FB_DEBUG=0 python tools/stage_times.py 4096 1.671837 3 2>&1 | tail -1
