#!/usr/bin/env python3
"""Static SASS instruction census per kernel of libmp3b200.so (cuobjdump -sass) -> profiles/<tag>_sass_census.txt: which pipes a
kernel is written for (FFMA2 / FP64 / tensor: UTC*MMA, LDTM / TMA unit: UBLKCP / LDGSTS = cp.async ...)."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
so = os.path.join(ROOT, "swift-mp3_b200", "libmp3b200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
keys = ["FFMA2", "FFMA", "DFMA", "DMUL", "UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UBLKCP", "UTMALDG", "SYNCS", "LDGSTS", "LDS", "STS", "LDG", "STG", "ATOMS", "MUFU", "REDUX", "SHFL"]
out = ["# cuobjdump -sass swift-mp3_b200/libmp3b200.so: static instruction census per kernel (sm_100a)"]
cur, cnt, n = None, None, 0
def flush():
    if cur:
        out.append("%-70s instr %6d  %s" % (cur, n, " ".join("%s=%d" % (k, cnt[k]) for k in keys if cnt[k])))
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        flush(); cur, cnt, n = m.group(1), collections.Counter(), 0
        continue
    m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        n += 1
        op = m.group(1)
        for k in keys:
            if op == k or (k in ("SYNCS", "UTCBAR", "UBLKCP", "UTCHMMA") and op.startswith(k)):
                cnt[k] += 1; break
flush()
open(os.path.join(ROOT, "profiles", tag + "_sass_census.txt"), "w").write("\n".join(out) + "\n")
print("\n".join(l for l in out if "filterbank" in l or "psy" in l or "outer" in l))
