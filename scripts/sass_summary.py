#!/usr/bin/env python
"""Mnemonic counts per kernel of the shipped library (`cuobjdump -sass`): the evidence that the tcgen05 / TMA / TMEM
instructions are really there.  Usage: python scripts/sass_summary.py [lib.so] > profiles/<name>.txt"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "sake_b200/libsake_b200.so"
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
cols = ["UTCHMMA", "LDTM", "STTM", "UBLKCP", "UTCBAR", "UTMALDG", "UTMASTG", "SYNCS", "MUFU", "BSSY", "LDG", "STG", "LDS", "STS", "SHFL", "RED"]
print(f"# cuobjdump -sass {lib} (product build, sm_100a): instruction-mnemonic counts per kernel")
print("# tcgen05.mma -> UTC*MMA, tcgen05.ld -> LDTM, cp.async.bulk -> UBLKCP, tcgen05.commit -> UTCBAR, tensor-map TMA would be UTMALDG/UTMASTG (none used:")
print("# every bulk copy here is the 1-D form), BSSY = divergent-branch regions (the first thing to check in a thread-per-row kernel)")
print(f"{'kernel':92s}" + "".join(f"{c:>9s}" for c in cols))
cur, cnt, i = None, None, 0
rows = []
for line in sass.split("\n"):
    m = re.search(r"Function : (\S+)", line)
    if m:
        if cur is not None:
            rows.append((cur, cnt))
        cur, cnt = names[i], collections.Counter()
        i += 1
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cnt is not None:
        op = m.group(1)
        for c in cols:
            if op.startswith(c) or (c == "UTCHMMA" and op.startswith("UTC") and "MMA" in op):
                cnt[c] += 1
                break
if cur is not None:
    rows.append((cur, cnt))
for name, cnt in rows:
    if "sake::" not in name:
        continue
    print(f"{name[:90]:92s}" + "".join(f"{cnt[c]:9d}" for c in cols))
