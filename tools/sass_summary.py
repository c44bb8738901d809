#!/usr/bin/env python
"""Per-kernel counts of the Blackwell-native SASS mnemonics in libsovits_b200.so (tcgen05.mma = UTC*MMA, tcgen05.ld/st =
LDTM/STTM, bulk TMA copies = UBLKCP, tensor-map TMA = UTMALDG, mbarrier = SYNCS) - the evidence B200_PROFILING.md asks for.
    python tools/sass_summary.py [so-vits-svc_b200/libsovits_b200.so] > profiles/r02/sass_summary.txt"""
import re
import subprocess
import sys
from collections import Counter, OrderedDict

so = sys.argv[1] if len(sys.argv) > 1 else "so-vits-svc_b200/libsovits_b200.so"
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n
kern = OrderedDict()
cur = None
for line in out.splitlines():
    m = re.match(r"\s+Function : (\S+)", line)
    if m:
        cur = m.group(1)
        kern[cur] = Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(2)
        kern[cur]["total"] += 1
        for key in ("UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "HMMA", "FFMA", "MUFU", "LDG", "STG", "RED", "LDS", "STS"):
            if op.startswith(key):
                kern[cur][key] += 1
cols = ["total", "UTCHMMA", "LDTM", "STTM", "UBLKCP", "UTMALDG", "SYNCS", "HMMA", "FFMA", "MUFU", "LDG", "STG", "RED"]
print(f"{'kernel':110s} " + " ".join(f"{c:>8s}" for c in cols))
for k, c in kern.items():
    name = re.sub(r"svb::\(anonymous namespace\)::|svb::", "", demangle(k))
    name = re.sub(r"\((?:[^()]|\([^()]*\))*\)\s*$", "", name).replace("(int)", "").replace("(bool)", "")
    print(f"{name[:110]:110s} " + " ".join(f"{c.get(col, 0):8d}" for col in cols))
