"""Per-kernel SASS opcode histogram of libct3_b200.so (cuobjdump -sass), the evidence DESIGN.md cites for
"Blackwell-native": UTCHMMA (tcgen05.mma), UTMALDG (TMA tensor loads), LDTM (tcgen05.ld), UTCBAR (tcgen05.commit),
UBLKCP (bulk copies), HMMA (legacy mma.sync).  Runs on the CPU-only build container.
    python scripts/sass_opcodes.py > profiles/r2_sass_opcodes.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "cotracker_b200", "lib", "libct3_b200.so")
KEY = ["UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "UTCBAR", "UTCATOM", "HMMA", "FFMA", "FFMA2", "SHFL",
       "LDS", "STS", "LDG", "STG", "BAR", "SYNCS", "BRX", "MUFU"]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
hist, name = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        d = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        d = d.replace("(anonymous namespace)::", "").replace("void ", "").replace("ct3::", "")
        name = re.sub(r"\(.*", "", d)
        while name in hist:
            name += "'"
        hist[name] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and name:
        hist[name][m.group(1)] += 1
print(f"# SASS opcode counts per kernel of {os.path.relpath(lib, ROOT)} (sm_100a); total = all instructions")
print("kernel".ljust(58) + "total".rjust(7) + "".join(k.rjust(8) for k in KEY))
tot = collections.Counter()
for n, h in hist.items():
    print(n[:57].ljust(58) + str(sum(h.values())).rjust(7) + "".join(str(h.get(k, 0)).rjust(8) for k in KEY))
    tot.update(h)
print("ALL".ljust(58) + str(sum(tot.values())).rjust(7) + "".join(str(tot.get(k, 0)).rjust(8) for k in KEY))
