"""Per-kernel SASS instruction summary of libtod.so (cuobjdump -sass): how many tcgen05 MMA (UTCHMMA / UTCQMMA), TMEM load
(LDTM), TMA load / store (UTMALDG / UTMASTG), tensor-memory management and cluster instructions each kernel contains --
the evidence that the hot kernels are Blackwell-native (B200_PROFILING.md lists the mnemonics).
usage: sass_summary.py [path/to/libtod.so] > profiles/rN_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "transparent_object_detection_b200", "libtod.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
KEYS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTCATOMSWS", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP", "SYNCS", "UCGABAR",
        "MUFU.TANH", "MUFU.EX2", "FFMA2", "HMMA", "LDG", "STG", "LDS", "STS", "BAR.SYNC", "ACQBULK", "ELECT"]
kern, counts, total = None, collections.OrderedDict(), {}
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = m.group(1)
        counts[kern] = collections.Counter()
        total[kern] = 0
        continue
    if kern is None:
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if not m:
        continue
    op = m.group(1)
    total[kern] += 1
    for k in KEYS:
        if op.startswith(k):
            counts[kern][k] += 1
            break


def demangle(n):
    try:
        return subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip().replace("tod::", "")
    except Exception:
        return n


print(f"SASS summary of {os.path.relpath(lib, ROOT)} (sm_100a), {len(counts)} kernels; columns = instruction counts in the kernel's code")
agg = collections.Counter()
for k, c in counts.items():
    name = demangle(k)
    name = re.sub(r"\(.*\)$", "", name)
    items = " ".join(f"{key}={c[key]}" for key in KEYS if c[key])
    print(f"{name[:70]:70s} {total[k]:6d} instr | {items}")
    agg.update(c)
print("TOTAL " + " ".join(f"{key}={agg[key]}" for key in KEYS if agg[key]))
