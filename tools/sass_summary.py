"""Per-kernel counts of the SASS mnemonics that prove (or disprove) a Blackwell-native kernel, from `cuobjdump -sass` of the product
library (B200_PROFILING.md "What proves a Blackwell-native kernel"). No GPU needed.

    python tools/sass_summary.py > profiles/sass_summary.md
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "audio_residual_b200", "libard_b200.so")
OPS = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "HMMA", "MUFU", "LDGSTS"]

sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
counts = collections.OrderedDict()
fn = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1)
        counts[fn] = collections.Counter()
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and fn:
        op = m.group(1)
        counts[fn]["_total"] += 1
        for o in OPS:
            if op.startswith(o):
                counts[fn][o] += 1
names = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
print("# SASS evidence: tcgen05 / TMEM / TMA instructions per kernel of libard_b200.so\n")
print("`cuobjdump -sass audio_residual_b200/libard_b200.so` (sm_100a), counted by `tools/sass_summary.py`. `UTCHMMA` = `tcgen05.mma`, `LDTM` / `STTM` ="
      " `tcgen05.ld` / `.st`, `UTMALDG` / `UTMASTG` = TMA load / store, `HMMA` = legacy `mma.sync`, `LDGSTS` = `cp.async`.\n")
print("| kernel | SASS instr | " + " | ".join(OPS) + " |")
print("|---|---|" + "---|" * len(OPS))
for (fn, c), name in sorted(zip(counts.items(), names), key=lambda t: -(t[0][1]["UTCHMMA"] * 1000 + t[0][1]["HMMA"])):
    if not any(c[o] for o in OPS):
        continue
    short = re.sub(r"\(.*", "", name).replace("ard::", "")
    print(f"| `{short}` | {c['_total']} | " + " | ".join(str(c[o]) if c[o] else "" for o in OPS) + " |")
