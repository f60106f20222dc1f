"""Static SASS instruction counts of the hot kernels (cuobjdump -sass of the built library), the evidence for TMA loads
(UTMALDG.3D), mbarriers (SYNCS), 16-byte shared-memory loads and global stores:
    python scripts/sass_counts.py [kernel-name-substring ...] > profiles/rNN_sass_counts.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ("UTMALDG.3D", "SYNCS", "LDS.128", "LDS.64", "STG.E.128", "STG.E.64", "LDG.E.128", "LDG.E.64", "DFMA", "DMUL", "DADD", "MUFU",
        "IMAD.WIDE.U32", "IMAD.HI.U32", "ATOMG", "ST.E", "LD.E", "BAR", "ELECT")
names = sys.argv[1:] or ["stage_pair_kernel", "stage_rows_kernel"]
txt = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "jams_b200", "libjams_b200.so")], capture_output=True, text=True, check=True).stdout
print("cuobjdump -sass jams_b200/libjams_b200.so (sm_100a): static instruction counts by mnemonic, per kernel instantiation")
for f in re.split(r"\n\s*Function : ", txt)[1:]:
    name = f.split("\n", 1)[0]
    if not any(n in name for n in names):
        continue
    c = collections.Counter()
    for m in re.finditer(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", f, re.M):
        op = m.group(1)
        for key in KEYS:
            if op.startswith(key):
                c[key] += 1
        c["total"] += 1
    demangled = subprocess.run(["cu++filt", name], capture_output=True, text=True).stdout.strip() or name
    print(demangled[:160])
    print("   " + "  ".join(f"{k}={v}" for k, v in sorted(c.items())))
