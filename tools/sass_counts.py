"""Developer tool: per-kernel counts of the tcgen05 / TMEM / TMA SASS mnemonics in the in-tree library (cuobjdump -sass).
argv: [library]  ->  stdout (profiles/*_sass_evidence.txt)"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "kws_b200", "lib", "libfastgrnn_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
KEYS = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR"]
counts, name = {}, None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1)
        counts[name] = dict.fromkeys(KEYS, 0)
        continue
    if name:
        for k in KEYS:
            if re.search(r"\b%s\b" % k, line):
                counts[name][k] += 1
print("cuobjdump -sass %s: tcgen05 / TMEM / TMA instruction counts per kernel" % os.path.relpath(lib, ROOT))
for n in sorted(counts):
    c = counts[n]
    if any(c.values()):
        print("%-110s " % n + " ".join("%s %4d" % (k, c[k]) for k in KEYS))
print("peer-memory kernel (no tcgen05): " + ", ".join(n for n in counts if "peer" in n))
