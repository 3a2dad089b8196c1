"""Per-kernel counts of the SASS mnemonics that prove (or disprove) a Blackwell-native kernel, from the in-tree library:
python tools/sass_summary.py > profiles/r02/sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(
    ROOT, "a-multimodal-diffusion-based-model-for-point-cloud-completion_b200", "libpcd_b200.so")
MN = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "MUFU.EX2", "MUFU.TANH", "HMMA", "FFMA2", "LDGSTS"]
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
counts, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur).replace("void ", "").replace("pcd::", "")
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    for k in MN:
        if re.search(r"\b" + re.escape(k) + r"\b", line) or (k in ("UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "SYNCS") and k in line):
            counts[cur][k] += 1
print(f"{'kernel':70s} " + " ".join(f"{k:>9s}" for k in MN))
for name, c in counts.items():
    print(f"{name[:70]:70s} " + " ".join(f"{c.get(k, 0):9d}" for k in MN))
print("\nHMMA (legacy mma.sync path) total:", sum(c.get("HMMA", 0) for c in counts.values()))
