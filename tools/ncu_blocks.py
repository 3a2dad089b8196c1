"""Summarise the SASS page of an ncu report: stall samples per block of instructions.

usage: python tools/ncu_blocks.py report.ncu-rep [block=25] [min_samples=150]"""
import csv, subprocess, sys, io
from collections import Counter
rep = sys.argv[1]
blk = int(sys.argv[2]) if len(sys.argv) > 2 else 25
thr = int(sys.argv[3]) if len(sys.argv) > 3 else 150
import os
extra = os.environ.get("NCU_SELECT", "").split()   # e.g. NCU_SELECT="--launch-skip 1 --launch-count 1"
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", *extra], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = rows[1]; ix = {k: i for i, k in enumerate(hdr)}
end = next((i for i in range(2, len(rows)) if rows[i] and rows[i][0] == 'Kernel Name'), len(rows))
data = [r for r in rows[2:end] if len(r) == len(hdr)]
print(rows[0][1][:90])
keys = [k for k in hdr if k.startswith('stall_') and 'Not Issued' not in k]
tot_s = sum(int(r[ix['# Samples']]) for r in data)
print("total samples", tot_s, "instructions", len(data))
def op(r):
    t = r[1].split()
    return (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
for lo in range(0, len(data), blk):
    rng = range(lo, min(lo + blk, len(data)))
    ns = sum(int(data[n][ix['# Samples']]) for n in rng)
    if ns < thr: continue
    ex = max(int(data[n][ix['Instructions Executed']]) for n in rng)
    c = Counter(op(data[n]) for n in rng)
    tot = Counter()
    for n in rng:
        for k in keys: tot[k] += int(data[n][ix[k]])
    print(lo, ns, ex, dict(c.most_common(3)), {k[6:]: v for k, v in tot.most_common(3)})
if len(sys.argv) > 4:
    lo, hi = map(int, sys.argv[4].split(':'))
    for n in range(lo, hi):
        r = data[n]
        print(n, r[1][:78].ljust(78), r[ix['Instructions Executed']], r[ix['# Samples']],
              {k[6:]: int(r[ix[k]]) for k in keys if int(r[ix[k]]) > 0.25 * max(1, int(r[ix['# Samples']])) and int(r[ix['# Samples']]) > 20})
