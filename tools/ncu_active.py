"""Per CUDA source line: executed warp instructions and average active threads (development tool).
usage: ncu_active.py <all.sass from nvdisasm -g -c> <ncu source csv> <kernel name substring> [kernel index] [top]"""
import collections
import csv
import re
import sys

sass_path, csv_path, kname = sys.argv[1:4]
kidx = int(sys.argv[4]) if len(sys.argv) > 4 else 0
top = int(sys.argv[5]) if len(sys.argv) > 5 else 60
lines = open(sass_path, errors="replace").read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and kname in l and l.rstrip().endswith(":"))
ins = []
cur = ("?", 0)
for l in lines[start + 1:]:
    if l.startswith("//--------------------- .text.") or l.startswith(".section"):
        if ins:
            break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        ins.append((m.group(2).strip(), cur))
rows = list(csv.reader(open(csv_path)))
idx = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
hdr = rows[idx[kidx] + 1]
data = rows[idx[kidx] + 2: idx[kidx + 1] if len(idx) > kidx + 1 else len(rows)]
iex, ith, ismp = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
n = min(len(ins), len(data))
print("nvdisasm instrs", len(ins), "ncu instrs", len(data))
ex = collections.Counter(); th = collections.Counter(); smp = collections.Counter()
tot = sum(int(r[iex]) for r in data)
for i in range(n):
    key = ins[i][1]
    ex[key] += int(data[i][iex]); th[key] += int(data[i][ith]); smp[key] += int(data[i][ismp])
byfile_ex = collections.Counter(); byfile_th = collections.Counter()
for key in ex:
    byfile_ex[key[0]] += ex[key]; byfile_th[key[0]] += th[key]
for f, c in byfile_ex.most_common(8):
    print(f"{f:28s} instr share {c / tot:6.3f}  avg active {byfile_th[f] / max(1, c):5.1f}")
# histogram of instruction share by active-thread bucket
buckets = collections.Counter()
for i in range(n):
    e_, t_ = int(data[i][iex]), int(data[i][ith])
    if e_:
        buckets[min(32, int(round(t_ / e_ / 4.0)) * 4)] += e_
print("share of executed warp instructions by avg active threads:", {b: round(c / tot, 3) for b, c in sorted(buckets.items())})
for key, c in ex.most_common(top):
    print(f"{key[0]}:{key[1]:5d}  instr {c / tot:6.3f}  active {th[key] / max(1, c):5.1f}  samples {smp[key]}")
if len(sys.argv) > 6:
    f = sys.argv[6]
    print(f"--- {f} by line")
    for key in sorted(k_ for k_ in ex if k_[0] == f):
        if ex[key] / tot > 0.0005:
            print(f"{key[1]:5d}  instr {ex[key] / tot:6.4f}  active {th[key] / max(1, ex[key]):5.1f}")
