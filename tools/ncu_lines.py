"""Attribute ncu per-SASS-instruction counts to CUDA source lines (development tool).

usage: ncu_lines.py <all.sass from `nvdisasm -g -c`> <ncu source csv> <mangled kernel name substring> [top]
"""
import collections
import csv
import re
import sys

sass_path, csv_path, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40

# ---- nvdisasm: list of (instruction text, file, line) for the kernel
lines = open(sass_path, errors="replace").read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and kname in l and l.rstrip().endswith(":"))
ins = []
cur = ("?", 0)
inl = ""
for l in lines[start + 1:]:
    if l.startswith("//--------------------- .text.") or l.startswith(".section"):
        if ins:
            break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        inl = m.group(3)
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        ins.append((m.group(2).strip(), cur, inl))

rows = list(csv.reader(open(csv_path)))
idx = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
hdr = rows[idx[0] + 1]
data = rows[idx[0] + 2: idx[1] if len(idx) > 1 else len(rows)]
iex, ismp = hdr.index("Instructions Executed"), hdr.index("# Samples")
print("nvdisasm instrs", len(ins), "ncu instrs", len(data))
n = min(len(ins), len(data))
by_line = collections.Counter()
smp_line = collections.Counter()
tot = sum(int(r[iex]) for r in data)
tots = sum(int(r[ismp]) for r in data)
for i in range(n):
    key = ins[i][1]
    by_line[key] += int(data[i][iex])
    smp_line[key] += int(data[i][ismp])
by_file = collections.Counter()
for (f, ln), c in by_line.items():
    by_file[f] += c
print("by file:", [(f, round(c / tot, 3)) for f, c in by_file.most_common(8)])
for (f, ln), c in by_line.most_common(top):
    print(f"{f}:{ln:5d}  instr {c / tot:6.3f}  samples {smp_line[(f, ln)] / max(1, tots):6.3f}")
