"""Summarise an .ncu-rep (development tool): headline metrics, opcode mix, stall ratios."""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
evals = float(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
srows = list(csv.reader(src.splitlines()))
kidx = [i for i, r in enumerate(srows) if r and r[0] == "Kernel Name"]
for kn, d in enumerate(rows[2:]):
    g = lambda name: d[hdr.index(name)] if name in hdr else "n/a"  # noqa: E731
    print("==", g("Kernel Name"))
    for name in ("gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
                 "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
                 "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
                 "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
                 "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"):
        print(f"   {name} = {g(name)}")
    if evals:
        print("   warp instructions per gradient evaluation =", float(g("smsp__inst_executed.sum")) / evals)
    st = []
    for i, h in enumerate(hdr):
        if "average_warps_issue_stalled" in h and h.endswith("per_issue_active.ratio") and "not_issued" not in h:
            try:
                st.append((float(d[i]), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
            except ValueError:
                pass
    print("   stalls (warps per issue):", ", ".join(f"{n} {v:.2f}" for v, n in sorted(st, reverse=True)[:8]))
    if kn < len(kidx):
        shdr = srows[kidx[kn] + 1]
        data = srows[kidx[kn] + 2: kidx[kn + 1] if kn + 1 < len(kidx) else len(srows)]
        isrc, iex = shdr.index("Source"), shdr.index("Instructions Executed")
        tot = sum(int(r[iex]) for r in data)
        ops = collections.Counter()
        for r in data:
            t = r[isrc].strip().split()
            if t:
                op = t[1] if t[0].startswith("@") else t[0]
                ops[op.split(".")[0]] += int(r[iex])
        print("   static SASS instructions:", len(data))
        print("   opcode mix:", ", ".join(f"{o} {c / tot:.3f}" for o, c in ops.most_common(16)))
