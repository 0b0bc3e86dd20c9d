"""Development tool: reads the per-chain start / end times written by MDG_CHAIN_CLOCK=<file> (one fit of n TaxIDs;
[n][6][2] uint64 nanoseconds of %globaltimer per chunk) and prints the timeline of the NUTS phase: per run kind the
first start / last end, and the number of live chains over time.
python tools/chain_timeline.py file n_taxids"""
import sys

import numpy as np

raw = np.fromfile(sys.argv[1], dtype=np.uint64)
n = int(sys.argv[2])
per = n * 6 * 2
fits = raw.reshape(-1, n, 6, 2)[-1:]  # the last fit of the file (the first one pays lazy module loading: launches serialise)
names = ["PMD all", "null all", "PMD fwd", "null fwd", "PMD rev", "null rev"]
for f, c in enumerate(fits):
    ok = c[:, :, 1] > 0
    t0 = c[:, :, 0][ok].min()
    s = (c[:, :, 0].astype(np.int64) - int(t0)) * 1e-6
    e = (c[:, :, 1].astype(np.int64) - int(t0)) * 1e-6
    print(f"fit {f}: NUTS phase {e[ok].max():.1f} ms")
    for r in range(6):
        d = (e - s)[:, r]
        print(f"  {names[r]:9s} first start {s[:, r].min():7.1f}  median start {np.median(s[:, r]):7.1f}  last start {s[:, r].max():7.1f}  "
              f"median end {np.median(e[:, r]):7.1f}  last end {e[:, r].max():7.1f}   chain ms: median {np.median(d):6.1f} p99 {np.percentile(d, 99):6.1f} max {d.max():6.1f}")
    edges = np.linspace(0, e[ok].max(), 41)
    live = [(int(((s <= t) & (e > t)).sum()), [int(((s[:, r] <= t) & (e[:, r] > t)).sum()) for r in range(6)]) for t in edges[:-1]]
    print("  t ms   live chains  per run kind")
    for t, (tot, by) in zip(edges[:-1], live):
        print(f"  {t:6.1f} {tot:6d}   {by}")
    area = (e - s)[ok].sum()
    print(f"  chain-ms {area:.0f}; mean live chains {area / e[ok].max():.0f}")
