"""Throughput of K8 select_top against the reference's pandas expression (development tool)."""
import os
import sys
import time

import numpy as np
import pandas as pd
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from metadamage_b200.backend import Context  # noqa: E402

n_tax = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
n_top = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
rng = np.random.default_rng(3)
tax = rng.permutation(n_tax).astype(np.int64)
nal = np.minimum(10 * (rng.pareto(1.1, n_tax) + 1), 6e7).astype(np.uint32)
tax_row, nal_row = np.repeat(tax, 30), np.repeat(nal, 30)
first = (np.arange(n_tax, dtype=np.int64) * 30)
dev = torch.device("cuda", 0)
ctx = Context(0)
ctx.set_stream(torch.cuda.current_stream(dev).cuda_stream)
cols = dict(tax_id=torch.from_numpy(tax_row).to(dev), n_alignments=torch.from_numpy(nal_row.view(np.int32)).to(dev))
outs = dict(tax_id=torch.from_numpy(tax).to(dev), first_row=torch.from_numpy(first).to(dev))
idx = torch.empty(n_top, dtype=torch.int64, device=dev)
for _ in range(3):
    n_sel = ctx.select_top_device(cols, outs, n_tax, n_top, idx)
ms = [0.0] * 5
for i in range(5):
    ctx.select_top_device(cols, outs, n_tax, n_top, idx)
    ms[i] = ctx.timings()["total_ms"]
got = np.sort(tax[idx[:n_sel].cpu().numpy()])
df = pd.DataFrame({"tax_id": pd.Series(tax_row).astype("category"), "N_alignments": nal_row})
t0 = time.time()
top = df.groupby("tax_id", observed=True)["N_alignments"].sum().nlargest(n_top).index
sel = df[df["tax_id"].isin(top)]
t_pd = time.time() - t0
exp = np.sort(pd.unique(sel["tax_id"]).astype(np.int64))
print(f"n_tax {n_tax} rows {len(tax_row)} n_top {n_top}: K8 {np.mean(ms):.3f} ms (device-resident, CUDA events, {ctx.timings()['n_launches']} launches) "
      f"= {(len(tax_row) * 12 + n_tax * 16 * 17) / np.mean(ms) / 1e6:.0f} GB/s of algorithmic traffic; pandas groupby.sum().nlargest()+isin: {t_pd * 1e3:.0f} ms; same set: {np.array_equal(got, exp)}")
