"""Dump per-run leapfrog counts of a fit batch with the TaxIDs' coverage (development tool: tail analysis)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from metadamage_b200 import _lib, synthetic as syn  # noqa: E402
from metadamage_b200.backend import Context  # noqa: E402

n_fit = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
rank = int(sys.argv[2]) if len(sys.argv) > 2 else 0  # bench.py's per-rank workload
ctx = Context(0)
tid, k, N, g = syn.dense_fit_batch(n_fit, seed=syn.SEEDS["cfg2"] + 1000 * rank, tax_id_start=1 + rank * 100_000_000)
out = ctx.fit_batch(tid, k, N, _lib.default_config())
res = out["result"]
np.savez_compressed(os.path.join("gpurun_out", f"leapfrogs_rank{rank}.npz"), n_leapfrog=res["run"]["n_leapfrog"], step=res["run"]["step_size"],
                    k=k, N=N, tax_id=tid)
print(ctx.timings())
