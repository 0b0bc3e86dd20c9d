"""How fast does ONE chain run when it has the GPU to itself? The collapsed-step-size TaxID of rank 1's
bench shard (tools/straggler_probe.py) fitted alone, then inside batches of low-coverage fillers
(development tool; on the GPU box)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from metadamage_b200 import _lib
from metadamage_b200.backend import Context
K0 = np.array([2, 1, 2, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 2, 1, 1, 2, 1, 1, 2, 0, 1, 0, 0, 0, 0, 1, 0], np.uint32)
N0 = np.array([5, 4, 6, 1, 1, 2, 1, 1, 2, 7, 2, 1, 3, 3, 3, 2, 3, 3, 6, 1, 2, 3, 2, 6, 3, 3, 3, 2, 6, 4], np.uint32)
TID0 = 100024416
ctx = Context(0)
cfg = _lib.default_config()
for n in (1, 1, 600, 2400, 9600):
    k = np.repeat(K0[None], n, 0); N = np.repeat(N0[None], n, 0)
    tid = np.arange(n, dtype=np.int64) + 7000
    tid[0] = TID0
    out = ctx.fit_batch(tid, k, N, cfg)
    t = ctx.timings()
    L = out["result"]["run"]["n_leapfrog"]
    print(f"batch {n}: nuts_ms {t['nuts_ms']:.1f}; straggler runs {L[0].tolist()}; "
          f"us per leapfrog of the longest if it spans the launch: {1e3 * t['nuts_ms'] / L[0].max():.2f}; "
          f"mean leapfrogs per chain {L.mean():.0f}", flush=True)
