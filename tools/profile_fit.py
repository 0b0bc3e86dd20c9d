"""Short fit-only program for ncu captures of the NUTS kernels (development tool):
python tools/profile_fit.py n_fit warmup samples. Prints the leapfrog totals per run kind."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from metadamage_b200 import _lib, synthetic as syn  # noqa: E402
from metadamage_b200.backend import Context  # noqa: E402

n_fit = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
warm = int(sys.argv[2]) if len(sys.argv) > 2 else 40
samp = int(sys.argv[3]) if len(sys.argv) > 3 else 40
tid, k, N, g = syn.dense_fit_batch(n_fit)
ctx = Context(0)
out = ctx.fit_batch(tid, k, N, _lib.default_config(num_warmup=warm, num_samples=samp, do_map=0))
print("fit", ctx.timings())
