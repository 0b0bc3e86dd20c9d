"""Short program for ncu captures: one counts_reduce launch on the 10M-row stress input and one
small fit batch (all fit kernels) — same kernels as bench.py, sized for ~40 replay passes."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from metadamage_b200 import _lib, synthetic as syn  # noqa: E402
from metadamage_b200.backend import Context  # noqa: E402

n_fit = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
warm = int(sys.argv[2]) if len(sys.argv) > 2 else 60
samp = int(sys.argv[3]) if len(sys.argv) > 3 else 60
dev = torch.device("cuda", 0)
ctx = Context(0)
ctx.set_stream(torch.cuda.current_stream(dev).cuda_stream)
r = bench.counts_stress(ctx, torch, dev, reps=int(os.environ.get("MDG_COUNTS_REPS", "1")))
print("counts", r["kernel_ms"], "ms", r["achieved"], "GB/s")
tid, k, N, g = syn.dense_fit_batch(n_fit)
out = ctx.fit_batch(tid, k, N, _lib.default_config(num_warmup=warm, num_samples=samp, do_fwd_rev=int(os.environ.get("MDG_PT_FWD_REV", "1"))))
print("fit", ctx.timings())
