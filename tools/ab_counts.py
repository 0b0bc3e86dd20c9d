"""Counts kernel A/B on the GPU box (development tool): the 10M-row stress input of bench.py, CUDA-event time of
one mdg_counts_reduce call (device resident), with the kernel selected by MDG_COUNTS_TILES (round-1 tile kernel)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from metadamage_b200.backend import Context  # noqa: E402

dev = torch.device("cuda", 0)
ctx = Context(0)
ctx.set_stream(torch.cuda.current_stream(dev).cuda_stream)
r = bench.counts_stress(ctx, torch, dev, reps=int(sys.argv[1]) if len(sys.argv) > 1 else 10)
print(f"tiles={os.environ.get('MDG_COUNTS_TILES', '')} kernel_ms={r['kernel_ms']:.4f} achieved={r['achieved']:.0f} GB/s frac={r['frac']:.3f} kept={r['kept_taxa']}")
