"""cProfile of the counts seam on a cfg2 TSV (development tool, GPU box)."""
import cProfile
import os
import pstats
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from metadamage_b200 import counts, fits, synthetic as syn, utils  # noqa: E402

g = syn.make_mismatch_matrix(0, n_fit=10000)
tmp = tempfile.mkdtemp()
path = os.path.join(tmp, "cfg2.txt")
syn.write_tsv(g, path)
cfg = utils.Config(out_dir=os.path.join(tmp, "out"), max_fits=None, max_cores=1, max_position=15, min_alignments=10, min_y_sum=10,
                   substitution_bases_forward="CT", substitution_bases_reverse="GA", forced=True, version="x")
cfg.add_filename(path)
counts.compute_counts(cfg)
t0 = time.perf_counter(); df = counts.compute_counts(cfg); print("compute_counts", time.perf_counter() - t0)
pr = cProfile.Profile(); pr.enable(); df = counts.compute_counts(cfg); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
