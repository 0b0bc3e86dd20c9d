"""First-contact GPU check (development tool): device vs oracle on every building block."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from metadamage_b200 import synthetic as syn  # noqa: E402
from metadamage_b200 import _lib  # noqa: E402
from metadamage_b200.backend import Context  # noqa: E402
from oracle import oracle as O  # noqa: E402

n_fit = int(sys.argv[1]) if len(sys.argv) > 1 else 512
ctx = Context(0)
print("fp64 peak TFLOP/s:", ctx.fp64_peak_tflops())

# philox
key = np.array([[0, 0], [0xFFFFFFFF, 0xFFFFFFFF], [0xA4093822, 0x299F31D0]], np.uint32)
ctr = np.array([[0, 0, 0, 0], [0xFFFFFFFF] * 4, [0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344]], np.uint32)
print("philox", [[hex(v) for v in row] for row in ctx.philox(key, ctr)])

# special functions
from scipy import special  # noqa: E402
x = np.concatenate([10.0 ** np.linspace(-9, 9, 2000), np.linspace(0.01, 30, 3000)])
lg, dg = ctx.lgamma_digamma(x)
print("lgamma max rel err", np.max(np.abs(lg - special.gammaln(x)) / np.maximum(1, np.abs(special.gammaln(x)))),
      "digamma max rel err", np.max(np.abs(dg - special.digamma(x)) / np.maximum(1, np.abs(special.digamma(x)))))

g = syn.make_mismatch_matrix(0, n_fit=n_fit)
sel = g["passes"]
# counts
t = time.time()
r = ctx.counts_reduce(g["tax_id"], g["n_alignments"], g["is_reverse"], g["pos0"], g["counts16"], want_noise=True)
print("counts e2e %.3fs" % (time.time() - t), ctx.timings())
ro = O.counts_reduce(g["tax_id"], g["n_alignments"], g["is_reverse"], g["pos0"], g["counts16"])
for key_ in ("n_fwd_ref", "n_rev_ref", "z", "y_sum_total", "keep", "tax_id", "n_alignments", "first_row", "k", "N"):
    print("  counts", key_, np.array_equal(r[key_], ro[key_]))
print("  counts f_fwd bit-exact", np.array_equal(r["f_fwd"].view(np.uint32), ro["f_fwd"].view(np.uint32)),
      np.array_equal(r["f_rev"].view(np.uint32), ro["f_rev"].view(np.uint32)))
print("  noise max abs diff", np.nanmax(np.abs(r["noise"] - ro["noise"])))

tid, k, N = g["tax_ids"][sel], np.ascontiguousarray(g["k"][sel]), np.ascontiguousarray(g["N"][sel])
# logp/grad
rng = np.random.default_rng(0)
u = rng.uniform(-2, 2, (64, 4))
u[:, 2] -= 2
u[:, 3] += 4
for model in (0, 1):
    for mask in (0, 1, 2):
        for jac in (True, False):
            a = ctx.logp_grad(k[0], N[0], u, model=model, lane_mask=mask, with_jacobian=jac)
            b = O.logp_grad(k[0], N[0], u, model=model, lane_mask=mask, with_jacobian=jac)
            ok = np.isfinite(b[0])
            print("  logp model", model, "mask", mask, "jac", jac, "nvalid", ok.sum(),
                  "max|dlogp|", np.max(np.abs(a[0][ok] - b[0][ok])), "max|dgrad|", np.max(np.abs(a[1][ok] - b[1][ok])),
                  "max|dll|", np.max(np.abs(a[2][ok] - b[2][ok])), "nan agree", np.array_equal(np.isnan(a[0]), np.isnan(b[0])))

cfg = _lib.default_config()
nsmall = min(8, len(tid))
t = time.time()
out = ctx.fit_batch(tid[:nsmall], k[:nsmall], N[:nsmall], cfg, want_trace=True, want_samples=True, want_waic=True)
print("fit small %.3fs" % (time.time() - t), ctx.timings())
oo = O.fit_batch(tid[:nsmall], k[:nsmall], N[:nsmall], O.default_config(), want_trace=True, want_samples=True, want_waic=True)
res, reso = out["result"], oo["result"]
for f in ("map_A", "map_q", "map_c", "map_phi", "map_logp", "map_null_q", "map_null_phi"):
    print("  ", f, np.max(np.abs(res[f] - reso[f]) / np.maximum(1e-300, np.abs(reso[f]))))
for i in range(nsmall):
    for rk in range(6):
        a, b = out["trace"][i, rk], oo["trace"][i, rk]
        d = np.nanmax(np.abs(a - b), axis=1)
        first_bad = int(np.argmax(d > 1e-6)) if (d > 1e-6).any() else len(d)
        print("  tax", i, "run", rk, "trace agrees for first", first_bad, "transitions; n_leap", res["run"][i, rk]["n_leapfrog"],
              reso["run"][i, rk]["n_leapfrog"], "eps %.4f %.4f" % (res["run"][i, rk]["step_size"], reso["run"][i, rk]["step_size"]))
for f in ("D_max", "n_sigma", "q_mean", "concentration_mean", "D_max_marginalized_mean", "asymmetry", "n_sigma_forward", "D_max_forward", "D_max_reverse"):
    print("  ", f, res[f][:4], reso[f][:4])

t = time.time()
out = ctx.fit_batch(tid, k, N, cfg)
dt = time.time() - t
tm = ctx.timings()
print("fit %d taxa: %.3fs -> %.1f fits/s" % (len(tid), dt, len(tid) / dt), tm)
print("status counts", np.unique(out["result"]["status"], return_counts=True))
