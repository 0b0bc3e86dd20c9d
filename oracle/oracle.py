"""ctypes wrapper of oracle/libmdg_oracle.so — the CPU restatement of the reference path.

TEST INFRASTRUCTURE ONLY (see the header of mdg_oracle.c): imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, never by the
product package `metadamage_b200`.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from metadamage_b200._abi import (
    FIT_RESULT_DTYPE,
    NUM_RUNS,
    FitConfig,
    ptr,
)

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libmdg_oracle.so")
_lib = None


def build(force=False):
    """Compile the oracle with gcc (recipe: oracle/Makefile)."""
    src = os.path.join(_HERE, "mdg_oracle.c")
    hdr = os.path.join(_HERE, "..", "include", "mdg.h")
    if not force and os.path.exists(_LIB_PATH):
        newest = max(os.path.getmtime(src), os.path.getmtime(hdr))
        if os.path.getmtime(_LIB_PATH) >= newest:
            return _LIB_PATH
    subprocess.run(["make", "-C", _HERE, "-B", "-s"], check=True)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_lgamma.restype = C.c_double
        _lib.orc_lgamma.argtypes = [C.c_double]
        _lib.orc_digamma.restype = C.c_double
        _lib.orc_digamma.argtypes = [C.c_double]
        _lib.orc_n_sigma.restype = C.c_double
    return _lib


def default_config(**changes):
    cfg = FitConfig()
    lib().orc_fit_config_default(C.byref(cfg))
    return cfg.copy(**changes) if changes else cfg


def num_threads():
    return int(lib().orc_num_threads())


def philox(key2, ctr4):
    key = np.ascontiguousarray(key2, dtype=np.uint32)
    ctr = np.ascontiguousarray(ctr4, dtype=np.uint32)
    out = np.zeros(4, dtype=np.uint32)
    lib().orc_philox4x32(ptr(key), ptr(ctr), ptr(out))
    return out


def make_key(seed, tax_id):
    out = np.zeros(2, dtype=np.uint32)
    lib().orc_make_key(C.c_uint64(seed), C.c_int64(tax_id), ptr(out))
    return out


def lgamma(x):
    return np.array([lib().orc_lgamma(float(v)) for v in np.atleast_1d(x)])


def digamma(x):
    return np.array([lib().orc_digamma(float(v)) for v in np.atleast_1d(x)])


def tsv_parse(text):
    """Restatement of the tokenising half of dd.read_csv (counts.py:229-235) on bytes."""
    text = bytes(text)
    cap = text.count(b"\n") + 1
    out = dict(tax_id=np.zeros(cap, np.int64), n_alignments=np.zeros(cap, np.uint32), is_reverse=np.zeros(cap, np.uint8),
               pos0=np.zeros(cap, np.uint8), counts16=np.zeros((16, cap), np.uint32))
    n_rows, n_cols = C.c_int64(0), C.c_int32(0)
    rc = lib().orc_tsv_parse(text, C.c_int64(len(text)), C.c_int64(cap), ptr(out["tax_id"]), ptr(out["n_alignments"]),
                             ptr(out["is_reverse"]), ptr(out["pos0"]), ptr(out["counts16"]), C.c_int64(cap),
                             C.byref(n_rows), C.byref(n_cols))
    if rc != 0:
        raise RuntimeError(f"orc_tsv_parse failed with {rc}")
    n = n_rows.value
    res = {k: v[:n] for k, v in out.items() if k != "counts16"}
    res["counts16"] = np.ascontiguousarray(out["counts16"][:, :n])
    res["n_rows"], res["n_cols"] = n, n_cols.value
    return res


def counts_reduce(tax_id, n_alignments, is_reverse, pos0, counts16, fwd="CT", rev="GA",
                  max_position=15, min_alignments=10, min_y_sum=10, want_noise=True):
    """Restatement of counts.py:237-256 on SoA arrays; returns a dict of numpy arrays."""
    base = "ACGT"
    n = len(tax_id)
    P = max_position
    tax_id = np.ascontiguousarray(tax_id, dtype=np.int64)
    n_alignments = np.ascontiguousarray(n_alignments, dtype=np.uint32)
    is_reverse = np.ascontiguousarray(is_reverse, dtype=np.uint8)
    pos0 = np.ascontiguousarray(pos0, dtype=np.uint8)
    counts16 = np.ascontiguousarray(counts16, dtype=np.uint32)
    assert counts16.shape == (16, n)
    out = dict(
        n_fwd_ref=np.zeros(n, np.uint32), n_rev_ref=np.zeros(n, np.uint32),
        f_fwd=np.zeros(n, np.float32), f_rev=np.zeros(n, np.float32),
        z=np.zeros(n, np.int8), y_sum_total=np.zeros(n, np.uint64), keep=np.zeros(n, np.uint8),
        tax_id=np.zeros(n, np.int64), n_alignments=np.zeros(n, np.uint32),
        first_row=np.zeros(n, np.int64),
        k=np.zeros((n, 2 * P), np.uint32), N=np.zeros((n, 2 * P), np.uint32),
        noise=np.zeros((n, 3), np.float64) if want_noise else None,
    )
    n_tax = C.c_int64(0)
    rc = lib().orc_counts_reduce(
        C.c_int64(n), ptr(tax_id), ptr(n_alignments), ptr(is_reverse), ptr(pos0),
        ptr(counts16), C.c_int64(n),
        base.index(fwd[0]), base.index(fwd[1]), base.index(rev[0]), base.index(rev[1]),
        C.c_int(P), C.c_uint32(min_alignments), C.c_uint64(min_y_sum),
        ptr(out["n_fwd_ref"]), ptr(out["n_rev_ref"]), ptr(out["f_fwd"]), ptr(out["f_rev"]),
        ptr(out["z"]), ptr(out["y_sum_total"]), ptr(out["keep"]),
        ptr(out["tax_id"]), ptr(out["n_alignments"]), ptr(out["first_row"]),
        ptr(out["k"]), ptr(out["N"]), ptr(out["noise"]), C.byref(n_tax))
    if rc != 0:
        raise RuntimeError(f"orc_counts_reduce failed with {rc}")
    m = n_tax.value
    for key in ("tax_id", "n_alignments", "first_row", "k", "N", "noise"):
        if out[key] is not None:
            out[key] = out[key][:m].copy()
    out["n_tax"] = m
    return out


def select_top(tax_id_row, n_alignments_row, keep_row, tax_id, first_row, n_top):
    """fits.py:736-744 extract_top_max_fits restated on the arrays of counts_reduce: positions (ascending) in
    the per-TaxID arrays of the n_top TaxIDs with the largest sum of N_alignments over their kept rows; the
    groupby result is indexed by sorted tax_id and nlargest keeps the first among ties, i.e. the smaller id.
    Returns (index, weight)."""
    tax_id_row = np.asarray(tax_id_row, dtype=np.int64)
    nal = np.asarray(n_alignments_row, dtype=np.uint64)
    keep = np.ones(len(tax_id_row), bool) if keep_row is None else np.asarray(keep_row).astype(bool)
    tax_id = np.asarray(tax_id, dtype=np.int64)
    first_row = np.asarray(first_row, dtype=np.int64)
    weight = np.zeros(len(tax_id), np.uint64)
    for t, r0 in enumerate(first_row):  # rows of a TaxID are contiguous
        r1 = r0
        while r1 < len(tax_id_row) and tax_id_row[r1] == tax_id_row[r0]:
            r1 += 1
        weight[t] = nal[r0:r1][keep[r0:r1]].sum()
    order = sorted(range(len(tax_id)), key=lambda t: (-int(weight[t]), int(tax_id[t])))
    return np.array(sorted(order[:max(0, int(n_top))]), dtype=np.int64), weight


def logp_grad(k, N, u, cfg=None, model=0, lane_mask=0, with_jacobian=True, max_position=None):
    """log joint + gradient wrt unconstrained u for one TaxID; u: [n_eval][4]."""
    cfg = cfg or default_config()
    k = np.ascontiguousarray(k, dtype=np.uint32).ravel()
    N = np.ascontiguousarray(N, dtype=np.uint32).ravel()
    P = max_position or len(k) // 2
    u = np.ascontiguousarray(np.atleast_2d(u), dtype=np.float64)
    if u.shape[1] != 4:
        u = np.ascontiguousarray(np.pad(u, ((0, 0), (0, 4 - u.shape[1]))))
    ne = u.shape[0]
    logp = np.zeros(ne)
    grad = np.zeros((ne, 4))
    ll = np.zeros((ne, 2 * P))
    lib().orc_test_logp_grad(C.c_int(P), ptr(k), ptr(N), C.byref(cfg), C.c_int(model), C.c_int(lane_mask),
                             C.c_int(int(with_jacobian)), C.c_int64(ne), ptr(u), ptr(logp), ptr(grad), ptr(ll))
    return logp, grad, ll


def map_fit(k, N, cfg=None, model=0):
    cfg = cfg or default_config()
    k = np.ascontiguousarray(k, dtype=np.uint32).ravel()
    N = np.ascontiguousarray(N, dtype=np.uint32).ravel()
    theta = np.zeros(4)
    logp = C.c_double(0)
    iters = C.c_int(0)
    conv = C.c_int(0)
    lib().orc_test_map(C.c_int(len(k) // 2), ptr(k), ptr(N), C.byref(cfg), C.c_int(model), ptr(theta),
                       C.byref(logp), C.byref(iters), C.byref(conv))
    return dict(q=theta[0], A=theta[1], c=theta[2], phi=theta[3], logp=logp.value,
                iters=iters.value, converged=bool(conv.value))


def nuts_run(k, N, tax_id, run_kind, cfg=None, want_trace=False):
    cfg = cfg or default_config()
    k = np.ascontiguousarray(k, dtype=np.uint32).ravel()
    N = np.ascontiguousarray(N, dtype=np.uint32).ravel()
    S, W = cfg.num_samples, cfg.num_warmup
    samples = np.zeros((S, 4))
    trace = np.zeros((W + S, 4)) if want_trace else None
    step = C.c_double(0)
    acc = C.c_double(0)
    ngrad = C.c_uint64(0)
    rc = lib().orc_test_nuts_run(C.c_int(len(k) // 2), ptr(k), ptr(N), C.byref(cfg), C.c_int64(tax_id),
                                 C.c_int(run_kind), ptr(samples), ptr(trace), C.byref(step), C.byref(acc),
                                 C.byref(ngrad))
    return dict(rc=rc, samples=samples, trace=trace, step_size=step.value, mean_accept=acc.value,
                n_grad=ngrad.value)


def fit_batch(tax_id, k, N, cfg=None, mism12=None, noise3=None, want_samples=False, want_trace=False,
              want_waic=False, n_threads=0):
    """Restatement of fits.py:428-469 for a dense batch; mirrors mdg_fit_batch."""
    cfg = cfg or default_config()
    tax_id = np.ascontiguousarray(tax_id, dtype=np.int64)
    k = np.ascontiguousarray(k, dtype=np.uint32)
    N = np.ascontiguousarray(N, dtype=np.uint32)
    n_tax, R = k.shape
    P = R // 2
    S, W = cfg.num_samples, cfg.num_warmup
    res = np.zeros(n_tax, dtype=FIT_RESULT_DTYPE)
    med = np.zeros((n_tax, R), np.float32)
    lo = np.zeros((n_tax, R), np.float32)
    hi = np.zeros((n_tax, R), np.float32)
    samples = np.full((n_tax, NUM_RUNS, S, 4), np.nan) if want_samples else None
    trace = np.full((n_tax, NUM_RUNS, W + S, 4), np.nan) if want_trace else None
    waic = np.zeros((n_tax, NUM_RUNS, 2, R)) if want_waic else None
    if mism12 is not None:
        mism12 = np.ascontiguousarray(mism12, dtype=np.uint32)
    if noise3 is not None:
        noise3 = np.ascontiguousarray(noise3, dtype=np.float64)
    rc = lib().orc_fit_batch(C.c_int64(n_tax), C.c_int(P), ptr(tax_id), ptr(k), ptr(N), ptr(mism12), ptr(noise3),
                             C.byref(cfg), ptr(res), ptr(med), ptr(lo), ptr(hi), ptr(samples), ptr(trace),
                             ptr(waic), C.c_int(n_threads))
    if rc != 0:
        raise RuntimeError(f"orc_fit_batch failed with {rc}")
    return dict(result=res, median=med, hpdi_lo=lo, hpdi_hi=hi, samples=samples, trace=trace, waic=waic)


def noise(mism12, max_position):
    m = np.ascontiguousarray(mism12, dtype=np.uint32)
    out = np.zeros(3)
    lib().orc_noise(C.c_int(max_position), ptr(m), ptr(out))
    return out


def n_sigma(waic_i_pmd, waic_i_null):
    a = np.ascontiguousarray(waic_i_pmd, dtype=np.float64)
    b = np.ascontiguousarray(waic_i_null, dtype=np.float64)
    return float(lib().orc_n_sigma(C.c_int(len(a)), ptr(a), ptr(b)))


def median_hpdi(values, prob=0.68):
    v = np.array(values, dtype=np.float64)
    med, lo, hi = C.c_double(0), C.c_double(0), C.c_double(0)
    lib().orc_median_hpdi(ptr(v), C.c_int(len(v)), C.c_double(prob), C.byref(med), C.byref(lo), C.byref(hi))
    return med.value, lo.value, hi.value


def betabinom_draws(alpha, beta, n, n_draws, seed=1):
    out = np.zeros(n_draws)
    lib().orc_test_betabinom_draws(C.c_uint64(seed), C.c_int64(n_draws), C.c_double(alpha), C.c_double(beta),
                                   C.c_double(n), ptr(out))
    return out


def adaptation_schedule(num_steps):
    s = (C.c_int * 16)()
    e = (C.c_int * 16)()
    n = lib().orc_test_schedule(C.c_int(num_steps), s, e)
    return [(s[i], e[i]) for i in range(n)]
