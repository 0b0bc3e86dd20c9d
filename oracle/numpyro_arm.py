"""Run-time probe for the REAL reference (numpyro + jax) and, if it is there, the reference arm itself.

TEST INFRASTRUCTURE ONLY, like everything under oracle/: used by `bench.py --impl reference`, by bench.py's
`cpu_baseline` leg and by tests/. BASELINE.md 3.1 / SURVEY.md 8(c),(d): numpyro 0.4.1 (pin `^0.4.1`,
/root/reference/pyproject.toml:17) and jax are not installable offline, so the normal outcome of `probe()`
is (None, reason) and the C restatement (mdg_oracle.c) is the CPU baseline. Should a box ever carry them
(site-packages or a driver-installed `baseline/_ref/`), this module runs the reference's OWN
`fits.fit_single_group_without_timeout` (fits.py:428-469) on dense (tax_id, k, N) rows through the
reference's own `MCMC(NUTS(model))` objects (fits.py:382-387, kwargs of fits.py:792-799) and returns rows
that the parity tests compare with the CUDA rows at 3 x MCSE.
"""
import importlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "baseline", "_ref")


def probe():
    """(numpyro module, None) if `import numpyro, jax` works (site-packages or baseline/_ref), else (None, why)."""
    added = False
    if os.path.isdir(REF_DIR) and REF_DIR not in sys.path:
        sys.path.append(REF_DIR)
        added = True
    try:
        import jax  # noqa: F401
        import numpyro

        return numpyro, None
    except Exception as exc:  # ImportError, or a jaxlib that does not load on this box
        if added:
            sys.path.remove(REF_DIR)
        return None, f"{type(exc).__name__}: {exc}"


def load_reference_fits():
    """The reference's own `metadamage.fits` module from baseline/_ref (or wherever `metadamage.fits` with a
    `fit_single_group_without_timeout` resolves), imported beside this repo's `metadamage` re-export package
    without disturbing it. Returns the module or raises ImportError."""
    saved = {k: v for k, v in sys.modules.items() if k == "metadamage" or k.startswith("metadamage.")}
    for k in saved:
        del sys.modules[k]
    old_path = list(sys.path)
    try:
        sys.path = [REF_DIR] + [p for p in old_path if os.path.abspath(p or ".") != ROOT]
        mod = importlib.import_module("metadamage.fits")
        if not hasattr(mod, "fit_single_group_without_timeout") or not hasattr(mod, "model_PMD"):
            raise ImportError("metadamage.fits found, but it is not the reference's module")
        return mod
    finally:
        sys.path = old_path
        for k in [k for k in sys.modules if k == "metadamage" or k.startswith("metadamage.")]:
            if k not in saved:
                # keep the reference's modules reachable under a private prefix only
                sys.modules["_reference_" + k] = sys.modules.pop(k)
        sys.modules.update(saved)


class _Cfg:
    """The two attributes of utils.Config that fits.group_to_numpyro_data reads (fits.py:398-419)."""

    def __init__(self, fwd="CT", rev="GA"):
        self.substitution_bases_forward = fwd
        self.substitution_bases_reverse = rev


def _group_frame(tax_id, k, N, fwd="CT", rev="GA"):
    """One TaxID's df_counts rows in the reference's layout (z = +1..+P then -1..-P, counts.py:167-172) with
    the columns fits.py:398-419, 272-283, 359-376 read. Only P = 15 reproduces the reference (fits.py:405-411)."""
    import pandas as pd

    P = len(k) // 2
    z = np.concatenate([np.arange(1, P + 1), -np.arange(1, P + 1)])
    cols = {"tax_id": tax_id, "tax_name": "", "tax_rank": "", "N_alignments": int(N.max()), "position": z}
    df = pd.DataFrame(cols)
    for r in "ACGT":
        for o in "ACGT":
            df[r + o] = 0
    df[fwd] = np.where(z > 0, k, 0)
    df[rev] = np.where(z < 0, k, 0)
    df[fwd[0]] = np.where(z > 0, N, 0)
    df[rev[0]] = np.where(z < 0, N, 0)
    return df


def fit_rows(tax_id, k, N, fwd="CT", rev="GA", mcmc_kwargs=None):
    """The reference's per-TaxID fit on dense rows. Returns (list of fit_result dicts, seconds of the first fit
    [includes the jit], seconds of all the others)."""
    fits = load_reference_fits()
    kw = dict(progress_bar=False, num_warmup=500, num_samples=1000, num_chains=1, chain_method="sequential")
    kw.update(mcmc_kwargs or {})
    mcmcs = [fits.init_mcmc(fits.model_PMD, **kw), fits.init_mcmc(fits.model_null, **kw),
             fits.init_mcmc(fits.model_PMD, **kw), fits.init_mcmc(fits.model_null, **kw)]
    cfg = _Cfg(fwd, rev)
    rows, t_first, t_rest = [], 0.0, 0.0
    for i in range(len(tax_id)):
        group = _group_frame(int(tax_id[i]), np.asarray(k[i], dtype=np.int64), np.asarray(N[i], dtype=np.int64), fwd, rev)
        t0 = time.perf_counter()
        d = fits.fit_single_group_without_timeout(group, cfg, *mcmcs)
        dt = time.perf_counter() - t0
        if i == 0:
            t_first = dt
        else:
            t_rest += dt
        rows.append(d["fit_result"])
    return rows, t_first, t_rest
