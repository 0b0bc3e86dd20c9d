import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU restatement (test infrastructure). Built on demand with gcc."""
    from oracle import oracle as O

    O.build()
    return O


@pytest.fixture(scope="session")
def sample_inputs():
    """SoA columns parsed from the reference's data/input/data_{ancient,control}.txt."""
    z = np.load(os.path.join(GOLDEN, "sample_inputs.npz"))
    out = {}
    for name in ("ancient", "control"):
        out[name] = {k: z[f"{name}_{k}"] for k in ("tax_id", "n_alignments", "is_reverse", "pos0", "counts16")}
    return out


@pytest.fixture(scope="session")
def counts_golden():
    return np.load(os.path.join(GOLDEN, "counts_golden.npz"))


@pytest.fixture(scope="session")
def fits_golden():
    return np.load(os.path.join(GOLDEN, "fits_golden.npz"))


@pytest.fixture(scope="session")
def ctx():
    """One GPU context for the -m gpu tests. Fails loudly if the CUDA library or GPU is missing."""
    from metadamage_b200.backend import Context

    c = Context(0)
    yield c
    c.close()


def mcse_batch_means(x, n_batches=20):
    """Monte-Carlo standard error of mean(x) by batch means."""
    x = np.asarray(x, dtype=np.float64)
    n = (len(x) // n_batches) * n_batches
    means = x[:n].reshape(n_batches, -1).mean(axis=1)
    return means.std(ddof=1) / np.sqrt(n_batches)


def null_posterior_quadrature(k, N, n_grid=500):
    """Ground truth for the 2-parameter null model (fits.py:62-67) by brute-force quadrature, no sampler:
    q ~ Beta(2,3), delta ~ Exponential(mean 1000), phi = delta + 2, y_z ~ BetaBinomial(q phi, (1-q) phi, N_z),
    with scipy's beta-binomial pmf. Returns posterior mean and variance of q and of log(delta)."""
    from scipy import stats
    from scipy.special import logsumexp

    k = np.asarray(k, dtype=np.int64)
    N = np.asarray(N, dtype=np.int64)

    def log_post(u, v):  # u = logit(q), v = log(delta); density in (u, v) coordinates
        q = 1.0 / (1.0 + np.exp(-u))
        d = np.exp(v)
        phi = d + 2.0
        lp = stats.beta.logpdf(q, 2, 3) + np.log(q) + np.log1p(-q)  # + log |dq/du|
        lp = lp + stats.expon.logpdf(d, scale=1000.0) + v           # + log |d delta/dv|
        ll = stats.betabinom.logpmf(k[:, None, None], N[:, None, None], (q * phi)[None], ((1 - q) * phi)[None]).sum(0)
        return lp + ll

    # two passes: coarse box, then +-8 sd around the mode
    lo_u, hi_u, lo_v, hi_v = -12.0, 6.0, -6.0, 14.0
    for _ in range(3):
        u = np.linspace(lo_u, hi_u, n_grid)
        v = np.linspace(lo_v, hi_v, n_grid)
        U, V = np.meshgrid(u, v, indexing="ij")
        lp = log_post(U, V)
        w = np.exp(lp - logsumexp(lp))
        mu, mv = (w * U).sum(), (w * V).sum()
        su, sv = np.sqrt((w * (U - mu) ** 2).sum()), np.sqrt((w * (V - mv) ** 2).sum())
        lo_u, hi_u = max(mu - 9 * su, -30.0), min(mu + 9 * su, 30.0)
        lo_v, hi_v = max(mv - 9 * sv, -30.0), min(mv + 9 * sv, 25.0)
    q = 1.0 / (1.0 + np.exp(-U))
    mean_q = (w * q).sum()
    return {"mean_q": mean_q, "var_q": (w * (q - mean_q) ** 2).sum(), "mean_logdelta": mv, "var_logdelta": sv ** 2}
