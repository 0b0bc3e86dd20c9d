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
    # exact lppd_i = log E[p(y_i | theta)] and pWAIC_i = Var[log p(y_i | theta)] (fits.py:147-165 in the limit S -> inf)
    phi = np.exp(V) + 2.0
    ll = stats.betabinom.logpmf(k[:, None, None], N[:, None, None], (q * phi)[None], ((1 - q) * phi)[None])
    lppd_i = logsumexp(ll + np.log(np.maximum(w, 1e-300))[None], axis=(1, 2))
    m1 = (w[None] * ll).sum(axis=(1, 2))
    pwaic_i = (w[None] * (ll - m1[:, None, None]) ** 2).sum(axis=(1, 2))
    return {"mean_q": mean_q, "var_q": (w * (q - mean_q) ** 2).sum(), "mean_logdelta": mv, "var_logdelta": sv ** 2,
            "lppd_i": lppd_i, "pwaic_i": pwaic_i}


_PMD_QUAD_CACHE = {}


def pmd_posterior_quadrature(k, N, n_grid=((16, 16, 16, 16, 20), (20, 28))):
    """Ground truth for the 4-parameter PMD model (fits.py:43-59) by brute-force quadrature, no sampler:
    q ~ Beta(2,3), A ~ Beta(2,3), c ~ Beta(1,9), delta ~ Exponential(mean 1000), phi = delta + 2,
    D_z = A (1-q)^(|z|-1) + c, y_z ~ BetaBinomial(D_z phi, (1-D_z) phi, N_z); positions are the first half
    (forward) then the second half (reverse) of k/N, |z| - 1 = 0..P-1 in each, or one half only.
    Tensor trapezoid grids in the unconstrained coordinates (logit q, logit A, logit c, log delta) with the
    Jacobians: axis-aligned passes refined around the mass, then passes along the principal axes; the beta-binomial is scipy.special.betaln, checked
    against scipy.stats.betabinom on a random subset. Returns posterior means and variances of q, A, c,
    log delta and D_max = A + c, and the mass on the grid's boundary."""
    from scipy import special, stats
    from scipy.special import logsumexp

    k = np.asarray(k, dtype=np.float64)
    N = np.asarray(N, dtype=np.float64)
    key = (k.tobytes(), N.tobytes(), repr(n_grid))
    if key in _PMD_QUAD_CACHE:
        return _PMD_QUAD_CACHE[key]
    n_pos = len(k)
    half = n_pos // 2 if n_pos % 2 == 0 and n_pos > 15 else n_pos
    x = np.arange(n_pos) % half  # |z| - 1
    log_c = special.gammaln(N + 1) - special.gammaln(k + 1) - special.gammaln(N - k + 1)

    def log_lik_i(U, i):
        q, A, c = (1.0 / (1.0 + np.exp(-u)) for u in U[:3])
        phi = np.exp(U[3]) + 2.0
        Dz = np.minimum(A * (1.0 - q) ** x[i] + c, 1.0 - 1e-16)
        a, b = Dz * phi, (1.0 - Dz) * phi
        return special.betaln(k[i] + a, N[i] - k[i] + b) - special.betaln(a, b) + log_c[i]

    def log_post(U):  # U: four broadcastable arrays
        q, A, c = (1.0 / (1.0 + np.exp(-u)) for u in U[:3])
        d = np.exp(U[3])
        phi = d + 2.0
        lp = stats.beta.logpdf(q, 2, 3) + np.log(q) + np.log1p(-q)
        lp = lp + stats.beta.logpdf(A, 2, 3) + np.log(A) + np.log1p(-A)
        lp = lp + stats.beta.logpdf(c, 1, 9) + np.log(c) + np.log1p(-c)
        lp = lp + stats.expon.logpdf(d, scale=1000.0) + U[3]
        lp = np.broadcast_to(lp, np.broadcast_shapes(*(u.shape for u in U))).copy()
        bad = np.zeros(lp.shape, bool)
        for i in range(n_pos):
            Dz = A * (1.0 - q) ** x[i] + c
            bad |= np.broadcast_to(Dz >= 1.0, lp.shape)  # clip(Dz, 0, 1) -> beta = 0 -> NaN in the reference: no mass
            Dz = np.minimum(Dz, 1.0 - 1e-16)
            a, b = Dz * phi, (1.0 - Dz) * phi
            lp += special.betaln(k[i] + a, N[i] - k[i] + b) - special.betaln(a, b) + log_c[i]
        lp[bad] = -np.inf
        return lp

    # the betaln form IS scipy's beta-binomial pmf
    rng = np.random.default_rng(0)
    for _ in range(20):
        i = rng.integers(n_pos)
        a, b = rng.uniform(0.01, 50), rng.uniform(0.5, 3000)
        ref = stats.betabinom.logpmf(k[i], N[i], a, b)
        assert abs(special.betaln(k[i] + a, N[i] - k[i] + b) - special.betaln(a, b) + log_c[i] - ref) < 1e-8 * max(1.0, abs(ref))

    # axis-aligned passes find the mass; the later passes use a grid along the principal axes of the
    # weighted covariance (q, A and c are strongly correlated at high coverage: an axis-aligned grid
    # would need its spacing below the CONDITIONAL widths)
    lo = np.array([-10.0, -12.0, -12.0, -5.0])
    hi = np.array([8.0, 6.0, 4.0, 13.0])
    mean, L = None, None
    for n in n_grid[0]:
        ax = [np.linspace(lo[j], hi[j], n) for j in range(4)]
        U = np.meshgrid(*ax, indexing="ij", sparse=True)
        lp = log_post(U)
        w = np.exp(lp - logsumexp(lp))
        mean = np.array([(w * U[j]).sum() for j in range(4)])
        dev = [U[j] - mean[j] for j in range(4)]
        cov = np.array([[(w * dev[i] * dev[j]).sum() for j in range(4)] for i in range(4)])
        space = (hi - lo) / (n - 1)
        sd = np.maximum(np.sqrt(np.diag(cov)), space)  # a peak between grid points: never shrink below the spacing
        lo = np.maximum(mean - 7.5 * sd, -30.0)
        hi = np.minimum(mean + 7.5 * sd, [30.0, 30.0, 30.0, 25.0])
    L = np.linalg.cholesky(cov + np.diag(space ** 2) * 0.25)
    for n in n_grid[1]:
        t = np.linspace(-7.5, 7.5, n)
        T = np.meshgrid(t, t, t, t, indexing="ij")
        U = [mean[j] + sum(L[j, i] * T[i] for i in range(j + 1)) for j in range(4)]
        lp = log_post(U)
        w = np.exp(lp - logsumexp(lp))
        mean = np.array([(w * U[j]).sum() for j in range(4)])
        dev = [U[j] - mean[j] for j in range(4)]
        cov = np.array([[(w * dev[i] * dev[j]).sum() for j in range(4)] for i in range(4)])
        L = np.linalg.cholesky(cov)
    edge = sum(w.take(0, axis=j).sum() + w.take(-1, axis=j).sum() for j in range(4))
    q, A, c = (1.0 / (1.0 + np.exp(-U[j])) for j in range(3))
    out = {"edge_mass": float(edge)}
    for name, v in (("q", q), ("A", A), ("c", c), ("logdelta", U[3]), ("D_max", A + c)):
        m = float((w * v).sum())
        out["mean_" + name] = m
        out["var_" + name] = float((w * (v - m) ** 2).sum())
    # exact lppd_i = log E[p(y_i | theta)] and pWAIC_i = Var[log p(y_i | theta)] (fits.py:147-165 in the limit S -> inf)
    logw = np.log(np.maximum(w, 1e-300))
    lppd_i, pwaic_i = np.empty(n_pos), np.empty(n_pos)
    for i in range(n_pos):
        ll = log_lik_i(U, i)
        lppd_i[i] = logsumexp(ll + logw)
        m1 = (w * ll).sum()
        pwaic_i[i] = (w * (ll - m1) ** 2).sum()
    out["lppd_i"], out["pwaic_i"] = lppd_i, pwaic_i
    out["_grid"] = (mean, L, log_post)  # for pmd_predictive_quadrature
    _PMD_QUAD_CACHE[key] = out
    return out


def n_sigma_by_quadrature(k, N):
    """The reference's D-max significance (fits.py:194-201) from the EXACT per-position WAIC terms of the
    PMD and the null posterior (quadrature, no sampler)."""
    pmd, null = pmd_posterior_quadrature(k, N), null_posterior_quadrature(k, N)
    waic_pmd = -2.0 * (pmd["lppd_i"] - pmd["pwaic_i"])
    waic_null = -2.0 * (null["lppd_i"] - null["pwaic_i"])
    n = len(waic_pmd)
    dse = np.sqrt(n * np.var(waic_pmd - waic_null))
    return {"n_sigma": (waic_null.sum() - waic_pmd.sum()) / dse, "waic_pmd": waic_pmd.sum(), "waic_null": waic_null.sum(),
            "D_max_mean": pmd["mean_D_max"], "D_max_std": np.sqrt(pmd["var_D_max"])}


def pmd_predictive_quadrature(k, N, pos=0, n=20, prob=0.68):
    """Exact posterior predictive of fits.py:89-120 at one position (default z = 1 forward, the reference's
    D_max): pmf(y) = E_posterior[BetaBinomial(y; N_pos, D phi, (1 - D) phi)] on a principal-axis grid of the
    PMD posterior; returns the median and the narrowest interval holding `prob` of the mass (what np.median
    and numpyro's hpdi converge to as the number of draws grows), as fractions y / N_pos, and the pmf."""
    from scipy import special
    from scipy.special import logsumexp

    post = pmd_posterior_quadrature(k, N)
    mean, L, log_post = post["_grid"]
    k = np.asarray(k, dtype=np.float64)
    N = np.asarray(N, dtype=np.float64)
    n_pos = len(k)
    half = n_pos // 2 if n_pos % 2 == 0 and n_pos > 15 else n_pos
    xz = pos % half
    t = np.linspace(-7.5, 7.5, n)
    T = np.meshgrid(t, t, t, t, indexing="ij")
    U = [mean[j] + sum(L[j, i] * T[i] for i in range(j + 1)) for j in range(4)]
    lp = log_post(U)
    logw = (lp - logsumexp(lp)).ravel()
    keep = logw > logw.max() - 40.0
    logw = logw[keep]
    q, A, c = (1.0 / (1.0 + np.exp(-U[j].ravel()[keep])) for j in range(3))
    phi = np.exp(U[3].ravel()[keep]) + 2.0
    D = np.minimum(A * (1.0 - q) ** xz + c, 1.0 - 1e-16)
    a, b = D * phi, (1.0 - D) * phi
    Np = int(N[pos])
    y = np.arange(Np + 1, dtype=np.float64)
    log_c = special.gammaln(Np + 1) - special.gammaln(y + 1) - special.gammaln(Np - y + 1)
    base = logw - special.betaln(a, b)
    log_pmf = lambda yy: logsumexp(base + special.betaln(yy + a, Np - yy + b)) + log_c[int(yy)]  # noqa: E731
    # support first (every `step`-th count), then every count inside it
    step = max(1, Np // 150)
    coarse = np.array([log_pmf(yy) for yy in y[::step]])
    live = np.flatnonzero(coarse > coarse.max() - 45.0)
    y_lo, y_hi = max(0, (live[0] - 1) * step), min(Np, (live[-1] + 1) * step)
    pmf = np.zeros(Np + 1)
    pmf[y_lo:y_hi + 1] = np.exp([log_pmf(yy) for yy in y[y_lo:y_hi + 1]])
    pmf /= pmf.sum()
    cdf = np.cumsum(pmf)
    median = int(np.searchsorted(cdf, 0.5))
    best = (Np + 1, 0, Np)
    c0 = np.r_[0.0, cdf]
    for lo in range(Np + 1):
        hi = int(np.searchsorted(cdf, c0[lo] + prob - 1e-12))
        if hi <= Np and hi - lo < best[0]:
            best = (hi - lo, lo, hi)
    m1 = (pmf * y).sum()
    return {"median": median / Np, "hpdi_lo": best[1] / Np, "hpdi_hi": best[2] / Np, "pmf": pmf, "N": Np,
            "sd_counts": float(np.sqrt((pmf * (y - m1) ** 2).sum()))}


def asymmetry_by_quadrature(k, N):
    """Exact n_sigma_forward, n_sigma_reverse (fits.py:317, 339) and asymmetry (fits.py:204-227, 352): the
    forward-only and reverse-only refits are the same models on one half of the positions."""
    k, N = np.asarray(k), np.asarray(N)
    P = len(k) // 2
    fwd, rev = n_sigma_by_quadrature(k[:P], N[:P]), n_sigma_by_quadrature(k[P:], N[P:])
    both = pmd_posterior_quadrature(k, N)
    pf, pr = pmd_posterior_quadrature(k[:P], N[:P]), pmd_posterior_quadrature(k[P:], N[P:])
    waic_all = -2.0 * (both["lppd_i"] - both["pwaic_i"])
    waic_fr = np.r_[-2.0 * (pf["lppd_i"] - pf["pwaic_i"]), -2.0 * (pr["lppd_i"] - pr["pwaic_i"])]
    dse = np.sqrt(len(waic_all) * np.var(waic_all - waic_fr))
    return {"n_sigma_forward": fwd["n_sigma"], "n_sigma_reverse": rev["n_sigma"], "asymmetry": (waic_fr.sum() - waic_all.sum()) / dse}
