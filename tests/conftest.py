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
