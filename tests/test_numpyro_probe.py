"""The run-time probe for the real reference: `import numpyro, jax` from
site-packages or a driver-installed baseline/_ref. Offline the probe must say why it failed (and bench.py then
labels the C port); if a box ever carries the modules, the reference's own fit_single_group_without_timeout is run
on 200 cfg2 TaxIDs and compared with the CUDA rows at 3 x MCSE."""
import numpy as np
import pytest

from oracle import numpyro_arm


def test_probe_reports_a_reason_or_a_module():
    mod, why = numpyro_arm.probe()
    assert (mod is None) != (why is None)
    if mod is None:
        assert "jax" in why or "numpyro" in why or "Error" in why


@pytest.mark.gpu
def test_cuda_rows_match_numpyro_when_it_is_installed(ctx):
    mod, why = numpyro_arm.probe()
    if mod is None:
        pytest.skip(f"numpyro / jax not importable here ({why}): the C restatement stays the checker")
    from metadamage_b200 import _lib, synthetic as syn

    tid, k, N, _ = syn.dense_fit_batch(200)
    rows, _, _ = numpyro_arm.fit_rows(tid, k, N)
    got = ctx.fit_batch(tid, k, N, _lib.default_config())["result"]
    ok = (got["status"] & 1) == 0
    for name, std in (("D_max_marginalized_mean", "D_max_marginalized_std"), ("q_mean", "q_std"), ("concentration_mean", "concentration_std")):
        ref = np.array([r[name] for r in rows])
        # MCSE of a mean from 1000 draws with an effective sample size of at least ~250 on either side
        mcse = np.hypot(got[std], got[std]) / np.sqrt(250.0)
        z = (got[name] - ref)[ok] / np.maximum(mcse[ok], 1e-12)
        assert np.mean(np.abs(z) > 3.0) <= 0.02 and abs(np.mean(z)) < 0.3, (name, np.mean(z), np.std(z))
