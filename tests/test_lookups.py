"""N4: dashboard-facing lookups. tests/golden/lookups_golden.npz was produced by the
REFERENCE's `dashboard.fit_results.FitResults` (its pandas half, loaded stand-alone by
tests/golden/make_golden.py) reading result files written by this package's `io.Parquet`; the
same files travel inside the fixture. `metadamage_b200.lookups.FitResults` must reproduce every
derived column, range, marker size, filter result and single-TaxID fetch."""
import io as _io
import json
import os
import zipfile

import numpy as np
import pytest

GOLDEN = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lookups_golden.npz"))


@pytest.fixture(scope="module")
def fit_results(tmp_path_factory):
    from metadamage_b200 import lookups

    folder = tmp_path_factory.mktemp("results")
    with zipfile.ZipFile(_io.BytesIO(GOLDEN["files_zip"].tobytes())) as z:
        z.extractall(folder)
    return lookups.FitResults(folder)


def test_derived_columns_and_marker_sizes(fit_results):
    d = fit_results.df_fit_results
    assert np.array_equal(d["tax_id"].to_numpy(np.int64), GOLDEN["tax_id_order"])
    for c in ("N_alignments_log10", "N_alignments_sqrt", "N_sum_total_log10", "size"):
        assert np.array_equal(d[c].to_numpy(np.float64), GOLDEN[f"col_{c}"], equal_nan=True), c
    assert fit_results.max_of_size == GOLDEN["max_of_size"] and fit_results.marker_size_max == 30
    for tr in ("identity", "log10", "constant"):
        fit_results.set_marker_size(tr, 12)
        assert np.array_equal(fit_results.df_fit_results["size"].to_numpy(np.float64), GOLDEN[f"size_{tr}"])
        assert fit_results.max_of_size == GOLDEN[f"max_of_size_{tr}"] and fit_results.marker_size_max == 12
    fit_results.set_marker_size("sqrt")
    assert fit_results.set_marker_size([], []) is None
    with pytest.raises(AssertionError):
        fit_results.set_marker_size("cube")
    assert sorted(fit_results.shortnames) == ["sampleA", "sampleB"] and len(fit_results.all_tax_ids) == 65


def test_ranges(fit_results):
    keys = sorted(fit_results.ranges)
    assert keys == [str(k) for k in GOLDEN["range_keys"]]
    got = np.array([fit_results.ranges[k] for k in keys], dtype=np.float64)
    assert np.array_equal(got, GOLDEN["range_values"])


def test_filters_and_single_taxid_fetches(fit_results):
    filters = json.loads(str(GOLDEN["filters_json"]))
    assert len(filters) == int(GOLDEN["n_filters"])
    for i, f in enumerate(filters):
        f = {k: (tuple(v) if isinstance(v, list) and k not in ("shortnames", "tax_ids", "tax_ranks", "tax_names") else v) for k, v in f.items()}
        assert np.array_equal(fit_results.filter(f).index.to_numpy(np.int64), GOLDEN[f"filter{i}_index"]), (i, f)
    with pytest.raises(AssertionError):
        fit_results.filter({"tax_id": 1}, df_type="df_counts")
    tax = int(GOLDEN["single_pred_tax"])
    pred = fit_results.get_single_fit_prediction("sampleA", tax)
    assert np.array_equal(pred["median"].to_numpy(np.float64), GOLDEN["single_pred_median"]) and len(pred) == 30
    assert list(pred["position"]) == list(range(1, 16)) + list(range(-1, -16, -1))
    assert len(fit_results.get_single_count_group("sampleA", tax)) == int(GOLDEN["single_count_rows"]) == 30
