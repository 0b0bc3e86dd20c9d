"""N4: dashboard-facing lookups served from arrays. tests/golden/lookups_golden.npz was produced by
the REFERENCE's `dashboard.fit_results.FitResults` (its pandas half, loaded stand-alone by
tests/golden/make_golden.py) reading result files written by this package's `io.Parquet`; the same files travel
inside the fixture. `metadamage_b200.lookups.ResultArrays` — a structure of arrays with mask filters and offset
slices, no DataFrame queries — must give every derived column, range, marker size, filter result and
single-TaxID fetch the reference's class gives."""
import io as _io
import json
import os
import zipfile

import numpy as np
import pytest

GOLDEN = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lookups_golden.npz"))


@pytest.fixture(scope="module")
def store(tmp_path_factory):
    from metadamage_b200 import lookups

    folder = tmp_path_factory.mktemp("results")
    with zipfile.ZipFile(_io.BytesIO(GOLDEN["files_zip"].tobytes())) as z:
        z.extractall(folder)
    return lookups.ResultArrays.from_folder(folder)


def test_derived_columns_and_marker_sizes(store):
    assert np.array_equal(store.categorical["tax_id"].values().astype(np.int64), GOLDEN["tax_id_order"])
    for c in ("N_alignments_log10", "N_alignments_sqrt", "N_sum_total_log10", "size"):
        assert np.array_equal(np.asarray(store.numeric[c], np.float64), GOLDEN[f"col_{c}"], equal_nan=True), c
    assert store.max_of_size == GOLDEN["max_of_size"] and store.marker_size_max == 30
    for tr in ("identity", "log10", "constant"):
        store.set_marker_size(tr, 12)
        assert np.array_equal(store.numeric["size"], GOLDEN[f"size_{tr}"])
        assert store.max_of_size == GOLDEN[f"max_of_size_{tr}"] and store.marker_size_max == 12
    store.set_marker_size("sqrt")
    assert store.set_marker_size([], []) is None
    with pytest.raises(AssertionError):
        store.set_marker_size("cube")
    assert sorted(store.categorical["shortname"].categories) == ["sampleA", "sampleB"] and len(store.categorical["tax_id"].categories) == 65


def test_ranges(store):
    ranges = store.ranges()
    keys = sorted(ranges)
    assert keys == [str(k) for k in GOLDEN["range_keys"]]
    got = np.array([ranges[k] for k in keys], dtype=np.float64)
    assert np.array_equal(got, GOLDEN["range_values"])


def test_filters_and_single_taxid_fetches(store):
    filters = json.loads(str(GOLDEN["filters_json"]))
    assert len(filters) == int(GOLDEN["n_filters"])
    for i, f in enumerate(filters):
        f = {k: (tuple(v) if isinstance(v, list) and k not in ("shortnames", "tax_ids", "tax_ranks", "tax_names") else v) for k, v in f.items()}
        assert np.array_equal(store.select(f), GOLDEN[f"filter{i}_index"]), (i, f)
    tax = int(GOLDEN["single_pred_tax"])
    pred = store.prediction("sampleA", tax)
    assert np.array_equal(np.asarray(pred["median"], np.float64), GOLDEN["single_pred_median"]) and len(pred["median"]) == 30
    assert list(pred["position"]) == list(range(1, 16)) + list(range(-1, -16, -1))
    grp = store.counts_group("sampleA", tax)
    assert len(grp["position"]) == int(GOLDEN["single_count_rows"]) == 30 and set(np.asarray(grp["tax_id"]).astype(int)) == {tax}
    assert len(store.prediction("sampleA", -12345)["median"]) == 0 and len(store.prediction("nope", tax)["median"]) == 0


def test_from_fit_rows_without_files():
    """The same lookups straight from mdg_fit_result rows (no parquet round trip)."""
    from metadamage_b200 import lookups
    from metadamage_b200._abi import FIT_RESULT_DTYPE

    n = 6
    res = np.zeros(n, FIT_RESULT_DTYPE)
    res["D_max"] = np.linspace(0.1, 0.6, n)
    res["n_sigma"] = np.arange(n)
    res["N_sum_total"] = 1000 * (1 + np.arange(n))
    res["status"][2] = 1  # failed fit: dropped
    dense = dict(tax_id=np.arange(10, 10 + n), tax_name=np.array(["a", "b", "c", "d", "e", "f"], dtype=object),
                 tax_rank=np.array(["species"] * n, dtype=object), N_alignments=np.array([10, 100, 1000, 10, 100, 1000], np.uint32))
    med = np.arange(n * 30, dtype=np.float32).reshape(n, 30)
    st = lookups.ResultArrays.from_fit(res, dense, "s1", med, med, med)
    assert st.n == 5 and 12 not in st.categorical["tax_id"].values()
    assert list(st.select({"N_alignments": (2, 3)})) == [1, 3, 4]         # exponents of ten: 100 .. 1000, the failed fit gone
    assert list(st.select({"tax_names": ["a", "f"], "n_sigma": (0, 10)})) == [0, 4]
    assert np.array_equal(st.prediction("s1", 13)["median"], med[3])
