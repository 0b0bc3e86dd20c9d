"""Oracle (CPU restatement) of the counts path vs goldens produced by the reference's own
counts.py functions (tests/golden/make_golden.py) and vs the SURVEY.md 8c known answers."""
import numpy as np
import pytest

from metadamage_b200.counts import reference_row_order

CASES = {
    "ancient_CT_GA": ("ancient", "CT", "GA", 10, 10),
    "control_CT_GA": ("control", "CT", "GA", 10, 10),
    "ancient_GA_CT": ("ancient", "GA", "CT", 10, 10),
    "control_cut": ("control", "CT", "GA", 200000, 7000),
    "ancient_same_ref": ("ancient", "CT", "CA", 10, 10),
}


def run_case(reduce_fn, sample_inputs, name):
    src, fwd, rev, min_al, min_y = CASES[name]
    s = sample_inputs[src]
    return reduce_fn(s["tax_id"], s["n_alignments"], s["is_reverse"], s["pos0"], s["counts16"], fwd=fwd, rev=rev,
                     max_position=15, min_alignments=min_al, min_y_sum=min_y), s


def check_against_golden(r, s, g, name, fwd, rev):
    """Compare kept rows, put into the reference's row order, with the reference DataFrame."""
    keep = r["keep"].astype(bool)
    order = reference_row_order(s["n_alignments"][keep], s["tax_id"][keep], r["z"][keep])
    pick = lambda a: np.asarray(a)[keep][order]  # noqa: E731
    assert np.array_equal(pick(s["tax_id"]), g[f"{name}__tax_id"])
    assert np.array_equal(pick(s["n_alignments"]), g[f"{name}__N_alignments"])
    assert np.array_equal(pick(r["z"]).astype(np.int64), g[f"{name}__position"])
    assert np.array_equal(pick(r["n_fwd_ref"]).astype(np.int64), g[f"{name}__n_fwd_ref"])
    assert np.array_equal(pick(r["n_rev_ref"]).astype(np.int64), g[f"{name}__n_rev_ref"])
    assert np.array_equal(pick(r["y_sum_total"]).astype(np.int64), g[f"{name}__y_sum_total"])
    # f is stored as float32 by the reference's downcast (utils.py:351-354): bit-exact
    assert np.array_equal(pick(r["f_fwd"]), g[f"{name}__f_fwd"].astype(np.float32))
    assert np.array_equal(pick(r["f_rev"]), g[f"{name}__f_rev"].astype(np.float32))
    bases = "ACGT"
    kf = s["counts16"][bases.index(fwd[0]) * 4 + bases.index(fwd[1])]
    kr = s["counts16"][bases.index(rev[0]) * 4 + bases.index(rev[1])]
    assert np.array_equal(pick(kf).astype(np.int64), g[f"{name}__k_fwd"])
    assert np.array_equal(pick(kr).astype(np.int64), g[f"{name}__k_rev"])


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_counts_match_reference_functions(oracle, sample_inputs, counts_golden, name):
    r, s = run_case(oracle.counts_reduce, sample_inputs, name)
    check_against_golden(r, s, counts_golden, name, CASES[name][1], CASES[name][2])


def test_oracle_counts_known_answers(oracle, sample_inputs):
    """SURVEY.md 8c table."""
    r, s = run_case(oracle.counts_reduce, sample_inputs, "ancient_CT_GA")
    assert r["n_tax"] == 3
    assert list(r["tax_id"]) == [0, 1, 2]
    assert list(r["n_alignments"]) == [57733890, 25055907, 24803642]
    assert [int(r["y_sum_total"][i * 30]) for i in range(3)] == [24017107, 11474326, 11860573]
    assert (r["k"][0, 0], r["N"][0, 0]) == (4784754, 11631013)
    assert (r["k"][0, 15], r["N"][0, 15]) == (4957870, 11735881)
    assert (r["k"][0, 14], r["N"][0, 14]) == (276617, 10458142)
    assert (r["k"][0, 29], r["N"][0, 29]) == (278800, 10483297)
    assert int(r["N"][0].sum()) == 315303241
    r, s = run_case(oracle.counts_reduce, sample_inputs, "control_CT_GA")
    assert [int(r["y_sum_total"][i * 30]) for i in range(3)] == [9282, 6817, 7511]
    assert (r["k"][2, 0], r["N"][2, 0]) == (1128, 35305)
    assert (r["k"][2, 29], r["N"][2, 29]) == (39, 11298)


def test_oracle_counts_cut_drops_taxa(oracle, sample_inputs, counts_golden):
    r, s = run_case(oracle.counts_reduce, sample_inputs, "control_cut")
    # N_alignments >= 200000 keeps only tax 0; y_sum >= 7000 holds for it (9282)
    assert r["n_tax"] == 1 and r["tax_id"][0] == 0
    assert r["keep"].sum() == 30


def test_oracle_noise_matches_reference(oracle, fits_golden):
    """fits.add_noise_estimates (fits.py:359-376) on all six sample TaxIDs."""
    for m12, expected in zip(fits_golden["noise_mism12"], fits_golden["noise_expected"]):
        got = oracle.noise(m12, 15)
        np.testing.assert_allclose(got, expected, rtol=1e-12)


def test_oracle_counts_edge_cases(oracle):
    """Empty input, single-row TaxIDs, ragged TaxIDs, positions beyond max_position, N = 0 rows."""
    e = np.zeros(0)
    r = oracle.counts_reduce(e, e, e, e, np.zeros((16, 0)), max_position=15)
    assert r["n_tax"] == 0
    rng = np.random.default_rng(3)
    tax = np.array([5, 5, 5, 7, 9, 9, 9, 9], np.int64)
    nal = np.array([50, 50, 50, 5, 80, 80, 80, 80], np.uint32)
    rev = np.array([0, 0, 1, 0, 0, 1, 1, 0], np.uint8)
    pos = np.array([0, 1, 0, 0, 0, 0, 20, 20], np.uint8)
    c16 = rng.integers(0, 40, (16, 8)).astype(np.uint32)
    c16[4:8, 1] = 0  # a C-reference row with N = 0 -> f = 0/0 -> 0 (counts.py:254)
    r = oracle.counts_reduce(tax, nal, rev, pos, c16, max_position=15, min_alignments=10, min_y_sum=1)
    assert r["f_fwd"][1] == 0.0 and r["n_fwd_ref"][1] == 0
    assert list(r["z"]) == [1, 2, -1, 1, 1, -1, -21, 21]
    assert list(r["keep"]) == [1, 1, 1, 0, 1, 1, 0, 0]  # tax 7 fails min_alignments; |z| > 15 rows dropped
    assert list(r["tax_id"]) == [5, 9]
    y5 = int(c16[7, 0]) + int(c16[7, 1]) + int(c16[8, 2])
    assert r["y_sum_total"][0] == y5
    assert r["k"][0, 0] == c16[7, 0] and r["k"][0, 15] == c16[8, 2] and r["k"][0, 2] == 0
    assert r["N"][1, 15] == c16[8:12, 5].sum()
