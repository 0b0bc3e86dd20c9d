"""N3: `--max-fits` selection. The reference's own fits.get_top_max_fits generated
tests/golden/topn_golden.npz (tests/golden/make_golden.py); the oracle restatement is checked against
it on CPU, the CUDA kernels (K8, mdg_select_top) against both on the GPU, plus the host seam."""
import os

import numpy as np
import pandas as pd
import pytest

GOLDEN = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "topn_golden.npz"))
CASES = sorted({k.split("_")[0] for k in GOLDEN.files})


def per_taxid(tax_row):
    head = np.r_[True, tax_row[1:] != tax_row[:-1]]
    first = np.flatnonzero(head).astype(np.int64)
    return tax_row[first], first


def expected_sets(case):
    return {int(k.split("top")[1]): GOLDEN[k] for k in GOLDEN.files if k.startswith(case + "_top")}


@pytest.mark.parametrize("case", CASES)
def test_oracle_select_top_matches_reference(oracle, case):
    tax_row, nal_row = GOLDEN[f"{case}_tax_id_row"], GOLDEN[f"{case}_n_alignments_row"]
    tax, first = per_taxid(tax_row)
    for n_top, want in expected_sets(case).items():
        idx, w = oracle.select_top(tax_row, nal_row, None, tax, first, n_top)
        assert np.array_equal(tax[idx], want), (case, n_top)
        assert w.sum() == nal_row.astype(np.uint64).sum()


def test_host_seam_matches_reference():
    from metadamage_b200 import fits

    for case in CASES:
        tax_row, nal_row = GOLDEN[f"{case}_tax_id_row"], GOLDEN[f"{case}_n_alignments_row"]
        df = pd.DataFrame({"tax_id": pd.Series(tax_row).astype("category"), "N_alignments": nal_row})
        for n_top, want in expected_sets(case).items():
            got = pd.unique(fits.get_top_max_fits(df, n_top)["tax_id"]).astype(np.int64)
            assert np.array_equal(got, want)
        assert len(fits.get_top_max_fits(df, None)) == len(df) and len(fits.get_top_max_fits(df, 0)) == len(df)


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_gpu_select_top_matches_reference(ctx, oracle, case):
    tax_row, nal_row = GOLDEN[f"{case}_tax_id_row"], GOLDEN[f"{case}_n_alignments_row"]
    tax, first = per_taxid(tax_row)
    for n_top, want in expected_sets(case).items():
        idx, w = ctx.select_top(tax_row, nal_row, None, tax, first, n_top, want_weight=True)
        assert np.array_equal(tax[idx], want), (case, n_top)
        assert np.array_equal(w, oracle.select_top(tax_row, nal_row, None, tax, first, n_top)[1])
        assert np.all(np.diff(idx) > 0)  # df_counts order


@pytest.mark.gpu
def test_gpu_select_top_with_cut_flags_and_edge_cases(ctx, oracle):
    rng = np.random.default_rng(5)
    n_tax = 5000
    tax = rng.permutation(n_tax).astype(np.int64) - 17  # includes negative ids
    rows = rng.integers(1, 70, n_tax)
    tax_row = np.repeat(tax, rows)
    nal_row = np.repeat(rng.choice([10, 11, 500, 2 ** 31], n_tax), rows).astype(np.uint32)
    keep = (rng.random(len(tax_row)) < 0.7).astype(np.uint8)
    first = np.r_[0, np.cumsum(rows)[:-1]].astype(np.int64)
    for n_top in (0, 1, 33, 2500, n_tax, n_tax + 1):
        got, w = ctx.select_top(tax_row, nal_row, keep, tax, first, n_top, want_weight=True)
        exp, we = oracle.select_top(tax_row, nal_row, keep, tax, first, n_top)
        assert np.array_equal(w, we)
        assert np.array_equal(got, exp), n_top
    # a subset of the TaxIDs (as after the cuts of counts_reduce): only the listed ones compete
    sub = np.sort(rng.choice(n_tax, 800, replace=False))
    got = ctx.select_top(tax_row, nal_row, keep, tax[sub], first[sub], 100)
    exp, _ = oracle.select_top(tax_row, nal_row, keep, tax[sub], first[sub], 100)
    assert np.array_equal(got, exp)
    # empty input
    e = np.empty(0, np.int64)
    assert len(ctx.select_top(e, np.empty(0, np.uint32), None, e, e, 10)) == 0
    # duplicate tax ids with equal weights cannot be ranked: loud error
    from metadamage_b200._lib import MdgError
    with pytest.raises(MdgError):
        ctx.select_top(np.array([7, 7, 7, 7], np.int64), np.array([5, 5, 5, 5], np.uint32), None,
                       np.array([7, 7], np.int64), np.array([0, 0], np.int64), 1)


@pytest.mark.gpu
def test_gpu_select_top_after_counts_reduce_on_device(ctx, oracle):
    """The device-resident chain counts_reduce -> select_top (no host round trip of the arrays)."""
    import torch

    from metadamage_b200 import synthetic as syn

    g = syn.make_mismatch_matrix(3000, seed=91)
    dev = torch.device("cuda", 0)
    cols = dict(tax_id=torch.from_numpy(g["tax_id"]).to(dev), n_alignments=torch.from_numpy(g["n_alignments"].view(np.int32)).to(dev),
                is_reverse=torch.from_numpy(g["is_reverse"]).to(dev), pos0=torch.from_numpy(g["pos0"]).to(dev),
                counts16=torch.from_numpy(g["counts16"].view(np.int32)).to(dev))
    n, m = len(g["tax_id"]), len(g["tax_ids"])
    outs = dict(keep=torch.empty(n, dtype=torch.uint8, device=dev), tax_id=torch.empty(m, dtype=torch.int64, device=dev),
                n_alignments=torch.empty(m, dtype=torch.int32, device=dev), first_row=torch.empty(m, dtype=torch.int64, device=dev),
                k=torch.empty((m, 30), dtype=torch.int32, device=dev), N=torch.empty((m, 30), dtype=torch.int32, device=dev))
    ctx.set_stream(torch.cuda.current_stream(dev).cuda_stream)
    try:
        n_fit = ctx.counts_reduce_device(cols, outs)
        idx = torch.empty(200, dtype=torch.int64, device=dev)
        n_sel = ctx.select_top_device(cols, outs, n_fit, 200, idx)
    finally:
        ctx.set_stream(0)
    r = oracle.counts_reduce(g["tax_id"], g["n_alignments"], g["is_reverse"], g["pos0"], g["counts16"])
    exp, _ = oracle.select_top(g["tax_id"], g["n_alignments"], r["keep"], r["tax_id"], r["first_row"], 200)
    assert n_sel == min(200, n_fit) == len(exp)
    assert np.array_equal(idx[:n_sel].cpu().numpy(), exp)
