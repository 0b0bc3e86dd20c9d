"""-m gpu: the reference-facing seams end to end (counts.compute_counts -> fits.compute_fits ->
parquet), on a TSV written in both of the reference's layouts."""
import numpy as np
import pytest

from test_host import make_cfg, write_tsv

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("legacy", [True, False])
def test_fit_pipeline_on_sample_file(tmp_path, sample_inputs, counts_golden, legacy):
    from metadamage_b200 import counts, fits, io
    from metadamage_b200.main import main

    path = tmp_path / "data_ancient.txt"
    write_tsv(path, sample_inputs["ancient"], legacy)
    cfg = make_cfg(tmp_path / "out", max_fits=10)
    cfg.add_filenames([path])
    cfg.add_filename(path)
    df = counts.compute_counts(cfg)
    g = counts_golden
    # df_counts equals the reference's DataFrame: order, values, dtypes
    assert list(df.columns) == [str(c) for c in g["ancient_CT_GA__columns"]] + ["shortname"]
    assert np.array_equal(df["position"].to_numpy(), g["ancient_CT_GA__position"])
    assert np.array_equal(df["tax_id"].to_numpy(np.int64), g["ancient_CT_GA__tax_id"])
    assert np.array_equal(df["C"].to_numpy(np.int64), g["ancient_CT_GA__n_fwd_ref"])
    assert np.array_equal(df["y_sum_total"].to_numpy(np.int64), g["ancient_CT_GA__y_sum_total"])
    assert np.array_equal(df["f_CT"].to_numpy(), g["ancient_CT_GA__f_fwd"].astype(np.float32))
    assert df["position"].dtype == np.int8 and df["C"].dtype == np.uint32 and df["f_GA"].dtype == np.float32
    assert str(df["strand"].dtype) == "category" and str(df["tax_id"].dtype) == "category"

    main([path], cfg)  # counts -> fits -> parquet, as `metadamage fit` does
    res = io.Parquet(cfg.filename_fit_results).load()
    pred = io.Parquet(cfg.filename_fit_predictions).load()
    fmap = io.Parquet(cfg.filename_fit_map).load()
    assert list(res.columns) == fits.FIT_RESULT_COLUMNS + ["shortname"] and len(res) == 3
    assert list(res["tax_id"]) == [0, 1, 2]  # df_counts order (N_alignments descending)
    assert len(pred) == 90 and list(pred.columns) == ["tax_id", "position", "median", "hdpi_lower", "hdpi_upper", "shortname"]
    # the sample is strongly damaged ancient DNA: D_max ~ 0.42, decisive n_sigma
    assert np.all(np.abs(res["D_max"].to_numpy() - np.array([0.418, 0.446, 0.446])) < 0.02)
    assert np.all(res["n_sigma"].to_numpy() > 5)
    assert abs(fmap["map_A"][0] - 0.39343) < 1e-3 and abs(fmap["map_q"][0] - 0.53304) < 1e-3
    assert abs(res["normalized_noise"][0] - 0.25177245) < 1e-6
    assert io.Parquet(cfg.filename_fit_results).load_metadata()["N_fits"] == 3
    # cached second run: nothing recomputed, same tables
    df2 = counts.load_counts(cfg)
    assert df2["position"].tolist() == df["position"].tolist()
    res2, pred2 = fits.get_fits(df2, cfg)
    assert res2["D_max"].tolist() == res["D_max"].tolist()


def test_multi_gpu_thread_partition_is_bit_identical(tmp_path):
    """fits.fit_dense over 2 GPUs (one host thread + ctx per GPU, contiguous TaxID ranges, no
    collective) gives the same bytes as 1 GPU. Skipped on single-GPU boxes."""
    from metadamage_b200 import _lib, fits, synthetic as syn

    if _lib.load().mdg_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    tid, k, N, _ = syn.dense_fit_batch(64, seed=77)
    dense = dict(tax_id=tid, k=k, N=N, mism12=None)
    cfg = _lib.default_config(num_warmup=60, num_samples=80)
    one = fits.fit_dense(dense, None, cfg, n_gpus=1)
    two = fits.fit_dense(dense, None, cfg, n_gpus=2)
    for key in ("result", "median", "hpdi_lo", "hpdi_hi"):
        assert one[key].tobytes() == two[key].tobytes()


def test_counts_order_kernel_matches_reference_order(ctx, oracle):
    """K1d (mdg_counts_order, C8): the row permutation of counts.py:167-172 from the device == the host lexsort on the
    same keys, on ragged TaxIDs with duplicate positions, dropped rows, |z| > P and ties in N_alignments."""
    from metadamage_b200 import counts
    from test_gpu_counts import ragged_input

    rng = np.random.default_rng(3)
    for max_rows, n_tax in ((1, 3000), (40, 1500), (700, 60)):
        tax, nal, rev, pos, c16 = ragged_input(rng, n_tax, max_rows, 15)
        nal = np.repeat(rng.integers(1, 8, n_tax).astype(np.uint32) * 30, np.diff(np.r_[np.flatnonzero(np.r_[True, tax[1:] != tax[:-1]]), len(tax)]))
        r = ctx.counts_reduce(tax, nal, rev, pos, c16, min_alignments=60, min_y_sum=3)
        tax_order = np.lexsort((-r["tax_id"], -r["n_alignments"].astype(np.int64)))
        perm = ctx.counts_order(tax, r["z"], r["keep"], r["first_row"], tax_order)
        keep = r["keep"].astype(bool)
        rows = np.flatnonzero(keep)
        want = rows[counts.reference_row_order(nal[keep], tax[keep], r["z"][keep])]
        assert np.array_equal(perm, want)


def test_compute_counts_pieces_and_dense_reuse(tmp_path, oracle):
    """counts.compute_counts on a synthetic 22-column file: (1) the df_counts of the one-piece and of the
    multi-piece path (file cut at TaxID boundaries, the way cfg.gpus > 1 spreads it) are identical and in the
    reference's order; (2) K1's dense k/N/noise ride along and equal a K1 run on the frame's own columns (the
    parquet-cache path) and the oracle; (3) --max-fits through mdg_select_top == the pandas expression."""
    from metadamage_b200 import counts, fits, synthetic as syn

    g = syn.make_mismatch_matrix(0, n_fit=400, seed=5)
    path = tmp_path / "synth.txt"
    syn.write_tsv(g, str(path))
    cfg = make_cfg(tmp_path / "out")
    cfg.add_filename(path)
    df = counts.compute_counts(cfg)
    o = oracle.counts_reduce(g["tax_id"], g["n_alignments"], g["is_reverse"], g["pos0"], g["counts16"])
    assert df["tax_id"].nunique() == o["n_tax"] == 400 and len(df) == 400 * 30
    keep = o["keep"].astype(bool)
    order = counts.reference_row_order(g["n_alignments"][keep], g["tax_id"][keep], o["z"][keep])
    assert np.array_equal(df["tax_id"].to_numpy(np.int64), g["tax_id"][keep][order])
    assert np.array_equal(df["position"].to_numpy(), o["z"][keep][order])
    assert np.array_equal(df["CT"].to_numpy(), g["counts16"][7][keep][order])
    assert np.array_equal(df["f_GA"].to_numpy().view(np.uint32), o["f_rev"][keep][order].view(np.uint32))
    assert set(df["tax_name"].unique()) == {"synthetic taxon"} and str(df["tax_name"].dtype) == "category"
    # (2) dense hand-off
    dense = counts.dense_from_df_counts(df, cfg)
    assert dense is counts._DENSE[id(df)]
    again = counts.dense_from_df_counts(df.copy(), cfg)   # a different object: K1 on the frame's columns
    for key in ("tax_id", "k", "N"):
        assert np.array_equal(dense[key], again[key]), key
    np.testing.assert_allclose(dense["noise"], again["noise"], rtol=1e-11)
    by_id = {t: i for i, t in enumerate(o["tax_id"])}
    idx = np.array([by_id[t] for t in dense["tax_id"]])
    assert np.array_equal(dense["k"], o["k"][idx]) and np.array_equal(dense["N"], o["N"][idx])
    # (1) multi-piece path
    text = open(path, "rb").read()
    spans = counts.split_text_at_taxid_boundaries(text, 4)
    assert len(spans) == 4 and spans[0][0] == 0 and spans[-1][1] == len(text)
    ctx = counts.get_context(0)
    pieces = [counts._counts_piece_on_gpu(ctx, text[a:b], cfg) for a, b in spans]
    df4 = counts._assemble_df_counts(pieces, cfg)
    assert df4.equals(df)
    # (3) --max-fits on the device
    top = fits.select_top_dense(df, dense, 37)
    want = fits.extract_top_max_fits(df, 37)
    assert np.array_equal(top["tax_id"], want["tax_id"].astype(np.int64).unique())
    assert fits.select_top_dense(df, dense, None) is dense


def test_compute_counts_on_two_gpus_equals_one(tmp_path):
    """cfg.gpus = 2: the file is cut at a TaxID boundary, each GPU tokenises and reduces its piece (own ctx, own host
    thread, nothing exchanged), the pieces are merged on the host: same df_counts as on one GPU, and the fits of the
    two-GPU run (TaxID ranges per GPU) equal the one-GPU ones bit for bit. Skipped on single-GPU boxes."""
    from metadamage_b200 import _lib, counts, fits, synthetic as syn

    if _lib.load().mdg_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    g = syn.make_mismatch_matrix(0, n_fit=300, seed=11)
    path = tmp_path / "synth2.txt"
    syn.write_tsv(g, str(path))
    cfg1 = make_cfg(tmp_path / "o1")
    cfg1.add_filename(path)
    cfg2 = make_cfg(tmp_path / "o2", )
    cfg2.gpus = 2
    cfg2.add_filename(path)
    df1, df2 = counts.compute_counts(cfg1), counts.compute_counts(cfg2)
    assert df2.equals(df1)
    kw = dict(progress_bar=False, num_warmup=60, num_samples=80, num_chains=1, chain_method="sequential")
    r1, p1 = fits.compute_fits(df1, cfg1, kw)
    r2, p2 = fits.compute_fits(df2, cfg2, kw)
    assert r2.drop(columns=["shortname"]).equals(r1.drop(columns=["shortname"])) and p2["median"].equals(p1["median"])
