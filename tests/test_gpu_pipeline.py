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
