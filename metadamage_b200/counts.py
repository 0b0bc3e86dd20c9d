"""Counts stage: mismatch-matrix file -> df_counts, with the arithmetic on the GPU (K1).

Keeps the reference's seam `compute_counts_with_dask(cfg)` / `load_counts(cfg)`
(counts.py:212-306): same columns, order and dtypes of the returned DataFrame, same parquet
cache rules. The reference's dask pipeline (counts.py:229-258) is replaced by one
`mdg_counts_reduce` call; row ordering (counts.py:167-172) and the DataFrame assembly stay on
the host.
"""
import logging

import numpy as np
import pandas as pd

from . import io, utils
from .backend import BASES, Context

logger = logging.getLogger(__name__)

REF_OBS_BASES = [r + o for r in BASES for o in BASES]
# the reference's 22 header-less columns (counts.py:37-45)
COLUMNS = ["tax_id", "tax_name", "tax_rank", "N_alignments", "strand", "position", *REF_OBS_BASES]
# the legacy 20-column layout WITH a header that the shipped data/input/*.txt use
LEGACY_COLUMNS = ["tax_id", "N_alignments", "strand", "position", *REF_OBS_BASES]

_ctx_cache = {}


def get_context(device=0):
    if device not in _ctx_cache:
        _ctx_cache[device] = Context(device)
    return _ctx_cache[device]


def read_mismatch_table(filename):
    """Parse the TSV into a DataFrame with the reference's 22 columns. Both layouts are
    accepted: 22 header-less columns (what counts.py:229-235 reads) and the legacy 20 columns
    with a '#taxid ...' header line (what data/input/data_ancient.txt actually contains)."""
    with open(filename, "r") as fh:
        first = fh.readline()
    n_fields = len(first.rstrip("\n").split("\t"))
    has_header = first.startswith("#") or not first.split("\t")[0].lstrip("-").isdigit()
    if n_fields == len(COLUMNS):
        df = pd.read_csv(filename, sep="\t", header=0 if has_header else None, names=COLUMNS)
    elif n_fields == len(LEGACY_COLUMNS):
        df = pd.read_csv(filename, sep="\t", header=0 if has_header else None, names=LEGACY_COLUMNS)
        df.insert(1, "tax_name", "")
        df.insert(2, "tax_rank", "")
    else:
        raise ValueError(f"{filename}: expected {len(COLUMNS)} or {len(LEGACY_COLUMNS)} tab-separated columns, got {n_fields}")
    return df


def group_rows_by_tax_id(df):
    """The kernel needs all rows of a TaxID contiguous (they are in the reference's input files);
    if they are not, a stable sort by first appearance makes them so."""
    tax = df["tax_id"].to_numpy()
    heads = np.flatnonzero(np.r_[True, tax[1:] != tax[:-1]])
    if len(np.unique(tax[heads])) == len(heads):
        return df
    order = pd.Series(np.arange(len(df))).groupby(tax, sort=False).transform("min").to_numpy()
    return df.iloc[np.argsort(order, kind="stable")].reset_index(drop=True)


def soa_columns(df):
    """DataFrame -> the SoA arrays of the C-ABI."""
    if len(df) and (df[REF_OBS_BASES].to_numpy().max() > np.iinfo(np.uint32).max or df["N_alignments"].max() > np.iinfo(np.uint32).max):
        raise AssertionError("Dataframe contains too large values.")
    return dict(
        tax_id=df["tax_id"].to_numpy(np.int64),
        n_alignments=df["N_alignments"].to_numpy(np.uint32),
        is_reverse=(df["strand"].to_numpy() != "5'").astype(np.uint8),
        pos0=df["position"].to_numpy(np.uint8),
        counts16=np.ascontiguousarray(df[REF_OBS_BASES].to_numpy(np.uint32).T),
    )


def reference_row_order(n_alignments, tax_id, z):
    """Row order of counts.py:167-172: N_alignments desc, tax_id desc, then z = +1..+P, -1..-P
    (sort key `order` = 1/z for z > 0 else z, descending)."""
    z = z.astype(np.float64)
    order = np.where(z > 0, 1.0 / np.where(z > 0, z, 1.0), z)
    return np.lexsort((-order, -tax_id.astype(np.int64), -n_alignments.astype(np.int64)))


def read_mismatch_table_gpu(filename, ctx):
    """Same table as `read_mismatch_table`, with the numeric columns tokenised on the GPU
    (mdg_tsv_parse, K0) instead of by pandas; the two string columns of the 22-column layout are
    cut out of the file bytes through the (offset, length) spans the kernel returns."""
    with open(filename, "rb") as fh:
        text = fh.read()
    r = ctx.tsv_parse(text, want_spans=True)
    logger.info("tsv: %d bytes -> %d rows (parse kernels %.3f ms)", len(text), r["n_rows"], ctx.timings()["counts_ms"])
    data = {"tax_id": r["tax_id"]}
    if r["n_cols"] == 22:
        for col, key in (("tax_name", "name_span"), ("tax_rank", "rank_span")):
            spans = r[key]
            # few distinct strings, many rows: decode each distinct (offset-independent) value once
            uniq, inv = np.unique(spans[:, 1], return_inverse=True) if len(spans) else (np.zeros(0, np.int64), np.zeros(0, np.int64))
            vals = np.empty(len(spans), dtype=object)
            cache = {}
            for i in range(len(spans)):
                o, n = int(spans[i, 0]), int(spans[i, 1])
                b = text[o:o + n]
                v = cache.get(b)
                if v is None:
                    v = cache[b] = b.decode("utf-8", "replace")
                vals[i] = v
            data[col] = vals
    else:
        data["tax_name"] = ""
        data["tax_rank"] = ""
    data["N_alignments"] = r["n_alignments"]
    data["strand"] = np.where(r["is_reverse"] == 1, "3'", "5'") if r["n_cols"] else np.zeros(0, dtype=object)
    data["position"] = r["pos0"].astype(np.int64)
    df = pd.DataFrame(data, columns=COLUMNS[:6])
    for i, name in enumerate(REF_OBS_BASES):
        df[name] = r["counts16"][i]
    return df


def compute_counts(cfg, df_in=None, ctx=None):
    """The GPU replacement of compute_counts_with_dask (counts.py:212-273)."""
    ctx = ctx or get_context(0)
    if df_in is None:
        df_in = read_mismatch_table_gpu(cfg.filename, ctx)
    df_in = group_rows_by_tax_id(df_in)
    fwd, rev = cfg.substitution_bases_forward, cfg.substitution_bases_reverse
    cols = soa_columns(df_in)
    r = ctx.counts_reduce(cols["tax_id"], cols["n_alignments"], cols["is_reverse"], cols["pos0"], cols["counts16"],
                          fwd=fwd, rev=rev, max_position=cfg.max_position, min_alignments=cfg.min_alignments,
                          min_y_sum=cfg.min_y_sum)
    logger.info("counts: %d rows -> %d TaxIDs kept (kernel %.3f ms)", len(df_in), r["n_tax"], ctx.timings()["counts_ms"])
    keep = r["keep"].astype(bool)
    df = df_in.loc[keep].copy()
    df["position"] = r["z"][keep].astype(np.int64)
    df[fwd[0]] = r["n_fwd_ref"][keep]
    df[rev[0]] = r["n_rev_ref"][keep]
    df[f"f_{fwd}"] = r["f_fwd"][keep]
    df[f"f_{rev}"] = r["f_rev"][keep]
    df["y_sum_total"] = r["y_sum_total"][keep]
    order = reference_row_order(df["N_alignments"].to_numpy(), df["tax_id"].to_numpy(), df["position"].to_numpy())
    df = df.iloc[order].reset_index(drop=True)
    df["shortname"] = cfg.shortname
    return utils.downcast_dataframe(df, ["tax_id", "tax_name", "tax_rank", "strand", "shortname"])


# the reference's name for the seam (counts.py:212); `use_processes` is accepted and ignored
def compute_counts_with_dask(cfg, use_processes=True):
    return compute_counts(cfg)


CACHE_KEYS = ["min_alignments", "min_y_sum", "substitution_bases_forward", "substitution_bases_reverse",
              "shortname", "filename", "max_position"]


def load_counts(cfg):
    """counts.py:276-306: reuse counts/<shortname>.parquet iff present, not --forced and its
    metadata match the configuration."""
    parquet = io.Parquet(cfg.filename_counts)
    if parquet.exists(cfg.forced):
        if utils.metadata_is_similar(parquet.load_metadata(), cfg.to_dict(), include=CACHE_KEYS):
            logger.info("Loading DataFrame from parquet-file.")
            df_counts = parquet.load()
            cfg.set_number_of_fits(df_counts)
            return df_counts
    logger.info("Creating DataFrame, please wait.")
    df_counts = compute_counts(cfg)
    parquet.save(df_counts, metadata=cfg.to_dict())
    cfg.set_number_of_fits(df_counts)
    return df_counts


def dense_from_df_counts(df_counts, cfg):
    """df_counts (reference order: per TaxID z = +1..+P then -1..-P) -> tax ids, names, ranks,
    N_alignments and the dense k/N [n_tax][2P] + mism12 [n_tax][2P][12] the fit kernels take
    (fits.py:398-419, 359-363). Missing positions stay zero."""
    P = int(cfg.max_position)
    fwd, rev = cfg.substitution_bases_forward, cfg.substitution_bases_reverse
    tax = df_counts["tax_id"].to_numpy(np.int64)
    uniq, first, inv = np.unique(tax, return_index=True, return_inverse=True)
    rank_of = np.argsort(np.argsort(first))      # order of first appearance (df_counts order)
    t_idx = rank_of[inv]
    n_tax = len(uniq)
    z = df_counts["position"].to_numpy(np.int64)
    ok = np.abs(z) <= P
    slot = np.where(z > 0, z - 1, P + np.abs(z) - 1)
    k = np.zeros((n_tax, 2 * P), np.uint32)
    N = np.zeros((n_tax, 2 * P), np.uint32)
    kcol = np.where(z > 0, df_counts[fwd].to_numpy(np.int64), df_counts[rev].to_numpy(np.int64))
    ncol = np.where(z > 0, df_counts[fwd[0]].to_numpy(np.int64), df_counts[rev[0]].to_numpy(np.int64))
    np.add.at(k, (t_idx[ok], slot[ok]), kcol[ok].astype(np.uint32))
    np.add.at(N, (t_idx[ok], slot[ok]), ncol[ok].astype(np.uint32))
    off = [c for c in REF_OBS_BASES if c[0] != c[1]]
    m12 = np.zeros((n_tax, 2 * P, 12), np.uint32)
    np.add.at(m12, (t_idx[ok], slot[ok]), df_counts[off].to_numpy(np.uint32)[ok])
    head = np.sort(first)
    meta = df_counts.iloc[head]
    return dict(tax_id=tax[head], tax_name=meta["tax_name"].to_numpy(), tax_rank=meta["tax_rank"].to_numpy(),
                N_alignments=meta["N_alignments"].to_numpy(np.uint32), k=k, N=N, mism12=m12)
