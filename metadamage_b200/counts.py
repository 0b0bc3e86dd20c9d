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
    if len(df) and not (0 <= df["position"].min() and df["position"].max() <= 254):
        raise ValueError("position must be in [0, 254] (the GPU tokeniser enforces the same range)")
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


def _as_byte_array(text):
    """bytes-like or uint8 ndarray -> uint8 ndarray (no copy)"""
    return text if isinstance(text, np.ndarray) else np.frombuffer(text, dtype=np.uint8)


def _decode_spans(text, spans):
    """(offset, length) byte spans of the file text -> list of str. The spans are gathered into one [n][longest]
    byte matrix and decoded in one go (ASCII) or string by string (anything else: UTF-8 with replacement)."""
    spans = np.asarray(spans, dtype=np.int64).reshape(-1, 2)
    if len(spans) == 0:
        return []
    buf = _as_byte_array(text)
    off, length = spans[:, 0], spans[:, 1]
    longest = int(length.max())
    if longest == 0:
        return [""] * len(spans)
    col = np.arange(longest, dtype=np.int64)[None, :]
    idx = np.minimum(off[:, None] + col, len(buf) - 1)
    mat = buf[idx]
    mat[col >= length[:, None]] = 0
    if mat.max() < 128 and not ((mat == 0) & (col < length[:, None])).any():
        return mat.view(f"S{longest}").ravel().astype(f"U{longest}").astype(object).tolist()
    return [buf[int(o):int(o) + int(n)].tobytes().decode("utf-8", "replace") for o, n in zip(off, length)]


_TEXT_STAGE = {}  # one reusable page-locked staging buffer per process for the file text


def read_file_bytes(filename):
    """The file as a uint8 array in a reusable PINNED buffer: no fresh 100 MB allocation per file (page faults were
    two thirds of the read time) and the copy to the GPU is a plain DMA. Valid until the next call."""
    import os

    size = os.path.getsize(filename)
    stage = _TEXT_STAGE.get("buf")
    if stage is None or len(stage) < size:
        try:
            import torch

            stage = torch.empty(max(size + size // 4, 1 << 20), dtype=torch.uint8, pin_memory=torch.cuda.is_available()).numpy()
        except Exception:  # no torch / no CUDA runtime: ordinary memory
            stage = np.empty(max(size + size // 4, 1 << 20), dtype=np.uint8)
        _TEXT_STAGE["buf"] = stage
    view = stage[:size]
    got = 0
    with open(filename, "rb", buffering=0) as fh:
        while got < size:
            n = fh.readinto(memoryview(view[got:]))
            if not n:
                break
            got += n
    return view[:got]


def read_mismatch_table_gpu(filename, ctx):
    """Same table as `read_mismatch_table`, with the numeric columns tokenised on the GPU
    (mdg_tsv_parse, K0) instead of by pandas; the two string columns of the 22-column layout are
    cut out of the file bytes through the (offset, length) spans the kernel returns, decoded once per
    run of equal spans' bytes (rows of a TaxID share name and rank)."""
    text = read_file_bytes(filename)
    r = ctx.tsv_parse(text, want_spans=True)
    logger.info("tsv: %d bytes -> %d rows (parse kernels %.3f ms)", len(text), r["n_rows"], ctx.timings()["counts_ms"])
    data = {"tax_id": r["tax_id"]}
    n = r["n_rows"]
    if r["n_cols"] == 22 and n:
        tax = r["tax_id"]
        heads = np.flatnonzero(np.r_[True, tax[1:] != tax[:-1]])
        seg = np.diff(np.r_[heads, n])
        for col, key in (("tax_name", "name_span"), ("tax_rank", "rank_span")):
            data[col] = np.repeat(np.array(_decode_spans(text, r[key][heads]), dtype=object), seg)
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


def _find_byte(buf, byte, start, stop=None):
    """index of the first `byte` in buf[start:stop], or -1 (scanned in 64 KB windows)"""
    stop = len(buf) if stop is None else stop
    while start < stop:
        end = min(start + (1 << 16), stop)
        hit = np.flatnonzero(buf[start:end] == byte)
        if len(hit):
            return start + int(hit[0])
        start = end
    return -1


def split_text_at_taxid_boundaries(text, n_parts):
    """Cut the file bytes into <= n_parts pieces whose first line starts a new TaxID (K1 is
    partitioned at segment boundaries, so no TaxID spans two GPUs and nothing is exchanged). A nominal cut
    at k/n of the bytes is moved forward to the next line whose first field differs from the line before."""
    buf = _as_byte_array(text)
    n = len(buf)
    if n_parts <= 1 or n < 1 << 16:
        return [(0, n)]
    NL, TAB = 10, 9
    cuts = [0]
    for k in range(1, n_parts):
        pos = _find_byte(buf, NL, max(cuts[-1], (n * k) // n_parts))
        if pos < 0:
            break
        pos += 1
        back = np.flatnonzero(buf[max(0, pos - 1 - (1 << 16)):pos - 1] == NL)  # start of the line before the cut
        prev_start = max(0, pos - 1 - (1 << 16)) + int(back[-1]) + 1 if len(back) else (0 if pos - 1 <= (1 << 16) else -1)
        if prev_start < 0:
            break  # a line longer than 64 KB: not a mismatch table
        prev_id = buf[prev_start:_find_byte(buf, TAB, prev_start)]
        while pos < n:
            tab = _find_byte(buf, TAB, pos, min(n, pos + (1 << 16)))
            if tab < 0 or not np.array_equal(buf[pos:tab], prev_id):
                break
            nxt = _find_byte(buf, NL, pos)
            if nxt < 0:
                pos = n
                break
            pos = nxt + 1
        if pos >= n:
            break
        if pos > cuts[-1]:
            cuts.append(pos)
    return [(a, b) for a, b in zip(cuts, cuts[1:] + [n]) if b > a]


def _counts_piece_on_gpu(ctx, text, cfg):
    """One piece of the file on one GPU: K0 (text -> device columns) -> K1 (device) -> K1d (row order); only the
    kept rows (already in the reference's order inside this piece: TaxIDs by N_alignments / tax_id descending) and
    the per-TaxID arrays come back to the host. Returns None if the rows of a TaxID are not contiguous."""
    import torch

    fwd, rev = cfg.substitution_bases_forward, cfg.substitution_bases_reverse
    P = int(cfg.max_position)
    cols, n, n_cols = ctx.tsv_parse_device(text, want_spans=True)
    t_parse = ctx.timings()["counts_ms"]
    if n == 0:
        return dict(n_rows=0, n_cols=n_cols)
    dev = cols["tax_id"].device
    tax_all = cols["tax_id"].cpu().numpy()
    heads = np.flatnonzero(np.r_[True, tax_all[1:] != tax_all[:-1]])
    if len(np.unique(tax_all[heads])) != len(heads):
        return None
    m_cap = len(heads)
    outs = dict(
        n_fwd_ref=torch.empty(n, dtype=torch.int32, device=dev), n_rev_ref=torch.empty(n, dtype=torch.int32, device=dev),
        f_fwd=torch.empty(n, dtype=torch.float32, device=dev), f_rev=torch.empty(n, dtype=torch.float32, device=dev),
        z=torch.empty(n, dtype=torch.int8, device=dev), y_sum_total=torch.empty(n, dtype=torch.int64, device=dev),
        keep=torch.empty(n, dtype=torch.uint8, device=dev), tax_id=torch.empty(m_cap, dtype=torch.int64, device=dev),
        n_alignments=torch.empty(m_cap, dtype=torch.int32, device=dev), first_row=torch.empty(m_cap, dtype=torch.int64, device=dev),
        k=torch.empty((m_cap, 2 * P), dtype=torch.int32, device=dev), N=torch.empty((m_cap, 2 * P), dtype=torch.int32, device=dev),
        noise=torch.empty((m_cap, 3), dtype=torch.float64, device=dev))
    n_tax = ctx.counts_reduce_device(cols, outs, fwd=fwd, rev=rev, max_position=P, min_alignments=cfg.min_alignments,
                                     min_y_sum=cfg.min_y_sum)
    t_counts = ctx.timings()["counts_ms"]
    logger.info("counts: %d rows -> %d TaxIDs kept (K0 %.3f ms, K1 %.3f ms)", n, n_tax, t_parse, t_counts)
    tax = dict(tax_id=outs["tax_id"][:n_tax].cpu().numpy(), n_alignments=outs["n_alignments"][:n_tax].cpu().numpy().view(np.uint32),
               first_row=outs["first_row"][:n_tax].cpu().numpy(), k=outs["k"][:n_tax].cpu().numpy().view(np.uint32),
               N=outs["N"][:n_tax].cpu().numpy().view(np.uint32), noise=outs["noise"][:n_tax].cpu().numpy())
    # C8: TaxID-level keys sorted on the host (n_tax of them), row-level permutation on the device
    tax_order = np.lexsort((-tax["tax_id"], -tax["n_alignments"].astype(np.int64)))
    n_keep = int(torch.count_nonzero(outs["keep"]).item()) if n_tax else 0
    perm = torch.empty(max(n_keep, 1), dtype=torch.int64, device=dev)
    n_out = ctx.counts_order_device(cols, outs, n_tax, torch.from_numpy(tax_order).to(dev), perm) if n_tax else 0
    perm = perm[:n_out]

    def take(t):  # kept rows in output order, to the host
        return t.index_select(0, perm).cpu().numpy()

    rows = dict(tax_id=take(cols["tax_id"]), n_alignments=take(cols["n_alignments"]).view(np.uint32),
                is_reverse=take(cols["is_reverse"]), z=take(outs["z"]),
                counts16=cols["counts16"].index_select(1, perm).cpu().numpy().view(np.uint32),
                n_fwd_ref=take(outs["n_fwd_ref"]).view(np.uint32), n_rev_ref=take(outs["n_rev_ref"]).view(np.uint32),
                f_fwd=take(outs["f_fwd"]), f_rev=take(outs["f_rev"]), y_sum_total=take(outs["y_sum_total"]).view(np.uint64))
    names = ranks = None
    if n_cols == 22 and n_tax:
        first = torch.from_numpy(tax["first_row"]).to(dev)
        names = _decode_spans(text, cols["name_span"].index_select(0, first).cpu().numpy())
        ranks = _decode_spans(text, cols["rank_span"].index_select(0, first).cpu().numpy())
    return dict(n_rows=n, n_cols=n_cols, rows=rows, tax=tax, tax_order=tax_order, names=names, ranks=ranks)


def _categorical_from_groups(values_per_group, group_of_row):
    """per-TaxID strings -> the row-level categorical astype("category") would build (sorted categories), hashing
    the n_tax strings once instead of sorting them or touching every row"""
    per_group = pd.Categorical(np.asarray(values_per_group, dtype=object))
    return pd.Categorical.from_codes(np.asarray(per_group.codes)[group_of_row], categories=per_group.categories)


def _assemble_df_counts(pieces, cfg):
    """The kept rows of every piece -> df_counts in the reference's order, columns and dtypes (counts.py:167-172,
    270-272). Pieces are internally ordered already; a k-way merge by (N_alignments, tax_id) descending joins them."""
    fwd, rev = cfg.substitution_bases_forward, cfg.substitution_bases_reverse
    P = int(cfg.max_position)
    if len(pieces) == 1:
        rows = pieces[0]["rows"]
        tax = {key: val[pieces[0]["tax_order"]] for key, val in pieces[0]["tax"].items() if key != "first_row"}
    else:
        rows = {key: np.concatenate([p["rows"][key] for p in pieces], axis=1 if key == "counts16" else 0) for key in pieces[0]["rows"]}
        tax = {key: np.concatenate([p["tax"][key][p["tax_order"]] for p in pieces]) for key in pieces[0]["tax"] if key != "first_row"}
    # rows per kept TaxID, in each piece's output order
    seg_len = np.concatenate([np.diff(np.r_[np.flatnonzero(np.r_[True, p["rows"]["tax_id"][1:] != p["rows"]["tax_id"][:-1]]),
                                            len(p["rows"]["tax_id"])]) if len(p["rows"]["tax_id"]) else np.zeros(0, np.int64)
                              for p in pieces]).astype(np.int64)
    names = [v for p in pieces for v in (np.asarray(p["names"], dtype=object)[p["tax_order"]] if p["names"] is not None else [""] * len(p["tax_order"]))]
    ranks = [v for p in pieces for v in (np.asarray(p["ranks"], dtype=object)[p["tax_order"]] if p["ranks"] is not None else [""] * len(p["tax_order"]))]
    if len(pieces) > 1:  # merge the pieces' TaxID lists; rows follow their TaxIDs
        order = np.lexsort((-tax["tax_id"], -tax["n_alignments"].astype(np.int64)))
        starts = np.r_[0, np.cumsum(seg_len)[:-1]]
        row_idx = np.concatenate([np.arange(starts[t], starts[t] + seg_len[t]) for t in order]) if len(order) else np.zeros(0, np.int64)
        rows = {key: (val[:, row_idx] if key == "counts16" else val[row_idx]) for key, val in rows.items()}
        tax = {key: val[order] for key, val in tax.items()}
        seg_len = seg_len[order]
        names = [names[t] for t in order]
        ranks = [ranks[t] for t in order]
    n_tax = len(seg_len)
    group_of_row = np.repeat(np.arange(n_tax), seg_len)
    n_rows = len(group_of_row)
    # categorical columns straight from integer codes (what astype("category") would build, without hashing
    # every row): tax_id and the strings are per-TaxID values, the strand has two
    uniq_tax, inv_tax = np.unique(tax["tax_id"], return_inverse=True)
    strand_present = np.unique(rows["is_reverse"] != 0)
    strand_cats = np.array(["3'", "5'"], dtype=object)[[0] if strand_present.tolist() == [True] else ([1] if strand_present.tolist() == [False] else [0, 1])]
    strand_codes = np.zeros(n_rows, np.int8) if len(strand_cats) == 1 else np.where(rows["is_reverse"] != 0, 0, 1).astype(np.int8)
    data = {"tax_id": pd.Categorical.from_codes(inv_tax[group_of_row], categories=uniq_tax) if n_tax else rows["tax_id"],
            "tax_name": _categorical_from_groups(names, group_of_row) if n_tax else [],
            "tax_rank": _categorical_from_groups(ranks, group_of_row) if n_tax else [], "N_alignments": rows["n_alignments"],
            "strand": pd.Categorical.from_codes(strand_codes, categories=strand_cats) if n_rows else [], "position": rows["z"]}
    for i, name in enumerate(REF_OBS_BASES):
        data[name] = rows["counts16"][i]
    # the reference's column layout (counts.py:88, 111, 201-203): a reference-base column shared by both
    # substitutions is written once
    data[fwd[0]] = rows["n_fwd_ref"]
    data[rev[0]] = rows["n_rev_ref"]
    data[f"f_{fwd}"] = rows["f_fwd"]
    data[f"f_{rev}"] = rows["f_rev"]
    if n_rows and rows["y_sum_total"].max() > np.iinfo(np.uint32).max:
        raise AssertionError("Dataframe contains too large values.")  # utils.py:338-339
    data["y_sum_total"] = rows["y_sum_total"].astype(np.uint32)
    data["shortname"] = pd.Categorical.from_codes(np.zeros(n_rows, np.int8), categories=[cfg.shortname])
    df = pd.DataFrame(data, copy=False)  # every column already has its final dtype (utils.downcast_dataframe would change nothing)
    # K1's dense per-TaxID outputs are remembered for this very DataFrame object (df_counts order), so that
    # compute_fits does not rebuild them from the rows
    first_row = np.r_[0, np.cumsum(seg_len)[:-1]].astype(np.int64) if n_tax else np.zeros(0, np.int64)
    _remember_dense(df, dict(tax_id=tax["tax_id"], tax_name=np.asarray(names, dtype=object), tax_rank=np.asarray(ranks, dtype=object),
                             N_alignments=tax["n_alignments"], k=tax["k"], N=tax["N"], noise=tax["noise"], mism12=None,
                             first_row=first_row, max_position=P, fwd=fwd, rev=rev))
    return df


def compute_counts(cfg, df_in=None, ctx=None):
    """The GPU replacement of compute_counts_with_dask (counts.py:212-273): file bytes -> K0 -> K1 -> K1d on the
    device(s), only the kept rows return to the host. With cfg.gpus > 1 the file is cut at TaxID boundaries and
    every GPU handles one piece (no exchange)."""
    if df_in is None:
        text = read_file_bytes(cfg.filename)
        n_gpus = max(1, min(int(getattr(cfg, "gpus", 1) or 1), _lib_device_count()))
        spans = split_text_at_taxid_boundaries(text, n_gpus) if ctx is None else [(0, len(text))]
        from .parallel import run_on_gpus

        def worker(rank, start, stop):
            out = []
            for a, b in spans[start:stop]:
                out.append(_counts_piece_on_gpu(ctx or get_context(rank), text[a:b] if (a, b) != (0, len(text)) else text, cfg))
            return out

        pieces = [p for part in run_on_gpus(len(spans), len(spans), worker) if part for p in part]
        if all(p is not None for p in pieces):
            pieces = [p for p in pieces if p.get("n_rows", 0) > 0 and len(p["rows"]["tax_id"]) >= 0]
            if pieces and len({p["n_cols"] for p in pieces}) == 1:
                tids = np.concatenate([p["tax"]["tax_id"] for p in pieces])
                if len(np.unique(tids)) == len(tids):
                    return _assemble_df_counts(pieces, cfg)
            if not pieces:
                df_in = read_mismatch_table_gpu(cfg.filename, ctx or get_context(0))
        # rows of a TaxID are not contiguous (or the pieces disagree): regroup on the host and take the host-buffer path
        if df_in is None:
            df_in = read_mismatch_table_gpu(cfg.filename, ctx or get_context(0))
    ctx = ctx or get_context(0)
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
    tax_order = np.lexsort((-r["tax_id"], -r["n_alignments"].astype(np.int64)))
    order = ctx.counts_order(cols["tax_id"], r["z"], r["keep"], r["first_row"], tax_order) if r["n_tax"] else np.zeros(0, np.int64)
    # `order` indexes input rows; df holds the kept rows only
    kept_pos = np.cumsum(keep) - 1
    df = df.iloc[kept_pos[order]].reset_index(drop=True)
    df["shortname"] = cfg.shortname
    return utils.downcast_dataframe(df, ["tax_id", "tax_name", "tax_rank", "strand", "shortname"])


def _lib_device_count():
    from . import _lib

    return _lib.load().mdg_device_count() or 1


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


_DENSE = {}


def _remember_dense(df, dense):
    import weakref

    key = id(df)
    _DENSE[key] = dense
    weakref.finalize(df, _DENSE.pop, key, None)


def dense_from_df_counts(df_counts, cfg, ctx=None):
    """df_counts -> what the fit kernels take (fits.py:398-419, 359-376): tax ids, names, ranks, N_alignments in
    df_counts order, dense k/N [n_tax][2P] and the noise statistic [n_tax][3]. If this DataFrame came out of
    `compute_counts` in this process, K1 has produced all of that already; otherwise (parquet cache, a filtered
    frame) K1 runs once more on the frame's own columns with the cuts switched off. Rows with |z| > max_position
    are ignored; missing positions stay zero."""
    P = int(cfg.max_position)
    fwd, rev = cfg.substitution_bases_forward, cfg.substitution_bases_reverse
    hit = _DENSE.get(id(df_counts))
    if hit is not None and hit["max_position"] == P and hit["fwd"] == fwd and hit["rev"] == rev and \
            int(hit["first_row"][-1] if len(hit["first_row"]) else 0) <= len(df_counts):
        return hit
    ctx = ctx or get_context(0)
    n = len(df_counts)
    tax = df_counts["tax_id"].to_numpy(np.int64)
    if n == 0:
        z0 = np.zeros((0, 2 * P), np.uint32)
        return dict(tax_id=tax, tax_name=np.zeros(0, object), tax_rank=np.zeros(0, object), N_alignments=np.zeros(0, np.uint32),
                    k=z0, N=z0.copy(), noise=np.zeros((0, 3)), mism12=None, first_row=np.zeros(0, np.int64), max_position=P, fwd=fwd, rev=rev)
    z = df_counts["position"].to_numpy(np.int64)
    pos0 = np.clip(np.abs(z) - 1, 0, 254).astype(np.uint8)
    strand = df_counts["strand"]
    is_rev = (strand.astype(object).to_numpy() != "5'").astype(np.uint8) if "strand" in df_counts else (z < 0).astype(np.uint8)
    c16 = np.ascontiguousarray(df_counts[REF_OBS_BASES].to_numpy(np.uint32).T)
    r = ctx.counts_reduce(tax, df_counts["N_alignments"].to_numpy(np.uint32), is_rev, pos0, c16, fwd=fwd, rev=rev,
                          max_position=P, min_alignments=0, min_y_sum=0, want_noise=True, want_rows=False)
    if len(np.unique(r["tax_id"])) != r["n_tax"]:
        raise ValueError("df_counts: the rows of a TaxID must be contiguous")
    meta = df_counts.iloc[r["first_row"]]
    return dict(tax_id=r["tax_id"], tax_name=meta["tax_name"].astype(object).to_numpy(), tax_rank=meta["tax_rank"].astype(object).to_numpy(),
                N_alignments=r["n_alignments"], k=r["k"], N=r["N"], noise=r["noise"], mism12=None, first_row=r["first_row"],
                max_position=P, fwd=fwd, rev=rev)
