"""Array-level lookups over the fit outputs (what the reference's dashboard asks of the result
files — dashboard/fit_results.py:82-97 derived columns, 107-148 axis ranges, 150-182 marker sizes, 184-229
slider / dropdown filters, 231-237 single-TaxID fetches; dashboard/figures.py:454-516 consumes them).

`ResultArrays` keeps the fit results as a structure of arrays in file order — numeric columns as numpy
arrays, categorical columns as integer codes + category tables — and answers every lookup with array
operations: filters are boolean masks, a TaxID's 2P prediction rows / count rows are a slice found
through per-(shortname, TaxID) offsets (`first_row`, the same offsets mdg_counts_reduce returns). It is
built from the parquet files this package (or the reference) wrote, or directly from the `mdg_fit_result`
rows of a fit that is still in memory (`ResultArrays.from_fit`). No Dash, no Plotly, no DataFrame queries.
"""
from pathlib import Path

import numpy as np

from . import io

LOG_SLIDER_COLUMNS = ("N_alignments", "y_sum_total", "N_sum_total")  # dashboard/utils.py:48
RANGE_PADDING = {"n_sigma": 1, "D_max": 0.1}                       # dashboard/fit_results.py:133-139
MARKER_TRANSFORMS = {
    "identity": lambda n: n.astype(np.float64),
    "sqrt": lambda n: np.sqrt(n.astype(np.float64)),
    "log10": lambda n: np.log10(n.astype(np.float64)),
    "constant": lambda n: np.ones(len(n), np.float64),
}


def slider_to_value(x):
    """dashboard/utils.py:51-52: slider positions of the log-scaled columns are exponents of ten."""
    x = np.asarray(x, dtype=np.float64)
    return np.where(x < 0, 0.0, 10.0 ** np.clip(x, 0, None))


class Categorical:
    """codes into a table of distinct values"""

    def __init__(self, values):
        self.categories, self.codes = np.unique(np.asarray(values, dtype=object).astype(str) if len(values) and
                                                not np.issubdtype(np.asarray(values).dtype, np.number) else np.asarray(values),
                                                return_inverse=True)

    def mask_eq(self, value):
        # the dashboard hands tax_rank / tax_name over already quoted for its query string (fit_results.py:203-212)
        if isinstance(value, str) and len(value) >= 2 and value[0] == value[-1] and value[0] in "'\"":
            value = value[1:-1]
        hit = np.flatnonzero(self.categories == (str(value) if self.categories.dtype.kind in "UO" else value))
        return self.codes == hit[0] if len(hit) else np.zeros(len(self.codes), bool)

    def mask_in(self, values):
        want = np.array([str(v) if self.categories.dtype.kind in "UO" else v for v in values], dtype=self.categories.dtype)
        return np.isin(self.codes, np.flatnonzero(np.isin(self.categories, want)))

    def values(self):
        return self.categories[self.codes]


def _segment_offsets(shortname_codes, tax_id):
    """{(shortname code, tax id): (first_row, n_rows)} for rows grouped by (shortname, TaxID)."""
    n = len(tax_id)
    if n == 0:
        return {}
    head = np.flatnonzero(np.r_[True, (tax_id[1:] != tax_id[:-1]) | (shortname_codes[1:] != shortname_codes[:-1])])
    length = np.diff(np.r_[head, n])
    return {(int(shortname_codes[h]), int(tax_id[h])): (int(h), int(m)) for h, m in zip(head, length)}


class ResultArrays:
    def __init__(self, numeric, categorical, predictions=None, counts_loader=None):
        self.numeric = dict(numeric)          # column -> float64 / int64 array [n_fits]
        self.categorical = dict(categorical)  # column -> Categorical
        self.n = len(next(iter(self.numeric.values()))) if self.numeric else 0
        n_al = self.numeric["N_alignments"].astype(np.float64)
        self.numeric["N_alignments_log10"] = np.log10(n_al)           # fit_results.py:84
        self.numeric["N_alignments_sqrt"] = np.sqrt(n_al)             # fit_results.py:85
        with np.errstate(divide="ignore", invalid="ignore"):
            self.numeric["N_sum_total_log10"] = np.log10(self.numeric["N_sum_total"].astype(np.float64))
        self.set_marker_size("sqrt", 30)
        self.predictions = predictions
        self._pred_offsets = None
        self._counts_loader = counts_loader
        self._counts = {}

    # ------------------------------------------------------------------ construction
    @classmethod
    def from_folder(cls, folder):
        """From out_dir/{fit_results, fit_predictions, counts}/<shortname>.parquet (io.py:19-91)."""
        folder = Path(folder)
        df = io.Parquet(folder / "fit_results").load()
        numeric, categorical = {}, {}
        for col in df.columns:
            if str(df[col].dtype) in ("category", "object", "str", "string"):
                vals = df[col].astype(object).to_numpy()
                categorical[col] = Categorical(vals.astype(np.int64) if col == "tax_id" else vals)
            else:
                numeric[col] = df[col].to_numpy()
        pred = io.Parquet(folder / "fit_predictions").load()
        short = Categorical(pred["shortname"].astype(object).to_numpy())
        predictions = dict(shortname=short, tax_id=pred["tax_id"].astype(object).to_numpy().astype(np.int64),
                           **{c: pred[c].to_numpy() for c in ("position", "median", "hdpi_lower", "hdpi_upper")})

        def counts_loader(shortname):
            d = io.Parquet(folder / "counts").load(shortname)
            return {c: (d[c].astype(object).to_numpy() if str(d[c].dtype) == "category" else d[c].to_numpy()) for c in d.columns}

        return cls(numeric, categorical, predictions, counts_loader)

    @classmethod
    def from_fit(cls, result, dense, shortname, median=None, hpdi_lo=None, hpdi_hi=None):
        """Straight from a fit still in memory: `result` = mdg_fit_result rows (backend.fit_batch), `dense` = the
        per-TaxID arrays of counts.dense_from_df_counts (tax_name, tax_rank, N_alignments); failed fits dropped."""
        from ._abi import FIT_FAILED
        from .fits import FIT_RESULT_COLUMNS

        ok = (result["status"] & FIT_FAILED) == 0
        numeric = {c: np.asarray(result[c][ok]) for c in FIT_RESULT_COLUMNS if c in result.dtype.names and c != "tax_id"}
        numeric["N_alignments"] = np.asarray(dense["N_alignments"])[ok]
        categorical = dict(tax_id=Categorical(np.asarray(dense["tax_id"])[ok]), tax_name=Categorical(np.asarray(dense["tax_name"])[ok]),
                           tax_rank=Categorical(np.asarray(dense["tax_rank"])[ok]), shortname=Categorical(np.array([shortname] * int(ok.sum()), dtype=object)))
        predictions = None
        if median is not None:
            R = median.shape[1]
            z = np.arange(R // 2) + 1
            predictions = dict(shortname=Categorical(np.array([shortname] * (int(ok.sum()) * R), dtype=object)),
                               tax_id=np.repeat(np.asarray(dense["tax_id"])[ok], R), position=np.tile(np.r_[z, -z], int(ok.sum())),
                               median=median[ok].ravel(), hdpi_lower=hpdi_lo[ok].ravel(), hdpi_upper=hpdi_hi[ok].ravel())
        return cls(numeric, categorical, predictions)

    # ------------------------------------------------------------------ marker sizes (fit_results.py:150-182)
    def set_marker_size(self, transformation="sqrt", size_max=30):
        if isinstance(transformation, list) and isinstance(size_max, list):
            if not transformation and not size_max:
                return None
            transformation, size_max = transformation[0], size_max[0]
        if transformation not in MARKER_TRANSFORMS:
            raise AssertionError(f"Did not recieve proper marker_transformation: {transformation}")
        self.numeric["size"] = MARKER_TRANSFORMS[transformation](self.numeric["N_alignments"])
        self.max_of_size = float(np.max(self.numeric["size"])) if self.n else 0.0
        self.marker_size_max = size_max
        return None

    # ------------------------------------------------------------------ axis ranges (fit_results.py:107-148)
    def ranges(self, spacing=20):
        out = {}
        for col, arr in self.numeric.items():
            if col == "size":  # the marker size is not an axis (the reference lists its columns before adding it)
                continue
            a = np.asarray(arr)
            a = a[np.isfinite(a)]
            lo, hi = (a.min(), a.max()) if len(a) else (np.nan, np.nan)
            delta = hi - lo  # in the column's own precision, as the reference's pandas arithmetic does
            out[col] = [lo - delta / spacing, hi + delta / spacing]
        for col, pad in RANGE_PADDING.items():  # the forward / reverse panels never reach further than the combined one + padding
            for side in (f"{col}_forward", f"{col}_reverse"):
                if col in out and side in out:
                    out[side][0] = max(out[side][0], out[col][0] - pad)
                    out[side][1] = min(out[side][1], out[col][1] + pad)
        return out

    # ------------------------------------------------------------------ filters (fit_results.py:184-229)
    def mask(self, filters):
        """Boolean mask of the fits that pass every filter: `shortname(s)`, `tax_id(s)`, `tax_rank(s)`, `tax_name(s)`
        (equality / membership) and (low, high) intervals on numeric columns, the log-scaled sliders in exponents."""
        m = np.ones(self.n, bool)
        for key, flt in filters.items():
            if flt is None:
                continue
            col = key[:-1] if key.endswith("s") and key[:-1] in self.categorical else key
            if col in self.categorical and key != col:
                m &= self.categorical[col].mask_in(flt)
            elif col in self.categorical:
                m &= self.categorical[col].mask_eq(flt)
            else:
                low, high = flt
                if key in LOG_SLIDER_COLUMNS:
                    low, high = slider_to_value(low), slider_to_value(high)
                v = self.numeric[key]
                m &= (v >= low) & (v <= high)
        return m

    def select(self, filters):
        """Row numbers (file order) of the fits that pass the filters."""
        return np.flatnonzero(self.mask(filters))

    # ------------------------------------------------------------------ single-TaxID fetches (fit_results.py:231-237)
    def prediction(self, shortname, tax_id):
        """The 2P rows of df_fit_predictions of one (sample, TaxID): a slice through the per-fit offsets."""
        p = self.predictions
        if self._pred_offsets is None:
            self._pred_offsets = _segment_offsets(p["shortname"].codes, p["tax_id"])
        code = np.flatnonzero(p["shortname"].categories == shortname)
        first, n = self._pred_offsets.get((int(code[0]) if len(code) else -1, int(tax_id)), (0, 0))
        return {c: p[c][first:first + n] for c in ("position", "median", "hdpi_lower", "hdpi_upper")}

    def counts_group(self, shortname, tax_id):
        """The rows of one TaxID in counts/<shortname>.parquet (rows of a TaxID are contiguous there)."""
        if shortname not in self._counts:
            cols = self._counts_loader(shortname)
            tax = np.asarray(cols["tax_id"]).astype(np.int64)
            self._counts[shortname] = (cols, _segment_offsets(np.zeros(len(tax), np.int64), tax))
        cols, offsets = self._counts[shortname]
        first, n = offsets.get((0, int(tax_id)), (0, 0))
        return {c: v[first:first + n] for c, v in cols.items()}
