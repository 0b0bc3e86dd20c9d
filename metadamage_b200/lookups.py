"""Dashboard-facing lookups on the result files: the data half of the reference's
`dashboard.fit_results.FitResults` (dashboard/fit_results.py:74-239) — derived columns, axis ranges,
marker sizes, slider / dropdown filters and the per-TaxID fetches the figures call
(dashboard/figures.py:454-516) — without Dash or Plotly. Host pandas on the parquet files this
package (or the reference) wrote; nothing here touches the GPU."""
from pathlib import Path

import numpy as np

from . import io

LOG_TRANSFORM_COLUMNS = ["N_alignments", "y_sum_total", "N_sum_total"]  # dashboard/utils.py:48


def log_transform_slider(x):
    """dashboard/utils.py:51-52"""
    return np.where(x < 0, 0, 10 ** np.clip(x, 0, a_max=None))


class FitResults:
    def __init__(self, folder):
        self.folder = Path(folder)
        self._load_df_fit_results()
        self._load_df_fit_predictions()
        self._compute_ranges()

    # ---- loading (dashboard/fit_results.py:74-105)
    def load_df_counts_shortname(self, shortname, columns=None):
        return io.Parquet(self.folder / "counts").load(shortname, columns=columns)

    def _load_df_fit_results(self):
        df = io.Parquet(self.folder / "fit_results").load()
        df["N_alignments_log10"] = np.log10(df["N_alignments"])
        df["N_alignments_sqrt"] = np.sqrt(df["N_alignments"])
        with np.errstate(divide="ignore", invalid="ignore"):
            df["N_sum_total_log10"] = np.log10(df["N_sum_total"])
        self.df_fit_results = df
        self.all_tax_ids = set(df.tax_id.unique())
        self.all_tax_names = set(df.tax_name.unique())
        self.all_tax_ranks = set(df.tax_rank.unique())
        self.shortnames = list(df.shortname.unique())
        self.columns = list(df.columns)
        self.set_marker_size(marker_transformation="sqrt")

    def _load_df_fit_predictions(self):
        self.df_fit_predictions = io.Parquet(self.folder / "fit_predictions").load()

    # ---- axis ranges (dashboard/fit_results.py:107-148)
    def _get_range_of_column(self, column, spacing):
        array = self.df_fit_results[column]
        array = array[np.isfinite(array) & array.notnull()]
        range_min, range_max = array.min(), array.max()
        delta = range_max - range_min
        return [range_min - delta / spacing, range_max + delta / spacing]

    def _compute_ranges(self, spacing=20):
        ranges = {}
        for column in self.columns:
            try:
                ranges[column] = self._get_range_of_column(column, spacing=spacing)
            except TypeError:  # categorical columns
                pass
        for column, range_ in ranges.items():
            if "_forward" in column or "_reverse" in column:
                continue
            fwd, rev = f"{column}_forward", f"{column}_reverse"
            if fwd in ranges and rev in ranges:
                padding = {"n_sigma": 1, "D_max": 0.1, "noise": 1}.get(column)
                if padding is None:
                    # the reference leaves `paddding` unbound here (UnboundLocalError, or the previous
                    # column's value); no other column has _forward/_reverse partners in its schema
                    continue
                for key in (fwd, rev):
                    r = ranges[key]
                    if r[0] < range_[0] - padding:
                        r[0] = range_[0] - padding
                    if r[1] > range_[1] + padding:
                        r[1] = range_[1] + padding
        self.ranges = ranges

    # ---- marker sizes (dashboard/fit_results.py:150-182)
    def set_marker_size(self, marker_transformation="sqrt", marker_size_max=30):
        df = self.df_fit_results
        if isinstance(marker_transformation, list) and isinstance(marker_size_max, list):
            if len(marker_transformation) == 0 and len(marker_size_max) == 0:
                return None
            marker_transformation, marker_size_max = marker_transformation[0], marker_size_max[0]
        if marker_transformation == "identity":
            df.loc[:, "size"] = df["N_alignments"]
        elif marker_transformation == "sqrt":
            df.loc[:, "size"] = np.sqrt(df["N_alignments"])
        elif marker_transformation == "log10":
            df.loc[:, "size"] = np.log10(df["N_alignments"])
        elif marker_transformation == "constant":
            df.loc[:, "size"] = np.ones_like(df["N_alignments"])
        else:
            raise AssertionError(f"Did not recieve proper marker_transformation: {marker_transformation}")
        self.max_of_size = np.max(df["size"])
        self.marker_size_max = marker_size_max
        return None

    # ---- filters (dashboard/fit_results.py:184-229)
    def filter(self, filters, df_type="df_fit_results"):
        query = ""
        for column, flt in filters.items():
            if flt is None:
                continue
            elif column == "shortnames":
                query += f"(shortname in {flt}) & "
            elif column == "shortname":
                query += f"(shortname == '{flt}') & "
            elif column == "tax_id":
                query += f"(tax_id == {flt}) & "
            elif column == "tax_ids":
                query += f"(tax_id in {flt}) & "
            elif column == "tax_rank":
                query += f"(tax_rank == {flt}) & "
            elif column == "tax_ranks":
                query += f"(tax_rank in {flt}) & "
            elif column == "tax_name":
                query += f"(tax_name == {flt}) & "
            elif column == "tax_names":
                query += f"(tax_name in {flt}) & "
            else:
                low, high = flt
                if column in LOG_TRANSFORM_COLUMNS:
                    low, high = log_transform_slider(low), log_transform_slider(high)
                query += f"({low} <= {column} <= {high}) & "
        query = query[:-2]
        if "fit_results" in df_type:
            return self.df_fit_results.query(query)
        raise AssertionError(f"df_type = {df_type} not implemented yet, only 'df_fit_results'")

    # ---- single-TaxID fetches (dashboard/fit_results.py:231-237)
    def get_single_count_group(self, shortname, tax_id):
        return self.load_df_counts_shortname(shortname).query(f"tax_id == {tax_id}")

    def get_single_fit_prediction(self, shortname, tax_id):
        return self.df_fit_predictions.query(f"shortname == '{shortname}' & tax_id == {tax_id}")
