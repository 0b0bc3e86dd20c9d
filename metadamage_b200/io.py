"""Parquet result files with the run configuration stored as JSON in the schema metadata
under the key b"metadamage" — the on-disk format of the reference (io.py:19-91), which its
dashboard reads (dashboard/fit_results.py:74-101)."""
import json
from pathlib import Path

import pyarrow as pa
import pyarrow.parquet as pq

META_KEY = b"metadamage"


class Parquet:
    def __init__(self, filename):
        self.filename = Path(filename)

    def __repr__(self):
        return f"Parquet('{self.filename}')"

    def exists(self, forced=False):
        return self.filename.exists() and not forced

    def load_metadata(self):
        schema = pq.read_schema(self.filename)
        return json.loads(schema.metadata[META_KEY])

    def load(self, shortname=None, tax_id=None, columns=None):
        filename = self.filename if shortname is None else self.filename / f"{shortname}.parquet"
        filters = None if tax_id is None else [("tax_id", "==", tax_id)]
        if isinstance(columns, str):
            columns = [columns]
        df = pq.read_table(filename, filters=filters, columns=columns).to_pandas()
        if "tax_id" in df.columns:
            df = df.astype({"tax_id": "category"})
        return df

    def save(self, df, metadata=None):
        self.filename.parent.mkdir(parents=True, exist_ok=True)
        table = pa.Table.from_pandas(df)
        merged = {META_KEY: json.dumps(metadata or {}).encode(), **(table.schema.metadata or {})}
        pq.write_table(table.replace_schema_metadata(merged), self.filename)
