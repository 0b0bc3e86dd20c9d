"""K0 tsv_parse (the step before the hot path): oracle vs pandas.read_csv — the
parser the reference uses (counts.py:229-235) — on CPU; CUDA vs oracle vs pandas under -m gpu."""
import io

import numpy as np
import pandas as pd
import pytest

from metadamage_b200 import counts
from test_host import write_tsv

KEYS = ("tax_id", "n_alignments", "is_reverse", "pos0", "counts16")


def pandas_parse(text, n_cols, header):
    names = counts.COLUMNS if n_cols == 22 else counts.LEGACY_COLUMNS
    df = pd.read_csv(io.BytesIO(text), sep="\t", header=0 if header else None, names=names)
    return dict(tax_id=df["tax_id"].to_numpy(np.int64), n_alignments=df["N_alignments"].to_numpy(np.uint32),
                is_reverse=(df["strand"].to_numpy() != "5'").astype(np.uint8), pos0=df["position"].to_numpy(np.uint8),
                counts16=np.ascontiguousarray(df[counts.REF_OBS_BASES].to_numpy(np.uint32).T)), df


def sample_text(tmp_path, sample, legacy):
    path = tmp_path / "x.txt"
    write_tsv(path, sample, legacy)
    return path.read_bytes()


def synthetic_text(n_tax, n_cols, seed=0, crlf=False, trailing_newline=True):
    rng = np.random.default_rng(seed)
    lines = []
    for t in range(n_tax):
        tax = int(rng.integers(-5, 10 ** 9))
        nal = int(rng.integers(1, 2 ** 32 - 1))
        for strand in ("5'", "3'"):
            for pos in range(int(rng.integers(1, 6))):
                cnt = [str(int(v)) for v in rng.integers(0, [10, 10 ** 5, 2 ** 32 - 1][int(rng.integers(0, 3))], 16)]
                head = [str(tax)] + (["Homo sapiens", "species"] if n_cols == 22 else []) + [str(nal), strand, str(pos)]
                lines.append("\t".join(head + cnt))
    nl = "\r\n" if crlf else "\n"
    return (nl.join(lines) + (nl if trailing_newline else "")).encode()


def assert_same(a, b):
    assert a["n_rows"] == len(b["tax_id"])
    for key in KEYS:
        assert np.array_equal(a[key], b[key]), key


@pytest.mark.parametrize("legacy", [True, False])
def test_oracle_tsv_matches_pandas_on_sample(oracle, tmp_path, sample_inputs, legacy):
    text = sample_text(tmp_path, sample_inputs["ancient"], legacy)
    got = oracle.tsv_parse(text)
    ref, _ = pandas_parse(text, 20 if legacy else 22, header=legacy)
    assert got["n_cols"] == (20 if legacy else 22)
    assert_same(got, ref)
    for key in KEYS:
        assert np.array_equal(got[key], sample_inputs["ancient"][key])


@pytest.mark.parametrize("n_cols,crlf,trailing", [(20, False, True), (22, False, False), (22, True, True), (20, True, False)])
def test_oracle_tsv_matches_pandas_on_synthetic(oracle, n_cols, crlf, trailing):
    text = synthetic_text(300, n_cols, seed=n_cols + crlf, crlf=crlf, trailing_newline=trailing)
    ref, _ = pandas_parse(text, n_cols, header=False)
    assert_same(oracle.tsv_parse(text), ref)


@pytest.mark.gpu
@pytest.mark.parametrize("legacy", [True, False])
def test_gpu_tsv_matches_oracle_and_pandas_on_sample(ctx, oracle, tmp_path, sample_inputs, legacy):
    text = sample_text(tmp_path, sample_inputs["control"], legacy)
    got = ctx.tsv_parse(text, want_spans=not legacy)
    assert got["n_cols"] == (20 if legacy else 22)
    assert_same(got, oracle.tsv_parse(text))
    assert_same(got, pandas_parse(text, got["n_cols"], header=legacy)[0])
    if not legacy:
        o, n = got["name_span"][7]
        assert text[o:o + n] == b"Homo sapiens"
        o, n = got["rank_span"][89]
        assert text[o:o + n] == b"species"


@pytest.mark.gpu
@pytest.mark.parametrize("n_cols,crlf,trailing", [(20, False, True), (22, False, False), (22, True, True), (20, True, False)])
def test_gpu_tsv_matches_oracle_on_synthetic(ctx, oracle, n_cols, crlf, trailing):
    text = synthetic_text(4000, n_cols, seed=10 + n_cols + crlf, crlf=crlf, trailing_newline=trailing)
    got = ctx.tsv_parse(text)
    assert_same(got, oracle.tsv_parse(text))
    assert got["n_rows"] == text.count(b"\n") + (0 if text.endswith(b"\n") else 1)


@pytest.mark.gpu
def test_gpu_tsv_edge_cases(ctx):
    from metadamage_b200._lib import MdgError

    assert ctx.tsv_parse(b"")["n_rows"] == 0
    assert ctx.tsv_parse(b"#taxid\tN\n")["n_rows"] == 0  # header only
    one = b"7\t12\t5'\t0\t" + b"\t".join(str(i).encode() for i in range(16))
    r = ctx.tsv_parse(one)
    assert r["n_rows"] == 1 and r["tax_id"][0] == 7 and r["counts16"][:, 0].tolist() == list(range(16)) and r["is_reverse"][0] == 0
    r = ctx.tsv_parse(one.replace(b"5'", b"3'") + b"\n")
    assert r["n_rows"] == 1 and r["is_reverse"][0] == 1
    with pytest.raises(MdgError, match="fields"):
        ctx.tsv_parse(one + b"\n" + one[:-3].rsplit(b"\t", 1)[0] + b"\n")        # a line with 19 fields
    with pytest.raises(MdgError, match="number"):
        ctx.tsv_parse(one.replace(b"\t12\t", b"\t1x2\t") + b"\n")
    with pytest.raises(MdgError, match="range"):
        ctx.tsv_parse(one.replace(b"\t12\t", b"\t99999999999\t") + b"\n")
    with pytest.raises(MdgError, match="columns"):
        ctx.tsv_parse(b"1\t2\t3\n")
    # blank lines at the end of the file are not rows (pandas.read_csv skips them)
    r = ctx.tsv_parse(one + b"\n" + one + b"\n\n\r\n")
    assert r["n_rows"] == 2 and r["tax_id"].tolist() == [7, 7]
    assert ctx.tsv_parse(b"\n\n")["n_rows"] == 0
    # the text is a host pointer whatever `mem` says; a device pointer is refused, not dereferenced
    import ctypes as C

    import torch

    from metadamage_b200._abi import MDG_DEVICE

    dev_text = torch.zeros(64, dtype=torch.uint8, device="cuda")
    cap = 4
    cols = [torch.empty(cap, dtype=torch.int64, device="cuda"), torch.empty(cap, dtype=torch.int32, device="cuda"),
            torch.empty(cap, dtype=torch.uint8, device="cuda"), torch.empty(cap, dtype=torch.uint8, device="cuda"),
            torch.empty((16, cap), dtype=torch.int32, device="cuda")]
    n_rows, n_cols = C.c_int64(0), C.c_int32(0)
    rc = ctx._lib.mdg_tsv_parse(ctx._h, MDG_DEVICE, C.cast(dev_text.data_ptr(), C.c_char_p), 64, cap, *[C.c_void_p(t.data_ptr()) for t in cols], cap,
                                None, None, C.byref(n_rows), C.byref(n_cols))
    assert rc == -1 and b"host pointer" in ctx._lib.mdg_last_error()


@pytest.mark.gpu
def test_gpu_tsv_large_file_roundtrip(ctx):
    """~100 MB of text: parse -> counts_reduce equals counts_reduce on the generator's own arrays."""
    from metadamage_b200 import synthetic as syn

    g = syn.make_mismatch_matrix(40_000, seed=5)
    n = len(g["tax_id"])
    cols = [g["tax_id"].astype(str), g["n_alignments"].astype(str), np.where(g["is_reverse"] == 1, "3'", "5'"), g["pos0"].astype(str)]
    cols += [g["counts16"][i].astype(str) for i in range(16)]
    text = ("\n".join("\t".join(row) for row in zip(*cols)) + "\n").encode()
    r = ctx.tsv_parse(text)
    assert r["n_rows"] == n
    for key in KEYS:
        assert np.array_equal(r[key], g[key]), key
