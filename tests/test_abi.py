"""The C-ABI shared library loads on a CPU-only box and exports every symbol include/mdg.h
declares; struct layouts seen from Python match the header. No compute call is made here."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def lib():
    from metadamage_b200 import build, _lib

    build.build()
    return _lib.load()


def declared_functions():
    text = open(os.path.join(ROOT, "include", "mdg.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mdg_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported(lib):
    from metadamage_b200 import _lib

    names = declared_functions()
    assert len(names) >= 15
    for name in names:
        assert hasattr(lib, name), f"{name} is declared in include/mdg.h but not exported"
    assert sorted(_lib.EXPORTED_SYMBOLS) == names


def test_version_and_default_config(lib):
    from metadamage_b200 import _lib

    assert lib.mdg_version() == 200
    cfg = _lib.default_config()
    # fits.py:792-799 and the priors of fits.py:43-67
    assert (cfg.num_warmup, cfg.num_samples, cfg.max_tree_depth) == (500, 1000, 10)
    assert (cfg.q_prior_a, cfg.q_prior_b, cfg.A_prior_a, cfg.A_prior_b, cfg.c_prior_a, cfg.c_prior_b) == (2, 3, 2, 3, 1, 9)
    assert cfg.phi_prior_rate == 1 / 1000 and cfg.phi_min == 2 and cfg.target_accept == 0.8 and cfg.hpdi_prob == 0.68


def test_struct_layouts_match_header(tmp_path):
    from metadamage_b200._abi import FIT_RESULT_DTYPE, RUN_DIAG_DTYPE, FitConfig, Timings

    src = tmp_path / "sz.c"
    src.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "mdg.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu\\n",'
        "sizeof(mdg_fit_config),sizeof(mdg_fit_result),sizeof(mdg_run_diag),sizeof(mdg_timings),"
        "offsetof(mdg_fit_result,run),offsetof(mdg_fit_result,map_A));return 0;}\n")
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    sizes = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert sizes == [ctypes.sizeof(FitConfig), FIT_RESULT_DTYPE.itemsize, RUN_DIAG_DTYPE.itemsize,
                     ctypes.sizeof(Timings), FIT_RESULT_DTYPE.fields["run"][1], FIT_RESULT_DTYPE.fields["map_A"][1]]


def test_no_cpu_fallback(lib):
    """Without a GPU the product path must fail loudly, never fall back to a CPU implementation."""
    from metadamage_b200 import _lib
    from metadamage_b200.backend import Context

    if lib.mdg_device_count() > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(_lib.MdgError):
        Context(0)


def test_product_does_not_import_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py may touch oracle/."""
    pkg = os.path.join(ROOT, "metadamage_b200")
    forbidden = re.compile(r"(^|\n)\s*(from|import)\s+oracle\b|libmdg_oracle|mdg_oracle\.c|orc_[a-z_]+\(")
    checked = 0
    for top in (pkg, os.path.join(ROOT, "metadamage")):
        for dirpath, _, files in os.walk(top):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h")):
                    text = open(os.path.join(dirpath, f)).read()
                    assert not forbidden.search(text), f"{f} reaches into oracle/"
                    checked += 1
    assert checked > 10


def test_oracle_and_library_share_struct_layout(oracle):
    cfg_o = oracle.default_config()
    from metadamage_b200 import _lib

    assert bytes(cfg_o) == bytes(_lib.default_config())
    assert np.dtype(oracle.FIT_RESULT_DTYPE).itemsize == 568
