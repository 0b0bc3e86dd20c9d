"""The N > 1 path on CPU: world_size-2 gloo processes partition a TaxID batch exactly like
bench.py / the CLI do on GPUs (contiguous ranges, no collective on the fit path) and gather the
per-rank result rows on rank 0."""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np

from conftest import ROOT
from metadamage_b200.parallel import partition, run_on_gpus


def test_partition_is_contiguous_and_balanced():
    for n in (0, 1, 7, 8, 1000, 1_000_003):
        for g in (1, 2, 4, 8):
            b = partition(n, g)
            assert len(b) == g and b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(g - 1))
            sizes = [e - s for s, e in b]
            assert max(sizes) - min(sizes) <= 1


def test_run_on_gpus_threads_and_order():
    out = run_on_gpus(10, 4, lambda rank, a, b: (rank, list(range(a, b))))
    assert [o[0] for o in out] == [0, 1, 2, 3]
    assert sum((o[1] for o in out), []) == list(range(10))
    try:
        run_on_gpus(4, 2, lambda rank, a, b: 1 / 0)
    except ZeroDivisionError:
        pass
    else:
        raise AssertionError("worker exceptions must propagate")


WORKER = textwrap.dedent("""
    import os, sys
    import numpy as np
    import torch.distributed as dist
    sys.path.insert(0, os.environ["MDG_ROOT"])
    from metadamage_b200._abi import FIT_RESULT_DTYPE
    from metadamage_b200.parallel import partition, gather_structured

    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    n_tax = 11
    start, stop = partition(n_tax, world)[rank]
    # every rank "fits" its own contiguous share: a deterministic function of the tax id only
    local = np.zeros(stop - start, dtype=FIT_RESULT_DTYPE)
    local["tax_id"] = np.arange(start, stop) + 1000
    local["D_max"] = np.sin(local["tax_id"])
    local["run"]["n_leapfrog"][:, 0] = local["tax_id"] % 97
    med = np.outer(local["tax_id"], np.ones(30)).astype(np.float32)
    dist.barrier()
    res = gather_structured(local, dist)
    allmed = gather_structured(med, dist)
    if rank == 0:
        assert res is not None and len(res) == n_tax
        assert list(res["tax_id"]) == list(range(1000, 1000 + n_tax))
        assert np.allclose(res["D_max"], np.sin(res["tax_id"]))
        assert list(res["run"]["n_leapfrog"][:, 0]) == [t % 97 for t in range(1000, 1000 + n_tax)]
        assert allmed.shape == (n_tax, 30) and allmed[5, 3] == 1005
        print("GATHER_OK")
    else:
        assert res is None
    dist.destroy_process_group()
""")


def test_world_size_2_gloo_partition_and_gather(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    env = dict(os.environ, MDG_ROOT=ROOT, OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(script)]
    out = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=240)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "GATHER_OK" in out.stdout
