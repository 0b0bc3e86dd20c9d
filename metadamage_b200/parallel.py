"""Multi-GPU partitioning of the TaxID batch. TaxIDs are independent (groupby("tax_id")
everywhere: counts.py:201, fits.py:484, 577), so the batch is cut into contiguous ranges, one per
GPU, with NO collective on the fit path; results are only gathered on the host. Because every
random stream is keyed by (seed, tax_id), any partition gives bit-identical per-TaxID results.

Two drivers share `partition`:
  * one process, one host thread per GPU (used by the CLI via fits.compute_fits);
  * one process per GPU under torchrun (used by bench.py), results gathered with
    torch.distributed (NCCL on GPUs; gloo in the CPU tests).
"""
import threading

import numpy as np


def partition(n_items, n_parts):
    """Contiguous, balanced ranges: [(start, stop)] * n_parts (empty ranges allowed)."""
    n_parts = max(1, int(n_parts))
    base, extra = divmod(int(n_items), n_parts)
    bounds, start = [], 0
    for r in range(n_parts):
        stop = start + base + (1 if r < extra else 0)
        bounds.append((start, stop))
        start = stop
    return bounds


def run_on_gpus(n_items, n_gpus, worker):
    """Call worker(rank, start, stop) on one host thread per GPU (ctypes releases the GIL during
    the C-ABI call) and return the list of results in rank order. Exceptions are re-raised."""
    bounds = partition(n_items, n_gpus)
    results = [None] * len(bounds)
    errors = []

    def _run(rank, start, stop):
        try:
            results[rank] = worker(rank, start, stop)
        except BaseException as exc:  # noqa: BLE001 - re-raised in the caller's thread
            errors.append(exc)

    if len(bounds) == 1:
        _run(0, *bounds[0])
    else:
        threads = [threading.Thread(target=_run, args=(r, *b)) for r, b in enumerate(bounds) if b[1] > b[0]]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
    if errors:
        raise errors[0]
    return results


def gather_structured(local, dist, dst=0):
    """Gather per-rank numpy arrays (same dtype, ragged first dim) to rank `dst` through
    torch.distributed. Returns the concatenated array on `dst`, None elsewhere. Works with the
    gloo backend (CPU tensors) and with NCCL (tensors are staged on the current CUDA device)."""
    import torch

    rank, world = dist.get_rank(), dist.get_world_size()
    use_cuda = dist.get_backend() == "nccl"
    dev = torch.device("cuda", torch.cuda.current_device()) if use_cuda else torch.device("cpu")
    item = local.dtype.itemsize
    trailing = local.shape[1:]
    flat = np.ascontiguousarray(local).view(np.uint8).reshape(-1)
    sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([flat.size], dtype=torch.int64, device=dev))
    sizes = [int(s.item()) for s in sizes]
    cap = max(max(sizes), 1)
    send = torch.zeros(cap, dtype=torch.uint8, device=dev)
    if flat.size:
        send[: flat.size] = torch.from_numpy(flat.copy()).to(dev)
    recv = [torch.zeros(cap, dtype=torch.uint8, device=dev) for _ in range(world)]
    dist.all_gather(recv, send)
    if rank != dst:
        return None
    parts = [recv[r][: sizes[r]].cpu().numpy().view(local.dtype).reshape((-1,) + trailing) for r in range(world)]
    assert all(p.size * item == s for p, s in zip(parts, sizes))
    return np.concatenate(parts)
