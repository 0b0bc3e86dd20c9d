#!/usr/bin/env python
"""bench.py — TaxID damage fits/sec on N B200s (BASELINE.json metric), one JSON line on rank 0.

    python bench.py --gpus 1 --steps 3 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...        # the CPU arm (restated reference, all host cores)

A "step" is one pass of the hot path over one batch: counts_reduce (K1) over the synthetic
mismatch matrix, then MAP + 6 NUTS runs + WAIC + posterior predictive + row assembly (K3-K7) for
every TaxID that passes the cuts. Per-GPU workload (weak scaling): BASELINE config 2, "synthetic
10k-TaxID mismatch matrix, +-15 positions" = 10 000 fitted TaxIDs per GPU (the generator keeps
the ~47 000 TaxIDs that fail the cuts in the input so the cut/compaction path of K1 runs too).

  value  whole-job fits/s with inputs resident in HBM (device pointers through the C-ABI)
  e2e    the same through the host-buffer C-ABI calls the Python seams make (pinned host inputs,
         H2D + D2H inside the timed region)
  roofline         the dominant kernel (NUTS): SURVEY.md 8d flop model / CUDA-event time, against
                   the FP64 FMA peak measured live on this GPU
  roofline_counts  K1 on the 10M-row stress input (BASELINE config 5): 111 algorithmic B/row
                   against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline     the oracle (C port of the reference algorithm; numpyro is not installable
                   offline) on all host cores, on a bounded sample of the same workload
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "taxid_damage_fits_per_sec"
UNIT = "fits/s"
ALGO_BYTES_PER_ROW = 111
# SURVEY.md 8d nominal FP64 flop model per log-density-gradient evaluation
FLOPS_PER_GRAD = {0: 300 * 30 + 55, 1: 190 * 30 + 165, 2: 300 * 15 + 55, 3: 190 * 15 + 165, 4: 300 * 15 + 55, 5: 190 * 15 + 165}


def model_flops(leapfrogs, max_position):
    scale = max_position / 15.0
    return sum(l * (FLOPS_PER_GRAD[r] * scale) for r, l in enumerate(leapfrogs))


def workload(args, rank):
    from metadamage_b200 import synthetic as syn

    seed = syn.SEEDS["cfg2"] + 1000 * rank
    g = syn.make_mismatch_matrix(0, max_position=args.max_position, seed=seed, n_fit=args.taxa_per_gpu,
                                 tax_id_start=1 + rank * 100_000_000)
    return g


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.device}", f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                smax.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def cpu_baseline(g, args, n_sample=None, threads=0):
    """The restated reference (oracle, C, OpenMP over TaxIDs) on a bounded sample."""
    from oracle import oracle as O

    O.build()
    cores = threads or (os.cpu_count() or 1)
    sel = np.flatnonzero(g["passes"])
    n_sample = n_sample or max(8, min(len(sel), 64 * cores))  # ~10 s of CPU work on all cores
    # an evenly spread sample of the same TaxIDs the GPU fits
    pick = sel[np.linspace(0, len(sel) - 1, n_sample).astype(int)]
    cfg = O.default_config()
    t0 = time.perf_counter()
    out = O.fit_batch(g["tax_ids"][pick], g["k"][pick], g["N"][pick], cfg, n_threads=cores)
    dt = time.perf_counter() - t0
    ok = int(((out["result"]["status"] & 1) == 0).sum())
    return {"value": n_sample / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{n_sample} of the {len(sel)} fitted TaxIDs of the same workload, evenly spread; "
                      f"full fit (MAP + 6 NUTS runs 500+1000 + WAIC + predictive); {dt:.1f} s; {ok} ok; "
                      "restated reference (numpyro unavailable offline)"}, dt


def workload_config(args, world, n_in_tax=None, n_rows=None):
    P = args.max_position
    shape = f" ({n_in_tax} input TaxIDs, {n_rows} rows)" if n_in_tax else ""
    return {
        "workload": f"cfg2: synthetic heavy-tailed mismatch matrix, {args.taxa_per_gpu} fitted TaxIDs per GPU{shape}, "
                    f"+-{P} positions, counts + MAP + 6 NUTS runs (500 warm-up + 1000 draws) + WAIC + predictive D_max",
        "taxa_per_gpu": args.taxa_per_gpu, "max_position": P,
        "partition": f"by TaxID over {world} GPU(s), no collective on the fit path",
        "l2": "flushed between steps (256 MiB write)"}


def run_reference(args):
    """The reference arm: the reference's algorithm on the host cores. numpyro/jax cannot be
    installed offline, so this times the C restatement (oracle) with all host threads, each step
    on a bounded, evenly spread sample of the same 10k-TaxID workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    g = workload(args, 0)
    cores = os.cpu_count() or 1
    n_fit = int(g["passes"].sum())
    n_sample = max(8, min(n_fit, 32 * cores))
    for _ in range(args.warmup):
        cpu_baseline(g, args, n_sample=max(4, 2 * cores))
    times, last = [], None
    for _ in range(args.steps):
        last, dt = cpu_baseline(g, args, n_sample=n_sample)
        times.append(dt)
    value = n_sample * len(times) / sum(times)
    last["value"] = value
    cfg = workload_config(args, args.gpus, len(g["tax_ids"]), len(g["tax_id"]))
    cfg["reference_sample_per_step"] = n_sample
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
        "cpu_baseline": last,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def counts_stress(ctx, torch, dev, n_rows=10_000_000, reps=5):
    """BASELINE config 5: 10M-row counts aggregation, device resident, HBM-bound."""
    P = 15
    gen = torch.Generator(device=dev)
    gen.manual_seed(20240004)
    n_tax = (n_rows + 2 * P - 1) // (2 * P)
    tax = torch.repeat_interleave(torch.arange(1, n_tax + 1, device=dev, dtype=torch.int64), 2 * P)[:n_rows].contiguous()
    u = torch.rand(n_tax, device=dev, generator=gen)
    n_al = torch.clamp((10.0 / (1.0 - u) ** (1 / 1.1)), 10, 6e7).to(torch.int32)
    n_al_row = torch.repeat_interleave(n_al, 2 * P)[:n_rows].contiguous()
    pos = torch.arange(n_rows, device=dev) % (2 * P)
    is_rev = (pos >= P).to(torch.uint8)
    pos0 = (pos % P).to(torch.uint8)
    frac = torch.rand((16, n_rows), device=dev, generator=gen) * 0.05
    frac[::5] = 0.2  # the diagonal carries the coverage
    counts16 = (frac * n_al_row.to(torch.float32)[None, :]).to(torch.int32).contiguous()
    cols = dict(tax_id=tax, n_alignments=n_al_row, is_reverse=is_rev, pos0=pos0, counts16=counts16)
    outs = dict(
        n_fwd_ref=torch.empty(n_rows, dtype=torch.int32, device=dev), n_rev_ref=torch.empty(n_rows, dtype=torch.int32, device=dev),
        f_fwd=torch.empty(n_rows, dtype=torch.float32, device=dev), f_rev=torch.empty(n_rows, dtype=torch.float32, device=dev),
        z=torch.empty(n_rows, dtype=torch.int8, device=dev), y_sum_total=torch.empty(n_rows, dtype=torch.int64, device=dev),
        keep=torch.empty(n_rows, dtype=torch.uint8, device=dev), tax_id=torch.empty(n_tax, dtype=torch.int64, device=dev),
        n_alignments=torch.empty(n_tax, dtype=torch.int32, device=dev), first_row=torch.empty(n_tax, dtype=torch.int64, device=dev),
        k=torch.empty((n_tax, 2 * P), dtype=torch.int32, device=dev), N=torch.empty((n_tax, 2 * P), dtype=torch.int32, device=dev),
    )
    times = []
    kept = 0
    for _ in range(reps + 2):
        kept = ctx.counts_reduce_device(cols, outs)
        times.append(ctx.timings()["counts_ms"])
    ms = float(np.mean(times[2:]))
    peak, which = hbm_peak()
    achieved = ALGO_BYTES_PER_ROW * n_rows / (ms * 1e-3) / 1e9
    # dram__bytes_read.sum + dram__bytes_write.sum of one launch on this 10M-row input, from the ncu --set full
    # capture summarised in profiles/r01_counts_ncu.md (460 MB + 255 MB); not re-measured live
    traffic = 715e6 if n_rows == 10_000_000 else None
    return {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
            "traffic_source": "ncu --set full capture (profiles/r01_counts_ncu.md), bytes per launch",
            "kernel": "counts_reduce_kernel", "rows": n_rows, "kept_taxa": int(kept), "kernel_ms": ms,
            "algorithmic_bytes_per_row": ALGO_BYTES_PER_ROW, "peak_source": which,
            "note": "inputs (780 MB) exceed L2; mean of %d launches after 2 warm-ups, CUDA events on the launch stream" % reps}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--taxa-per-gpu", type=int, default=10_000, help="fitted TaxIDs per GPU per step (cfg2: 10k; cfg3: 125k at 8 GPUs)")
    ap.add_argument("--max-position", type=int, default=15)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-counts-stress", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist

    from metadamage_b200 import _lib
    from metadamage_b200._abi import FIT_RESULT_DTYPE
    from metadamage_b200.backend import Context

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    ctx = Context(local_rank)
    stream = torch.cuda.current_stream(dev)
    ctx.set_stream(stream.cuda_stream)
    cfg = _lib.default_config()
    P = args.max_position
    R = 2 * P

    g = workload(args, rank)
    n_rows = len(g["tax_id"])
    n_in_tax = len(g["tax_ids"])

    # ---------------- device-resident buffers (value) ----------------
    cols = dict(
        tax_id=torch.from_numpy(g["tax_id"]).to(dev), n_alignments=torch.from_numpy(g["n_alignments"].view(np.int32)).to(dev),
        is_reverse=torch.from_numpy(g["is_reverse"]).to(dev), pos0=torch.from_numpy(g["pos0"]).to(dev),
        counts16=torch.from_numpy(g["counts16"].view(np.int32)).to(dev))
    outs = dict(
        n_fwd_ref=torch.empty(n_rows, dtype=torch.int32, device=dev), n_rev_ref=torch.empty(n_rows, dtype=torch.int32, device=dev),
        f_fwd=torch.empty(n_rows, dtype=torch.float32, device=dev), f_rev=torch.empty(n_rows, dtype=torch.float32, device=dev),
        z=torch.empty(n_rows, dtype=torch.int8, device=dev), y_sum_total=torch.empty(n_rows, dtype=torch.int64, device=dev),
        keep=torch.empty(n_rows, dtype=torch.uint8, device=dev), tax_id=torch.empty(n_in_tax, dtype=torch.int64, device=dev),
        n_alignments=torch.empty(n_in_tax, dtype=torch.int32, device=dev), first_row=torch.empty(n_in_tax, dtype=torch.int64, device=dev),
        k=torch.empty((n_in_tax, R), dtype=torch.int32, device=dev), N=torch.empty((n_in_tax, R), dtype=torch.int32, device=dev),
        noise=torch.empty((n_in_tax, 3), dtype=torch.float64, device=dev))
    res_dev = torch.empty(n_in_tax * FIT_RESULT_DTYPE.itemsize, dtype=torch.uint8, device=dev)
    med_dev = torch.empty((3, n_in_tax, R), dtype=torch.float32, device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    stats = {"counts_ms": [], "map_ms": [], "nuts_ms": [], "ppc_ms": [], "assemble_ms": [], "launches": 0, "leapfrogs": np.zeros(6)}

    def step_device(record):
        flush.zero_()  # evict L2 between steps
        n_fit = ctx.counts_reduce_device(cols, outs)
        t1 = ctx.timings()
        ctx.fit_batch_device(outs["tax_id"][:n_fit], outs["k"][:n_fit], outs["N"][:n_fit], res_dev, cfg,
                             median=med_dev[0], hpdi_lo=med_dev[1], hpdi_hi=med_dev[2], noise3=outs["noise"][:n_fit])
        t2 = ctx.timings()
        if record:
            stats["counts_ms"].append(t1["counts_ms"])
            for key in ("map_ms", "nuts_ms", "ppc_ms", "assemble_ms"):
                stats[key].append(t2[key])
            stats["launches"] += t1["n_launches"] + t2["n_launches"]
            stats["leapfrogs"] += np.array(t2["leapfrogs"], dtype=np.float64)
        return n_fit

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(step_fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        n = 0
        for _ in range(steps):
            n += step_fn(True)
        e1.record(stream)
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        cnt = torch.tensor([float(n)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        return float(ms.item()), float(cnt.item())

    for _ in range(args.warmup):
        step_device(False)
    sampler = ClockSampler(local_rank)
    sampler.start()
    dev_ms, dev_fits = timed(step_device, args.steps)

    # ---------------- host-buffer path (e2e) ----------------
    def pinned(a):
        view = a.view(np.int32) if a.dtype == np.uint32 else a
        t = torch.empty(view.shape, dtype=torch.from_numpy(view[:0].copy()).dtype, pin_memory=True)
        t.numpy()[...] = view
        return t.numpy().view(a.dtype)

    h = {key: pinned(g[key]) for key in ("tax_id", "n_alignments", "is_reverse", "pos0", "counts16")}
    e2e_bytes = {"h2d": 0, "d2h": 0}

    def pinned_empty(shape, dtype):
        nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        t = torch.empty(max(nbytes, 1), dtype=torch.uint8, pin_memory=True)
        return t.numpy()[:nbytes].view(dtype).reshape(shape)

    # result buffers of the host API, pinned and reused from step to step (Context.counts_reduce(out=...))
    host_out = dict(
        n_fwd_ref=pinned_empty(n_rows, np.uint32), n_rev_ref=pinned_empty(n_rows, np.uint32), f_fwd=pinned_empty(n_rows, np.float32),
        f_rev=pinned_empty(n_rows, np.float32), z=pinned_empty(n_rows, np.int8), y_sum_total=pinned_empty(n_rows, np.uint64),
        keep=pinned_empty(n_rows, np.uint8), tax_id=pinned_empty(n_in_tax, np.int64), n_alignments=pinned_empty(n_in_tax, np.uint32),
        first_row=pinned_empty(n_in_tax, np.int64), k=pinned_empty((n_in_tax, R), np.uint32), N=pinned_empty((n_in_tax, R), np.uint32),
        noise=pinned_empty((n_in_tax, 3), np.float64))

    def step_host(record):
        flush.zero_()
        r = ctx.counts_reduce(h["tax_id"], h["n_alignments"], h["is_reverse"], h["pos0"], h["counts16"],
                              max_position=P, want_noise=True, out=host_out)
        t1 = ctx.timings()
        out = ctx.fit_batch(r["tax_id"], r["k"], r["N"], cfg, noise3=r["noise"])
        t2 = ctx.timings()
        if record:
            stats["launches"] += t1["n_launches"] + t2["n_launches"]
            stats["chain_leapfrogs"] = out["result"]["run"]["n_leapfrog"]
            n_fit = r["n_tax"]
            e2e_bytes["h2d"] = sum(h[key].nbytes for key in h) + n_fit * (8 + 2 * R * 4 + 24)
            e2e_bytes["d2h"] = (n_rows * (4 + 4 + 4 + 4 + 1 + 8 + 1) + n_fit * (8 + 4 + 8 + 2 * R * 4 + 24)
                                + out["result"].nbytes + 3 * out["median"].nbytes)
        return r["n_tax"]

    step_host(False)
    e2e_ms, e2e_fits = timed(step_host, args.steps)
    clocks = sampler.stop()

    # ---------------- the north star's reduced unit: no forward-only / reverse-only refits ----------------
    cfg_reduced = cfg.copy(do_fwd_rev=0)

    def step_reduced(record):
        flush.zero_()
        n_fit = ctx.counts_reduce_device(cols, outs)
        ctx.fit_batch_device(outs["tax_id"][:n_fit], outs["k"][:n_fit], outs["N"][:n_fit], res_dev, cfg_reduced,
                             median=med_dev[0], hpdi_lo=med_dev[1], hpdi_hi=med_dev[2], noise3=outs["noise"][:n_fit])
        if record:
            stats["launches"] += 3 + ctx.timings()["n_launches"]
        return n_fit

    step_reduced(False)
    red_ms, red_fits = timed(step_reduced, 1)

    # ---------------- longest chain (the tail of a step is one sequential Markov chain) ----------------
    chains = stats["chain_leapfrogs"].astype(np.float64)
    chain_max = torch.tensor([chains.max()], dtype=torch.float64, device=dev)
    chain_sum = torch.tensor([chains.sum(), float(chains.size)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(chain_max, op=dist.ReduceOp.MAX)
        dist.all_reduce(chain_sum, op=dist.ReduceOp.SUM)
    chain_stats = {"mean_leapfrogs_per_chain": float(chain_sum[0].item() / chain_sum[1].item()),
                   "max_leapfrogs_of_one_chain": float(chain_max.item()),
                   "note": "a step cannot end before its longest chain does: a chain is sequential (one warp; measured 1.9 us per leapfrog "
                           "with the GPU to itself, ~3.4 us averaged over a full batch); rarely (one chain in the 480 000 of the eight bench "
                           "shards: 464 068 leapfrogs, in rank 1's) a chain adapts to a collapsed step size and runs ~50x the mean "
                           "(DESIGN.md section 7, profiles/r01_shard_probe.log)"}

    # ---------------- rooflines, CPU baseline (rank 0 only) ----------------
    if rank == 0:
        fp64_peak = ctx.fp64_peak_tflops()
        nuts_ms = float(np.sum(stats["nuts_ms"]))
        flops = model_flops(list(stats["leapfrogs"]), P)
        achieved = flops / (nuts_ms * 1e-3) / 1e12 if nuts_ms > 0 else 0.0
        roofline = {
            "bound": "fp64", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s", "frac": achieved / fp64_peak,
            "traffic": 948.0, "traffic_source": "ncu --set full capture of the PMD/all launch (profiles/r01_nuts_ncu.md): bytes of DRAM traffic per launch, i.e. none",
            "kernel": "nuts_kernel<PMD|null, 32|16 lanes> (4 launches per step)",
            "flop_model": "SURVEY.md 8d nominal: PMD 300*n_obs+55, null 190*n_obs+165 flop per gradient evaluation",
            "gradient_evaluations_per_step": float(stats["leapfrogs"].sum() / max(1, args.steps)),
            "gradient_evaluations_per_s": float(stats["leapfrogs"].sum() / (nuts_ms * 1e-3)) if nuts_ms > 0 else 0.0,
            "peak_source": "FP64 FMA peak measured live on this GPU (mdg_measure_fp64_peak); not in MEASURED_PEAKS.json",
            "ncu": {"source": "profiles/r01_nuts_ncu.md (ncu --set full, PMD/all launch; static, not re-measured by this run)",
                    "warp_instructions_per_evaluation": 1096, "fp64_instruction_share": 0.414,
                    "executed_fp64_flop_per_evaluation": 20850,
                    "executed_fp64_tflops_at_this_rate": 20850 * float(stats["leapfrogs"].sum() / (nuts_ms * 1e-3)) / 1e12 if nuts_ms > 0 else 0.0,
                    "fp64_pipe_busy_full_size": 0.48, "ipc_per_scheduler_full_size": 0.58,
                    "pipe_busy_pct_in_capture": {"fp64": 33.3, "alu": 14.8, "xu_sfu": 7.4, "fma_fp32": 5.5, "lsu": 25.5, "issue_slots": 46.5}},
            "note": "HBM traffic of the fit kernels is negligible (240 B in, ~1 KB out per TaxID): compute bound, no tensor cores",
        }
        roofline_counts = None if args.no_counts_stress else counts_stress(ctx, torch, dev)
        cpu = None
        if not args.no_cpu_baseline:
            cpu, _ = cpu_baseline(g, args)
        per_step = lambda key: float(np.mean(stats[key])) if stats[key] else 0.0  # noqa: E731
        line = {
            "metric": METRIC, "value": dev_fits / (dev_ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, world, n_in_tax, n_rows),
            "e2e": {"value": e2e_fits / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(e2e_bytes["h2d"]),
                    "d2h_bytes_per_step": int(e2e_bytes["d2h"]), "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": int(stats["launches"]),
            "clocks": clocks,
            "roofline": roofline,
            "roofline_counts": roofline_counts,
            "chains": chain_stats,
            "reduced_unit": {"value": red_fits / (red_ms * 1e-3), "unit": UNIT, "ms_per_step": red_ms, "steps": 1,
                             "what": "counts + MAP + PMD/null NUTS on all positions + WAIC + predictive D_max, WITHOUT the forward-only / "
                                     "reverse-only refits of fits.py:298-356 (the north star's reduced unit; `value` above is the full fit)"},
            "cpu_baseline": cpu,
            "kernel_ms_per_step": {"counts": per_step("counts_ms"), "map": per_step("map_ms"), "nuts": per_step("nuts_ms"),
                                   "ppc": per_step("ppc_ms"), "assemble": per_step("assemble_ms")},
        }
        print(json.dumps(line), flush=True)
    barrier()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def _json_only_stdout():
    """stdout carries ONE JSON line: library chatter written to file descriptor 1 (NCCL prints its version
    there under torchrun) is sent to stderr; print() keeps the real stdout."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = real


if __name__ == "__main__":
    _json_only_stdout()
    main()
