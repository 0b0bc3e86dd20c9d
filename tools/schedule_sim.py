"""Development tool (CPU only): chain lengths from the oracle and a processor-sharing model of the NUTS schedule.

    python tools/schedule_sim.py lengths 3000 /tmp/chainlen.npz    # n_leapfrog of the 6 runs of 3 000 cfg2 TaxIDs (oracle, all host cores)
    python tools/schedule_sim.py orders /tmp/chainlen.npz           # queue orders: random, by coverage, oracle longest-first
    python tools/schedule_sim.py granularity /tmp/chainlen.npz      # slots released per chain / per CTA of 4, 8, 16 chains

Model: 9 472 chain slots (148 SMs x 4 CTAs x 16 chains); PMD queue (all-position runs, then forward / reverse runs), then
the null queue; a live chain advances at base speed (leapfrogs per ms under full load: 83 PMD all, 111 PMD half, 128 / 161
null) times min(3, slots / live chains) — a chain alone on the GPU runs about three times faster than under full load.
What it was used for (profiles/r02_chain_timeline.md): with aggregate throughput conserved the makespan is set by the
longest chain's own length, so neither the queue order nor the release granularity moves it by more than 2-3 %."""
import os
import sys
from concurrent.futures import ThreadPoolExecutor

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

SLOTS, CAP = 9472, 3.0
SPEED = {0: 83.0, 2: 111.0, 4: 111.0, 1: 128.0, 3: 161.0, 5: 161.0}


def lengths(n, path):
    from metadamage_b200 import synthetic as syn
    from oracle import oracle as orc

    tid, k, N, _ = syn.dense_fit_batch(n)
    cfg = orc.default_config()

    def one(i):
        return [orc.nuts_run(k[i], N[i], int(tid[i]), rk, cfg)["n_grad"] for rk in range(6)]

    with ThreadPoolExecutor(os.cpu_count()) as ex:
        L = np.array(list(ex.map(one, range(n))), dtype=np.int64)
    np.savez(path, L=L, k=k, N=N, tid=tid)
    for r in range(6):
        print("run", r, "leapfrogs: median, 90 %, 99 %, 99.9 %, max =", np.percentile(L[:, r], [50, 90, 99, 99.9, 100]).astype(int))


def simulate(L, order, group_size=1, n=10000, seed=0, dt=1.0):
    rng = np.random.default_rng(seed)
    idx = rng.integers(0, len(L), n)
    Ls = L[idx].astype(float)
    o = order(Ls, rng)
    pmd = [(Ls[i, 0], SPEED[0]) for i in o] + [(Ls[i, r], SPEED[r]) for i in o for r in (2, 4)]
    nul = [(Ls[i, 1], SPEED[1]) for i in o] + [(Ls[i, r], SPEED[r]) for i in o for r in (3, 5)]
    n_cta = SLOTS // group_size
    rem, spd, busy = np.zeros(SLOTS), np.zeros(SLOTS), np.zeros(SLOTS, bool)
    kind = np.ones(n_cta, int)  # 1: PMD CTA, 2: null CTA
    cta_of = np.arange(SLOTS) // group_size
    ip = inn = 0
    t = 0.0
    while True:
        for s in np.flatnonzero(~busy):
            if kind[cta_of[s]] == 1 and ip < len(pmd):
                rem[s], spd[s] = pmd[ip]; ip += 1; busy[s] = True
            elif kind[cta_of[s]] == 2 and inn < len(nul):
                rem[s], spd[s] = nul[inn]; inn += 1; busy[s] = True
        if ip >= len(pmd):
            idle_cta = (kind == 1) & ~busy.reshape(n_cta, group_size).any(1)
            if idle_cta.any():
                kind[idle_cta] = 2
                continue
        live = int(busy.sum())
        if live == 0:
            return t
        rem[busy] -= spd[busy] * min(CAP, SLOTS / live) * dt
        busy &= rem > 0
        t += dt


if __name__ == "__main__":
    what = sys.argv[1]
    if what == "lengths":
        lengths(int(sys.argv[2]), sys.argv[3])
    else:
        d = np.load(sys.argv[2])
        L, N = d["L"], d["N"].astype(float)
        keep = L.max(1) < 60000  # the bench batch has no chain beyond 48 k
        L, Nsum = L[keep], N[keep].sum(1)
        if what == "orders":
            orders = {"random": lambda Ls, r: r.permutation(len(Ls)), "oracle longest-first (PMD all)": lambda Ls, r: np.argsort(-Ls[:, 0])}
            for name, fn in orders.items():
                print(f"{name:32s} makespan {np.mean([simulate(L, fn, seed=s) for s in range(2)]):.0f} ms")
        else:
            for g in (16, 8, 4, 1):
                print(f"chains per CTA {g:2d}: makespan {np.mean([simulate(L, lambda Ls, r: r.permutation(len(Ls)), g, seed=s) for s in range(2)]):.0f} ms")
