#!/bin/bash
# Round-2 measurement recipe (run on the GPU box through gpurun): the bench line, the launch list of the same
# command under ncu (per-launch times: cold-cache and serialised, shares only), and ncu --set full captures of
# the dominant kernel (NUTS, at full occupancy) and of the counts kernel.
set -x
cd "$(dirname "$0")/.."
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-seam > gpurun_out/r02_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:nuts_group -c 2 -o gpurun_out/r02_nuts -f \
    python tools/profile_fit.py 32768 40 40 > gpurun_out/r02_ncu_nuts.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:counts_stream_kernel -c 1 -s 2 -o gpurun_out/r02_counts -f \
    python tools/profile_target.py 256 5 5 > gpurun_out/r02_ncu_counts.log 2>&1
ls -la gpurun_out/*.ncu-rep
