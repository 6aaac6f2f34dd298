#!/bin/bash
# One GPU call of the round: tests, bench (both arms), launch list, full ncu capture of the headline kernel.
# usage (on the GPU box, from the repo root): bash tools/gpu_round.sh <tag>
TAG=${1:-r2}
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/${TAG}_gputest.log 2>&1; echo "pytest exit $?" >> $O/${TAG}_gputest.log
tail -3 $O/${TAG}_gputest.log
timeout 600 python bench.py --steps 50 --warmup 3 > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench_n1.err; echo "bench exit $?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_reference.json 2>> $O/${TAG}_bench_n1.err
for k in 1 4 16 64; do timeout 120 python tools/quick_bench.py 4096 $k; done > $O/${TAG}_ksweep.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary > $O/${TAG}_ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sifs128r_kernel -s 4 -c 1 -f -o $O/${TAG}_sifs128r_full \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary > $O/${TAG}_ncu_full.log 2>&1
ls -la $O
