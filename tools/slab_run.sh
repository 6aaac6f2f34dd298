#!/bin/bash
# usage: bash tools/slab_run.sh <ngpus> <tag>   (on the GPU box): bit parity vs one GPU at 128^3 / 256^3, timing at 512^3
N=$1; TAG=$2; O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
{
for n in 128 256; do $TR --master-port 2961$((n/128)) tools/bench_configs.py --only c5slab --n3 $n --check 2>/dev/null | grep '^{'; done
for tr in push peer nccl; do $TR --master-port 29620 tools/bench_configs.py --only c5slab --n3 512 --transport $tr 2>/dev/null | grep '^{'; done
PDEOPT_SLAB_TIMING=1 $TR --master-port 29621 tools/bench_configs.py --only c5slab --n3 512 2>/dev/null | grep '^{'
} > $O/${TAG}_slab_n$N.log 2>&1
cat $O/${TAG}_slab_n$N.log | cut -c1-420
