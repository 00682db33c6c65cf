#!/bin/bash
# 8-GPU legs: multi-rank parity check, bench n=200 (configs[4]) and n=100 (configs[3]).   usage: multi8.sh <tag> [ngpu]
tag=${1:-r1}; N=${2:-8}; out=gpurun_out; mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29701 tests/multi_gpu_check.py > $out/${tag}_check_${N}gpu.log 2>&1; echo "check rc=$?" >> $out/${tag}_check_${N}gpu.log
timeout 600 $TR --master-port 29702 bench.py --gpus $N --orbitals 200 --steps 2 --warmup 3 > $out/${tag}_bench_n200_${N}gpu.json 2> $out/${tag}_bench_n200_${N}gpu.err
timeout 300 $TR --master-port 29703 bench.py --gpus $N --orbitals 100 --steps 3 --warmup 3 > $out/${tag}_bench_n100_${N}gpu.json 2> $out/${tag}_bench_n100_${N}gpu.err
tail -2 $out/${tag}_check_${N}gpu.log; cut -c1-300 $out/${tag}_bench_n200_${N}gpu.json; tail -3 $out/${tag}_bench_n200_${N}gpu.err
