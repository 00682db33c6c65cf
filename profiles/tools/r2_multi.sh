#!/bin/bash
# Round 2 multi-GPU call: torchrun parity check (sharded vs single GPU, epoch wrap, mailbox stress) + bench at N GPUs.
# usage: bash profiles/tools/r2_multi.sh <tag> <N>
tag=${1:-r2m}; N=${2:-2}
out=gpurun_out
mkdir -p $out
nvidia-smi topo -m > $out/${tag}_topo.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29641 \
    tests/multi_gpu_check.py > $out/${tag}_check.log 2>&1; echo "check rc=$?" >> $out/${tag}_check.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29642 \
    bench.py --gpus $N --steps 3 --warmup 3 > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?" >> $out/${tag}_bench.err
grep -v "NCCL INFO" $out/${tag}_check.log | tail -8; cut -c1-300 $out/${tag}_bench.json; grep -v "NCCL INFO" $out/${tag}_bench.err | tail -8
