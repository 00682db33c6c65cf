#!/bin/bash
# quick N-GPU bench (no parity check, no north star): usage r2_multi_quick.sh <tag> <N>
tag=${1:-mq}; N=${2:-2}
out=gpurun_out; mkdir -p $out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29643 \
    bench.py --gpus $N --steps 3 --warmup 3 --no-north-star --no-parity --no-e2e > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?" >> $out/${tag}_bench.err
python - <<PY
import json
p=json.loads(open('$out/${tag}_bench.json').read().splitlines()[-1])
print('ms_per_step',p['ms_per_step'],'sweep',p['phases']['jacobi_cycle_kernel_ms'])
print(p['phases']['sweep_phase_us_per_pair'])
print(p['phases'].get('sweep_chain_us'))
PY
tail -3 $out/${tag}_bench.err
