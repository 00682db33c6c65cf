#!/bin/bash
# N-GPU bench (no parity, no north star, no e2e) for several library builds: r2_multi_ab.sh <N> lib1 lib2 ...
N=$1; shift
out=gpurun_out; mkdir -p $out
for lib in "$@"; do
  tag=$(basename $lib .so)
  QIO_B200_LIB=$PWD/$lib timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29644 \
    bench.py --gpus $N --steps 3 --warmup 3 --no-north-star --no-parity --no-e2e > $out/ab_${tag}_$N.json 2> $out/ab_${tag}_$N.err
  python - <<PY
import json
p=json.loads(open('$out/ab_${tag}_$N.json').read().splitlines()[-1])
c=p['phases'].get('sweep_chain_us') or {}
print('$tag N=$N ms_per_step %.2f sweep %.2f' % (p['ms_per_step'], p['phases']['jacobi_cycle_kernel_ms']), {k: round(v,2) for k,v in c.items() if k not in ('raw_us',)})
PY
done
