#!/bin/bash
# quick regression on one B200: GPU parity tests + a short bench (no CPU baseline).   usage: quick.sh <tag> [n]
tag=${1:-q}; n=${2:-100}; out=gpurun_out; mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest.log
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --orbitals $n > $out/${tag}_bench.json 2> $out/${tag}_bench.err
tail -4 $out/${tag}_pytest.log
python - <<PY
import json
d=json.loads([l for l in open('$out/${tag}_bench.json') if l.startswith('{')][-1])
print('ms/step',d['ms_per_step'],'e2e',d['e2e']['ms_per_step'],'sweep',d['phases']['jacobi_cycle_kernel_first_cycle_ms'],'frac',d['roofline']['frac'],'acc',d['rotations_accepted'],'cost',d['cost_after_cycle'])
print(d['phases']['sweep_phase_us_per_pair'])
print(d['phases']['linesearch_cycles_per_pair_cta0'])
print({k:v for k,v in d['phases'].items() if k.endswith('_ms')})
PY
tail -3 $out/${tag}_bench.err
