#!/bin/bash
# quick iteration: sweep-related GPU tests + n=100 bench without the north-star leg / cpu baseline
tag=${1:-q}
out=gpurun_out
mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_large.py tests/test_gpu_multi.py -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest.log
timeout 300 python bench.py --no-north-star --no-cpu-baseline > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?" >> $out/${tag}_bench.err
tail -4 $out/${tag}_pytest.log; python - <<PY
import json
p=json.loads(open('$out/${tag}_bench.json').read().splitlines()[-1])
print('ms_per_step',p['ms_per_step'],'e2e',p['e2e']['ms_per_step'],'sweep',p['phases']['jacobi_cycle_kernel_ms'])
print(p['phases']['sweep_phase_us_per_pair'])
print(p['phases']['linesearch_cycles_per_pair_cta0'])
print(p['decision_audit'])
PY
tail -3 $out/${tag}_bench.err
