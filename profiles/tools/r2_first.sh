#!/bin/bash
# Round 2, first GPU call: GPU tests, smoke, bench with the north-star leg.  outputs under gpurun_out/<tag>_*
tag=${1:-r2a}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,memory.total --format=csv > $out/${tag}_gpu.txt
free -g >> $out/${tag}_gpu.txt; nproc >> $out/${tag}_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest.log
timeout 300 python __graft_entry__.py smoke > $out/${tag}_smoke.log 2>&1; echo "smoke rc=$?" >> $out/${tag}_smoke.log
timeout 600 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?" >> $out/${tag}_bench.err
tail -5 $out/${tag}_pytest.log; tail -2 $out/${tag}_smoke.log; cut -c1-400 $out/${tag}_bench.json; tail -5 $out/${tag}_bench.err
