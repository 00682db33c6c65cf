#!/bin/bash
# Round capture on one B200: tests, smoke, bench (both arms), ncu launch list, ncu --set full of the sweep kernel.
# usage: bash profiles/tools/capture.sh <tag>      outputs under gpurun_out/<tag>_*
tag=${1:-r1}
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest.log
python __graft_entry__.py smoke > $out/${tag}_smoke.log 2>&1; echo "smoke rc=$?" >> $out/${tag}_smoke.log
python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err
python bench.py --impl reference --steps 1 --warmup 0 > $out/${tag}_bench_ref.json 2> $out/${tag}_bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $out/${tag}_ncu_launch.log 2>&1
python profiles/tools/sweep_profile.py 100 4950 > $out/${tag}_sp_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:sweep3_kernel -o $out/${tag}_sweep_full -f \
    python profiles/tools/sweep_profile.py 100 4950 > $out/${tag}_ncu_sweep.log 2>&1
tail -3 $out/${tag}_pytest.log; cat $out/${tag}_smoke.log | tail -2; cat $out/${tag}_bench.json | cut -c1-600
