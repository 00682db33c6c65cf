#!/bin/bash
# Round-2 ncu captures on one B200 (each command first runs plain, then under ncu): launch list of the bench, ncu --set full of the
# sweep kernel (n = 100 and n = 200, full cycles), the TMA prep kernel and the TMA rotation kernel.  outputs: gpurun_out/<tag>_*
tag=${1:-r2cap}; out=gpurun_out; mkdir -p $out
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-north-star"
$B > $out/${tag}_plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/${tag}_launches_n100.csv $B > $out/${tag}_ncu_launch.log 2>&1
S1="python profiles/tools/sweep_profile.py 100 4950"
$S1 > $out/${tag}_plain_s100.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:sweep3_kernel -c 1 -f -o $out/${tag}_sweep3_n100 $S1 > $out/${tag}_ncu_s100.log 2>&1
S2="python profiles/tools/sweep_profile.py 200 19900"
$S2 > $out/${tag}_plain_s200.log 2>&1 && \
ncu --section MemoryWorkloadAnalysis --section SpeedOfLight --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_bytes.sum --clock-control none -k regex:sweep3_kernel -c 1 -f -o $out/${tag}_sweep3_n200 $S2 > $out/${tag}_ncu_s200.log 2>&1
P1="python profiles/tools/prep_profile.py 100 2"
$P1 > $out/${tag}_plain_prep.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:prep_tma_kernel -s 1 -c 1 -f -o $out/${tag}_prep_n100 $P1 > $out/${tag}_ncu_prep.log 2>&1
R1="python profiles/tools/rotate_profile.py 100 1"
$R1 > $out/${tag}_plain_rot.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:rotate_tma_kernel -s 4 -c 1 -f -o $out/${tag}_rotate_n100 $R1 > $out/${tag}_ncu_rot.log 2>&1
for f in s100 s200 prep rot; do tail -n 2 $out/${tag}_plain_$f.log; done; ls -la $out/${tag}_*.ncu-rep
