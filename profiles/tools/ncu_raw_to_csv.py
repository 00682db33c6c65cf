"""ncu -i <report>.ncu-rep --page raw --csv  ->  a (metric, unit, value) table of the metrics that matter for the rooflines.
    python profiles/tools/ncu_raw_to_csv.py gpurun_out/x.ncu-rep profiles/x_ncu_raw.csv"""
import csv
import re
import subprocess
import sys

KEEP = re.compile(r'^(Kernel Name|Block Size|Grid Size|dram__bytes_(read|write)\.sum|dram__throughput|gpu__dram_throughput|gpu__time_duration|'
                  r'lts__t_bytes\.sum$|lts__t_sector_hit_rate|lts__throughput|l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum$|'
                  r'launch__(registers_per_thread|occupancy_limit|shared_mem_per_block|grid_size|block_size|waves)|'
                  r'sm__throughput|sm__issue_active|sm__inst_executed\.sum$|sm__warps_active\.avg\.pct|smsp__cycles_active\.avg$|'
                  r'sm__pipe_fp64_cycles_active\.avg\.pct|sm__inst_executed_pipe_tensor_subpipe_dmma\.avg\.pct|sm__pipe_tensor_subpipe_dmma|'
                  r'sm__ops_path_tensor_src_fp64\.avg|smsp__average_warps?_issue_stalled_[a-z_]+_per_issue_active|smsp__warp_issue_stalled|'
                  r'smsp__pcsamp_warps_issue_stalled_[a-z_]+$)')
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
with open(sys.argv[2], 'w', newline='') as f:
    w = csv.writer(f)
    w.writerow(['metric', 'unit', 'value'] + (['value_launch_%d' % k for k in range(2, len(rows) - 1)] if len(rows) > 3 else []))
    for c, (h, u) in enumerate(zip(hdr, units)):
        if KEEP.match(h):
            w.writerow([h, u] + [r[c] for r in rows[2:]])
print('wrote', sys.argv[2])
