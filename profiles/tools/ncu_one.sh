#!/usr/bin/env bash
# ncu --set full of ONE launch of one kernel during the default bench; prints the key lines of the details page and keeps the raw CSV.
# usage: ncu_one.sh <kernel regex> <tag> [launch-skip]
set -uo pipefail
K="$1"; TAG="$2"; SKIP="${3:-3}"
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on --kernel-name "regex:$K" --launch-skip $SKIP --launch-count 1 -f -o /tmp/$TAG \
    python bench.py --steps 1 --warmup 2 --no-tool-e2e --no-cpu-baseline --no-e2e > /tmp/ncu_$TAG.log 2>&1
tail -2 /tmp/ncu_$TAG.log
ncu -i /tmp/$TAG.ncu-rep --page details > gpurun_out/${TAG}_details.txt 2>/dev/null
ncu -i /tmp/$TAG.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv 2>/dev/null
grep -E "Duration|Executed Ipc Active|Issue Slots Busy|Registers Per|Achieved Occupancy|Theoretical Occupancy|L1/TEX Hit|Mem Busy|Max Bandwidth|Mem Pipes Busy|DRAM Throughput|Warp Cycles Per Issued|No Eligible|Shared Memory Configuration|bank conflict|Est. Speedup" gpurun_out/${TAG}_details.txt | head -40
python - gpurun_out/${TAG}_raw.csv <<'PY'
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
h, r = rows[0], rows[2]
for i, n in enumerate(h):
    if any(k in n for k in ("warp_issue_stalled", "pipe_lsu", "pipe_alu", "pipe_fp64", "pipe_fma", "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared", "lsu_mem_shared_op")):
        if "pct" in n or n.endswith(".sum") or "ratio" in n:
            print(n, r[i])
PY
# per-source-line instruction profile when a mangled-name key is given as 4th argument
if [ -n "${4:-}" ]; then
  python profiles/tools/sass_by_line.py /tmp/$TAG.ncu-rep 0 bedops_b200/lib/libbedkit.so "$4" 400 > gpurun_out/${TAG}_by_line.txt 2>&1
  head -45 gpurun_out/${TAG}_by_line.txt
fi
