#!/usr/bin/env bash
# One `ncu --set full` capture of the launches of ONE bench step, reduced on the GPU box to what fits gpurun_out/:
# the raw metrics page as CSV and per-source-line instruction profiles of the hot kernels (the .ncu-rep itself is
# hundreds of MB and stays in /tmp).  usage: ncu_step.sh <tag>
set -uo pipefail
TAG="$1"
OUT=gpurun_out
mkdir -p "$OUT"
FLAGS="--no-tool-e2e --no-cpu-baseline --no-e2e"
python bench.py --steps 1 --warmup 2 $FLAGS > "$OUT/${TAG}_plain.json" 2> "$OUT/${TAG}_plain.err" || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name regex:^k_ -c 2000 --csv --log-file "$OUT/${TAG}_launches.csv" \
    python bench.py --steps 1 --warmup 2 $FLAGS > /tmp/ncu_l.log 2>&1
# the library's launches in order (the synthetic generator's come first); a load begins with k_efflen, a step has two loads:
# the third step begins at the fifth k_efflen
read SKIP COUNT < <(python - "$OUT/${TAG}_launches.csv" <<'PY'
import csv, sys
names = []
hdr = None
for r in csv.reader(open(sys.argv[1])):
    if len(r) > 5 and r[0] == "ID":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        names.append(r[hdr.index("Kernel Name")])
eff = [i for i, n in enumerate(names) if "k_efflen" in n]
print(eff[4], len(names) - eff[4])
PY
)
echo "launch-skip $SKIP launch-count $COUNT"
ncu --set full --clock-control none --import-source on --kernel-name regex:^k_ --launch-skip $SKIP --launch-count $COUNT -f -o /tmp/${TAG} \
    python bench.py --steps 1 --warmup 2 $FLAGS > /tmp/ncu_f.log 2>&1
tail -2 /tmp/ncu_f.log
ncu -i /tmp/${TAG}.ncu-rep --page raw --csv > "$OUT/${TAG}_ncu_raw_full.csv" 2>/dev/null
python - "$OUT/${TAG}_ncu_raw_full.csv" "$TAG" <<'PY'
import csv, sys, subprocess
rows = list(csv.reader(open(sys.argv[1])))
h = rows[0]
ki = h.index("Kernel Name")
di = h.index("gpu__time_duration.sum")
best = {}
for idx, r in enumerate(rows[2:]):
    name = r[ki]
    for want in ("k_parse", "k_map_group", "k_emit<", "k_count_rows", "k_emit_len"):
        if want in name or (want == "k_emit<" and name.startswith("k_emit") and "len" not in name):
            d = float(r[di].replace(",", ""))
            if want not in best or d > best[want][1]:
                best[want] = (idx, d, name)
for want, (idx, d, name) in best.items():
    print(want, idx, d, name[:60])
    key = {"k_parse": "k_parse_wsILi5ELb1", "k_map_group": "k_map_groupILi0ELj3E", "k_emit<": "k_emitINS_9BedmapRowILi0E", "k_count_rows": "k_count_rows",
           "k_emit_len": "k_emit_lenINS_9BedmapRowILi0E"}[want]
    out = open("gpurun_out/%s_%s_by_line.txt" % (sys.argv[2], want.strip("<")), "w")
    subprocess.run([sys.executable, "profiles/tools/sass_by_line.py", "/tmp/%s.ncu-rep" % sys.argv[2], str(idx), "bedops_b200/lib/libbedkit.so", key, "400"], stdout=out, stderr=subprocess.STDOUT)
PY
ls -la "$OUT" | tail -12
