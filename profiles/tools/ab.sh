#!/usr/bin/env bash
# A/B of library variants on the default bench (device-resident leg only): ab.sh <name|default> ...
# prints ms/step and the per-kernel split of each; lines kept in gpurun_out/ab_<name>.json
set -uo pipefail
mkdir -p gpurun_out
for v in "$@"; do
  if [ "$v" = default ]; then unset BEDKIT_LIB; else export BEDKIT_LIB="$PWD/bedops_b200/lib/variants/$v.so"; fi
  python bench.py --steps 5 --warmup 3 --no-tool-e2e --no-cpu-baseline --no-e2e > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err || { echo "$v FAILED"; tail -3 gpurun_out/ab_$v.err; continue; }
  python - "$v" gpurun_out/ab_$v.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
k = d["roofline"]["kernel_ms_per_step"]
print(sys.argv[1], "step %.3f" % d["ms_per_step"], " ".join("%s %.3f" % (n.replace("k_", ""), v) for n, v in k.items()))
PY
done
