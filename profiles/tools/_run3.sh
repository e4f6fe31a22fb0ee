export BEDKIT_MAP_KERNEL=group
profiles/tools/ab.sh default c2 c4 c8
BEDKIT_CONFIGS=5 python profiles/tools/bench_configs.py 2>&1 | tail -1 | sed 's/^/group: /'
unset BEDKIT_MAP_KERNEL
BEDKIT_CONFIGS=5 python profiles/tools/bench_configs.py 2>&1 | tail -1 | sed 's/^/row: /'
