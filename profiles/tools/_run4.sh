profiles/tools/ab.sh default
BEDKIT_MAP_KERNEL=row profiles/tools/ab.sh default | sed 's/^default/row/'
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_random_differential.py tests/test_bedmap_more_ops.py -m gpu -x -q -k "not closest and not ec_validation and not parser" > gpurun_out/mg_tests.log 2>&1; tail -5 gpurun_out/mg_tests.log
timeout 300 python -m pytest tests/test_gpu_scale.py -m gpu -x -q -k "config2 or config5" > gpurun_out/mg_scale.log 2>&1; tail -3 gpurun_out/mg_scale.log
