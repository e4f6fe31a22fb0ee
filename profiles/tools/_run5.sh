timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "both_window_kernels" > gpurun_out/mg_t2.log 2>&1; tail -4 gpurun_out/mg_t2.log
profiles/tools/ncu_one.sh k_map_group mg2 2 k_map_groupILi0ELj3E 2>&1 | grep -v "^l1tex__t_\|^sm__sass\|^smsp__sass\|dshared\|\.min\.\|\.max\.\|\.sum\.pct"
