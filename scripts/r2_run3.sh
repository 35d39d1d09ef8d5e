#!/bin/bash
# round 2, GPU call 3: group-per-row block kernels: parity + A/B; transpose with prefetch; wgrad accumulation experiment
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2c_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/r2c_tests.log
tail -5 gpurun_out/r2c_tests.log
for cfg in "groups:GCNN_BLOCK_GROUPS=1" "rows:GCNN_BLOCK_GROUPS=0" "auto:X=1" "generic:GCNN_BLOCKS=0"; do
  name=${cfg%%:*}; env=${cfg#*:}
  env $env python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2c_bench_$name.json 2> gpurun_out/r2c_bench_$name.err
done
GCNN_CHAIN_PARTS=1024 python tests/grad_error_report.py --big > gpurun_out/r2c_grad_err_parts1024.log 2>&1
python tests/grad_error_report.py --big > gpurun_out/r2c_grad_err_parts148.log 2>&1
GCNN_LIB=gcnn_cut_selector_b200/build/act2/libgcnn_b200.so python tests/grad_error_report.py --big > gpurun_out/r2c_grad_err_act2.log 2>&1
tail -1 gpurun_out/r2c_grad_err_parts1024.log; tail -1 gpurun_out/r2c_grad_err_parts148.log; tail -1 gpurun_out/r2c_grad_err_act2.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r2c_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_ncu_list.log 2>&1
python scripts/show_bench.py gpurun_out/r2c_bench_*.json | grep -E "==|edge_|csr_|sum of"
