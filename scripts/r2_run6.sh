#!/bin/bash
# round 2, GPU call 6: block kernels (padded ring, poison patching), dgrad accumulator split accuracy
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2f_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/r2f_tests.log
tail -5 gpurun_out/r2f_tests.log
python tests/grad_error_report.py --big > gpurun_out/r2f_grad_err_split.log 2>&1
GCNN_LIB=gcnn_cut_selector_b200/build/nosplit/libgcnn_b200.so python tests/grad_error_report.py --big > gpurun_out/r2f_grad_err_nosplit.log 2>&1
tail -1 gpurun_out/r2f_grad_err_split.log; tail -1 gpurun_out/r2f_grad_err_nosplit.log
for cfg in "auto:X=1" "generic:GCNN_BLOCKS=0" "nosplit:GCNN_LIB=gcnn_cut_selector_b200/build/nosplit/libgcnn_b200.so"; do
  name=${cfg%%:*}; env=${cfg#*:}
  env $env python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-extra-configs > gpurun_out/r2f_bench_$name.json 2> gpurun_out/r2f_bench_$name.err
done
python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extra-configs --graphs-per-gpu 128 > gpurun_out/r2f_bench_auto_g128.json 2> gpurun_out/r2f_bench_auto_g128.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r2f_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra-configs > gpurun_out/r2f_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'edge_block|transpose_blocks' -s 12 -c 8 -o gpurun_out/r2f_blocks -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra-configs > gpurun_out/r2f_ncu_full.log 2>&1
python scripts/show_bench.py gpurun_out/r2f_bench_auto.json gpurun_out/r2f_bench_generic.json gpurun_out/r2f_bench_nosplit.json gpurun_out/r2f_bench_auto_g128.json | grep -E "==|edge_|csr_|chain|sum of"
