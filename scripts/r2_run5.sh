#!/bin/bash
# round 2, GPU call 5: block kernels with the async pair ring, range-split transpose; new bench.py
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2e_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/r2e_tests.log
tail -5 gpurun_out/r2e_tests.log
for cfg in "auto:X=1" "generic:GCNN_BLOCKS=0"; do
  name=${cfg%%:*}; env=${cfg#*:}
  env $env python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-extra-configs > gpurun_out/r2e_bench_$name.json 2> gpurun_out/r2e_bench_$name.err
done
python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extra-configs --graphs-per-gpu 128 > gpurun_out/r2e_bench_auto_g128.json 2> gpurun_out/r2e_bench_auto_g128.err
python bench.py --steps 30 --warmup 5 > gpurun_out/r2e_bench_full.json 2> gpurun_out/r2e_bench_full.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r2e_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra-configs > gpurun_out/r2e_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'edge_block|transpose_blocks' -s 12 -c 8 -o gpurun_out/r2e_blocks -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra-configs > gpurun_out/r2e_ncu_full.log 2>&1
python scripts/show_bench.py gpurun_out/r2e_bench_auto.json gpurun_out/r2e_bench_generic.json gpurun_out/r2e_bench_auto_g128.json | grep -E "==|edge_|csr_|sum of"
tail -c 3000 gpurun_out/r2e_bench_full.json; tail -5 gpurun_out/r2e_bench_full.err
