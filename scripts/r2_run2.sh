#!/bin/bash
# round 2, GPU call 2: rest of the parity suite + ncu launch list and full capture of the block kernels
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2b_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/r2b_tests.log
tail -15 gpurun_out/r2b_tests.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2b_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'edge_block|transpose_blocks' -s 12 -c 10 -o gpurun_out/r2b_blocks -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_ncu_full.log 2>&1
ls -la gpurun_out/r2b_*
