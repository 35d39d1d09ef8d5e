#!/bin/bash
# round 2, GPU call 1: parity suite, measured gradient errors (3 vs 2 activation pieces), A/B bench of the block kernels
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2a_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/r2a_tests.log
tail -5 gpurun_out/r2a_tests.log
python tests/grad_error_report.py --big --out gpurun_out/r2a_grad_err_act3.json > gpurun_out/r2a_grad_err_act3.log 2>&1
GCNN_LIB=gcnn_cut_selector_b200/build/act2/libgcnn_b200.so python tests/grad_error_report.py --big --out gpurun_out/r2a_grad_err_act2.json > gpurun_out/r2a_grad_err_act2.log 2>&1
tail -1 gpurun_out/r2a_grad_err_act3.log; tail -1 gpurun_out/r2a_grad_err_act2.log
for rep in 1 2; do
python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2a_bench_blocks_$rep.json 2> gpurun_out/r2a_bench_blocks_$rep.err
GCNN_BLOCKS=0 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2a_bench_generic_$rep.json 2> gpurun_out/r2a_bench_generic_$rep.err
GCNN_LIB=gcnn_cut_selector_b200/build/act2/libgcnn_b200.so python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2a_bench_act2_$rep.json 2> gpurun_out/r2a_bench_act2_$rep.err
done
python bench.py --steps 50 --warmup 5 --no-cpu-baseline --graphs-per-gpu 128 > gpurun_out/r2a_bench_blocks_g128.json 2> gpurun_out/r2a_bench_blocks_g128.err
GCNN_BLOCKS=0 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --graphs-per-gpu 128 > gpurun_out/r2a_bench_generic_g128.json 2> gpurun_out/r2a_bench_generic_g128.err
python scripts/show_bench.py gpurun_out/r2a_bench_*.json
