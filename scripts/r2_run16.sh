# run 16: masked block pair -- tests, then same-box A/B over the per-convolution mask (0 = off, 7 = all)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "masked_block or per_sample_counts or config2 or problem_class" 2>&1 | tail -8
for m in 0 7 2 6 3; do
  GCNN_BLOCK_MASKS=$m timeout 300 python bench.py --steps 40 --warmup 5 --no-extra-configs --no-cpu-baseline > gpurun_out/r2q_bench_m$m.json 2> gpurun_out/r2q_bench_m$m.err
  echo "masks=$m exit $?"
  python scripts/show_bench.py gpurun_out/r2q_bench_m$m.json | grep -E "edge_forward|edge_backward|sum of kernel"
done
GCNN_BLOCK_MASKS=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-extra-configs --no-cpu-baseline --graphs-per-gpu 128 > gpurun_out/r2q_bench_g128_m0.json 2>/dev/null
GCNN_BLOCK_MASKS=7 timeout 300 python bench.py --steps 20 --warmup 5 --no-extra-configs --no-cpu-baseline --graphs-per-gpu 128 > gpurun_out/r2q_bench_g128_m7.json 2>/dev/null
python scripts/show_bench.py gpurun_out/r2q_bench_g128_m0.json gpurun_out/r2q_bench_g128_m7.json | grep -E "==|edge_forward|edge_backward|sum of kernel"
