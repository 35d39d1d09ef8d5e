# tests + bench classes + inference profile in one box
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s3_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/s3_tests.log
tail -5 gpurun_out/s3_tests.log
for lr in 512; do
echo "== long_row $lr"
GCNN_LONG_ROW=$lr python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/ab_new$lr.json 2> gpurun_out/ab_new$lr.err; python scripts/show_bench.py gpurun_out/ab_new$lr.json | grep -E "edge_|sum of"
GCNN_LONG_ROW=$lr python scripts/forward_profile.py capfac:1 capfac:4 miplib:1 setcov:32 indset:4 combauc:4 > gpurun_out/fwd_profile_$lr.jsonl 2> gpurun_out/fwd_profile_$lr.err; tail -3 gpurun_out/fwd_profile_$lr.err
done
