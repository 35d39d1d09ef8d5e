# round-end sequence on one B200: GPU tests, bench (both arms), inference tables
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r1_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/r1_tests.log
tail -3 gpurun_out/r1_tests.log
python bench.py > gpurun_out/r1_bench.json 2> gpurun_out/r1_bench.err; echo "bench exit $?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r1_bench_reference.json 2> gpurun_out/r1_bench_reference.err; echo "reference exit $?"
python scripts/score_latency.py > gpurun_out/r1_score_latency.jsonl 2> gpurun_out/r1_score_latency.err
python scripts/forward_profile.py capfac:1 capfac:4 miplib:1 setcov:32 indset:4 combauc:4 > gpurun_out/r1_forward_profile.jsonl 2> gpurun_out/r1_forward_profile.err
python __graft_entry__.py --smoke > gpurun_out/r1_smoke.log 2>&1; echo "smoke exit $?"
python scripts/show_bench.py gpurun_out/r1_bench.json
