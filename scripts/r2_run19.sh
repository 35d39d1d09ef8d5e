# row-pointer host batches + params_epoch: the GPU suite, smoke, the bench line (value / e2e / serving latencies)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2v_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/r2v_tests.log
tail -4 gpurun_out/r2v_tests.log
python __graft_entry__.py --smoke > gpurun_out/r2v_smoke.log 2>&1; echo "smoke exit $?"
python bench.py > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err; echo "bench exit $?"
python scripts/show_bench.py gpurun_out/r2v_bench.json | tail -30
