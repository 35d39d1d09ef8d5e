# run 15: final tree -- product + alt suites, smoke, default bench line, serving latency (round-1 method), 2-GPU dp check
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/r2p_tests_product.log; cat gpurun_out/r2p_tests_product.log
GCNN_LIB_VARIANT=alt python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/r2p_tests_alt.log; cat gpurun_out/r2p_tests_alt.log
python __graft_entry__.py --smoke 2>&1 | tail -2
python bench.py > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err; echo "bench exit $?"
python scripts/show_bench.py gpurun_out/r2p_bench.json | tail -4
python scripts/score_latency.py > gpurun_out/r2p_score_latency.jsonl 2> gpurun_out/r2p_score_latency.err; cat gpurun_out/r2p_score_latency.jsonl | cut -c1-250
python -m pytest tests/test_driver_trace.py -m gpu -q -s 2>&1 | grep "max_score_err" | cut -c1-400
