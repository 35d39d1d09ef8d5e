# head fused into the last forward chain: GPU suite (both libraries), smoke, same-box A/B of option head_in_chain
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2x_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/r2x_tests.log
tail -4 gpurun_out/r2x_tests.log
python __graft_entry__.py --smoke > gpurun_out/r2x_smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/r2x_smoke.log
python scripts/ab_option.py head_in_chain 0 1 > gpurun_out/r2x_ab_head.json 2> gpurun_out/r2x_ab_head.err; cat gpurun_out/r2x_ab_head.json; tail -3 gpurun_out/r2x_ab_head.err
GCNN_LIB_VARIANT=alt timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2x_tests_alt.log 2>&1; echo "alt tests exit $?" >> gpurun_out/r2x_tests_alt.log
tail -4 gpurun_out/r2x_tests_alt.log
