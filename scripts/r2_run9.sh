# run 9: full GPU suite on a 2-GPU box (bf16 modes, graph capture diagnostics) + the graph test alone with one visible GPU
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s 2>&1 | grep -v "^$" > gpurun_out/r2i_tests.log; echo "tests exit ${PIPESTATUS[0]}"
grep "precision=" gpurun_out/r2i_tests.log | sort | uniq | head -40
tail -5 gpurun_out/r2i_tests.log
CUDA_VISIBLE_DEVICES=0 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "graph_scoring" 2>&1 | tail -5
python -m pytest tests/test_gpu_parity.py -m gpu -q -k "graph_scoring" 2>&1 | tail -5
