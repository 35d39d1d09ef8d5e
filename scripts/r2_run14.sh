# run 14: product library (alternates compiled out) and the -DGCNN_ALT_PATHS variant through the whole GPU suite; smoke
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/r2n_tests_product.log; cat gpurun_out/r2n_tests_product.log
GCNN_LIB_VARIANT=alt python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/r2n_tests_alt.log; cat gpurun_out/r2n_tests_alt.log
python __graft_entry__.py --smoke 2>&1 | tail -2
python tests/grad_error_report.py --big --out gpurun_out/r2n_grad_err_product.json > gpurun_out/r2n_grad_err_product.log 2>&1; tail -1 gpurun_out/r2n_grad_err_product.log
GCNN_LIB_VARIANT=alt python tests/grad_error_report.py --big --out gpurun_out/r2n_grad_err_alt.json > gpurun_out/r2n_grad_err_alt.log 2>&1; tail -1 gpurun_out/r2n_grad_err_alt.log
