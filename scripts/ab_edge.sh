mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s3_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/s3_tests.log
tail -5 gpurun_out/s3_tests.log
python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/s3_bench.json 2> gpurun_out/s3_bench.err; tail -3 gpurun_out/s3_bench.err; python scripts/show_bench.py gpurun_out/s3_bench.json | grep -E "edge_|sum of"
python -c "
import json; d=json.load(open('gpurun_out/s3_bench.json')); print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['h2d_bytes_per_step']); print('e2e_host', d['e2e_records']['value'], d['e2e_records']['ms_per_step'], d['e2e_records']['h2d_bytes_per_step'])"
