#!/bin/bash
# round 2, GPU call 7: full parity suite (select_cuts, graph scoring), full bench line, reference arm
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2g_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/r2g_tests.log
tail -15 gpurun_out/r2g_tests.log
python bench.py --steps 50 --warmup 5 > gpurun_out/r2g_bench_full.json 2> gpurun_out/r2g_bench_full.err
tail -3 gpurun_out/r2g_bench_full.err
python scripts/show_bench.py gpurun_out/r2g_bench_full.json
python -c "
import json; d=json.load(open('gpurun_out/r2g_bench_full.json'))
print(json.dumps(d.get('configs'), indent=1)[:6000]); print(d.get('cpu_baseline'))"
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2g_smoke.log 2>&1; tail -2 gpurun_out/r2g_smoke.log
