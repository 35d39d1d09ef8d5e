# run 11: resident shards (tests + e2e_resident), host binding report, serving latency with the packed upload
mkdir -p gpurun_out
python -m pytest tests/test_shards.py tests/test_gpu_parity.py -m gpu -q -k "shard or records or assembly or graph_scoring" 2>&1 | tail -5
python bench.py --steps 30 --warmup 5 --no-extra-configs --no-cpu-baseline > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err; echo "bench exit $?"
python scripts/show_bench.py gpurun_out/r2k_bench.json | tail -4
python - <<'P'
import json
d = json.load(open("gpurun_out/r2k_bench.json"))
for k in ("e2e", "e2e_records", "e2e_resident", "host_binding"):
    print(k, d.get(k) if k == "host_binding" else {x: d[k][x] for x in ("value", "ms_per_step", "h2d_bytes_per_step")})
P
nvidia-smi topo -m 2>&1 | head -20
lscpu | grep -i "numa\|socket\|model name" | head
