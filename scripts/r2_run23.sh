# 8 GPUs: host batches packed in 1 MB pieces, with / without row pointers + local columns
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533"
for rp in 1 0; do
GCNN_HOST_ROW_POINTERS=$rp timeout 300 $TR bench.py --gpus 8 --steps 40 --warmup 5 --no-extra-configs > gpurun_out/r2ze_bench_8gpu_rp$rp.json 2> gpurun_out/r2ze_bench_8gpu_rp$rp.err; echo "rp=$rp exit $?"
done
python - <<'P'
import json
for f in ("r2ze_bench_8gpu_rp1", "r2ze_bench_8gpu_rp0"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
    except Exception as e:
        print(f, "unreadable", e); continue
    print(f, "value", round(d["value"]), "ms", round(d["ms_per_step"], 4))
    for k in ("e2e", "e2e_records", "e2e_resident"):
        print("   ", k, round(d[k]["value"]), round(d[k]["ms_per_step"], 4), d[k]["h2d_bytes_per_step"])
    if "config4" in d:
        c = d["config4"]
        print("    config4", round(c["value"]), round(c["ms_per_step"], 4), {k: (round(c[k]["value"]), round(c[k]["ms_per_step"], 4)) for k in ("e2e", "e2e_records", "e2e_resident")})
P
