# run 21 (8 GPUs) -- the full bench line of the current tree (row-pointer host batches: 10.2 MB instead of 14.2 MB per rank and step)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533"
timeout 400 $TR bench.py --gpus 8 --steps 40 --warmup 5 > gpurun_out/r2y_bench_8gpu.json 2> gpurun_out/r2y_bench_8gpu.err; echo "full exit $?"
python - <<'P'
import json
for f in ("r2y_bench_8gpu",):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
    except Exception as e:
        print(f, "unreadable", e); continue
    print(f, "value", round(d["value"]), "ms", round(d["ms_per_step"], 4), "binding", d.get("host_binding"))
    for k in ("e2e", "e2e_records", "e2e_resident"):
        print("   ", k, round(d[k]["value"]), round(d[k]["ms_per_step"], 4), d[k]["h2d_bytes_per_step"])
    if "config4" in d:
        c = d["config4"]
        print("    config4", round(c["value"]), round(c["ms_per_step"], 4), {k: (round(c[k]["value"]), round(c[k]["ms_per_step"], 4)) for k in ("e2e", "e2e_records", "e2e_resident")})
P
tail -3 gpurun_out/r2y_bench_8gpu.err
