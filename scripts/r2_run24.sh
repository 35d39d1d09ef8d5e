# final tree, 8 GPUs: data-parallel check, then the full bench line (with config 4)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533"
timeout 300 $TR tests/dp_gpu_check.py > gpurun_out/r2g_dp_check_8gpu.log 2>&1; echo "dp check exit $?"; grep "^{" gpurun_out/r2g_dp_check_8gpu.log | cut -c1-400
timeout 400 $TR bench.py --gpus 8 --steps 40 --warmup 5 > gpurun_out/r2g_bench_8gpu.json 2> gpurun_out/r2g_bench_8gpu.err; echo "full exit $?"
python - <<'P'
import json
d = json.load(open("gpurun_out/r2g_bench_8gpu.json"))
print("value", round(d["value"]), "ms", round(d["ms_per_step"], 4))
for k in ("e2e", "e2e_records", "e2e_resident"):
    print("   ", k, round(d[k]["value"]), round(d[k]["ms_per_step"], 4), d[k]["h2d_bytes_per_step"])
c = d["config4"]
print("    config4", round(c["value"]), round(c["ms_per_step"], 4), {k: (round(c[k]["value"]), round(c[k]["ms_per_step"], 4)) for k in ("e2e", "e2e_records", "e2e_resident")})
P
