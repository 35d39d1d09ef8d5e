# run 12 (8 GPUs) -- peer-memory all-reduce vs NCCL, host / record / resident end-to-end, then the full line with config 4
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533"
timeout 300 $TR tests/dp_gpu_check.py > gpurun_out/r2l_dp_check_8gpu.log 2>&1; echo "dp check exit $?"; grep "^{" gpurun_out/r2l_dp_check_8gpu.log
for peer in 1 0; do
  GCNN_DP_PEER=$peer timeout 300 $TR bench.py --gpus 8 --steps 40 --warmup 5 --no-extra-configs > gpurun_out/r2l_bench_8gpu_peer$peer.json 2> gpurun_out/r2l_bench_8gpu_peer$peer.err
  echo "peer=$peer exit $?"
done
timeout 400 $TR bench.py --gpus 8 --steps 40 --warmup 5 > gpurun_out/r2l_bench_8gpu_full.json 2> gpurun_out/r2l_bench_8gpu_full.err; echo "full exit $?"
python - <<'P'
import json
for f in ("r2l_bench_8gpu_peer1", "r2l_bench_8gpu_peer0", "r2l_bench_8gpu_full"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
    except Exception as e:
        print(f, "unreadable", e); continue
    print(f, "value", round(d["value"]), "ms", round(d["ms_per_step"], 4), "binding", d.get("host_binding"))
    for k in ("e2e", "e2e_records", "e2e_resident"):
        print("   ", k, round(d[k]["value"]), round(d[k]["ms_per_step"], 4), d[k]["h2d_bytes_per_step"])
    if "config4" in d:
        c = d["config4"]
        print("    config4", round(c["value"]), round(c["ms_per_step"], 4), {k: round(c[k]["value"]) for k in ("e2e", "e2e_records", "e2e_resident")})
P
lscpu | grep -i "numa\|socket\|^CPU(s)" | head
nvidia-smi topo -m 2>&1 | head -12
