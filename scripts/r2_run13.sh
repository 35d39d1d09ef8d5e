# run 13 (2 GPUs): dp check incl. the staged one-call path, 2-GPU bench peer vs NCCL
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544"
timeout 300 $TR tests/dp_gpu_check.py > gpurun_out/r2m_dp_check.log 2>&1; echo "dp check exit $?"; tail -3 gpurun_out/r2m_dp_check.log
for peer in 1 0; do
  GCNN_DP_PEER=$peer timeout 300 $TR bench.py --gpus 2 --steps 40 --warmup 5 --no-extra-configs > gpurun_out/r2m_bench_2gpu_peer$peer.json 2> gpurun_out/r2m_bench_2gpu_peer$peer.err
  echo "peer=$peer exit $?"
done
python - <<'P'
import json
for f in ("r2m_bench_2gpu_peer1", "r2m_bench_2gpu_peer0"):
    d = json.load(open(f"gpurun_out/{f}.json"))
    print(f, "value", round(d["value"]), "ms", round(d["ms_per_step"], 4))
    for k in ("e2e", "e2e_records", "e2e_resident"):
        print("   ", k, round(d[k]["value"]), round(d[k]["ms_per_step"], 4), d[k]["h2d_bytes_per_step"])
P
