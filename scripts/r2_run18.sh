# run 18 (4 GPUs): the 4-rank instantiation of the peer exchange (dp check + bench line), GPU suite on a multi-GPU box
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29577"
timeout 300 $TR tests/dp_gpu_check.py > gpurun_out/r2s_dp_check_4gpu.log 2>&1; echo "dp check exit $?"; grep "^{" gpurun_out/r2s_dp_check_4gpu.log | cut -c1-400
timeout 400 $TR bench.py --gpus 4 --steps 40 --warmup 5 > gpurun_out/r2s_bench_4gpu.json 2> gpurun_out/r2s_bench_4gpu.err; echo "bench exit $?"
python - <<'P'
import json
d = json.load(open("gpurun_out/r2s_bench_4gpu.json"))
print("value", round(d["value"]), "ms", round(d["ms_per_step"], 4), {k: (round(d[k]["value"]), round(d[k]["ms_per_step"], 4)) for k in ("e2e", "e2e_records", "e2e_resident")})
c = d["config4"]; print("config4", round(c["value"]), round(c["ms_per_step"], 4), {k: round(c[k]["value"]) for k in ("e2e", "e2e_records", "e2e_resident")})
P
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -3
