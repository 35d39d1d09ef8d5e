#!/bin/bash
# round 2, GPU call 8 (2 GPUs): parity suite, peer-memory exchange check, 2-GPU bench peer vs NCCL
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2h_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/r2h_tests.log
tail -8 gpurun_out/r2h_tests.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dp_gpu_check.py > gpurun_out/r2h_dp_check.log 2>&1; echo "dp check exit $?" >> gpurun_out/r2h_dp_check.log
grep -E "^\{|exit|Error|error" gpurun_out/r2h_dp_check.log | tail -8
for mode in 1 0; do
timeout 600 env GCNN_DP_PEER=$mode python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 50 --warmup 5 --no-extra-configs > gpurun_out/r2h_bench_2gpu_peer$mode.json 2> gpurun_out/r2h_bench_2gpu_peer$mode.err
python -c "
import json; d=json.load(open('gpurun_out/r2h_bench_2gpu_peer$mode.json')); print('peer=$mode', d['value'], d['ms_per_step'], d['e2e']['value'])"
done
