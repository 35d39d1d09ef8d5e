"""Where the data-parallel step spends its time (run under torchrun, 2+ GPUs): CUDA-event time of the step with its
parts enabled one after another.  Usage: torchrun --nproc-per-node 2 scripts/dp_breakdown.py"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from gcnn_cut_selector_b200 import GCNN, DataParallelTrainer, batching  # noqa: E402

rank, local = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
model = GCNN(device=dev, seed=0)
model.check_indices = False
tr = DataParallelTrainer(model, 1e-4)
tr.broadcast_parameters()
batches = bench.make_batches(4, 32, seed0=10_000 * rank)
ins = [model.prepare_inputs(batching.model_inputs(b, per_sample_counts=True)) for b in batches]
tgt = [torch.from_numpy(b[10]).to(dev) for b in batches]
N = tr.N


def variant(level):
    def f(i):
        j = i % 4
        model.loss_and_grads(ins[j], tgt[j], seed_scale=1.0, loss_out=tr.bucket[N + 1:N + 2])  # count + loss on device
        if level >= 2:
            dist.all_reduce(tr.bucket)
        if level >= 3:
            model.apply_gradients(1e-4, grad_divisor=tr.bucket[N:N + 1])
        if level >= 4:
            return tr.bucket[N + 1:N + 2] / tr.bucket[N:N + 1]
    return f


names = ["forward+backward (+ bucket tail written by the library)", "(same)", "+ all_reduce", "+ adam", "+ mean-loss tensor"]
for level, name in enumerate(names):
    f = variant(level)
    for i in range(10):
        f(i)
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(100):
        f(i)
    e1.record()
    torch.cuda.synchronize()
    if rank == 0:
        print(f"{name:32s} {e0.elapsed_time(e1) / 100:.4f} ms per step", flush=True)
dist.destroy_process_group()
