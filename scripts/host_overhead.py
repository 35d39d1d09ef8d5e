"""Host-side cost of enqueueing one staged training step (no waiting on the GPU)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from gcnn_cut_selector_b200 import GCNN, HostBatch
dev = torch.device("cuda:0")
model = GCNN(device=dev, seed=0)
model.check_indices = False
host = [HostBatch(b) for b in bench.make_batches(4, 32, seed0=0)]
model.stage_host(host[0], 0)
for i in range(20):
    model.stage_host(host[(i + 1) % 4], (i + 1) & 1)
    model.train_step_staged_async(i & 1, 1e-4)
    if i: model.train_step_result((i - 1) & 1)
torch.cuda.synchronize()
ts, tt = 0.0, 0.0
n = 200
for i in range(20, 20 + n):
    t0 = time.perf_counter()
    model.stage_host(host[(i + 1) % 4], (i + 1) & 1)
    t1 = time.perf_counter()
    model.train_step_staged_async(i & 1, 1e-4)
    t2 = time.perf_counter()
    torch.cuda.synchronize()  # isolate the enqueue cost from back-pressure
    ts += t1 - t0
    tt += t2 - t1
print(f"stage_host enqueue {1e6 * ts / n:.1f} us, train_step_staged_async enqueue {1e6 * tt / n:.1f} us per step")
