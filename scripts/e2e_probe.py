"""e2e step time with and without the per-step host-to-device copies (same staged entry points)."""
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


def run(copies: bool, lag: bool, n=300):
    model.stage_host(host[0], 0)
    model.stage_host(host[1], 1)
    torch.cuda.synchronize()
    pend = []
    t0 = None
    for i in range(n + 20):
        if i == 20:
            torch.cuda.synchronize()
            t0 = time.perf_counter()
        if copies:
            model.stage_host(host[(i + 1) % 4], (i + 1) & 1)
        model.train_step_staged_async(i & 1, 1e-4)
        pend.append(i & 1)
        if not lag or len(pend) > 1:
            model.train_step_result(pend.pop(0))
    while pend:
        model.train_step_result(pend.pop(0))
    torch.cuda.synchronize()
    return 1e3 * (time.perf_counter() - t0) / n


for copies in (True, False):
    for lag in (True, False):
        print(f"copies={copies} loss_lag={lag}: {run(copies, lag):.4f} ms per step")
