"""Where the time of the records-fed loop goes: host cost of a stage call, device time of staging alone, loop times."""
import os, sys, time, tempfile
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gcnn_cut_selector_b200 import GCNN, HostBatch, batching, shards, synth
dev = torch.device("cuda:0")
model = GCNN(device=dev, seed=0); model.check_indices = False
graphs, n_rot = 32, 4
sets = [synth.make_samples("setcov", graphs, seed0=1000 * b, n_structures=8) for b in range(n_rot)]
host = [HostBatch(batching.concat_samples(s)) for s in sets]
with tempfile.TemporaryDirectory() as tmp:
    shards.write_shard(os.path.join(tmp, "a.shard"), [x for s in sets for x in s])
    reader = shards.ShardReader(os.path.join(tmp, "a.shard"))
ids = [list(range(b * graphs, (b + 1) * graphs)) for b in range(n_rot)]
stagers = {"records": lambda i: model.stage_records(reader, ids[i % n_rot], i & 1), "host": lambda i: model.stage_host(host[i % n_rot], i & 1)}
for name, stage in stagers.items():
    for i in range(4): stage(i); model.train_step_staged(i & 1, 1e-4)
    torch.cuda.synchronize()
    t_host, t_dev = [], []
    for i in range(20):
        torch.cuda.synchronize(); t0 = time.perf_counter(); stage(i); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
        t_host.append(t1 - t0); t_dev.append(t2 - t0)
        model.train_step_staged(i & 1, 1e-4)
    # pipelined loop
    pend = []
    def step(i):
        stage(i + 1); model.train_step_staged_async(i & 1, 1e-4); pend.append(i & 1)
        if len(pend) > 1: model.train_step_result(pend.pop(0))
    stage(0)
    for i in range(5): step(i)
    while pend: model.train_step_result(pend.pop(0))
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(5, 205): step(i)
    while pend: model.train_step_result(pend.pop(0))
    torch.cuda.synchronize(); loop = (time.perf_counter() - t0) / 200
    # serial (no overlap): stage, sync, step
    t0 = time.perf_counter()
    for i in range(50): stage(i); model.train_step_staged(i & 1, 1e-4)
    serial = (time.perf_counter() - t0) / 50
    print(f"{name}: stage call host {1e6*np.median(t_host):.0f} us, stage alone to completion {1e6*np.median(t_dev):.0f} us, pipelined loop {1e3*loop:.4f} ms/step, serial {1e3*serial:.4f} ms/step", flush=True)
