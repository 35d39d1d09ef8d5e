"""Same-box A/B of a library option (gcnn_set_option) or of two builds: device-timed 32-graph training step (L2 flushed
between steps), device forward of a single combauc graph, graph-replayed host scoring.
    python scripts/ab_option.py head_in_chain 0 1 [--graphs 32] [--rounds 3]
prints one JSON line {option, value -> [step ms per round], ...}."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gcnn_cut_selector_b200 import GCNN, HostBatch, batching, synth  # noqa: E402

name, values = sys.argv[1], [int(x) for x in sys.argv[2:] if not x.startswith("--") and x.lstrip("-").isdigit()][:2]
arg = lambda k, d: int(sys.argv[sys.argv.index(k) + 1]) if k in sys.argv else d
graphs, rounds = arg("--graphs", 32), arg("--rounds", 3)
dev = torch.device("cuda:0")
torch.cuda.set_device(0)
model = GCNN(device=dev, seed=0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
batches = [batching.concat_samples(synth.make_samples("setcov", graphs, seed0=1000 * i)) for i in range(4)]
inputs = [model.prepare_inputs(batching.model_inputs(b, per_sample_counts=True)) for b in batches]
targets = [torch.from_numpy(b[10]).to(dev) for b in batches]
small = batching.concat_samples(synth.make_samples("combauc", 1, seed0=300))
small_in = model.prepare_inputs(batching.model_inputs(small, per_sample_counts=True))
small_hb = HostBatch(small)


def timed(fn, reps):
    for i in range(5):
        fn(i)
    torch.cuda.synchronize()
    tot = 0.0
    for i in range(reps):
        flush.fill_(i & 0xFF)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn(i)
        e1.record()
        e1.synchronize()
        tot += e0.elapsed_time(e1)
    return round(tot / reps, 5)


def step(i):
    model.loss_and_grads(inputs[i % 4], targets[i % 4])
    model.apply_gradients(1e-4)


def fwd(i):
    with torch.no_grad():
        model._forward(small_in, save_activations=False)


def host_graph():
    t = []
    for _ in range(10):
        model.score_host(small_hb, graph=True)
    for _ in range(200):
        t0 = time.perf_counter()
        model.score_host(small_hb, graph=True)
        t.append(time.perf_counter() - t0)
    return round(1e3 * float(np.median(t)), 5)


out = {"option": name, "graphs": graphs}
for r in range(rounds):
    for v in values:
        model.set_option(name, v)
        d = out.setdefault(str(v), {"step_ms": [], "combauc_forward_ms": [], "combauc_score_host_graph_ms": []})
        d["step_ms"].append(timed(step, 40))
        d["combauc_forward_ms"].append(timed(fwd, 40))
        d["combauc_score_host_graph_ms"].append(host_graph())
print(json.dumps(out))
