"""Cut-scoring (inference) latency and throughput, BASELINE configs 3 and 5: host batch in, host scores out
(`GCNN.score_host`, the `get_improvements(state, False).numpy()` path of model_benchmarker.py:91-106) and the
device-resident forward.  Prints one JSON line per case.  Usage (GPU box): python scripts/score_latency.py"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gcnn_cut_selector_b200 import GCNN, HostBatch, batching, synth  # noqa: E402

dev = torch.device("cuda:0")
model = GCNN(device=dev, seed=0)
model.check_indices = False
cases = [("setcov", 1), ("setcov", 4), ("combauc", 1), ("combauc", 4), ("capfac", 1), ("capfac", 4), ("indset", 1),
         ("indset", 4), ("miplib", 1)]
for shape, n in cases:
    batch = batching.concat_samples(synth.make_samples(shape, n, seed0=300))
    hb = HostBatch(batch)
    inputs = model.prepare_inputs(batching.model_inputs(batch))
    reps = 30 if shape == "miplib" else 200
    for _ in range(5):
        model.score_host(hb)
    torch.cuda.synchronize()
    lat = []
    for _ in range(reps):
        t0 = time.perf_counter()
        model.score_host(hb)
        lat.append(time.perf_counter() - t0)
    with torch.no_grad():
        for _ in range(5):
            model._forward(inputs, save_activations=False)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            model._forward(inputs, save_activations=False)
        e1.record()
        torch.cuda.synchronize()
    nc, nv, nk = int(np.sum(batch[7])), int(np.sum(batch[8])), int(np.sum(batch[9]))
    e_c, e_k = batch[1].shape[1], batch[5].shape[1]
    dev_ms = e0.elapsed_time(e1) / reps
    print(json.dumps({"shape": shape, "graphs": n, "n_cons": nc, "n_vars": nv, "n_cuts": nk, "edges": e_c + e_k,
                      "score_host_ms_p50": 1e3 * float(np.median(lat)), "score_host_ms_p95": 1e3 * float(np.percentile(lat, 95)),
                      "device_forward_ms": dev_ms, "cuts_per_s_device": nk / (dev_ms * 1e-3),
                      "edge_messages_per_s_device": (2 * e_c + e_k) / (dev_ms * 1e-3)}), flush=True)
