"""Same-box A/B of the host-side paths: host batches plain (13 copies) / packed (one copy) / row pointers + uint16 local
columns / both; option "params_epoch" on / off.
Prints one JSON line.  (python scripts/ab_host_paths.py [--e2e-graphs 32])"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gcnn_cut_selector_b200 import GCNN, HostBatch, batching, synth  # noqa: E402

dev = torch.device("cuda:0")
torch.cuda.set_device(0)
model = GCNN(device=dev, seed=0)
out = {}


def lat(fn, n=200, warm=10):
    for _ in range(warm):
        fn()
    t = []
    for _ in range(n):
        t0 = time.perf_counter()
        fn()
        t.append(time.perf_counter() - t0)
    return round(1e3 * float(np.median(t)), 4)


VARIANTS = {"plain": (False, False), "packed": (False, True), "rowptr": (True, False), "rowptr_packed": (True, True)}
announce = GCNN._announce_params


def epoch(on: bool):
    if on:
        model._announce_params = announce.__get__(model)
        model._params_seen = None
    else:
        model._announce_params = lambda: None
        model.set_option("params_epoch", 0)


for shape, n in (("combauc", 1), ("indset", 1), ("capfac", 1)):
    batch = batching.concat_samples(synth.make_samples(shape, n, seed0=300))
    hb = {v: HostBatch(batch, row_pointers=v[0], packed=v[1]) for v in VARIANTS.values()}
    res = {}
    for rnd in range(2):
        for ep in (False, True):
            epoch(ep)
            for name, rp in VARIANTS.items():
                k = f"epoch{int(ep)}_{name}"
                res.setdefault(k + "_eager", []).append(lat(lambda: model.score_host(hb[rp])))
                res.setdefault(k + "_graph", []).append(lat(lambda: model.score_host(hb[rp], graph=True)))
    out[f"score_{shape}"] = {k: min(v) for k, v in res.items()}

epoch(True)
batch = batching.concat_samples(synth.make_samples("setcov", 1, seed0=77))
hb = {v: HostBatch(batch, row_pointers=v[0], packed=v[1]) for v in VARIANTS.values()}
res = {}
for rnd in range(2):
    for name, rp in VARIANTS.items():
        res.setdefault(name, []).append(lat(lambda: model.train_step_host(hb[rp], 1e-4), 100))
out["config1_train_step_host"] = {k: min(v) for k, v in res.items()}

graphs = int(sys.argv[sys.argv.index("--e2e-graphs") + 1]) if "--e2e-graphs" in sys.argv else 32
batches = [batching.concat_samples(synth.make_samples("setcov", graphs, seed0=1000 * i)) for i in range(4)]
hosts = {v: [HostBatch(b, row_pointers=v[0], packed=v[1]) for b in batches] for v in VARIANTS.values()}
for h in hosts[(False, False)]:
    model.reserve(h.batch, True)


def e2e(host, K=60, W=5):
    pending = []
    model.stage_host(host[0], 0)

    def step(i):
        model.stage_host(host[(i + 1) % 4], (i + 1) & 1)
        model.train_step_staged_async(i & 1, 1e-4)
        pending.append(i & 1)
        if len(pending) > 1:
            model.train_step_result(pending.pop(0))
    for i in range(W):
        step(i)
    while pending:
        model.train_step_result(pending.pop(0))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(K):
        step(W + i)
    while pending:
        model.train_step_result(pending.pop(0))
    torch.cuda.synchronize()
    return 1e3 * (time.perf_counter() - t0) / K


res = {}
for rnd in range(3):
    for name, rp in VARIANTS.items():
        res.setdefault(name, []).append(round(e2e(hosts[rp]), 4))
out[f"e2e_{graphs}_graphs_ms"] = res
out["h2d_bytes"] = {name: hosts[rp][0].h2d_bytes for name, rp in VARIANTS.items()}
print(json.dumps(out))
