# run 10: graph-capture fix check + serving latency eager vs graph
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q -k "graph_scoring" 2>&1 | tail -15
python - <<'P' 2>&1 | tee gpurun_out/r2j_serve_latency.jsonl
import json, time, numpy as np, torch
from gcnn_cut_selector_b200 import GCNN, HostBatch
from gcnn_cut_selector_b200 import batching, synth
m = GCNN(device="cuda:0", seed=0)
for shape in ("combauc", "capfac", "indset", "setcov"):
    hb = HostBatch(batching.concat_samples(synth.make_samples(shape, 1, seed0=3)))
    res = {"shape": shape}
    for mode in (False, True):
        for _ in range(10): m.score_host(hb, graph=mode)
        ts = []
        for _ in range(300):
            t = time.perf_counter(); m.score_host(hb, graph=mode); ts.append(time.perf_counter() - t)
        ts = np.sort(ts) * 1e3
        res["graph" if mode else "eager"] = {"p50_ms": float(ts[150]), "p95_ms": float(ts[285])}
    res["graphs_cached"] = int(m._lib.gcnn_serve_graph_count(m._ws))
    print(json.dumps(res))
P
