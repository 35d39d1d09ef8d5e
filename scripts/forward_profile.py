"""Per-kernel-class CUDA-event times of the inference forward (BASELINE configs 3 and 5) -- the same profiling hooks
bench.py uses for the training step.  Usage (GPU box): python scripts/forward_profile.py [shape:graphs ...]"""
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gcnn_cut_selector_b200 import GCNN, batching, synth  # noqa: E402
from gcnn_cut_selector_b200._lib import check, load  # noqa: E402

lib = load()

dev = torch.device("cuda:0")
model = GCNN(device=dev, seed=0)
model.check_indices = False
cases = [a.split(":") for a in sys.argv[1:]] or [("capfac", "4"), ("miplib", "1")]
ncls = lib.gcnn_profile_num_classes()
for shape, n in cases:
    batch = batching.concat_samples(synth.make_samples(shape, int(n), seed0=300))
    inputs = model.prepare_inputs(batching.model_inputs(batch))
    reps = 20
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
        total = e0.elapsed_time(e1) / reps
        ms, ln, by = (C.c_double * ncls)(), (C.c_int64 * ncls)(), (C.c_double * ncls)()
        model.set_option("streams", 0)
        lib.gcnn_profile_begin()
        for _ in range(reps):
            model._forward(inputs, save_activations=False)
        check(lib.gcnn_profile_end(ms, ln, by, ncls))
        model.set_option("streams", 1)
    out = {"shape": shape, "graphs": int(n), "forward_ms": total, "classes": {}}
    for c in range(ncls):
        if ln[c]:
            out["classes"][lib.gcnn_profile_class_name(c).decode()] = {
                "launches": ln[c] / reps, "ms": ms[c] / reps, "gbs": by[c] / (ms[c] * 1e-3) / 1e9 if ms[c] > 0 else 0.0}
    print(json.dumps(out), flush=True)
