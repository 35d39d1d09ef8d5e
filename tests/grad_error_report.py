"""Measured gradient / score deviation of the CUDA path from the fp64 oracle, per fixture (run on a GPU box):

    python tests/grad_error_report.py [--out profiles/r2_grad_errors.json]

Prints, for every whole-model fixture of tests/test_gpu_parity.py, the worst per-tensor relative L2 and max-abs error of
the parameter gradients and the score error -- the numbers behind GRAD_TOL / TOL.  GCNN_LIB selects an A/B build.
Test infrastructure (imports the oracle); not part of the product path.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import gcnn_oracle as orc  # noqa: E402
from gcnn_cut_selector_b200 import GCNN, batching, synth  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def grad_errors(flat_got, grads_ref):
    flat_got = np.asarray(flat_got, np.float64)
    gmax = max(float(g.abs().max()) for g in grads_ref.values())
    gl2 = max(float(g.norm()) for g in grads_ref.values())
    worst_l2, worst_inf, o = (0.0, ""), (0.0, ""), 0
    for name, shape in orc.TRAINABLE:
        k = int(np.prod(shape))
        ref = grads_ref[name].reshape(-1).numpy().astype(np.float64)
        diff = flat_got[o:o + k] - ref
        l2 = np.linalg.norm(diff) / max(np.linalg.norm(ref), 1e-3 * gl2)
        linf = np.abs(diff).max() / max(np.abs(ref).max(), 1e-3 * gmax)
        if l2 > worst_l2[0]:
            worst_l2 = (float(l2), name)
        if linf > worst_inf[0]:
            worst_inf = (float(linf), name)
        o += k
    return worst_l2, worst_inf


def main():
    out_path = None
    if "--out" in sys.argv:
        out_path = sys.argv[sys.argv.index("--out") + 1]
    state = os.path.join(GOLDEN, "state_stream.pkl")
    oracle = orc.OracleGCNN(orc.restore_state(state, dtype=torch.float64), dtype=torch.float64)
    paths = {"tc_chains": dict(tensor_cores=1, fused=1, fused_backward=1, bf16_forward=1),
             "fp32": dict(tensor_cores=0, fused=1, fused_backward=1, bf16_forward=0)}
    cases = [("setcov", 1), ("setcov", 3), ("combauc", 4), ("indset", 4), ("capfac", 1), ("skewed", 3), ("mini", 3)]
    if "--big" in sys.argv:
        cases.append(("setcov", 32))
    rows = []
    for pname, opts in paths.items():
        m = GCNN(device="cuda:0", seed=0)
        if pname != "tc_chains" and not m._lib.gcnn_has_alt_paths():
            continue  # the fp32 SIMT path is an A/B alternate: GCNN_LIB_VARIANT=alt (-DGCNN_ALT_PATHS build)
        m.restore_state(state)
        for k, v in opts.items():
            m.set_option(k, v)
        for shape, n in cases:
            batch = batching.concat_samples(synth.make_samples(shape, n, seed0=1000 + n))
            for counts in (False, True):
                inputs, targets = batching.model_inputs(batch, per_sample_counts=counts), batch[10]
                loss_sum, scores = m.loss_and_grads(inputs, targets)
                torch.cuda.synchronize()
                if not counts:
                    loss, pred, grads = orc.loss_and_grads(oracle, batching.model_inputs(batch), targets)
                s_err = float(np.abs(scores.cpu().numpy() - pred.numpy()).max() / np.abs(pred.numpy()).max())
                l2, linf = grad_errors(m.flat_grads.cpu().numpy(), grads)
                row = {"path": pname, "shape": shape, "graphs": n, "per_sample_counts": counts, "score_err": s_err,
                       "grad_l2": l2[0], "grad_l2_tensor": l2[1], "grad_linf": linf[0], "grad_linf_tensor": linf[1]}
                rows.append(row)
                print(json.dumps(row), flush=True)
    worst = {}
    for r in rows:
        w = worst.setdefault(r["path"], {"score_err": 0.0, "grad_l2": 0.0, "grad_linf": 0.0})
        for k in w:
            w[k] = max(w[k], r[k])
    print(json.dumps({"worst": worst, "lib": os.environ.get("GCNN_LIB", "default")}))
    if out_path:
        with open(out_path, "w") as fh:
            json.dump({"rows": rows, "worst": worst, "lib": os.environ.get("GCNN_LIB", "default")}, fh, indent=1)


if __name__ == "__main__":
    main()
