"""Per-tensor gradient error of the CUDA path against the fp64 oracle for a few library configurations (debug aid)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import gcnn_oracle as orc
from gcnn_cut_selector_b200 import GCNN, batching, synth

path = os.path.join(ROOT, "tests", "golden", "state_stream.pkl")
m = GCNN(device="cuda:0", seed=0); m.restore_state(path)
o64 = orc.OracleGCNN(orc.restore_state(path, dtype=torch.float64), dtype=torch.float64)
shape, n = (sys.argv[1], int(sys.argv[2])) if len(sys.argv) > 2 else ("setcov", 3)
batch = batching.concat_samples(synth.make_samples(shape, n, seed0=4242))
totals, vectors = batching.model_inputs(batch), batching.model_inputs(batch, per_sample_counts=True)
loss, pred, grads = orc.loss_and_grads(o64, totals, batch[10])
gmax = max(float(g.abs().max()) for g in grads.values())
for name, opts, inp in [("tc+tiles", {"tensor_cores": 1, "tiles": 1}, vectors), ("tc", {"tensor_cores": 1, "tiles": 0}, vectors),
                        ("simt", {"tensor_cores": 0, "tiles": 0}, vectors), ("simt+tiles", {"tensor_cores": 0, "tiles": 1}, vectors),
                        ("tc unfused", {"tensor_cores": 1, "tiles": 0, "fused": 0}, vectors)]:
    for k, v in {"tensor_cores": 1, "tiles": 1, "fused": 1, **opts}.items():
        m.set_option(k, v)
    _, scores = m.loss_and_grads(inp, batch[10]); torch.cuda.synchronize()
    got = m.flat_grads.cpu().numpy().astype(np.float64)
    off, worst = 0, []
    for n, shape in orc.TRAINABLE:
        k = int(np.prod(shape)); ref = grads[n].reshape(-1).numpy()
        worst.append((np.abs(got[off:off + k] - ref).max() / max(np.abs(ref).max(), 1e-3 * gmax), n)); off += k
    worst.sort(reverse=True)
    s_err = np.abs(scores.cpu().numpy() - pred.numpy()).max() / np.abs(pred.numpy()).max()
    print(f"{name:12s} scores {s_err:.2e}  worst grads:", ", ".join(f"{n} {e:.2e}" for e, n in worst[:4]))
