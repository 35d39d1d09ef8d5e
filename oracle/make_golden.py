"""Generate tests/golden/*.npz by executing the REFERENCE'S OWN SOURCES (test infrastructure only).

Run in the build container (needs /root/reference, which does not exist on the GPU box):

    python oracle/make_golden.py

``/root/reference/model.py`` and ``utils.py`` are imported unmodified; TensorFlow / Keras / pyscipopt (absent here)
are replaced by the torch-backed stand-ins under ``oracle/tf_shim``.  What the fixtures therefore pin is everything
the reference's Python source decides -- model wiring (which tensor feeds which convolution, receiving side, concat
order), variable creation order and shapes (the ``save_state`` stream), the pre-norm pretraining protocol and
Chan-merge arithmetic, and ``load_batch``'s offset arithmetic and casts.  The arithmetic *inside* each TF op is the
shim's restatement of TF semantics (see oracle/gcnn_oracle.py header: parity against TensorFlow itself is unpinned).

Outputs (all small):
  batch_tiny.npz       3 'tiny' samples -> reference ``load_batch`` outputs (bit-exact integer contract)
  fwd_<case>.npz       inputs, fp64 scores / loss / flat gradient (stored as fp32), fp32-run scores
  pretrain_tiny.npz    batches + pre-norm parameters after the reference's pretraining loop
  metric_process.npz   preset predictions / targets -> mean loss and ranking accuracy from the reference's own
                       ``model_trainer.process`` (evaluation branch)
  selector.npz         preset qualities / parallelism matrices -> final cut order and nselectedcuts from the reference's own
                       ``CustomCutsel.cutselselect`` (model_benchmarker.py:70-157)
  driver_trace.pkl     the call sequence of the reference's own drivers -- ``model_trainer.pretrain``,
                       ``model_trainer.process`` (training and evaluation branch), ``model_tester.process`` -- run
                       unmodified over the reference's own GCNN (fp64): every call on the model, every tape.gradient /
                       apply_gradients, and what each returned (``--drivers-only``; replayed by tests/driver_replay.py)
  state_stream.pkl     the weights every fixture uses, written by the reference's own ``save_state`` (62 arrays)
"""
import gzip
import os
import pickle
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REFERENCE = os.environ.get("GCNN_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(HERE, "tf_shim"))
sys.path.insert(0, REFERENCE)
sys.path.insert(0, ROOT)

import tensorflow as tf  # noqa: E402  (the shim)
from tensorflow import keras  # noqa: E402

import gcnn_oracle as orc  # noqa: E402
from gcnn_cut_selector_b200 import synth  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def reference_model(params: dict, dtype):
    """Instantiate the reference GCNN over the shim and load ``params`` positionally (like restore_state)."""
    tf.set_float_dtype(dtype)
    keras.reset_name_counters()
    import model as ref_model
    m = ref_model.GCNN()
    variables = m.variables
    assert len(variables) == len(orc.PARAM_SPECS), (len(variables), len(orc.PARAM_SPECS))
    for v, (name, shape, trainable) in zip(variables, orc.PARAM_SPECS):
        assert tuple(v.shape) == tuple(shape) and v.trainable == trainable, (v.name, v.shape, name, shape)
        v.assign(params[name].to(dtype))
    assert [v.name for v in variables] == m.variables_topological_order
    return m, ref_model


def write_sample_files(samples, folder):
    files = []
    for i, (state, imp) in enumerate(samples):
        path = os.path.join(folder, f"sample_{i}.pkl")
        with gzip.open(path, "wb") as fh:
            pickle.dump({"data": [state, imp]}, fh)
        files.append(path)
    return files


def run_reference_batch(samples):
    import utils as ref_utils
    tf.set_float_dtype(torch.float32)
    with tempfile.TemporaryDirectory() as tmp:
        files = write_sample_files(samples, tmp)
        out = ref_utils.load_batch(files)
    return [o.numpy() for o in out]


def fwd_case(name, samples, params):
    batch = run_reference_batch(samples)
    inputs = tuple(batch[:7]) + (int(batch[7].sum()), int(batch[8].sum()), int(batch[9].sum()))
    targets = batch[10]
    res = {}
    for tag, dtype in (("f64", torch.float64), ("f32", torch.float32)):
        m, _ = reference_model(params, dtype)
        tin = tuple(tf.convert_to_tensor(x) if isinstance(x, np.ndarray) else tf.convert_to_tensor(np.int32(x))
                    for x in inputs)
        pred = m(tin, tf.convert_to_tensor(True))
        y = torch.as_tensor(targets).to(dtype)
        loss = ((y - pred) ** 2).mean()  # Keras MeanSquaredError on 1-D tensors (model_trainer.py:271)
        tv = m.trainable_variables
        grads = torch.autograd.grad(loss, [v._t for v in tv])
        res[tag] = (pred.detach().numpy(), float(loss), torch.cat([g.reshape(-1) for g in grads]).numpy())
    np.savez_compressed(
        os.path.join(GOLDEN, f"fwd_{name}.npz"),
        cons=batch[0], cons_ei=batch[1], cons_ef=batch[2], var=batch[3], cut=batch[4], cut_ei=batch[5],
        cut_ef=batch[6], n_cons=batch[7], n_vars=batch[8], n_cuts=batch[9], targets=targets,
        scores_f64=res["f64"][0], loss_f64=res["f64"][1], grad_f64_as_f32=res["f64"][2].astype(np.float32),
        scores_f32=res["f32"][0].astype(np.float32))
    print(f"fwd_{name}: {len(samples)} samples, N_k={batch[4].shape[0]}, loss={res['f64'][1]:.6e}, "
          f"f32-vs-f64 score dev={np.abs(res['f32'][0] - res['f64'][0]).max():.2e}")


def pretrain_case(params):
    """The reference's own pretrain loop (model_trainer.py:207-234) over two tiny batches, in fp64 and fp32."""
    batches = [run_reference_batch(synth.make_samples("tiny", 3, seed0=100 + 10 * b)) for b in range(2)]
    out = {}
    for tag, dtype in (("f64", torch.float64), ("f32", torch.float32)):
        m, ref_model = reference_model(params, dtype)
        m.pretrain_init()
        n_layers, order = 0, []
        torch.set_grad_enabled(False)  # TF eager tensors carry no tape here; keeps np.sqrt(tensor) legal
        while True:
            for batch in batches:
                inputs = tuple(tf.convert_to_tensor(x) for x in batch[:7]) + tuple(
                    tf.convert_to_tensor(np.int32(batch[i].sum())) for i in (7, 8, 9))
                if not m.pretrain(inputs, tf.convert_to_tensor(True)):
                    break
            res = m.pretrain_next()
            if res is None:
                break
            order.append(res[1])
            n_layers += 1
        torch.set_grad_enabled(True)
        assert n_layers == 11, n_layers
        non_trainable = [v for v in m.variables if not v.trainable]
        out[tag] = np.concatenate([v.numpy().reshape(-1) for v in non_trainable])
        pin = tuple(tf.convert_to_tensor(x) for x in batches[0][:7]) + tuple(
            tf.convert_to_tensor(np.int32(batches[0][i].sum())) for i in (7, 8, 9))
        out[tag + "_scores"] = m(pin, tf.convert_to_tensor(False)).detach().numpy()
    flat = {}
    for b, batch in enumerate(batches):
        for i, key in enumerate(["cons", "cons_ei", "cons_ef", "var", "cut", "cut_ei", "cut_ef", "n_cons", "n_vars",
                                 "n_cuts", "targets"]):
            flat[f"b{b}_{key}"] = batch[i]
    np.savez_compressed(os.path.join(GOLDEN, "pretrain_tiny.npz"), prenorm_f64=out["f64"],
                        prenorm_f32=out["f32"].astype(np.float32), scores_f64=out["f64_scores"],
                        layer_order=np.asarray(order), **flat)
    print("pretrain_tiny: 11 layers; order:", order)


def metric_case():
    """Ranking accuracy + mean loss of the reference's own ``process`` (model_trainer.py:239-316) on preset predictions
    (ties included), evaluation branch.  The model is a stand-in that returns the preset predictions of each batch."""
    import model_trainer as ref_trainer
    rng = np.random.default_rng(77)
    fractions = np.array([0.25, 0.5, 0.75, 1])
    batches, preds = [], []
    for b in range(3):
        n_cuts = rng.integers(1, 12, size=4 + b).astype(np.int32)
        total = int(n_cuts.sum())
        true = np.round(rng.uniform(0, 0.1, total), 2).astype(np.float32)            # coarse grid -> ties
        noise = rng.normal(0, 0.02, total) * (rng.random(total) < 0.5)
        pred = np.round(true + noise, 2).astype(np.float32)
        preds.append(pred)
        dummy = tf.convert_to_tensor(np.zeros((1, 1), np.float32))
        n = tf.convert_to_tensor(n_cuts)
        batches.append((dummy, dummy, dummy, dummy, dummy, dummy, dummy, n, n, n, tf.convert_to_tensor(true)))

    class Preset:
        def __init__(self):
            self.i = 0

        def __call__(self, batched_states, training):
            out = tf.convert_to_tensor(preds[self.i])
            self.i += 1
            return out

    mean_loss, mean_acc = ref_trainer.process(Preset(), batches, fractions, ref_trainer.MeanSquaredError())
    out = {"fractions": fractions, "mean_loss": np.float64(mean_loss), "mean_acc": np.asarray(mean_acc, np.float64),
           "n_batches": np.int64(len(batches))}
    for b, (batch, pred) in enumerate(zip(batches, preds)):
        out[f"pred{b}"] = pred
        out[f"true{b}"] = batch[10].numpy()
        out[f"n_cuts{b}"] = batch[9].numpy()
    np.savez_compressed(os.path.join(GOLDEN, "metric_process.npz"), **out)
    print("metric_process: mean_loss", mean_loss, "mean_acc", mean_acc)


def selector_case():
    """The reference's own ``CustomCutsel.cutselselect`` (model_benchmarker.py:70-157, hybrid branch: no GCNN, so the
    quality comes from the stand-in SCIP model) on preset qualities and parallelism matrices: the final order of the cuts
    and ``nselectedcuts``.  float32-representable values, so the device kernel sees exactly these numbers."""
    import model_benchmarker as ref_bench
    rng = np.random.default_rng(2024)
    cases = {}

    class Cut:
        def __init__(self, i):
            self.i = i

        def getNNonz(self):
            return 1

    class FakeScip:
        def __init__(self, q, par, par_forced):
            self.q, self.par, self.par_forced = q, par, par_forced

        def getCutEfficacy(self, cut):
            return float(self.q[cut.i])

        def getRowNumIntCols(self, cut):
            return 0

        def getRowObjParallelism(self, cut):
            return 0.0

        def getRowParallelism(self, a, b):
            if a.i < 0:  # forced cuts carry negative ids
                return float(self.par_forced[-a.i - 1][b.i])
            return float(self.par[a.i][b.i])

    specs = [("small", 12, 2, 5), ("ties", 40, 0, 40), ("forced", 64, 5, 30), ("dense", 150, 3, 150), ("one", 1, 1, 1),
             ("none_removed", 30, 2, 10)]
    for name, n, nf, max_sel in specs:
        q = np.round(rng.uniform(0.0, 1.0, n), 2 if name == "ties" else 6).astype(np.float32)
        hi = 0.05 if name == "none_removed" else 1.0
        par = rng.uniform(0.0, hi, (n, n)).astype(np.float32)
        par = np.maximum(par, par.T)
        mask = rng.random((n, n)) < (0.6 if name != "dense" else 0.1)
        par = np.where(np.maximum(mask, mask.T), np.float32(0.0), par)  # most pairs are orthogonal
        par_forced = rng.uniform(0.0, hi, (nf, n)).astype(np.float32)
        par_forced[rng.random((nf, n)) < 0.7] = 0.0
        sel = ref_bench.CustomCutsel()
        sel.model = FakeScip(q, par, par_forced)
        cuts = [Cut(i) for i in range(n)]
        forced = [Cut(-f - 1) for f in range(nf)]
        res = sel.cutselselect(cuts, forced, True, max_sel)
        cases[f"{name}_quality"] = q
        cases[f"{name}_par"] = par
        cases[f"{name}_par_forced"] = par_forced
        cases[f"{name}_max_selected"] = np.int64(max_sel)
        cases[f"{name}_order"] = np.array([c.i for c in res["cuts"]], dtype=np.int32)
        cases[f"{name}_n_selected"] = np.int64(res["nselectedcuts"])
        print(f"selector {name}: n={n} forced={nf} kept={res['nselectedcuts']}")
    cases["names"] = np.array([s[0] for s in specs])
    np.savez_compressed(os.path.join(GOLDEN, "selector.npz"), **cases)


def driver_case(params):
    """The reference's OWN drivers -- model_trainer.pretrain (model_trainer.py:194-236), model_trainer.process in its
    training and evaluation branches (:239-316) and model_tester.process (model_tester.py:173-237) -- run UNMODIFIED
    over the reference's own GCNN (fp64, TF stand-in) behind a recording proxy.  The trace lists, in order, every call
    the drivers make on the model (``pretrain_init`` / ``pretrain`` / ``pretrain_next`` / ``__call__`` and any other
    attribute they touch), every ``tape.gradient`` and ``optimizer.apply_gradients`` in between, and what each returned.
    tests/test_driver_trace.py replays exactly this call sequence against gcnn_cut_selector_b200.GCNN on the GPU (the
    reference tree does not exist on the GPU box and this container has no GPU, so the drivers themselves cannot be
    executed over the CUDA class anywhere; the recorded sequence is what travels)."""
    import model_tester as ref_tester
    import model_trainer as ref_trainer
    tf.set_float_dtype(torch.float64)
    raw = [run_reference_batch(synth.make_samples(shape, n, seed0=seed))
           for shape, n, seed in (("tiny", 3, 900), ("mini", 2, 910), ("tiny", 4, 920))]
    tf.set_float_dtype(torch.float64)
    batches = [tuple(tf.convert_to_tensor(x) for x in b) for b in raw]
    m, _ = reference_model(params, torch.float64)
    events = []
    index_of = {id(b[0]): i for i, b in enumerate(batches)}

    class Traced:
        def __call__(self, batched_states, training):
            out = m(batched_states, training)
            events.append({"op": "call", "batch": index_of[id(batched_states[0])], "training": builtins_bool(training),
                           "totals": [int(x) for x in batched_states[7:10]], "out": out.detach().numpy().copy()})
            return out

        def pretrain_init(self):
            events.append({"op": "pretrain_init"})
            return m.pretrain_init()

        def pretrain(self, batched_states, training):
            ret = m.pretrain(batched_states, training)
            events.append({"op": "pretrain", "batch": index_of[id(batched_states[0])], "training": builtins_bool(training),
                           "ret": builtins_bool(ret)})
            return ret

        def pretrain_next(self):
            res = m.pretrain_next()
            events.append({"op": "pretrain_next", "ret": None if res is None else str(res[1])})
            return res

        @property
        def trainable_variables(self):
            events.append({"op": "trainable_variables"})
            return m.trainable_variables

        def __getattr__(self, name):  # anything else the drivers reach for shows up in the trace
            events.append({"op": "getattr", "name": name})
            return getattr(m, name)

    traced = Traced()
    fractions = np.array([0.25, 0.5, 0.75, 1])
    lr = 1e-4  # model_trainer.py:53, the reference's default
    tf.trace = events
    torch.set_grad_enabled(False)  # TF eager tensors carry no tape outside tf.GradientTape
    n_layers = ref_trainer.pretrain(traced, batches[:2])
    events.append({"op": "phase", "name": "pretrain", "result": int(n_layers)})
    optimizer = ref_trainer.Adam(learning_rate=lambda: lr)
    loss, acc = ref_trainer.process(traced, batches, fractions, ref_trainer.MeanSquaredError(), optimizer)
    events.append({"op": "phase", "name": "train", "result": (float(loss), np.asarray(acc, np.float64))})
    loss, acc = ref_trainer.process(traced, batches[::-1], fractions, ref_trainer.MeanSquaredError())
    events.append({"op": "phase", "name": "valid", "result": (float(loss), np.asarray(acc, np.float64))})
    loss, acc = ref_tester.process(traced, batches)
    events.append({"op": "phase", "name": "test", "result": (float(loss), float(acc))})
    torch.set_grad_enabled(True)
    tf.trace = None
    assert n_layers == 11
    # gradients: the whole flat vector of the first step, per-array L2 norms of the later ones (fixture size)
    first = True
    for e in events:
        if e["op"] == "tape_gradient":
            flat = np.concatenate([g.reshape(-1) for g in e.pop("grads")])
            e["norms"] = np.array([np.linalg.norm(flat[o:o + k]) for o, k in _offsets()], np.float64)
            if first:
                e["flat"] = flat.astype(np.float32)
                first = False
    final = np.concatenate([v.numpy().reshape(-1) for v in m.variables if v.trainable]).astype(np.float32)
    prenorm = np.concatenate([v.numpy().reshape(-1) for v in m.variables if not v.trainable])
    with open(os.path.join(GOLDEN, "driver_trace.pkl"), "wb") as fh:
        pickle.dump({"batches": raw, "events": events, "fractions": fractions, "lr": lr, "final_trainable": final,
                     "final_prenorm": prenorm.astype(np.float64)}, fh)
    ops = [e["op"] for e in events]
    print("driver_trace:", len(events), "events;", {o: ops.count(o) for o in dict.fromkeys(ops)})
    print("  phases:", [(e["name"], e["result"]) for e in events if e["op"] == "phase"])


def builtins_bool(x):
    return bool(x.item()) if torch.is_tensor(x) else bool(x)


def _offsets():
    off = 0
    for _name, shape in orc.TRAINABLE:
        k = int(np.prod(shape))
        yield off, k
        off += k


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    if "--drivers-only" in sys.argv:
        driver_case(orc.restore_state(os.path.join(GOLDEN, "state_stream.pkl"), dtype=torch.float32))
        return
    if "--selector-only" in sys.argv:
        selector_case()
        return
    if "--metric-only" in sys.argv:
        metric_case()
        return
    params = orc.init_params(seed=12345, dtype=torch.float32)  # weights live in fp32, like the reference's
    # the weights fixture IS the save_state stream written by the reference's own method (model.py:47-56)
    m, _ = reference_model(params, torch.float32)
    m.save_state(os.path.join(GOLDEN, "state_stream.pkl"))

    # batching contract: raw samples + what reference load_batch makes of them
    samples = synth.make_samples("tiny", 3, seed0=0)
    samples[1] = synth.shuffle_edges(samples[1], 5)
    batch = run_reference_batch(samples)
    with open(os.path.join(GOLDEN, "batch_tiny_samples.pkl"), "wb") as fh:
        pickle.dump(samples, fh)
    np.savez_compressed(os.path.join(GOLDEN, "batch_tiny.npz"), **{f"out{i}": o for i, o in enumerate(batch)})
    print("batch_tiny dtypes:", [str(o.dtype) for o in batch])

    fwd_case("tiny3", samples, params)
    fwd_case("mini2", synth.make_samples("mini", 2, seed0=40), params)
    iso = synth.make_samples("tiny", 2, seed0=60)
    # a sample with an edge-less cut side and isolated nodes: scatter must leave zero rows (model.py:568-569)
    (c, ce, v, k, ke), imp = iso[0]
    keep = ke["indices"][0] != 1
    iso[0] = ((c, ce, v, k, {"indices": ke["indices"][:, keep], "values": ke["values"][keep]}), imp)
    fwd_case("isolated", iso, params)
    pretrain_case(params)
    metric_case()
    selector_case()
    driver_case(params)


if __name__ == "__main__":
    main()
