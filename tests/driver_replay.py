"""Replays the recorded call sequence of the reference's own drivers against a model (test infrastructure).

``tests/golden/driver_trace.pkl`` (made by ``oracle/make_golden.py --drivers-only``) holds every call that the reference's
UNMODIFIED ``model_trainer.pretrain`` (model_trainer.py:194-236), ``model_trainer.process`` -- training and evaluation
branch (:239-316) -- and ``model_tester.process`` (model_tester.py:173-237) made on the model while they ran over the
reference's own GCNN (fp64, TF stand-in), every ``tape.gradient`` / ``optimizer.apply_gradients`` in between, and what
each of them returned.  The reference tree does not exist on the GPU box and the build container has no GPU, so the
drivers cannot be executed over the CUDA class anywhere; ``replay`` walks the recorded sequence instead, issuing the same
call with the same inputs on a *subject* and comparing every return value with what the reference's model returned:
the ``pretrain`` booleans, the order of the layers ``pretrain_next`` freezes, the predictions of every ``model(...)``
call, the loss and the gradients behind ``tape.gradient``, the state after ``apply_gradients`` (through the following
calls), and -- recomputed from the subject's own predictions -- the mean loss / ranking accuracies the drivers returned.

Two subjects: ``OracleSubject`` (the CPU oracle in fp64; runs in the CPU suite and validates the replayer itself) and
``GcnnSubject`` (gcnn_cut_selector_b200.GCNN through its public, reference-named methods; ``-m gpu``).
"""
import os
import pickle
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import gcnn_oracle as orc  # noqa: E402


def load_trace():
    with open(os.path.join(ROOT, "tests", "golden", "driver_trace.pkl"), "rb") as fh:
        return pickle.load(fh)


def model_inputs(raw):
    """What the drivers hand to the model (model_trainer.py:254-259): seven arrays and the three TOTALS."""
    return tuple(raw[:7]) + (int(raw[7].sum()), int(raw[8].sum()), int(raw[9].sum()))


def layer_number(recorded_name: str) -> int:
    """'gcnn/sequential_3/pre_norm_layer_3' -> 3 ('pre_norm_layer' -> 0): Keras' creation order of the pre-norm layers."""
    tail = recorded_name.rsplit("/", 1)[-1]
    return int(tail.rsplit("_", 1)[-1]) if tail != "pre_norm_layer" else 0


class OracleSubject:
    def __init__(self, state_path):
        self.m = orc.OracleGCNN(orc.restore_state(state_path, dtype=torch.float64), dtype=torch.float64)
        self.adam = orc.AdamState()
        self.last = None

    def pretrain_init(self):
        self.m.pretrain_init()

    def pretrain(self, raw, training):
        return self.m.pretrain(model_inputs(raw), training)

    def pretrain_next(self):
        res = self.m.pretrain_next()
        return None if res is None else orc.PRENORM_LAYERS.index(res[1])

    def call(self, raw, training):
        self.last = raw
        with torch.no_grad():
            return self.m(model_inputs(raw), training).numpy()

    def n_trainable(self):
        return len(orc.TRAINABLE)

    def gradient(self):
        loss, _, grads = orc.loss_and_grads(self.m, model_inputs(self.last), self.last[10])
        self.grads = grads
        return float(loss), torch.cat([grads[n].reshape(-1) for n, _ in orc.TRAINABLE]).numpy()

    def apply(self, lr, iteration):
        orc.adam_step(self.m, self.adam, self.grads, lr)
        assert self.adam.step == iteration

    def accuracy(self, pred, raw, fractions):
        return orc.ranking_accuracy(pred, raw[10], raw[9], fractions)[0]

    def deviations(self, pred, raw):
        return orc.ranking_accuracy(pred, raw[10], raw[9], [1.0])[1]

    def final_trainable(self):
        return orc.flatten_trainable(self.m.params).numpy()


class GcnnSubject:
    """gcnn_cut_selector_b200.GCNN, used the way the reference's drivers use the reference's GCNN."""

    def __init__(self, state_path, device="cuda:0"):
        from gcnn_cut_selector_b200 import GCNN
        self.m = GCNN(device=device, seed=0)
        self.m.restore_state(state_path)
        self.pred = None

    def pretrain_init(self):
        self.m.pretrain_init()

    def pretrain(self, raw, training):
        return self.m.pretrain(model_inputs(raw), training)

    def pretrain_next(self):
        res = self.m.pretrain_next()
        return None if res is None else self.m._prenorm_layers.index(res[0])

    def call(self, raw, training):
        self.last = raw
        if training:  # inside `with tf.GradientTape()` (model_trainer.py:269-271): the autograd bridge keeps the tape
            self.pred = self.m(model_inputs(raw), True)
            return self.pred.detach().cpu().numpy()
        with torch.no_grad():
            return self.m(model_inputs(raw), False).cpu().numpy()

    def n_trainable(self):
        return len(self.m.trainable_variables)

    def gradient(self):
        # loss_fn(improvements, predictions) + tape.gradient(loss, model.trainable_variables) (model_trainer.py:271-272)
        target = torch.as_tensor(self.last[10], dtype=torch.float32, device=self.pred.device)
        loss = ((self.pred - target) ** 2).mean()
        (flat,) = torch.autograd.grad(loss, self.m.flat_params)
        assert torch.equal(flat, self.m.flat_grads)  # the views `trainable_gradients` hands out
        return float(loss.detach()), flat.cpu().numpy()

    def apply(self, lr, iteration):
        self.m.apply_gradients(lr)  # optimizer.apply_gradients(zip(grads, model.trainable_variables)) (:273)
        assert self.m.adam_step == iteration

    def accuracy(self, pred, raw, fractions):
        from gcnn_cut_selector_b200 import metrics
        return metrics.ranking_accuracy(torch.as_tensor(pred, device=self.m.device), raw[10], raw[9], fractions)

    def deviations(self, pred, raw):
        from gcnn_cut_selector_b200 import metrics
        return metrics.ranking_deviation(torch.as_tensor(pred, device=self.m.device), raw[10], raw[9]).cpu().numpy()

    def final_trainable(self):
        return self.m.flat_params.detach().cpu().numpy()


def rel_err(got, want):
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    return float(np.abs(got - want).max() / max(np.abs(want).max(), 1e-30))


def replay(trace, subject, tol_scores, tol_grads, tol_trained):
    """Returns a report dict; raises AssertionError at the first event whose result differs.  ``tol_scores`` /
    ``tol_grads``: before any parameter update (max-abs-relative for predictions, relative L2 per parameter array for
    gradients); ``tol_trained``: for everything after the first ``apply_gradients`` (fp32 Adam steps compound)."""
    batches, events, fractions = trace["batches"], trace["events"], trace["fractions"]
    offsets, off = [], 0
    for _name, shape in orc.TRAINABLE:
        k = int(np.prod(shape))
        offsets.append((off, k))
        off += k
    report = {"events": 0, "max_score_err": 0.0, "max_grad_err": 0.0, "max_loss_err": 0.0, "phases": {}}
    updated = False
    next_layer = 0
    phase_calls = []  # (batch index, predictions) of the model calls since the last phase marker
    for i, e in enumerate(events):
        op, where = e["op"], f"event {i} ({e['op']})"
        if op == "pretrain_init":
            subject.pretrain_init()
        elif op == "pretrain":
            assert subject.pretrain(batches[e["batch"]], e["training"]) == e["ret"], where
        elif op == "pretrain_next":
            got = subject.pretrain_next()
            if e["ret"] is None:
                assert got is None, where
            else:
                assert got == layer_number(e["ret"]) == next_layer, f"{where}: froze layer {got}, reference {e['ret']}"
                next_layer += 1
        elif op == "call":
            raw = batches[e["batch"]]
            assert list(model_inputs(raw)[7:]) == e["totals"], where
            out = subject.call(raw, e["training"])
            err = rel_err(out, e["out"])
            report["max_score_err"] = max(report["max_score_err"], err)
            assert err <= (tol_trained if updated else tol_scores), f"{where}: predictions differ by {err:.3e}"
            phase_calls.append((e["batch"], out))
        elif op == "trainable_variables":
            assert subject.n_trainable() == 46, where
        elif op == "tape_gradient":
            assert e["n_sources"] == 46
            loss, flat = subject.gradient()
            lerr = abs(loss - e["target"]) / abs(e["target"])
            report["max_loss_err"] = max(report["max_loss_err"], lerr)
            tol = tol_trained if updated else tol_grads
            assert lerr <= tol, f"{where}: loss {loss} vs {e['target']}"
            norms = np.array([np.linalg.norm(flat[o:o + k].astype(np.float64)) for o, k in offsets])
            scale = np.maximum(e["norms"], 1e-30)
            nerr = float((np.abs(norms - e["norms"]) / scale).max())
            assert nerr <= 10 * tol, f"{where}: gradient norms differ by {nerr:.3e}"
            if "flat" in e:
                for (o, k), (name, _) in zip(offsets, orc.TRAINABLE):
                    want = e["flat"][o:o + k].astype(np.float64)
                    gerr = float(np.linalg.norm(flat[o:o + k] - want) / max(np.linalg.norm(want), 1e-30))
                    report["max_grad_err"] = max(report["max_grad_err"], gerr)
                    assert gerr <= tol, f"{where}: {name}: relative L2 error {gerr:.3e}"
        elif op == "apply_gradients":
            assert e["n_vars"] == 46
            subject.apply(e["lr"], e["iteration"])
            updated = True
        elif op == "phase":
            if e["name"] == "pretrain":
                assert next_layer == e["result"] == 11, where
            else:
                # the drivers' own bookkeeping (model_trainer.py:303-316, model_tester.py:199-237), recomputed from the
                # SUBJECT's predictions: cut-weighted mean loss, and the ranking accuracy through the subject's metric
                cuts = sum(int(batches[b][9].sum()) for b, _ in phase_calls)
                samples = sum(len(batches[b][9]) for b, _ in phase_calls)
                loss = sum(float(np.mean((np.asarray(p, np.float64) - batches[b][10]) ** 2)) * int(batches[b][9].sum())
                           for b, p in phase_calls) / cuts
                want_loss, want_acc = e["result"]
                assert abs(loss - want_loss) <= tol_trained * abs(want_loss), f"{where}: mean loss {loss} vs {want_loss}"
                if e["name"] == "test":  # model_tester: mean over samples of deviation / n_cuts
                    acc = sum(float(np.sum(subject.deviations(p, batches[b]) / batches[b][9])) for b, p in phase_calls) / samples
                    assert abs(acc - want_acc) <= 1e-12, f"{where}: accuracy {acc} vs {want_acc}"
                else:
                    acc = sum(np.asarray(subject.accuracy(p, batches[b], fractions), np.float64) for b, p in phase_calls) / samples
                    np.testing.assert_allclose(acc, want_acc, rtol=0, atol=1e-12, err_msg=where)
                report["phases"][e["name"]] = (loss, np.asarray(acc).tolist())
            phase_calls = []
        else:
            raise AssertionError(f"{where}: the drivers touched model.{e.get('name')} -- not part of the replayed interface")
        report["events"] += 1
    final = subject.final_trainable()
    report["final_param_err"] = float(np.abs(final - trace["final_trainable"]).max())
    return report
