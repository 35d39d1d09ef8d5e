"""CPU: pin the oracle restatement against fixtures produced by executing the reference's own model.py / utils.py
(oracle/make_golden.py).  Covers wiring, variable order (save_state stream), batching offsets/casts and the pre-norm
pretraining protocol.  Tolerances: integer work bit-exact; fp64 scores/gradients 1e-9 relative to max-abs."""
import os
import pickle

import numpy as np
import pytest
import torch

import gcnn_oracle as orc

CASES = ["tiny3", "mini2", "isolated"]


def _params(golden_dir, dtype):
    return orc.restore_state(os.path.join(golden_dir, "state_stream.pkl"), dtype=dtype)


def _inputs(z, prefix=""):
    g = lambda k: z[prefix + k]
    return (g("cons"), g("cons_ei"), g("cons_ef"), g("var"), g("cut"), g("cut_ei"), g("cut_ef"),
            int(g("n_cons").sum()), int(g("n_vars").sum()), int(g("n_cuts").sum()))


def test_state_stream_order_and_shapes(golden_dir):
    with open(os.path.join(golden_dir, "state_stream.pkl"), "rb") as fh:
        arrays = [pickle.load(fh) for _ in orc.PARAM_SPECS]
        with pytest.raises(EOFError):
            pickle.load(fh)
    assert [a.shape for a in arrays] == [tuple(s) for _, s, _ in orc.PARAM_SPECS]
    assert all(a.dtype == np.float32 for a in arrays)
    # same weights the oracle's own initialiser makes (seeded), i.e. stream order == PARAM_SPECS order
    mine = orc.init_params(seed=12345, dtype=torch.float32)
    for a, (name, _, _) in zip(arrays, orc.PARAM_SPECS):
        np.testing.assert_array_equal(a, mine[name].numpy())


def test_save_restore_roundtrip(golden_dir, tmp_path):
    p = _params(golden_dir, torch.float32)
    orc.save_state(p, tmp_path / "s.pkl")
    assert open(tmp_path / "s.pkl", "rb").read() == open(os.path.join(golden_dir, "state_stream.pkl"), "rb").read()


def test_batching_bit_exact(golden_dir, tmp_path):
    with open(os.path.join(golden_dir, "batch_tiny_samples.pkl"), "rb") as fh:
        samples = pickle.load(fh)
    want = np.load(os.path.join(golden_dir, "batch_tiny.npz"))
    got = orc.concat_samples(samples)
    assert len(got) == 11
    for i, g in enumerate(got):
        w = want[f"out{i}"]
        assert g.dtype == w.dtype and g.shape == w.shape
        np.testing.assert_array_equal(g, w)
    # and through gzip-pickled files, like utils.load_batch
    import gzip
    files = []
    for i, (state, imp) in enumerate(samples):
        path = str(tmp_path / f"sample_{i}.pkl")
        with gzip.open(path, "wb") as fh:
            pickle.dump({"data": [state, imp]}, fh)
        files.append(path)
    for i, g in enumerate(orc.load_batch(files)):
        np.testing.assert_array_equal(g, want[f"out{i}"])


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("faithful", [True, False])
def test_forward_and_grads_fp64(golden_dir, case, faithful):
    z = np.load(os.path.join(golden_dir, f"fwd_{case}.npz"))
    model = orc.OracleGCNN(_params(golden_dir, torch.float64), dtype=torch.float64, faithful=faithful)
    loss, pred, grads = orc.loss_and_grads(model, _inputs(z), z["targets"])
    tol = 1e-9 if faithful else 1e-7  # the hoisted form re-associates the per-edge Dense
    ref = z["scores_f64"]
    assert np.abs(pred.numpy() - ref).max() <= tol * np.abs(ref).max()
    assert abs(float(loss) - float(z["loss_f64"])) <= tol * float(z["loss_f64"])
    flat = torch.cat([grads[n].reshape(-1) for n, _ in orc.TRAINABLE]).numpy()
    gref = z["grad_f64_as_f32"].astype(np.float64)  # stored rounded to fp32: 6e-8 relative
    o = 0
    for name, shape in orc.TRAINABLE:
        k = int(np.prod(shape))
        scale = max(np.abs(gref[o:o + k]).max(), 1e-30)
        assert np.abs(flat[o:o + k] - gref[o:o + k]).max() <= 2e-7 * scale + 1e-12, name
        o += k


@pytest.mark.parametrize("case", CASES)
def test_forward_fp32_within_tolerance(golden_dir, case):
    """The fp32 oracle (TF-CPU stand-in) agrees with fp64 truth to the north-star tolerance (1e-5 relative)."""
    z = np.load(os.path.join(golden_dir, f"fwd_{case}.npz"))
    for faithful in (True, False):
        model = orc.OracleGCNN(_params(golden_dir, torch.float32), dtype=torch.float32, faithful=faithful)
        pred = model.call(_inputs(z)).numpy()
        assert np.abs(pred - z["scores_f64"]).max() <= 1e-5 * np.abs(z["scores_f64"]).max()


def test_isolated_nodes_get_zero_conv_rows(golden_dir):
    z = np.load(os.path.join(golden_dir, "fwd_isolated.npz"))
    model = orc.OracleGCNN(_params(golden_dir, torch.float64), dtype=torch.float64)
    model.call(_inputs(z))
    cut_rows = np.unique(z["cut_ei"][0])
    missing = np.setdiff1d(np.arange(int(z["n_cuts"].sum())), cut_rows)
    assert missing.size > 0
    assert torch.all(model.trace["cut_conv/conv"][missing] == 0)


@pytest.mark.parametrize("tag,dtype,tol", [("f64", torch.float64, 1e-10), ("f32", torch.float32, 2e-5)])
def test_pretrain_protocol(golden_dir, tag, dtype, tol):
    z = np.load(os.path.join(golden_dir, "pretrain_tiny.npz"))
    model = orc.OracleGCNN(_params(golden_dir, dtype), dtype=dtype)
    batches = [_inputs(z, f"b{b}_") for b in range(2)]
    model.pretrain_init()
    n = 0
    while True:
        for b in batches:
            if not model.pretrain(b, True):
                break
        if model.pretrain_next() is None:
            break
        n += 1
    assert n == 11 == len(z["layer_order"])
    got = orc.flatten_prenorm(model.params).numpy().astype(np.float64)
    want = z["prenorm_" + tag].astype(np.float64)
    assert np.abs(got - want).max() <= tol * np.abs(want).max()
    if tag == "f64":
        pred = model.call(batches[0]).numpy()
        assert np.abs(pred - z["scores_f64"]).max() <= 1e-9 * np.abs(z["scores_f64"]).max()


def test_adam_matches_closed_form():
    """Keras Adam (epsilon 1e-7 outside the bias correction) on a scalar, two steps, by hand."""
    model = orc.OracleGCNN(dtype=torch.float64)
    name = "out_2/bias"
    model.params[name] = torch.tensor([0.5], dtype=torch.float64)
    st = orc.AdamState()
    g1, g2, lr, b1, b2, eps = 0.3, -0.2, 1e-2, 0.9, 0.999, 1e-7
    orc.adam_step(model, st, {name: torch.tensor([g1], dtype=torch.float64)}, lr)
    m1, v1 = (1 - b1) * g1, (1 - b2) * g1 * g1
    th1 = 0.5 - lr * np.sqrt(1 - b2) / (1 - b1) * m1 / (np.sqrt(v1) + eps)
    assert abs(float(model.params[name]) - th1) < 1e-15
    orc.adam_step(model, st, {name: torch.tensor([g2], dtype=torch.float64)}, lr)
    m2, v2 = b1 * m1 + (1 - b1) * g2, b2 * v1 + (1 - b2) * g2 * g2
    th2 = th1 - lr * np.sqrt(1 - b2 ** 2) / (1 - b1 ** 2) * m2 / (np.sqrt(v2) + eps)
    assert abs(float(model.params[name]) - th2) < 1e-15


def test_gradcheck_fp64_micrograph():
    """Autograd of the oracle vs central differences on a hand-sized graph (2 cons x 3 vars x 1 cut)."""
    p = orc.init_params(seed=3, dtype=torch.float64)
    model = orc.OracleGCNN(p, dtype=torch.float64)
    rng = np.random.default_rng(0)
    inputs = (rng.standard_normal((2, 4)), np.array([[0, 0, 1], [0, 2, 1]], np.int32), rng.standard_normal((3, 1)),
              rng.standard_normal((3, 14)), rng.standard_normal((1, 6)), np.array([[0, 0], [1, 2]], np.int32),
              rng.standard_normal((2, 1)), 2, 3, 1)
    y = np.array([0.05])
    loss, _, grads = orc.loss_and_grads(model, inputs, y)
    for name in ["cons_conv_feat_edge/kernel", "var_conv_feat_final/bias", "cut_emb_1/kernel", "out_2/kernel"]:
        w = model.params[name]
        idx = tuple(int(i) for i in np.unravel_index(int(torch.argmax(grads[name].abs())), w.shape))
        h = 1e-6
        w[idx] += h
        lp = orc.loss_and_grads(model, inputs, y)[0]
        w[idx] -= 2 * h
        lm = orc.loss_and_grads(model, inputs, y)[0]
        w[idx] += h
        fd = float(lp - lm) / (2 * h)
        assert abs(fd - float(grads[name][idx])) <= 1e-5 * max(abs(fd), 1e-8), name


def test_ranking_accuracy_matches_reference_process(golden_dir):
    """oracle.ranking_accuracy against the mean loss / accuracy the reference's own model_trainer.process returned on
    preset predictions with ties (fixture generated by oracle/make_golden.py from the unmodified reference source)."""
    z = np.load(os.path.join(golden_dir, "metric_process.npz"))
    acc, n_samples, loss_num, cuts = np.zeros(len(z["fractions"])), 0, 0.0, 0
    for b in range(int(z["n_batches"])):
        pred, true, n_cuts = z[f"pred{b}"], z[f"true{b}"], z[f"n_cuts{b}"]
        a, dev = orc.ranking_accuracy(pred, true, n_cuts, z["fractions"])
        assert dev.shape == n_cuts.shape and (dev <= n_cuts).all()
        acc += a
        n_samples += len(n_cuts)
        loss_num += float(np.mean((pred.astype(np.float32) - true) ** 2, dtype=np.float32)) * int(n_cuts.sum())
        cuts += int(n_cuts.sum())
    np.testing.assert_allclose(acc / n_samples, z["mean_acc"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(loss_num / cuts, float(z["mean_loss"]), rtol=1e-6)


def test_ranking_accuracy_hand_cases():
    fr = [0.25, 0.5, 0.75, 1]
    # identical order -> deviation = n; ties keep the original order in both rankings (stable sort)
    acc, dev = orc.ranking_accuracy([3, 2, 2, 1], [30, 20, 20, 10], [4], fr)
    assert dev.tolist() == [4] and acc.tolist() == [1, 1, 1, 1]
    # first position already wrong
    acc, dev = orc.ranking_accuracy([1, 2], [2, 1], [2], fr)
    assert dev.tolist() == [0] and acc.tolist() == [0, 0, 0, 0]
    # two samples: the second deviates at position 2 of 4 -> frac 0.5
    acc, dev = orc.ranking_accuracy([5, 4, 4, 3, 2, 1], [1, 0, 9, 8, 6, 7], [2, 4], fr)
    assert dev.tolist() == [2, 2] and acc.tolist() == [2, 2, 1, 1]


def test_select_cuts_matches_reference_cutsel(golden_dir):
    """The oracle's restatement of the ranking + parallelism filter against the reference's own
    ``CustomCutsel.cutselselect`` (tests/golden/selector.npz, generated by oracle/make_golden.py over the SCIP stand-in)."""
    z = np.load(os.path.join(golden_dir, "selector.npz"))
    for name in z["names"]:
        pf = z[f"{name}_par_forced"]
        order, n_sel = orc.select_cuts(z[f"{name}_quality"], z[f"{name}_par"], pf if len(pf) else None,
                                       max_selected=int(z[f"{name}_max_selected"]))
        np.testing.assert_array_equal(order, z[f"{name}_order"])
        assert n_sel == int(z[f"{name}_n_selected"])
