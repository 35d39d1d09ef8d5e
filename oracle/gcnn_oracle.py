"""CPU oracle for the GCNN hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import this module.  Nothing under ``gcnn_cut_selector_b200/`` imports it; the product path is CUDA only.

It restates, op for op and in the reference's op order, what ``/root/reference/model.py`` computes through
TensorFlow 2.7.1 (an un-vendored dependency, ``environment.yml:11``), on torch CPU tensors in fp32 or fp64:

* model wiring                      -- model.py:257-300
* pre-norm apply / statistics       -- model.py:365-382, 394-437
* partial graph convolution         -- model.py:533-575 (project -> gather -> add -> scale -> relu ->
                                       per-edge Dense -> scatter-sum in edge order -> scale -> concat -> MLP)
* parameter inventory and order     -- model.py:164-226, 480-508 (Keras layer-tracking order)
* weights stream                    -- model.py:47-67
* batching contract                 -- utils.py:339-426
* train step (MSE, grads, Adam)     -- model_trainer.py:259-273 (Keras Adam, epsilon 1e-7, TF bias-correction form)

Pinning status: TensorFlow cannot run in the build container, and the reference ships no tests or golden vectors
(SURVEY.md section 4), so the *arithmetic* of the TF ops is restated from TF's published semantics: **parity against
TensorFlow itself is unpinned**.  What IS pinned: ``oracle/make_golden.py`` executes the reference's own, unmodified
``model.py`` and ``utils.py`` sources over a torch-backed stand-in for the handful of TF/Keras symbols they use
(``oracle/tf_shim``) and this oracle is checked against those outputs (wiring, parameter order, batching offsets,
pre-norm protocol) in ``tests/test_oracle_golden.py``.
"""
from __future__ import annotations

import gzip
import pickle
from dataclasses import dataclass, field

import numpy as np
import torch

EMB = 64
CONS_F, EDGE_F, VAR_F, CUT_F = 4, 1, 14, 6


# ----------------------------------------------------------------------------------------------------------------
# Parameter inventory, in the order Keras enumerates ``model.variables`` (model.py:215): layer by layer in
# attribute-assignment order (model.py:174-208), each layer's own weights in creation order.
# ----------------------------------------------------------------------------------------------------------------
def _embedding(prefix: str, n_in: int):
    return [(f"{prefix}/prenorm/shift", (n_in,), False), (f"{prefix}/prenorm/scale", (n_in,), False),
            (f"{prefix}_1/kernel", (n_in, EMB), True), (f"{prefix}_1/bias", (EMB,), True),
            (f"{prefix}_2/kernel", (EMB, EMB), True), (f"{prefix}_2/bias", (EMB,), True)]


def _edge_embedding(prefix: str):
    return [(f"{prefix}/prenorm/shift", (EDGE_F,), False), (f"{prefix}/prenorm/scale", (EDGE_F,), False)]


def _conv(name: str):
    # model.py:486-508
    return [(f"{name}_feat_left/kernel", (EMB, EMB), True), (f"{name}_feat_left/bias", (EMB,), True),
            (f"{name}_feat_edge/kernel", (EDGE_F, EMB), True),
            (f"{name}_feat_right/kernel", (EMB, EMB), True),
            (f"{name}_final/prenorm/scale", (1,), False),
            (f"{name}_feat_final/kernel", (EMB, EMB), True), (f"{name}_feat_final/bias", (EMB,), True),
            (f"{name}_post/prenorm/scale", (1,), False),
            (f"{name}_out_1/kernel", (2 * EMB, EMB), True), (f"{name}_out_1/bias", (EMB,), True),
            (f"{name}_out_2/kernel", (EMB, EMB), True), (f"{name}_out_2/bias", (EMB,), True)]


PARAM_SPECS: list[tuple[str, tuple, bool]] = (
    _embedding("cons_emb", CONS_F) + _edge_embedding("cons_edge") + _embedding("var_emb", VAR_F)
    + _embedding("cut_emb", CUT_F) + _edge_embedding("cut_edge")
    + _conv("cons_conv") + _conv("var_conv") + _conv("cut_conv")
    + [("out_1/kernel", (EMB, EMB), True), ("out_1/bias", (EMB,), True),
       ("out_2/kernel", (EMB, 1), True), ("out_2/bias", (1,), True)])

assert len(PARAM_SPECS) == 62
TRAINABLE = [(n, s) for n, s, t in PARAM_SPECS if t]
NON_TRAINABLE = [(n, s) for n, s, t in PARAM_SPECS if not t]
N_TRAINABLE = sum(int(np.prod(s)) for _, s in TRAINABLE)
assert len(TRAINABLE) == 46 and N_TRAINABLE == 93121 and sum(int(np.prod(s)) for _, s in NON_TRAINABLE) == 58

# Pre-norm layers in the order ``pretrain_next_rec`` walks them (model.py:100-117) == execution order of ``call``.
PRENORM_LAYERS = ["cons_emb/prenorm", "cons_edge/prenorm", "var_emb/prenorm", "cut_emb/prenorm", "cut_edge/prenorm",
                  "cons_conv_final/prenorm", "cons_conv_post/prenorm", "var_conv_final/prenorm",
                  "var_conv_post/prenorm", "cut_conv_final/prenorm", "cut_conv_post/prenorm"]


def init_params(seed: int = 12345, dtype=torch.float64, identity_prenorm: bool = False) -> dict[str, torch.Tensor]:
    """Synthetic weights: orthogonal kernels (model.py:175 'orthogonal'), N(0, 0.1) biases
    (non-zero so the bias paths are exercised), pre-norm shift ~ N(0,1), scale ~ U(0.5, 2)."""
    rng = np.random.default_rng(seed)
    out = {}
    for name, shape, trainable in PARAM_SPECS:
        if name.endswith("kernel"):
            a = rng.standard_normal((max(shape), max(shape)))
            q, r = np.linalg.qr(a)
            q = q * np.sign(np.diag(r))
            w = q[:shape[0], :shape[1]]
        elif name.endswith("bias"):
            w = 0.1 * rng.standard_normal(shape)
        elif name.endswith("shift"):
            w = np.zeros(shape) if identity_prenorm else rng.standard_normal(shape)
        else:
            w = np.ones(shape) if identity_prenorm else rng.uniform(0.5, 2.0, shape)
        out[name] = torch.tensor(np.asarray(w, dtype=np.float64), dtype=dtype)
    return out


def flatten_trainable(params: dict) -> torch.Tensor:
    return torch.cat([params[n].reshape(-1) for n, _ in TRAINABLE])


def flatten_prenorm(params: dict) -> torch.Tensor:
    return torch.cat([params[n].reshape(-1) for n, _ in NON_TRAINABLE])


def unflatten(flat_trainable, flat_prenorm, dtype=None) -> dict:
    out, o = {}, 0
    for n, s in TRAINABLE:
        k = int(np.prod(s))
        out[n] = flat_trainable[o:o + k].reshape(s)
        o += k
    o = 0
    for n, s in NON_TRAINABLE:
        k = int(np.prod(s))
        out[n] = flat_prenorm[o:o + k].reshape(s)
        o += k
    if dtype is not None:
        out = {k: v.to(dtype) for k, v in out.items()}
    return {n: out[n] for n, _, _ in PARAM_SPECS}


# ----------------------------------------------------------------------------------------------------------------
# Weights stream (model.py:47-67): one pickle.dump(ndarray) per variable, in PARAM_SPECS order.
# ----------------------------------------------------------------------------------------------------------------
def save_state(params: dict, path: str):
    with open(path, "wb") as fh:
        for name, _, _ in PARAM_SPECS:
            pickle.dump(params[name].detach().cpu().numpy().astype(np.float32), fh)


def restore_state(path: str, dtype=torch.float32) -> dict:
    out = {}
    with open(path, "rb") as fh:
        for name, shape, _ in PARAM_SPECS:
            arr = np.asarray(pickle.load(fh))
            if tuple(arr.shape) != tuple(shape):
                raise ValueError(f"weights stream mismatch at {name}: {arr.shape} vs {shape}")
            out[name] = torch.tensor(arr, dtype=dtype)
    return out


# ----------------------------------------------------------------------------------------------------------------
# Pre-norm layer (model.py:303-437)
# ----------------------------------------------------------------------------------------------------------------
class PreNormException(Exception):
    """model.py:440"""


@dataclass
class _PreNormState:
    n_units: int
    waiting: bool = False
    received: bool = False
    mean: object = 0.0
    var: object = 0.0
    m2: object = 0.0
    count: float = 0.0


def _prenorm_update(st: _PreNormState, x: torch.Tensor):
    """model.py:394-423: batch mean / population variance, Chan merge.  Statistics are kept in the working dtype of
    ``x`` like the reference keeps them in fp32 TF scalars."""
    x = x.reshape(-1, st.n_units)
    sample_mean = x.mean(0)
    sample_var = ((x - sample_mean) ** 2).mean(0)
    sample_count = float(x.numel() / st.n_units)
    delta = sample_mean - st.mean
    st.m2 = st.var * st.count + sample_var * sample_count + delta ** 2 * st.count * sample_count / (
        st.count + sample_count)
    st.count += sample_count
    st.mean = st.mean + delta * sample_count / st.count
    st.var = st.m2 / st.count if st.count > 0 else 1


class OracleGCNN:
    """Mirror of the reference ``GCNN`` (model.py:136-300) on torch CPU."""

    def __init__(self, params: dict | None = None, dtype=torch.float32, faithful: bool = True):
        self.dtype = dtype
        self.faithful = faithful
        self.params = {k: v.detach().clone().to(dtype) for k, v in
                       (params or init_params(dtype=dtype, identity_prenorm=True)).items()}
        self._pn = {name: _PreNormState(int(np.prod(self.params[name + "/scale"].shape))) for name in PRENORM_LAYERS}
        self.trace: dict[str, torch.Tensor] = {}

    # -- helpers ------------------------------------------------------------------------------------------------
    def _prenorm(self, name: str, x: torch.Tensor) -> torch.Tensor:
        st = self._pn[name]
        if st.waiting:  # model.py:372-375
            _prenorm_update(st, x.detach())
            st.received = True
            raise PreNormException
        p = self.params
        if name + "/shift" in p:  # model.py:377-381
            x = x + p[name + "/shift"]
        return x * p[name + "/scale"]

    def _dense(self, name: str, x, relu=False, bias=True):
        y = x @ self.params[name + "/kernel"]
        if bias:
            y = y + self.params[name + "/bias"]
        return torch.relu(y) if relu else y

    def _embed(self, prefix: str, x):  # model.py:174-195
        x = self._prenorm(prefix + "/prenorm", x)
        return self._dense(prefix + "_2", self._dense(prefix + "_1", x, relu=True), relu=True)

    def _conv(self, name: str, from_v: bool, left, ei, ef, var, out_size: int):
        """model.py:533-575."""
        recv_side, recv_feats = (0, left) if from_v else (1, var)
        ei = ei.long()
        a = self._dense(name + "_feat_left", left)
        b = self._dense(name + "_feat_right", var, bias=False)
        joint = a[ei[0]] + ef @ self.params[name + "_feat_edge/kernel"] + b[ei[1]]  # model.py:564-565
        self.trace[name + "/z"] = joint
        joint = self._prenorm(name + "_final/prenorm", joint)
        joint = torch.relu(joint)
        if self.faithful:
            msg = self._dense(name + "_feat_final", joint)  # per-edge Dense, model.py:499-500
            conv = torch.zeros(out_size, EMB, dtype=joint.dtype).index_add_(0, ei[recv_side], msg)  # model.py:568
        else:  # hoisted: sum first, then one Dense per receiving node plus deg * bias (SURVEY.md header fact 2)
            h = torch.zeros(out_size, EMB, dtype=joint.dtype).index_add_(0, ei[recv_side], joint)
            deg = torch.zeros(out_size, dtype=joint.dtype).index_add_(
                0, ei[recv_side], torch.ones(ei.shape[1], dtype=joint.dtype))
            conv = h @ self.params[name + "_feat_final/kernel"] + deg[:, None] * self.params[name + "_feat_final/bias"]
        self.trace[name + "/conv"] = conv
        conv = self._prenorm(name + "_post/prenorm", conv)
        cat = torch.cat([conv, recv_feats], dim=1)  # model.py:573
        return self._dense(name + "_out_2", self._dense(name + "_out_1", cat, relu=True), relu=True)

    # -- model.py:257-300 ---------------------------------------------------------------------------------------
    def call(self, inputs, training=False):
        (cons, cons_ei, cons_ef, var, cut, cut_ei, cut_ef, n_cons, n_vars, n_cuts) = inputs
        t = lambda x: torch.as_tensor(np.asarray(x) if not torch.is_tensor(x) else x).to(self.dtype)
        cons, cons_ef, var, cut, cut_ef = t(cons), t(cons_ef), t(var), t(cut), t(cut_ef)
        cons_ei = torch.as_tensor(np.asarray(cons_ei) if not torch.is_tensor(cons_ei) else cons_ei)
        cut_ei = torch.as_tensor(np.asarray(cut_ei) if not torch.is_tensor(cut_ei) else cut_ei)
        n_cons, n_vars, n_cuts = int(n_cons), int(n_vars), int(n_cuts)

        c = self._embed("cons_emb", cons)
        cons_ef = self._prenorm("cons_edge/prenorm", cons_ef)
        v = self._embed("var_emb", var)
        k = self._embed("cut_emb", cut)
        cut_ef = self._prenorm("cut_edge/prenorm", cut_ef)
        c = self._conv("cons_conv", True, c, cons_ei, cons_ef, v, n_cons)
        v = self._conv("var_conv", False, c, cons_ei, cons_ef, v, n_vars)
        k = self._conv("cut_conv", True, k, cut_ei, cut_ef, v, n_cuts)
        self.trace["cut_out"] = k
        out = self._dense("out_2", self._dense("out_1", k, relu=True))
        return out.reshape(-1)

    __call__ = call

    # -- pretraining protocol, model.py:69-133 ------------------------------------------------------------------
    def pretrain_init(self):
        for st in self._pn.values():  # start_updates, model.py:384-392
            st.mean, st.var, st.m2, st.count = 0.0, 0.0, 0.0, 0.0
            st.waiting, st.received = True, False

    def pretrain(self, inputs, training=True) -> bool:
        try:
            with torch.no_grad():
                self.call(inputs, training)
            return False
        except PreNormException:
            return True

    def pretrain_next(self):
        for name in PRENORM_LAYERS:
            st = self._pn[name]
            if st.waiting and st.received:
                self._stop_updates(name, st)
                return st, name
        return None

    def _stop_updates(self, name, st):  # model.py:425-437
        as_t = lambda x: torch.as_tensor(x, dtype=self.dtype).reshape(st.n_units)
        if name + "/shift" in self.params:
            self.params[name + "/shift"] = -as_t(st.mean)
        var = as_t(st.var)
        var = torch.where(var == 0, torch.ones_like(var), var)
        self.params[name + "/scale"] = 1 / torch.sqrt(var)
        st.waiting = False


# ----------------------------------------------------------------------------------------------------------------
# Train step (model_trainer.py:259-273): MSE over all cuts of the batch, tape.gradient, Keras Adam.
# ----------------------------------------------------------------------------------------------------------------
@dataclass
class AdamState:
    step: int = 0
    m: dict = field(default_factory=dict)
    v: dict = field(default_factory=dict)


def loss_and_grads(model: OracleGCNN, inputs, targets, normaliser: float | None = None):
    """MSE = mean((y - p)^2) (Keras MeanSquaredError on 1-D tensors, model_trainer.py:271) and its gradient w.r.t.
    every trainable variable.  ``normaliser`` overrides the 1/N_k factor (used by the data-parallel tests, which
    seed un-normalised gradients and divide by the global cut count)."""
    names = [n for n, _ in TRAINABLE]
    leaves = [model.params[n].detach().clone().requires_grad_(True) for n in names]
    saved = dict(model.params)
    model.params.update(dict(zip(names, leaves)))
    try:
        pred = model.call(inputs, True)
        y = torch.as_tensor(np.asarray(targets) if not torch.is_tensor(targets) else targets).to(model.dtype)
        sq = ((y - pred) ** 2).sum()
        loss = sq / (pred.numel() if normaliser is None else normaliser)
        grads = torch.autograd.grad(loss, leaves, allow_unused=False)
    finally:
        model.params = saved
    return loss.detach(), pred.detach(), dict(zip(names, grads))


def adam_step(model: OracleGCNN, state: AdamState, grads: dict, lr: float,
              beta1=0.9, beta2=0.999, eps=1e-7):
    """Keras 2.7 ``Adam._resource_apply_dense`` (non-amsgrad): lr_t = lr * sqrt(1-b2^t) / (1-b1^t);
    m += (g-m)(1-b1); v += (g^2-v)(1-b2); theta -= lr_t * m / (sqrt(v) + eps)."""
    state.step += 1
    t = state.step
    lr_t = lr * np.sqrt(1 - beta2 ** t) / (1 - beta1 ** t)
    for name, g in grads.items():
        m = state.m.get(name, torch.zeros_like(g))
        v = state.v.get(name, torch.zeros_like(g))
        m = m + (g - m) * (1 - beta1)
        v = v + (g * g - v) * (1 - beta2)
        state.m[name], state.v[name] = m, v
        model.params[name] = model.params[name] - lr_t * m / (torch.sqrt(v) + eps)


def train_step(model: OracleGCNN, state: AdamState, inputs, targets, lr: float):
    loss, pred, grads = loss_and_grads(model, inputs, targets)
    adam_step(model, state, grads, lr)
    return loss, pred


# ----------------------------------------------------------------------------------------------------------------
# Batching contract (utils.py:339-426)
# ----------------------------------------------------------------------------------------------------------------
def concat_samples(samples):
    """``samples``: list of (state, improvements) with ``state`` the 5-tuple of dicts ``get_state`` produces
    (utils.py:236-238).  Follows utils.py:379-423 on in-memory samples."""
    cons, cons_ei, cons_ef, var, cut, cut_ei, cut_ef, imp = [], [], [], [], [], [], [], []
    for (s_cons, s_cons_e, s_var, s_cut, s_cut_e), s_imp in samples:
        cons.append(s_cons["values"]); cons_ei.append(s_cons_e["indices"]); cons_ef.append(s_cons_e["values"])
        var.append(s_var["values"]); cut.append(s_cut["values"])
        cut_ei.append(s_cut_e["indices"]); cut_ef.append(s_cut_e["values"]); imp.append(s_imp)
    n_cons = [x.shape[0] for x in cons]
    n_vars = [x.shape[0] for x in var]
    n_cuts = [x.shape[0] for x in cut]
    # utils.py:403-407: exclusive prefix sums per side, added column-wise to each sample's [2, E_s] block.
    cons_shift = np.cumsum([[0] + n_cons[:-1], [0] + n_vars[:-1]], axis=1)
    cut_shift = np.cumsum([[0] + n_cuts[:-1], [0] + n_vars[:-1]], axis=1)
    cons_ei = np.concatenate([e + cons_shift[:, j:j + 1] for j, e in enumerate(cons_ei)], axis=1)
    cut_ei = np.concatenate([e + cut_shift[:, j:j + 1] for j, e in enumerate(cut_ei)], axis=1)
    f32 = lambda xs: np.concatenate(xs, axis=0).astype(np.float32)
    return (f32(cons), cons_ei.astype(np.int32), f32(cons_ef), f32(var), f32(cut), cut_ei.astype(np.int32),
            f32(cut_ef), np.asarray(n_cons, np.int32), np.asarray(n_vars, np.int32), np.asarray(n_cuts, np.int32),
            np.concatenate(imp).astype(np.float32))


def load_batch(sample_files):
    """utils.py:371-376: gzip-pickled ``{'data': [state, improvements]}`` files."""
    samples = []
    for filename in sample_files:
        if isinstance(filename, bytes):
            filename = filename.decode()
        with gzip.open(filename, "rb") as fh:
            samples.append(tuple(pickle.load(fh)["data"]))
    return concat_samples(samples)


def model_inputs(batch):
    """model_trainer.py:259-263: only the totals reach the model."""
    return tuple(batch[:7]) + (int(np.sum(batch[7])), int(np.sum(batch[8])), int(np.sum(batch[9])))


def ranking_accuracy(predictions, improvements, n_cuts, fractions):
    """The accuracy bookkeeping of ``process`` (model_trainer.py:279-302), restated line by line: per sample, sort the
    cut indices by predicted and by true bound improvement (descending, Python ``sorted``: stable), find the first
    deviating position, and count the sample for every fraction its correctly ranked prefix reaches.
    Returns (acc [len(fractions)], deviations [n_samples])."""
    import numpy as np
    predictions = np.asarray(predictions, dtype=np.float32)
    improvements = np.asarray(improvements, dtype=np.float32)
    fractions = np.asarray(fractions, dtype=np.float64)
    acc = np.zeros(len(fractions))
    deviations = []
    start = 0
    for n in np.asarray(n_cuts).reshape(-1):
        pred = predictions[start:start + n]
        true = improvements[start:start + n]
        start += n
        pred_ranking = np.array(sorted(range(len(pred)), key=lambda x: pred[x], reverse=True))  # model_trainer.py:289
        true_ranking = np.array(sorted(range(len(true)), key=lambda x: true[x], reverse=True))  # model_trainer.py:290
        differences = (pred_ranking != true_ranking)
        deviation = int(np.argmax(differences)) if np.any(differences) else len(pred)          # :293-298
        deviations.append(deviation)
        acc += (deviation / len(pred) >= fractions)                                             # :300-301
    return acc, np.asarray(deviations, dtype=np.int32)


def select_cuts(quality, parallelism, parallelism_forced=None, p_max=0.1, p_max_ub=0.5, max_selected=None):
    """The ranking + parallelism filter of ``CustomCutsel.cutselselect`` (model_benchmarker.py:112-157), restated line by
    line on index arrays: ``parallelism[i, j]`` stands for ``model.getRowParallelism(cuts[i], cuts[j])`` and
    ``parallelism_forced[f, j]`` for ``getRowParallelism(forcedcuts[f], cuts[j])``.  Returns (order, n_selected) where
    ``order[pos]`` is the index of the cut at position ``pos`` of the reference's ``sorted_cuts``."""
    quality = np.asarray(quality)
    n = len(quality)
    parallelism = np.asarray(parallelism)
    n_forced = 0 if parallelism_forced is None else len(parallelism_forced)
    rankings = sorted(range(n), key=lambda x: quality[x], reverse=True)                      # :113
    sorted_cuts = np.array(rankings, dtype=np.int64)                                          # :114 (indices for objects)
    quality = -np.sort(-quality)                                                              # :117
    n_selected = n                                                                            # :120

    # `quality < 0.9 * quality[0]` (:128 / :149) under the reference's pinned numpy 1.22.3 (environment.yml:6): the product
    # of a Python float and a numpy scalar is a float64 scalar, and comparing a float32 ARRAY (the GCNN's scores) with it
    # uses value-based casting, i.e. the threshold is rounded to float32 and the comparison runs in float32; a float64
    # array (the hybrid rule's scores) compares in float64.  Spelled out so that the result does not depend on the numpy
    # installed here (numpy >= 2 would multiply in float32).
    threshold = 0.9 * float(quality[0]) if n else 0.0
    if quality.dtype == np.float32:
        threshold = np.float32(threshold)

    def remove(parallelism_row, n_selected, sorted_cuts):
        marked = (parallelism_row > p_max)                                                    # :125 / :146
        low_quality = np.logical_or(quality < threshold, parallelism_row > p_max_ub)          # :128 / :149
        to_remove = np.logical_and(marked, low_quality)                                       # :129 / :150
        removed = sorted_cuts[to_remove]                                                      # :132 / :153
        sorted_cuts = np.delete(sorted_cuts, to_remove)
        sorted_cuts = np.concatenate((sorted_cuts, removed))
        return sorted_cuts, n_selected - removed.size

    for f in range(n_forced):                                                                 # :121
        par = [parallelism_forced[f][sorted_cuts[j]] for j in range(n_selected)]              # :123
        par = np.pad(par, (0, n - n_selected), constant_values=0)                             # :124
        sorted_cuts, n_selected = remove(par, n_selected, sorted_cuts)
    i = 0                                                                                     # :141
    while i < n_selected - 1:
        par = [parallelism[sorted_cuts[i]][sorted_cuts[j]] for j in range(i + 1, n_selected)]  # :144
        par = np.pad(par, (i + 1, n - n_selected), constant_values=0)                         # :146
        sorted_cuts, n_selected = remove(par, n_selected, sorted_cuts)
        i += 1
    if max_selected is not None:
        n_selected = min(n_selected, max_selected)                                            # :157
    return sorted_cuts.astype(np.int32), int(n_selected)
