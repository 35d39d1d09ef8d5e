"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle and the golden fixtures.

Tolerances (BASELINE.json north_star: 1e-5 relative for the fp32 path, 1e-2 for a reduced-precision MLP path), all
relative to the max-abs of the tensor, against the fp64 oracle on identical inputs and weights:
  * edge indexing / offsets: bit-exact;
  * logits: 1e-5 on BOTH numerics paths;
  * parameter gradients (relative L2 per tensor; max-abs held to 10x, see assert_grads_close): 1e-5 on the default path
    (tcgen05 chains on bf16x3 tiles, all six products in the weight-gradient MMAs too) and on the exact-fp32 path
    (dense layers on fp32 SIMT FMAs; measured ~3e-7); 3e-5 on the two 3xTF32 alternates (the two-way split keeps 22 of
    24 operand mantissa bits).  tests/grad_error_report.py prints the measured worst case per fixture
    (profiles/r2_grad_errors*.json).
Every whole-model test below runs on all paths (fixture ``model``).
"""
import ctypes as C
import os

import numpy as np
import pytest
import torch

import gcnn_oracle as orc
from gcnn_cut_selector_b200 import batching, synth
from gcnn_cut_selector_b200._lib import check

pytestmark = pytest.mark.gpu

TOL = 1e-5


LINF_FACTOR = 4
GRAD_TOL = {"tc_chains": 1e-5, "tc_tf32_fwd": 3e-5, "tc_unfused": 3e-5, "fp32": 1e-5}


@pytest.fixture(scope="module", params=["tc_chains", "tc_tf32_fwd", "tc_unfused", "fp32"])
def model(request, golden_dir):
    """tc_chains: the default path (fused tcgen05 chains on bf16x3 tiles, forward and backward); tc_tf32_fwd: the same
    with the 3xTF32 forward chains; tc_unfused: one tensor-core launch per dense layer and per gradient (3xTF32); fp32:
    exact-fp32 SIMT dense layers.  The product library carries only tc_chains; the other three are compiled with
    -DGCNN_ALT_PATHS (``python -m gcnn_cut_selector_b200.build -DGCNN_ALT_PATHS --variant=alt``) and run with
    ``GCNN_LIB_VARIANT=alt python -m pytest tests -m gpu``."""
    from gcnn_cut_selector_b200 import GCNN
    m = GCNN(device="cuda:0", seed=0)
    if request.param != "tc_chains" and not m._lib.gcnn_has_alt_paths():
        pytest.skip("A/B alternate: needs the -DGCNN_ALT_PATHS build (GCNN_LIB_VARIANT=alt)")
    m.restore_state(os.path.join(golden_dir, "state_stream.pkl"))
    m.set_option("tensor_cores", 0 if request.param == "fp32" else 1)
    m.set_option("fused", 0 if request.param == "tc_unfused" else 1)
    m.set_option("fused_backward", 0 if request.param == "tc_unfused" else 1)
    m.set_option("bf16_forward", 1 if request.param == "tc_chains" else 0)
    m.grad_tol = GRAD_TOL[request.param]
    return m


@pytest.fixture(scope="module")
def oracle64(golden_dir):
    return orc.OracleGCNN(orc.restore_state(os.path.join(golden_dir, "state_stream.pkl"), dtype=torch.float64),
                          dtype=torch.float64)


def rel_err(got, want):
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    return np.abs(got - want).max() / max(np.abs(want).max(), 1e-30)


def golden_inputs(z, prefix=""):
    g = lambda k: z[prefix + k]
    return (g("cons"), g("cons_ei"), g("cons_ef"), g("var"), g("cut"), g("cut_ei"), g("cut_ef"),
            int(g("n_cons").sum()), int(g("n_vars").sum()), int(g("n_cuts").sum()))


def assert_grads_close(flat_got, grads_ref: dict, tol=TOL, grads_f32=None):
    """Per parameter tensor, against fp64 truth: relative L2 error <= tol and max-abs error <= LINF_FACTOR x tol.

    Why two norms: a ReLU network's gradient is discontinuous where a pre-activation is ~0.  Among the ~10^7
    pre-activations of a batch a few always sit within fp32 rounding of zero, and ANY fp32 evaluation (this one, the
    reference's TF kernels, the torch restatement -- which deviates up to 1.6e-5 in max-abs on these batches) may put
    them on the other side of the kink than fp64 does.  One such flip moves single gradient elements by ~1e-5 of the
    tensor's max-abs but leaves the tensor's L2 error far below 1e-5, so the tight bound is on the L2 norm and the
    max-abs bound catches real bugs (a wrong index or a dropped term shows up at 1e-2 .. 1).  LINF_FACTOR = 4: the worst
    max-abs deviation measured over all shapes is 2.6e-5 on the product path and 1.6e-5 on the exact-fp32 path
    (profiles/r2_grad_errors.md); round 1 allowed 10 x."""
    flat_got = np.asarray(flat_got, np.float64)
    gmax = max(float(g.abs().max()) for g in grads_ref.values())
    gl2 = max(float(g.norm()) for g in grads_ref.values())
    o = 0
    for name, shape in orc.TRAINABLE:
        k = int(np.prod(shape))
        ref = grads_ref[name].reshape(-1).numpy().astype(np.float64)
        diff = flat_got[o:o + k] - ref
        # tensors whose whole gradient is tiny are held to the global scale
        l2 = np.linalg.norm(diff) / max(np.linalg.norm(ref), 1e-3 * gl2)
        linf = np.abs(diff).max() / max(np.abs(ref).max(), 1e-3 * gmax)
        assert l2 <= tol, f"{name}: relative L2 error {l2:.3e} > {tol:.1e}"
        assert linf <= LINF_FACTOR * tol, f"{name}: max-abs error {linf:.3e} > {LINF_FACTOR * tol:.1e}"
        o += k


# ---- F1: CSR / CSC build, bit-exact -------------------------------------------------------------------------------
def _csr_case(model, ei, ef, n_left, n_vars):
    from gcnn_cut_selector_b200._lib import check
    lib, dev = model._lib, model.device
    E = ei.shape[1]
    check(lib.gcnn_workspace_reserve(model._ws, n_left, n_vars, 1, E, 1, 1))
    d_ei = torch.from_numpy(np.ascontiguousarray(ei.astype(np.int32))).to(dev)
    d_ef = torch.from_numpy(np.ascontiguousarray(ef.astype(np.float32))).to(dev)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    check(lib.gcnn_build_csr(model._ws, 0, d_ei.data_ptr(), d_ef.data_ptr(), E, n_left, n_vars, 1, st))
    for side, n_owner in ((0, n_left), (1, n_vars)):
        ptr = torch.empty(n_owner + 1, dtype=torch.int32, device=dev)
        other = torch.empty(E, dtype=torch.int32, device=dev)
        val = torch.empty(E, dtype=torch.float32, device=dev)
        perm = torch.empty(E, dtype=torch.int32, device=dev)
        check(lib.gcnn_csr_export(model._ws, 0, side, ptr.data_ptr(), other.data_ptr(), val.data_ptr(),
                                  perm.data_ptr(), st))
        check(lib.gcnn_check(model._ws, st))
        keys = ei[side].astype(np.int64)
        want_perm = np.argsort(keys, kind="stable")
        want_ptr = np.concatenate([[0], np.cumsum(np.bincount(keys, minlength=n_owner))])
        np.testing.assert_array_equal(ptr.cpu().numpy(), want_ptr.astype(np.int32))
        np.testing.assert_array_equal(perm.cpu().numpy(), want_perm.astype(np.int32))
        np.testing.assert_array_equal(other.cpu().numpy(), ei[1 - side][want_perm].astype(np.int32))
        np.testing.assert_array_equal(val.cpu().numpy(), ef.astype(np.float32)[want_perm])


@pytest.mark.parametrize("case", ["sorted", "shuffled", "empty", "single", "one_owner", "isolated", "big", "hub"])
def test_csr_build_bit_exact(model, case):
    rng = np.random.default_rng(11)
    if case == "sorted":
        (c, ce, v, k, ke), _ = synth.make_sample("setcov", 3)
        ei, ef, nl, nv = ce["indices"], ce["values"][:, 0], 500, 1000
    elif case == "shuffled":
        (c, ce, v, k, ke), _ = synth.shuffle_edges(synth.make_sample("setcov", 4), 1)
        ei, ef, nl, nv = ce["indices"], ce["values"][:, 0], 500, 1000
    elif case == "empty":
        ei, ef, nl, nv = np.zeros((2, 0), np.int64), np.zeros(0), 5, 7
    elif case == "single":
        ei, ef, nl, nv = np.array([[3], [2]]), np.array([0.5]), 5, 7
    elif case == "one_owner":
        ei, ef, nl, nv = np.vstack([np.zeros(300, np.int64), rng.integers(0, 50, 300)]), rng.standard_normal(300), 1, 50
    elif case == "isolated":  # many owners without edges, long pointer gaps
        ei = np.vstack([np.sort(rng.choice(100000, 200)), rng.integers(0, 70000, 200)])
        ef, nl, nv = rng.standard_normal(200), 100000, 70000
    elif case == "big":  # > 16 bits of key, unsorted on both sides, multiple CTAs and radix passes
        E = 1_000_003
        ei, ef, nl, nv = np.vstack([rng.integers(0, 100_000, E), rng.integers(0, 70_001, E)]), rng.standard_normal(E), 100_000, 70_001
    else:  # one variable adjacent to every row
        ei = np.vstack([np.repeat(np.arange(3000), 2), np.tile([0, 1], 3000)])
        ei[1, 1::2] = rng.integers(1, 9, 3000)
        ef, nl, nv = rng.standard_normal(6000), 3000, 9
    _csr_case(model, np.asarray(ei), np.asarray(ef), nl, nv)


@pytest.mark.parametrize("n_owner", [2, 255, 256, 257, 511, 512, 513, 65_535, 65_536, 65_537, 131_072, 262_144, 262_145])
def test_csr_digit_width_boundaries(model, n_owner):
    """Key widths around the points where the radix sort changes its pass count or digit width (8-bit digits; 9-bit ones
    for 9-, 17- and 18-bit keys), unsorted on both sides, keys hitting 0 and n_owner - 1, several tiles per pass."""
    rng = np.random.default_rng(n_owner)
    E = 20_011
    rows, cols = rng.integers(0, n_owner, E), rng.integers(0, n_owner + 3, E)
    rows[:4], cols[:4] = [0, n_owner - 1, 0, n_owner - 1], [n_owner + 2, 0, 0, n_owner + 2]
    _csr_case(model, np.vstack([rows, cols]), rng.standard_normal(E), n_owner, n_owner + 3)


def test_out_of_range_index_raises(model, golden_dir):
    from gcnn_cut_selector_b200 import InvalidArgumentError
    z = np.load(os.path.join(golden_dir, "fwd_tiny3.npz"))
    inputs = list(golden_inputs(z))
    bad = inputs[1].copy()
    bad[1, 3] = inputs[8]  # == n_vars
    inputs[1] = bad
    with pytest.raises(InvalidArgumentError):
        model(tuple(inputs), False)
    # the workspace stays usable
    out = model(golden_inputs(z), False)
    assert rel_err(out.cpu().numpy(), z["scores_f64"]) <= TOL


def test_violated_sorted_promise_is_reported(model):
    """gcnn_batch.flags may promise row-0-sorted edges (what the reference's loader produces); the device verifies."""
    from gcnn_cut_selector_b200 import InvalidArgumentError, _lib
    sample = synth.shuffle_edges(synth.make_sample("mini", 3), 1)
    batch = batching.concat_samples([sample])
    dev = model.prepare_inputs(batching.model_inputs(batch))
    assert dev[0].flags == 0  # the host check saw the shuffle
    dev[0].flags = _lib.BATCH_CONS_EDGES_SORTED | _lib.BATCH_CUT_EDGES_SORTED
    with pytest.raises(InvalidArgumentError, match="sorted"):
        model._forward(dev, save_activations=False)
    good = model.prepare_inputs(batching.model_inputs(batching.concat_samples([synth.make_sample("mini", 3)])))
    assert good[0].flags == 3
    assert torch.isfinite(model._forward(good, save_activations=False)).all()


def test_device_tensor_inputs_take_the_unhinted_path(model, oracle64):
    batch = batching.concat_samples(synth.make_samples("mini", 2, seed0=11))
    inputs = batching.model_inputs(batch)
    dev_inputs = tuple(torch.from_numpy(np.ascontiguousarray(x)).to(model.device) if isinstance(x, np.ndarray) else x
                       for x in inputs)
    assert model.prepare_inputs(dev_inputs)[0].flags == 0
    with torch.no_grad():
        out = model(dev_inputs, False)
    assert rel_err(out.cpu().numpy(), oracle64.call(inputs).numpy()) <= TOL


# ---- per-op parity ----------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("s_f", [0.7, -1.3, 0.0])  # the sign of the pre-norm scale selects the active half-line of z
# (5, 3, 3000) and (300, 2, 4000): rows of 600 / 2000 edges take the whole-CTA path of the forward / backward kernels
@pytest.mark.parametrize("n_recv,n_send,E", [(50, 70, 600), (1, 3, 5), (300, 2, 4000), (64, 64, 0), (5, 3, 3000), (2000, 900, 5000)])
def test_edge_forward_backward_ops(model, n_recv, n_send, E, s_f):
    from gcnn_cut_selector_b200._lib import check
    lib, dev = model._lib, model.device
    rng = np.random.default_rng(E + n_recv)
    recv = np.sort(rng.integers(0, n_recv, E))
    send = rng.integers(0, n_send, E)
    f = rng.standard_normal(E).astype(np.float32)
    R = rng.standard_normal((n_recv, 64)).astype(np.float32)
    S = rng.standard_normal((n_send, 64)).astype(np.float32)
    w = rng.standard_normal(64).astype(np.float32)
    G = rng.standard_normal((n_recv, 64)).astype(np.float32)
    f_shift, f_scale = 0.25, 1.5
    # oracle in fp64
    fe = (f.astype(np.float64) + f_shift) * f_scale
    z = R.astype(np.float64)[recv] + fe[:, None] * w.astype(np.float64) + S.astype(np.float64)[send]
    y = np.maximum(s_f * z, 0)
    H = np.zeros((n_recv, 64)); np.add.at(H, recv, y)
    cnt = np.zeros((n_recv, 64)); np.add.at(cnt, recv, (y > 0).astype(np.float64))
    dz = s_f * (s_f * z > 0) * G.astype(np.float64)[recv]
    dS = np.zeros((n_send, 64)); np.add.at(dS, send, dz)
    dw = (fe[:, None] * dz).sum(0)

    t = lambda a, dt=torch.float32: torch.from_numpy(np.ascontiguousarray(a)).to(dev, dt)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    ptr = np.concatenate([[0], np.cumsum(np.bincount(recv, minlength=n_recv))]).astype(np.int32)
    d_ptr, d_src, d_val = t(ptr, torch.int32), t(send.astype(np.int32), torch.int32), t(f)
    dR_, dS_, dw_, dG = t(R), t(S), t(w), t(G)
    dH = torch.empty(n_recv, 64, device=dev); dcnt = torch.empty(n_recv, 64, device=dev)
    check(lib.gcnn_edge_forward(d_ptr.data_ptr(), d_src.data_ptr(), d_val.data_ptr(), n_recv, dR_.data_ptr(),
                                dS_.data_ptr(), dw_.data_ptr(), f_shift, f_scale, s_f, dH.data_ptr(), dcnt.data_ptr(),
                                st))
    torch.cuda.synchronize()
    if E and s_f != 0.0:
        assert rel_err(dH.cpu().numpy(), H) <= 2e-6
    else:
        assert np.all(dH.cpu().numpy() == 0)
    np.testing.assert_array_equal(dcnt.cpu().numpy(), cnt.astype(np.float32))

    # transposed layout for the backward
    order = np.argsort(send, kind="stable")
    ptr_s = np.concatenate([[0], np.cumsum(np.bincount(send, minlength=n_send))]).astype(np.int32)
    check(lib.gcnn_workspace_reserve(model._ws, 8, 8, 8, 8, 8, 1))
    d_ptr_s, d_oth, d_val_s = t(ptr_s, torch.int32), t(recv[order].astype(np.int32), torch.int32), t(f[order])
    out_dS = torch.empty(n_send, 64, device=dev); out_dw = torch.empty(64, device=dev)
    check(lib.gcnn_edge_backward(model._ws, d_ptr_s.data_ptr(), d_oth.data_ptr(), d_val_s.data_ptr(), n_send,
                                 dR_.data_ptr(), dS_.data_ptr(), dG.data_ptr(), dw_.data_ptr(), f_shift, f_scale, s_f,
                                 out_dS.data_ptr(), out_dw.data_ptr(), st))
    torch.cuda.synchronize()
    if E and s_f != 0.0:
        tol = 2e-6 if E <= 500 * n_send else 1e-5  # one warp sums a 1,000-edge row serially in fp32 (as the reference does)
        assert rel_err(out_dS.cpu().numpy(), dS) <= tol
        assert rel_err(out_dw.cpu().numpy(), dw) <= tol
    else:
        assert np.all(out_dS.cpu().numpy() == 0) and np.all(out_dw.cpu().numpy() == 0)


@pytest.mark.parametrize("m", [1, 127, 128, 129, 1000])
def test_linear_forward_op(model, m):
    from gcnn_cut_selector_b200._lib import check
    if not model._lib.gcnn_has_alt_paths():
        pytest.skip("the fp32 SIMT dense kernel is an A/B alternate (-DGCNN_ALT_PATHS build)")
    rng = np.random.default_rng(m)
    X, W, b = rng.standard_normal((m, 64)), rng.standard_normal((64, 64)), rng.standard_normal(64)
    dev = model.device
    t = lambda a: torch.from_numpy(a.astype(np.float32)).to(dev)
    dX, dW, db = t(X), t(W), t(b)
    Y = torch.empty(m, 64, device=dev)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    check(model._lib.gcnn_linear_forward(dX.data_ptr(), dW.data_ptr(), db.data_ptr(), m, 64, 1, Y.data_ptr(), st))
    torch.cuda.synchronize()
    ref = np.maximum(X.astype(np.float32).astype(np.float64) @ W.astype(np.float32).astype(np.float64)
                     + b.astype(np.float32).astype(np.float64), 0)
    assert rel_err(Y.cpu().numpy(), ref) <= 2e-6


# ---- the tensor-core node chains one at a time (gcnn_conv_forward / backward, gcnn_embed_forward / backward,
#      gcnn_head_forward / backward) against numpy fp64 restatements of model.py:174-208, 563-573 and their adjoints -------
CONV_NAMES = ("cons_conv", "var_conv", "cut_conv")
NEXT_LAYER = (("var_conv_feat_left", True, False), ("cut_conv_feat_right", False, False), ("out_1", True, True))  # name, bias, relu


class _ChainOps:
    """A model with the golden weights, non-trivial pre-norm values and a workspace reserved for training."""

    def __init__(self, golden_dir):
        from gcnn_cut_selector_b200 import GCNN
        self.m = m = GCNN(device="cuda:0", seed=3)
        m.restore_state(os.path.join(golden_dir, "state_stream.pkl"))
        rng = np.random.default_rng(5)
        with torch.no_grad():
            m.flat_prenorm.copy_(torch.from_numpy(rng.uniform(0.6, 1.4, m.flat_prenorm.numel()).astype(np.float32)))
        check(m._lib.gcnn_workspace_reserve(m._ws, 512, 512, 512, 1024, 1024, 1))
        self.p64 = m.flat_params.detach().cpu().numpy().astype(np.float64)
        self.pn64 = m.flat_prenorm.cpu().numpy().astype(np.float64)
        self.table = {name: (shape, trainable, off) for name, shape, trainable, off in m._table}

    def get(self, name):
        shape, trainable, off = self.table[name]
        return (self.p64 if trainable else self.pn64)[off:off + int(np.prod(shape))].reshape(shape)

    def grad(self, flat, name):
        shape, _, off = self.table[name]
        return flat[off:off + int(np.prod(shape))].reshape(shape)

    def dev(self, a, dtype=torch.float32):
        return torch.from_numpy(np.ascontiguousarray(a)).to(device="cuda:0", dtype=dtype)

    def call(self, fn, *args):
        m = self.m
        ptr = lambda x: None if x is None else (x.data_ptr() if torch.is_tensor(x) else x)
        with torch.cuda.device(m.device):
            check(getattr(m._lib, fn)(*[ptr(a) for a in args], m._stream()))


@pytest.fixture(scope="module")
def chain_ops(golden_dir):
    return _ChainOps(golden_dir)


def l2_err(got, want):
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    return np.linalg.norm(got - want) / max(np.linalg.norm(want), 1e-30)


@pytest.mark.parametrize("M", [300, 128, 1])
@pytest.mark.parametrize("conv", [0, 1, 2])
def test_conv_chain_ops_match_numpy(chain_ops, conv, M):
    o, rng = chain_ops, np.random.default_rng(10 * conv + M)
    m, name = o.m, CONV_NAMES[conv]
    Wf, bf = o.get(f"{name}_feat_final/kernel"), o.get(f"{name}_feat_final/bias")
    Wo1, bo1 = o.get(f"{name}_out_1/kernel"), o.get(f"{name}_out_1/bias")
    Wo2, bo2 = o.get(f"{name}_out_2/kernel"), o.get(f"{name}_out_2/bias")
    nname, nbias, nrelu = NEXT_LAYER[conv]
    Wn, bn = o.get(f"{nname}/kernel"), (o.get(f"{nname}/bias") if nbias else np.zeros(64))
    s_p, s_f = float(o.get(f"{name}_post/prenorm/scale")[0]), float(o.get(f"{name}_final/prenorm/scale")[0])
    H, Xt = rng.normal(size=(M, 64)).astype(np.float32), rng.normal(size=(M, 64)).astype(np.float32)
    deg = rng.integers(0, 9, M)
    deg_ptr = np.concatenate(([0], np.cumsum(deg))).astype(np.int32)
    # forward
    H64, Xt64 = H.astype(np.float64), Xt.astype(np.float64)
    C = H64 @ Wf + deg[:, None] * bf
    U1 = np.maximum(np.concatenate([s_p * C, Xt64], 1) @ Wo1 + bo1, 0)
    Y = np.maximum(U1 @ Wo2 + bo2, 0)
    Pn = Y @ Wn + bn
    if nrelu:
        Pn = np.maximum(Pn, 0)
    dH, dXt_in, dptr = o.dev(H), o.dev(Xt), o.dev(deg_ptr, torch.int32)
    outs = [torch.full((M, 64), float("nan"), device="cuda:0") for _ in range(4)]
    scores = torch.full((M,), float("nan"), device="cuda:0") if conv == 2 else None
    o.call("gcnn_conv_forward", m._ws, m.flat_params, m.flat_prenorm, conv, dH, dXt_in, dptr, M, *outs, scores)
    for got, want, what in zip(outs, (C, U1, Y, Pn), ("C", "U1", "Y", "Pn")):
        assert l2_err(got.cpu().numpy(), want) <= TOL, what
    if conv == 2:
        want = Pn @ o.get("out_2/kernel")[:, 0] + o.get("out_2/bias")[0]
        assert l2_err(scores.cpu().numpy(), want) <= TOL
    # backward: saved activations as the forward kernel wrote them (the ReLU masks are read from them)
    Cs, U1s, Ys = (x.cpu().numpy().astype(np.float64) for x in outs[:3])
    dP = rng.normal(size=(M, 64)).astype(np.float32)
    cnt = rng.integers(0, 9, (M, 64)).astype(np.float32)
    dP64 = dP.astype(np.float64)
    dU2 = (dP64 @ Wn.T) * (Ys > 0)
    dU1 = (dU2 @ Wo2.T) * (U1s > 0)
    dcat = dU1 @ Wo1.T
    dC, dXt = s_p * dcat[:, :64], dcat[:, 64:]
    G = dC @ Wf.T
    dR = s_f * G * cnt
    want_g = {f"{nname}/kernel": Ys.T @ dP64, f"{name}_out_2/kernel": U1s.T @ dU2, f"{name}_out_2/bias": dU2.sum(0),
              f"{name}_out_1/kernel": np.concatenate([s_p * Cs, Xt64], 1).T @ dU1, f"{name}_out_1/bias": dU1.sum(0),
              f"{name}_feat_final/kernel": H64.T @ dC, f"{name}_feat_final/bias": (deg[:, None] * dC).sum(0)}
    if nbias:
        want_g[f"{nname}/bias"] = dP64.sum(0)
    bouts = [torch.full((M, 64), float("nan"), device="cuda:0") for _ in range(3)]
    grads = torch.zeros_like(m.flat_grads)
    o.call("gcnn_conv_backward", m._ws, m.flat_params, m.flat_prenorm, conv, o.dev(dP), outs[2], outs[1], outs[0], dXt_in, dH,
           o.dev(cnt), dptr, M, *bouts, grads)
    for got, want, what in zip(bouts, (dXt, G, dR), ("dXt", "G", "dR")):
        assert l2_err(got.cpu().numpy(), want) <= TOL, what
    flat = grads.cpu().numpy()
    for k, want in want_g.items():
        assert l2_err(o.grad(flat, k), want) <= TOL, k


@pytest.mark.parametrize("M", [300, 128, 1])
@pytest.mark.parametrize("node_type", [0, 1, 2])
def test_embedding_chain_ops_match_numpy(chain_ops, node_type, M):
    o, rng = chain_ops, np.random.default_rng(100 + 10 * node_type + M)
    m, name = o.m, ("cons", "var", "cut")[node_type]
    K = (4, 14, 6)[node_type]
    shift, scale = o.get(f"{name}_emb/prenorm/shift"), o.get(f"{name}_emb/prenorm/scale")
    W1, b1, W2, b2 = (o.get(f"{name}_emb_{i}/{w}") for i in (1, 2) for w in ("kernel", "bias"))
    proj = ((("cons_conv_feat_left", True),), (("cons_conv_feat_right", False), ("var_conv_feat_right", False)),
            (("cut_conv_feat_left", True),))[node_type]
    x = rng.normal(size=(M, K)).astype(np.float32)
    xn = (x.astype(np.float64) + shift) * scale
    h1 = np.maximum(xn @ W1 + b1, 0)
    out = np.maximum(h1 @ W2 + b2, 0)
    Ps = [out @ o.get(f"{n}/kernel") + (o.get(f"{n}/bias") if bias else 0.0) for n, bias in proj]
    dx = o.dev(x)
    outs = [torch.full((M, 64), float("nan"), device="cuda:0") for _ in range(2 + len(Ps))]
    o.call("gcnn_embed_forward", m._ws, m.flat_params, m.flat_prenorm, node_type, dx, M, outs[0], outs[1], outs[2],
           outs[3] if len(Ps) == 2 else None)
    for got, want, what in zip(outs, [h1, out] + Ps, ("h1", "out", "P0", "P1")):
        assert l2_err(got.cpu().numpy(), want) <= TOL, what
    h1s, outs64 = outs[0].cpu().numpy().astype(np.float64), outs[1].cpu().numpy().astype(np.float64)
    dPs = [rng.normal(size=(M, 64)).astype(np.float32) for _ in Ps]
    dXt = rng.normal(size=(M, 64)).astype(np.float32)
    dout = sum(d.astype(np.float64) @ o.get(f"{n}/kernel").T for d, (n, _) in zip(dPs, proj)) + dXt
    g0 = dout * (outs64 > 0)
    dh1 = (g0 @ W2.T) * (h1s > 0)
    want_g = {f"{name}_emb_2/kernel": h1s.T @ g0, f"{name}_emb_2/bias": g0.sum(0), f"{name}_emb_1/kernel": xn.T @ dh1,
              f"{name}_emb_1/bias": dh1.sum(0)}
    for d, (n, bias) in zip(dPs, proj):
        want_g[f"{n}/kernel"] = outs64.T @ d.astype(np.float64)
        if bias:
            want_g[f"{n}/bias"] = d.astype(np.float64).sum(0)
    grads = torch.zeros_like(m.flat_grads)
    o.call("gcnn_embed_backward", m._ws, m.flat_params, m.flat_prenorm, node_type, o.dev(dPs[0]),
           o.dev(dPs[1]) if len(dPs) == 2 else None, o.dev(dXt), outs[1], outs[0], dx, M, grads)
    flat = grads.cpu().numpy()
    for k, want in want_g.items():
        assert l2_err(o.grad(flat, k), want) <= TOL, k


@pytest.mark.parametrize("M", [300, 1])
def test_head_ops_match_numpy(chain_ops, M):
    o, rng = chain_ops, np.random.default_rng(200 + M)
    m = o.m
    w, b = o.get("out_2/kernel")[:, 0], o.get("out_2/bias")
    g = np.maximum(rng.normal(size=(M, 64)), 0).astype(np.float32)
    ds = rng.normal(size=M).astype(np.float32)
    dw, db = m.flat_params.detach()[o.table["out_2/kernel"][2]:], m.flat_params.detach()[o.table["out_2/bias"][2]:]
    scores = torch.full((M,), float("nan"), device="cuda:0")
    dg = o.dev(g)
    o.call("gcnn_head_forward", dg, dw, db, M, scores)
    assert l2_err(scores.cpu().numpy(), g.astype(np.float64) @ w + b[0]) <= TOL
    dg_pre = torch.full((M, 64), float("nan"), device="cuda:0")
    dw_db = torch.full((65,), float("nan"), device="cuda:0")
    o.call("gcnn_head_backward", m._ws, dg, dw, o.dev(ds), M, dg_pre, dw_db)
    ds64 = ds.astype(np.float64)
    assert l2_err(dg_pre.cpu().numpy(), ds64[:, None] * w[None, :] * (g > 0)) <= TOL
    assert l2_err(dw_db.cpu().numpy(), np.concatenate([g.astype(np.float64).T @ ds64, [ds64.sum()]])) <= TOL


# ---- whole model against the golden fixtures (reference source over the TF shim) ----------------------------------------
@pytest.mark.parametrize("case", ["tiny3", "mini2", "isolated"])
def test_forward_matches_golden(model, golden_dir, case):
    z = np.load(os.path.join(golden_dir, f"fwd_{case}.npz"))
    for training in (False, True):
        with torch.no_grad():
            out = model(golden_inputs(z), training)
        assert out.shape == z["scores_f64"].shape and out.dtype == torch.float32
        assert rel_err(out.cpu().numpy(), z["scores_f64"]) <= TOL


@pytest.mark.parametrize("case", ["tiny3", "mini2", "isolated"])
def test_gradients_match_golden(model, golden_dir, case):
    z = np.load(os.path.join(golden_dir, f"fwd_{case}.npz"))
    loss_sum, scores = model.loss_and_grads(golden_inputs(z), z["targets"])
    torch.cuda.synchronize()
    n = scores.numel()
    assert abs(float(loss_sum) / n - float(z["loss_f64"])) <= TOL * float(z["loss_f64"])
    ref = torch.from_numpy(z["grad_f64_as_f32"].astype(np.float64))
    grads_ref, o = {}, 0
    for name, shape in orc.TRAINABLE:
        k = int(np.prod(shape))
        grads_ref[name] = ref[o:o + k]
        o += k
    assert_grads_close(model.flat_grads.cpu().numpy(), grads_ref, tol=model.grad_tol)


def test_autograd_bridge_matches_fused_call(model, golden_dir):
    z = np.load(os.path.join(golden_dir, "fwd_mini2.npz"))
    model.loss_and_grads(golden_inputs(z), z["targets"])
    fused = model.flat_grads.clone()
    model.flat_params.grad = None
    pred = model(golden_inputs(z), True)
    y = torch.from_numpy(z["targets"]).to(model.device)
    loss = ((y - pred) ** 2).mean()
    loss.backward()
    # the bridge seeds 2 (p - y) / n in torch, the fused call inside head_loss_kernel: the seeds differ in the last bit, so
    # gradient elements agree to ~1e-7 of the gradient's scale (elements that are sums with heavy cancellation -- the
    # degree-weighted bias sums -- not to 1e-6 of their own tiny value)
    scale = float(fused.abs().max())
    torch.testing.assert_close(model.flat_params.grad, fused, rtol=1e-6, atol=1e-6 * scale)
    model.flat_params.grad = None


# ---- the four problem classes at BASELINE shapes, against the fp64 oracle run here -------------------------------------
# "skewed": heavy rows next to each other, hub columns, empty rows (weight-balanced CTA ranges); "skewed-coop" lowers the
# long-row threshold so that those rows, the hub columns and the cut rows are reduced by whole CTAs (forward and backward)
# counts=True passes load_batch's per-sample count vectors (utils.py:420-422): the shared-memory block edge kernels and
# the per-sample transposed layouts (csrc/edge_block.cu); capfac samples are too large for them and fall back
@pytest.mark.parametrize("counts", [False, True], ids=["totals", "per_sample_counts"])
@pytest.mark.parametrize("shape,n", [("setcov", 1), ("setcov", 3), ("combauc", 4), ("indset", 4), ("capfac", 1), ("skewed", 3),
                                     ("skewed-coop", 3)])
def test_problem_classes_forward_backward(model, oracle64, shape, n, counts):
    if shape == "skewed-coop":
        model.set_option("long_row", 64)
        try:
            test_problem_classes_forward_backward(model, oracle64, "skewed", n, counts)
        finally:
            model.set_option("long_row", 512)
        return
    batch = batching.concat_samples(synth.make_samples(shape, n, seed0=1000 + n))
    inputs, targets = batching.model_inputs(batch, per_sample_counts=counts), batch[10]
    loss_sum, scores = model.loss_and_grads(inputs, targets)
    torch.cuda.synchronize()
    loss, pred, grads = orc.loss_and_grads(oracle64, batching.model_inputs(batch), targets)
    assert rel_err(scores.cpu().numpy(), pred.numpy()) <= TOL
    assert abs(float(loss_sum) / scores.numel() - float(loss)) <= TOL * float(loss)
    assert_grads_close(model.flat_grads.cpu().numpy(), grads, tol=model.grad_tol)


def test_negative_and_zero_prenorm_scales(model, golden_dir):
    """Pre-norm scales are data (restore_state can load anything): a negative feature_module_final scale flips which
    half-line of the joint pre-activation is active, a zero one switches the convolution's message path off, negative
    post-conv and input scales are plain multipliers.  Scores and gradients against the fp64 oracle with the same values."""
    z = np.load(os.path.join(golden_dir, "fwd_mini2.npz"))
    inputs, targets = golden_inputs(z), z["targets"]
    saved = model.flat_prenorm.clone()
    names = [n for n, _, tr in orc.PARAM_SPECS if not tr]
    try:
        pn = orc.flatten_prenorm(orc.restore_state(os.path.join(golden_dir, "state_stream.pkl"), dtype=torch.float64)).clone()
        offs, o = {}, 0
        for n, shape, tr in orc.PARAM_SPECS:
            if not tr:
                offs[n] = o
                o += int(np.prod(shape))
        pn[offs["cons_conv_final/prenorm/scale"]] = -0.8
        pn[offs["cons_conv_post/prenorm/scale"]] = -1.7
        pn[offs["var_conv_final/prenorm/scale"]] = 0.0
        pn[offs["cut_conv_final/prenorm/scale"]] = -0.4
        pn[offs["cons_edge/prenorm/scale"]] = -2.0
        assert len(names) and pn.numel() == saved.numel()
        model.flat_prenorm.copy_(pn.to(torch.float32))
        params64 = orc.unflatten(model.flat_params.detach().cpu().double(), pn.double(), dtype=torch.float64)
        oracle = orc.OracleGCNN(params64, dtype=torch.float64)
        loss, pred, grads = orc.loss_and_grads(oracle, inputs, targets)
        vectors = inputs[:7] + (z["n_cons"], z["n_vars"], z["n_cuts"])  # per-sample counts: the block kernels
        for inp in (inputs, vectors):
            loss_sum, scores = model.loss_and_grads(inp, targets)
            torch.cuda.synchronize()
            assert rel_err(scores.cpu().numpy(), pred.numpy()) <= TOL
            assert_grads_close(model.flat_grads.cpu().numpy(), grads, tol=model.grad_tol)
    finally:
        model.flat_prenorm.copy_(saved)


def test_edge_order_invariance_and_determinism(model):
    samples = synth.make_samples("setcov", 2, seed0=5)
    base = batching.concat_samples(samples)
    shuf = batching.concat_samples([synth.shuffle_edges(s, 3 + i) for i, s in enumerate(samples)])
    with torch.no_grad():
        a = model(batching.model_inputs(base), False).cpu().numpy()
        a2 = model(batching.model_inputs(base), False).cpu().numpy()
        b = model(batching.model_inputs(shuf), False).cpu().numpy()
    np.testing.assert_array_equal(a, a2)  # fixed reduction order: bit-reproducible
    # stable sort restores (row, original order) grouping; with distinct columns inside a row the shuffled batch
    # sums the same terms in a different order only inside segments -> 1e-6
    assert rel_err(b, a) <= 1e-6
    model.loss_and_grads(batching.model_inputs(base), base[10])
    g1 = model.flat_grads.clone()
    model.loss_and_grads(batching.model_inputs(base), base[10])
    assert torch.equal(g1, model.flat_grads)


def test_batch_equals_single_graph_calls(model):
    samples = synth.make_samples("combauc", 3, seed0=77)
    whole = batching.concat_samples(samples)
    with torch.no_grad():
        out = model(batching.model_inputs(whole), False).cpu().numpy()
        parts = [model(batching.model_inputs(batching.concat_samples([s])), False).cpu().numpy() for s in samples]
    assert rel_err(out, np.concatenate(parts)) <= 1e-6


@pytest.mark.parametrize("shape,n", [("setcov", 3), ("combauc", 4), ("mini", 5), ("indset", 3)])
def test_per_sample_counts_select_block_kernels_same_results(model, shape, n):
    """Per-sample count vectors (what load_batch returns) select the shared-memory block edge kernels and the per-sample
    transposed layouts; scores and gradients must agree with the generic path (different summation order only)."""
    batch = batching.concat_samples(synth.make_samples(shape, n, seed0=4242))
    totals, vectors = batching.model_inputs(batch), batching.model_inputs(batch, per_sample_counts=True)
    assert model.prepare_inputs(vectors)[0].n_samples == n and model.prepare_inputs(totals)[0].n_samples == 0
    _, s_blocks = model.loss_and_grads(vectors, batch[10])
    g_blocks = model.flat_grads.clone()
    _, s_generic = model.loss_and_grads(totals, batch[10])
    g_generic = model.flat_grads.clone()
    assert rel_err(s_blocks.cpu().numpy(), s_generic.cpu().numpy()) <= 2e-6
    assert rel_err(g_blocks.cpu().numpy(), g_generic.cpu().numpy()) <= 1e-5
    # bit-reproducible, and switching the option off takes the generic path for the same inputs
    model.loss_and_grads(vectors, batch[10])
    assert torch.equal(g_blocks, model.flat_grads)
    model.set_option("blocks", 0)
    try:
        model.loss_and_grads(vectors, batch[10])
        assert torch.equal(g_generic, model.flat_grads)
    finally:
        model.set_option("blocks", 1)


def test_wrong_per_sample_counts_are_reported(model):
    from gcnn_cut_selector_b200 import InvalidArgumentError
    batch = list(batching.concat_samples(synth.make_samples("setcov", 2, seed0=1)))
    batch[8] = np.array([1500, 500], np.int32)  # sums still match, but sample 1's edges now leave its variable range
    with pytest.raises(InvalidArgumentError, match="sample"):
        model(batching.model_inputs(batch, per_sample_counts=True), False)
    with torch.no_grad():
        assert torch.isfinite(model(batching.model_inputs(batch), False)).all()
    # counts that do not add up to the totals are ignored (generic path), not an error
    good = batching.concat_samples(synth.make_samples("setcov", 2, seed0=1))
    ok = batching.model_inputs(good)
    dev, _keep = model.prepare_inputs(ok)
    wrong = [np.ascontiguousarray(x, np.int32) for x in (good[7], np.array([1500, 400]), good[9])]
    dev.sample_n_cons, dev.sample_n_vars, dev.sample_n_cuts = (x.ctypes.data for x in wrong)
    dev.n_samples = 2
    with torch.no_grad():
        a = model._forward((dev, _keep), save_activations=False)
        b = model(ok, False)
    assert torch.equal(a, b)


@pytest.mark.parametrize("case", ["setcov", "ragged", "empty_samples", "one_sample", "cuts"])
def test_block_transpose_bit_exact(model, case):
    """The per-sample counting sort (gcnn_build_csr_blocks) against numpy's stable argsort -- the same oracle as the radix
    sort: ptr, other, val and perm of both layouts bit for bit."""
    from gcnn_cut_selector_b200._lib import check
    rng = np.random.default_rng(3)
    if case == "setcov":
        samples = synth.make_samples("setcov", 5, seed0=60)
    elif case == "ragged":
        samples = synth.make_samples("combauc", 3, seed0=61) + synth.make_samples("mini", 4, seed0=62) + synth.make_samples("indset", 2, seed0=63)
    elif case == "one_sample":
        samples = synth.make_samples("setcov", 1, seed0=64)
    elif case == "cuts":
        samples = synth.make_samples("setcov", 4, seed0=65)
    else:  # samples without edges / without rows in between
        samples = synth.make_samples("mini", 3, seed0=66)
        (c, ce, v, k, ke), imp = samples[1]
        empty = {"indices": np.zeros((2, 0), np.int64), "values": np.zeros((0, 1))}
        samples[1] = ((c, empty, v, k, empty), imp)
    batch = batching.concat_samples(samples)
    which = 1 if case == "cuts" else 0
    ei, ef = (batch[5], batch[6][:, 0]) if which else (batch[1], batch[2][:, 0])
    n_left_s, n_vars_s = (batch[9] if which else batch[7]), batch[8]
    n_left, n_vars, E = int(n_left_s.sum()), int(n_vars_s.sum()), ei.shape[1]
    lib, dev = model._lib, model.device
    check(lib.gcnn_workspace_reserve(model._ws, n_left if not which else 1, n_vars, n_left if which else 1,
                                     E if not which else 1, E if which else 1, 1))
    d_ei = torch.from_numpy(np.ascontiguousarray(ei.astype(np.int32))).to(dev)
    d_ef = torch.from_numpy(np.ascontiguousarray(ef.astype(np.float32))).to(dev)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    cl, cv = np.ascontiguousarray(n_left_s, np.int32), np.ascontiguousarray(n_vars_s, np.int32)
    check(lib.gcnn_build_csr_blocks(model._ws, which, d_ei.data_ptr(), d_ef.data_ptr(), E, n_left, n_vars,
                                    cl.ctypes.data, cv.ctypes.data, len(cl), st))
    for side, n_owner in ((0, n_left), (1, n_vars)):
        ptr = torch.empty(n_owner + 1, dtype=torch.int32, device=dev)
        other = torch.empty(E, dtype=torch.int32, device=dev)
        val = torch.empty(E, dtype=torch.float32, device=dev)
        perm = torch.empty(E, dtype=torch.int32, device=dev)
        check(lib.gcnn_csr_export(model._ws, which, side, ptr.data_ptr(), other.data_ptr(), val.data_ptr(),
                                  perm.data_ptr(), st))
        check(lib.gcnn_check(model._ws, st))
        keys = ei[side].astype(np.int64)
        want_perm = np.argsort(keys, kind="stable")
        want_ptr = np.concatenate([[0], np.cumsum(np.bincount(keys, minlength=n_owner))])
        np.testing.assert_array_equal(ptr.cpu().numpy(), want_ptr.astype(np.int32))
        np.testing.assert_array_equal(perm.cpu().numpy(), want_perm.astype(np.int32))
        np.testing.assert_array_equal(other.cpu().numpy(), ei[1 - side][want_perm].astype(np.int32))
        np.testing.assert_array_equal(val.cpu().numpy(), ef.astype(np.float32)[want_perm])


def test_empty_cut_set_and_empty_edges(model, oracle64):
    (c, ce, v, k, ke), imp = synth.make_sample("mini", 21)
    empty = {"indices": np.zeros((2, 0), np.int64), "values": np.zeros((0, 1))}
    batch = batching.concat_samples([((c, ce, v, k, empty), imp)])
    inputs = batching.model_inputs(batch)
    loss_sum, scores = model.loss_and_grads(inputs, batch[10])
    loss, pred, grads = orc.loss_and_grads(oracle64, inputs, batch[10])
    assert rel_err(scores.cpu().numpy(), pred.numpy()) <= TOL
    assert_grads_close(model.flat_grads.cpu().numpy(), grads, tol=model.grad_tol)


# ---- train step, pretraining, weights stream -------------------------------------------------------------------------
def test_adam_train_steps_match_oracle(golden_dir):
    """(a) the fused Adam kernel against the oracle's Keras-Adam on IDENTICAL gradients (fp32 rounding only);
    (b) three full train steps against an independent oracle run: losses agree to 1e-5."""
    from gcnn_cut_selector_b200 import GCNN
    path = os.path.join(golden_dir, "state_stream.pkl")
    m = GCNN(device="cuda:0", seed=1)
    m.restore_state(path)
    o_same = orc.OracleGCNN(orc.restore_state(path, dtype=torch.float64), dtype=torch.float64)
    o_free = orc.OracleGCNN(orc.restore_state(path, dtype=torch.float64), dtype=torch.float64)
    st_same, st_free = orc.AdamState(), orc.AdamState()
    lr = 1e-3
    for step in range(3):
        batch = batching.concat_samples(synth.make_samples("mini", 2, seed0=300 + step))
        inputs, targets = batching.model_inputs(batch), batch[10]
        loss, _ = m.train_step(inputs, targets, lr)
        g = m.flat_grads.cpu().numpy().astype(np.float64)
        grads, off = {}, 0
        for name, shape in orc.TRAINABLE:
            k = int(np.prod(shape))
            grads[name] = torch.from_numpy(g[off:off + k].reshape(shape))
            off += k
        orc.adam_step(o_same, st_same, grads, lr)
        got = m.flat_params.detach().cpu().numpy().astype(np.float64)
        want = orc.flatten_trainable(o_same.params).numpy()
        assert np.abs(got - want).max() <= 5e-7, f"step {step}: {np.abs(got - want).max():.3e}"
        oloss, _ = orc.train_step(o_free, st_free, inputs, targets, lr)
        assert abs(float(loss) - float(oloss)) <= 1e-5 * float(oloss)


def test_data_parallel_trainer_single_rank_equals_train_step(golden_dir):
    """world size 1: un-normalised seed + division by the (reduced) cut count inside Adam == the plain train step."""
    from gcnn_cut_selector_b200 import GCNN, DataParallelTrainer
    path = os.path.join(golden_dir, "state_stream.pkl")
    a, b = GCNN(device="cuda:0", seed=1), GCNN(device="cuda:0", seed=2)
    a.restore_state(path); b.restore_state(path)
    trainer = DataParallelTrainer(b, lr=1e-3)
    for step in range(2):
        batch = batching.concat_samples(synth.make_samples("mini", 2, seed0=500 + step))
        inputs, targets = batching.model_inputs(batch), batch[10]
        loss_a, _ = a.train_step(inputs, targets, 1e-3)
        loss_b = trainer.step(inputs, targets)
        assert abs(float(loss_a) - float(loss_b)) <= 1e-6 * abs(float(loss_a))
        ga, gb = a.flat_grads.cpu().numpy(), (b.flat_grads / trainer.bucket[trainer.N]).cpu().numpy()
        assert rel_err(gb, ga) <= 1e-6
    # identical gradients up to one rounding -> parameters agree to a few ulps of the update
    assert (a.flat_params.detach() - b.flat_params.detach()).abs().max().item() <= 2e-6


def test_pretrain_protocol_matches_golden(golden_dir):
    from gcnn_cut_selector_b200 import GCNN
    z = np.load(os.path.join(golden_dir, "pretrain_tiny.npz"))
    m = GCNN(device="cuda:0", seed=2)
    m.restore_state(os.path.join(golden_dir, "state_stream.pkl"))
    batches = [golden_inputs(z, f"b{b}_") for b in range(2)]
    m.pretrain_init()
    n = 0
    while True:  # model_trainer.py:207-234
        for b in batches:
            if not m.pretrain(b, True):
                break
        if m.pretrain_next() is None:
            break
        n += 1
    assert n == 11
    got = m.flat_prenorm.cpu().numpy().astype(np.float64)
    assert rel_err(got, z["prenorm_f64"]) <= 2e-6  # measured 1.6e-7 (the reference's own fp32 run: 5.8e-8)
    with torch.no_grad():
        out = m(batches[0], False)
    assert rel_err(out.cpu().numpy(), z["scores_f64"]) <= TOL  # measured 8.4e-7


def test_fused_pretraining_equals_the_eleven_pass_protocol(golden_dir):
    """Seven passes (input layers together) freeze exactly the values the reference's one-layer-per-pass loop does."""
    from gcnn_cut_selector_b200 import GCNN
    z = np.load(os.path.join(golden_dir, "pretrain_tiny.npz"))
    batches = [golden_inputs(z, f"b{b}_") for b in range(2)]
    out = []
    for fused in (False, True, "dp"):
        m = GCNN(device="cuda:0", seed=2)
        m.restore_state(os.path.join(golden_dir, "state_stream.pkl"))
        if fused == "dp":  # world size 1: the data-parallel schedule degenerates to the fused one
            from gcnn_cut_selector_b200 import DataParallelTrainer
            assert DataParallelTrainer(m, 1e-3).pretrain_fused(batches) == 7
        elif fused:
            assert m.pretrain_fused(batches) == 7
        else:
            m.pretrain_init()
            passes = 0
            while True:  # model_trainer.py:207-234
                for b in batches:
                    if not m.pretrain(b, True):
                        break
                if m.pretrain_next() is None:
                    break
                passes += 1
            assert passes == 11
        out.append(m.flat_prenorm.clone())
    assert torch.equal(out[0], out[1]) and torch.equal(out[0], out[2])


def test_save_restore_roundtrip(model, golden_dir, tmp_path):
    path = str(tmp_path / "w.pkl")
    model.save_state(path)
    assert open(path, "rb").read() == open(os.path.join(golden_dir, "state_stream.pkl"), "rb").read()
    assert model.variables_topological_order == [n for n, _, _ in orc.PARAM_SPECS]
    assert len(model.trainable_variables) == 46 and len(model.variables) == 62


def test_host_entry_points_match_device_path(golden_dir):
    from gcnn_cut_selector_b200 import GCNN, HostBatch
    path = os.path.join(golden_dir, "state_stream.pkl")
    batch = batching.concat_samples(synth.make_samples("setcov", 2, seed0=9))
    a, b = GCNN(device="cuda:0", seed=3), GCNN(device="cuda:0", seed=4)
    a.restore_state(path); b.restore_state(path)
    hb = HostBatch(batch)
    same_inputs = batching.model_inputs(batch, per_sample_counts=True)  # HostBatch carries the per-sample counts too
    with torch.no_grad():
        dev_scores = a(same_inputs, False).cpu().numpy()
    np.testing.assert_array_equal(b.score_host(hb), dev_scores)
    loss_a, _ = a.train_step(same_inputs, batch[10], 1e-3)
    loss_b = b.train_step_host(hb, 1e-3)
    assert abs(float(loss_a) - loss_b) <= 1e-6 * abs(loss_b)
    torch.testing.assert_close(a.flat_params.detach(), b.flat_params.detach(), rtol=0, atol=0)


def test_prefetching_host_path_matches_synchronous_path(golden_dir):
    """Three train steps and a scoring call with one batch kept in flight (stage slot i + 1, then step on slot i) give
    bit-identical parameters, losses and scores to the one-batch-at-a-time host calls."""
    from gcnn_cut_selector_b200 import GCNN, HostBatch
    path = os.path.join(golden_dir, "state_stream.pkl")
    batches = [batching.concat_samples(synth.make_samples("setcov", 2 + i, seed0=20 + i)) for i in range(3)]
    a, b = GCNN(device="cuda:0", seed=3), GCNN(device="cuda:0", seed=4)
    a.restore_state(path); b.restore_state(path)
    ha, hb = [HostBatch(x) for x in batches], [HostBatch(x) for x in batches]
    for h in hb:  # reserve for the largest batch first: growing the workspace invalidates staged slots
        b.reserve(h.batch, True)
    losses_a = [a.train_step_host(h, 1e-3) for h in ha]
    b.stage_host(hb[0], 0)
    losses_b = []
    for i in range(3):
        if i + 1 < 3:
            b.stage_host(hb[i + 1], (i + 1) & 1)
        losses_b.append(b.train_step_staged(i & 1, 1e-3))
    assert losses_a == losses_b
    torch.testing.assert_close(a.flat_params.detach(), b.flat_params.detach(), rtol=0, atol=0)
    # same again with the loss of step i read only after step i + 1 is enqueued
    c = GCNN(device="cuda:0", seed=5)
    c.restore_state(path)
    for h in hb:
        c.reserve(h.batch, True)
    c.stage_host(hb[0], 0)
    losses_c = []
    for i in range(3):
        if i + 1 < 3:
            c.stage_host(hb[i + 1], (i + 1) & 1)
        c.train_step_staged_async(i & 1, 1e-3)
        if i > 0:
            losses_c.append(c.train_step_result((i - 1) & 1))
    losses_c.append(c.train_step_result(2 & 1))
    assert losses_a == losses_c
    torch.testing.assert_close(a.flat_params.detach(), c.flat_params.detach(), rtol=0, atol=0)
    b.stage_host(hb[1], 0, training=False)
    np.testing.assert_array_equal(b.score_staged(0), a.score_host(ha[1]))


def test_row_pointer_host_batches_are_bit_identical(golden_dir):
    """A host batch whose sorted edge lists travel as row pointer + columns (gcnn_batch::cons_row_ptr / cut_row_ptr, 8 instead
    of 12 bytes per edge over PCIe) gives the same scores, losses and parameters, bit for bit, as one that copies [2, E];
    an unsorted list keeps the full copy; a pointer that does not span the list is refused."""
    from gcnn_cut_selector_b200 import GCNN, HostBatch, InvalidArgumentError
    path = os.path.join(golden_dir, "state_stream.pkl")
    batches = [list(batching.concat_samples(synth.make_samples("setcov", 2 + i, seed0=30 + i))) for i in range(2)]
    perm = np.random.default_rng(0).permutation(batches[1][5].shape[1])  # second batch: cut edges in random order
    batches[1][5], batches[1][6] = batches[1][5][:, perm], batches[1][6][perm]
    a, b = GCNN(device="cuda:0", seed=3), GCNN(device="cuda:0", seed=4)
    a.restore_state(path); b.restore_state(path)
    ha = [HostBatch(tuple(x), row_pointers=False, packed=False) for x in batches]  # thirteen plain copies
    hb = [HostBatch(tuple(x), row_pointers=True) for x in batches]  # one packed buffer: features, pointers, local columns
    assert HostBatch(tuple(batches[0])).row_ptrs == [None, None]  # default: only lists of >= 128 k edges
    assert hb[0].row_ptrs[0] is not None and hb[0].row_ptrs[1] is not None and ha[0].row_ptrs == [None, None]
    assert hb[1].row_ptrs[0] is not None and hb[1].row_ptrs[1] is None
    assert hb[0].col16[0] is not None and hb[0].col16[1] is not None and hb[1].col16[1] is None  # uint16 local columns
    assert HostBatch(tuple(batches[0]), row_pointers=True, packed=False).h2d_bytes == ha[0].h2d_bytes - 6 * (
        batches[0][1].shape[1] + batches[0][5].shape[1]) + 4 * (hb[0].batch.n_cons + 1 + hb[0].batch.n_cuts + 1)
    assert hb[0].batch.packed_bytes == hb[0].h2d_bytes and hb[0].h2d_bytes < ha[0].h2d_bytes
    hp = [HostBatch(tuple(x), row_pointers=False) for x in batches]  # packed, full index tensors
    for x, y in zip(ha, hp):
        np.testing.assert_array_equal(a.score_host(x), b.score_host(y))
    totals = list(batches[0])
    totals[7:10] = [int(np.sum(x)) for x in totals[7:10]]  # no per-sample counts: row pointers, full column indices
    ht = HostBatch(tuple(totals), row_pointers=True)
    assert ht.row_ptrs[0] is not None and ht.col16 == [None, None]
    np.testing.assert_array_equal(a.score_host(HostBatch(tuple(totals), row_pointers=False)), b.score_host(ht))
    for x, y in zip(ha, hb):
        np.testing.assert_array_equal(a.score_host(x), b.score_host(y))
        assert a.train_step_host(x, 1e-3) == b.train_step_host(y, 1e-3)
    torch.testing.assert_close(a.flat_params.detach(), b.flat_params.detach(), rtol=0, atol=0)
    bad = HostBatch(tuple(batches[0]), row_pointers=True)
    bad.row_ptrs[0][-1] -= 1
    with pytest.raises(InvalidArgumentError):
        b.score_host(bad)
    bad = HostBatch(tuple(batches[0]), row_pointers=True)  # ends fine, not monotone inside: found on the device
    bad.row_ptrs[0][5], bad.row_ptrs[0][6] = int(bad.row_ptrs[0][6]) + 3, int(bad.row_ptrs[0][5])
    with pytest.raises(InvalidArgumentError):
        b.score_host(bad)
    np.testing.assert_array_equal(a.score_host(ha[0]), b.score_host(hb[0]))  # (the sticky error word was cleared)


def test_head_in_the_last_chain_equals_the_separate_head_kernels(golden_dir):
    """Option "head_in_chain": the head's Dense(1), the MSE seed and that layer's backward computed in the epilogue of the
    cut convolution's forward chain (default) against the launches of their own -- same scores, loss and gradients to
    fp32 rounding (another summation order inside the 64-term dot product), a launch fewer; training and inference scores
    bit-identical in either mode; a cut count that is not a multiple of the 128-node tile and more than one tile."""
    from gcnn_cut_selector_b200 import GCNN
    path = os.path.join(golden_dir, "state_stream.pkl")
    m = GCNN(device="cuda:0", seed=3)
    m.restore_state(path)
    for shape, n in (("setcov", 3), ("combauc", 5), ("mini", 2)):
        batch = batching.concat_samples(synth.make_samples(shape, n, seed0=70 + n))
        inputs, targets = batching.model_inputs(batch, per_sample_counts=True), batch[10]
        got = {}
        m.loss_and_grads(inputs, targets)  # (packs the weight images of this parameter epoch: not counted below)
        for mode in (1, 0):
            m.set_option("head_in_chain", mode)
            n0 = m._lib.gcnn_kernel_launches()
            loss_sum, scores = m.loss_and_grads(inputs, targets)
            launches = m._lib.gcnn_kernel_launches() - n0
            with torch.no_grad():
                inference = m(inputs, False)
            if mode == 1:  # one code path computes the score in training and inference
                torch.testing.assert_close(inference, scores, rtol=0, atol=0)
            else:
                assert rel_err(inference.cpu().numpy(), scores.cpu().numpy()) <= 1e-6
            got[mode] = (float(loss_sum), scores.cpu().numpy().copy(), m.flat_grads.cpu().numpy().copy(), launches)
        m.set_option("head_in_chain", 1)
        assert got[0][3] == got[1][3] + 1, f"launches {got[0][3]} vs {got[1][3]}"
        assert abs(got[0][0] - got[1][0]) <= 1e-5 * abs(got[0][0]), f"loss {got[0][0]} vs {got[1][0]}"
        assert rel_err(got[1][1], got[0][1]) <= 1e-6, f"scores {rel_err(got[1][1], got[0][1])}"
        g_err = np.linalg.norm(got[1][2] - got[0][2]) / np.linalg.norm(got[0][2])
        assert g_err <= 1e-5, f"gradients {g_err}"


def test_params_epoch_reuses_and_refreshes_weight_images(golden_dir):
    """Option "params_epoch": scoring with frozen weights packs the weight images once (one launch fewer per call); every
    way the parameters can change -- a torch write, restore_state, the library's fused step, apply_gradients, a raw write
    announced with params_changed() -- is followed by scores that equal, bit for bit, those of a fresh model holding the
    same parameters (eager and graph-replayed scoring alike)."""
    from gcnn_cut_selector_b200 import GCNN, HostBatch
    path = os.path.join(golden_dir, "state_stream.pkl")
    batch = batching.concat_samples(synth.make_samples("setcov", 2, seed0=41))
    hb, hb2 = HostBatch(batch), HostBatch(batch)
    m = GCNN(device="cuda:0", seed=3)
    m.restore_state(path)

    def fresh_scores():
        f = GCNN(device="cuda:0", seed=9)
        with torch.no_grad():
            f.flat_params.copy_(m.flat_params)
            f.flat_prenorm.copy_(m.flat_prenorm)
        return f.score_host(hb2).copy()

    def check_both():
        want = fresh_scores()
        np.testing.assert_array_equal(m.score_host(hb), want)
        np.testing.assert_array_equal(m.score_host(hb, graph=True), want)

    launches = []
    for _ in range(3):
        n0 = m._lib.gcnn_kernel_launches()
        m.score_host(hb)
        launches.append(m._lib.gcnn_kernel_launches() - n0)
    assert launches[1] == launches[0] - 1 and launches[2] == launches[1]  # the pack ran once
    for _ in range(3):
        m.score_host(hb, graph=True)
    assert m._lib.gcnn_serve_graph_count(m._ws) >= 1
    check_both()
    with torch.no_grad():
        m.flat_params.mul_(1.01)  # a torch write (version counter)
    check_both()
    with torch.no_grad():
        m.trainable_variables[3].add_(0.01)  # ... through a view
    check_both()
    m.train_step_host(hb, 1e-3)  # the library's own Adam
    check_both()
    m.train_step(batching.model_inputs(batch, per_sample_counts=True), batch[10], 1e-3)  # gcnn_adam_step via apply_gradients
    check_both()
    m.restore_state(path)
    check_both()
    n0 = m._lib.gcnn_kernel_launches()
    m.score_host(hb)  # (frozen since check_both: no pack)
    m.params_changed()
    n1 = m._lib.gcnn_kernel_launches()
    m.score_host(hb)
    assert m._lib.gcnn_kernel_launches() - n1 == (n1 - n0) + 1  # announced raw write: packed again
    m.set_option("params_epoch", 0)  # promise withdrawn: every call packs
    m._announce_params = lambda: None
    n2 = m._lib.gcnn_kernel_launches()
    m.score_host(hb)
    assert m._lib.gcnn_kernel_launches() - n2 == (n1 - n0) + 1


# ---- full BASELINE sizes: size-independent properties -----------------------------------------------------------------
@pytest.mark.parametrize("n_graphs", [32, 128])  # BASELINE config 2, and config 4's per-GPU share at 8 GPUs (3.2 M edges)
def test_config2_properties(model, n_graphs):
    """Full-size setcov batches: block-diagonality -> the batch equals its two halves; bit-reproducible."""
    h = n_graphs // 2
    samples = synth.make_samples("setcov", n_graphs, seed0=2000, n_structures=8)
    whole = batching.concat_samples(samples)
    with torch.no_grad():
        out = model(batching.model_inputs(whole), False)
        out2 = model(batching.model_inputs(whole), False)
        halves = [model(batching.model_inputs(batching.concat_samples(samples[i:i + h])), False) for i in (0, h)]
    assert torch.equal(out, out2)
    assert rel_err(out.cpu().numpy(), torch.cat(halves).cpu().numpy()) <= 1e-6
    assert torch.isfinite(out).all()
    # gradient of the whole batch = cut-count-weighted mean of the halves' gradients (linearity of the MSE mean)
    model.loss_and_grads(batching.model_inputs(whole), whole[10])
    g = model.flat_grads.clone()
    acc = torch.zeros_like(g)
    for i in (0, h):
        half = batching.concat_samples(samples[i:i + h])
        model.loss_and_grads(batching.model_inputs(half), half[10], seed_scale=1.0 / whole[4].shape[0])
        acc += model.flat_grads
    assert rel_err(acc.cpu().numpy(), g.cpu().numpy()) <= 1e-5
    model.loss_and_grads(batching.model_inputs(whole), whole[10])
    assert torch.equal(g, model.flat_grads)  # gradients are bit-reproducible too (fixed reduction orders everywhere)


def test_miplib_scale_forward(model):
    """BASELINE config 5 (100k x 100k, 1M edges, heavy-tailed rows): finite, reproducible, edge-order invariant."""
    sample = synth.make_sample("miplib", 5)
    batch = batching.concat_samples([sample])
    with torch.no_grad():
        a = model(batching.model_inputs(batch), False)
        b = model(batching.model_inputs(batching.concat_samples([synth.shuffle_edges(sample, 1)])), False)
    assert a.shape == (5000,) and torch.isfinite(a).all()
    assert rel_err(b.cpu().numpy(), a.cpu().numpy()) <= 1e-5


# ---- full BASELINE sizes against the fp64 oracle (run here, on the box's host cores) -------------------------------------
_ORACLE_CACHE = {}


def _oracle_at_size(golden_dir, key, inputs, targets=None):
    """fp64 oracle results at BASELINE sizes, computed once per session (seconds to a minute of host time each).  The
    hoisted form (sum, then one Dense per receiving node) is the same function as the reference's per-edge Dense in exact
    arithmetic and is pinned to the reference-generated fixtures in fp64 like it (tests/test_oracle_golden.py); it keeps the [E, 64] fp64
    intermediates of a 3.2 M-edge batch within a few GB."""
    if key not in _ORACLE_CACHE:
        torch.set_num_threads(os.cpu_count() or 1)
        o = orc.OracleGCNN(orc.restore_state(os.path.join(golden_dir, "state_stream.pkl"), dtype=torch.float64),
                           dtype=torch.float64, faithful=False)
        if targets is None:
            with torch.no_grad():
                _ORACLE_CACHE[key] = (None, o.call(inputs).numpy(), None)
        else:
            loss, pred, grads = orc.loss_and_grads(o, inputs, targets)
            _ORACLE_CACHE[key] = (float(loss), pred.numpy(), grads)
    return _ORACLE_CACHE[key]


@pytest.mark.parametrize("counts", [False, True], ids=["totals", "per_sample_counts"])
def test_config2_scores_and_gradients_match_oracle(model, golden_dir, counts):
    """BASELINE config 2 at full size (32 setcov graphs: 16,000 x 32,000 nodes, 800,000 + 204,800 edges): scores, loss and
    every parameter gradient against the fp64 oracle on the same batch."""
    batch = batching.concat_samples(synth.make_samples("setcov", 32, seed0=2000, n_structures=8))
    loss, pred, grads = _oracle_at_size(golden_dir, "config2", batching.model_inputs(batch), batch[10])
    loss_sum, scores = model.loss_and_grads(batching.model_inputs(batch, per_sample_counts=counts), batch[10])
    torch.cuda.synchronize()
    assert rel_err(scores.cpu().numpy(), pred) <= TOL
    assert abs(float(loss_sum) / scores.numel() - loss) <= TOL * loss
    assert_grads_close(model.flat_grads.cpu().numpy(), grads, tol=model.grad_tol)


@pytest.mark.parametrize("counts", [False, True], ids=["totals", "per_sample_counts"])
def test_config4_share_scores_match_oracle(model, golden_dir, counts):
    """BASELINE config 4's per-GPU share at 8 GPUs (128 setcov graphs, 3.2 M + 0.8 M edges): scores against the fp64 oracle."""
    batch = batching.concat_samples(synth.make_samples("setcov", 128, seed0=4000, n_structures=8))
    _, pred, _ = _oracle_at_size(golden_dir, "config4", batching.model_inputs(batch))
    with torch.no_grad():
        out = model(batching.model_inputs(batch, per_sample_counts=counts), False)
    assert rel_err(out.cpu().numpy(), pred) <= TOL


def test_config5_miplib_scores_match_oracle(model, golden_dir):
    """BASELINE config 5 (100,000 x 100,000 nodes, 1 M + 0.5 M edges, heavy-tailed rows reduced by whole CTAs, 9-bit radix
    digits): the 5,000 cut scores against the fp64 oracle."""
    batch = batching.concat_samples([synth.make_sample("miplib", 5)])
    _, pred, _ = _oracle_at_size(golden_dir, "config5", batching.model_inputs(batch))
    with torch.no_grad():
        out = model(batching.model_inputs(batch), False)
        out_counts = model(batching.model_inputs(batch, per_sample_counts=True), False)  # one block, too large: generic path
    assert rel_err(out.cpu().numpy(), pred) <= TOL
    assert torch.equal(out, out_counts)


@pytest.mark.parametrize("shape", ["combauc", "capfac", "indset"])
@pytest.mark.parametrize("n", [1, 4])
def test_config3_scoring_matches_oracle(model, oracle64, shape, n):
    """BASELINE config 3: inference cut scoring on the three other problem classes, batch 1 (the plugin path,
    model_benchmarker.py:106) and batch 4 (model_tester.py:51), device and host entry points, with and without counts."""
    from gcnn_cut_selector_b200 import HostBatch
    batch = batching.concat_samples(synth.make_samples(shape, n, seed0=3000 + n))
    with torch.no_grad():
        pred = oracle64.call(batching.model_inputs(batch)).numpy()
        a = model(batching.model_inputs(batch), False).cpu().numpy()
        b = model(batching.model_inputs(batch, per_sample_counts=True), False).cpu().numpy()
    c = model.score_host(HostBatch(batch)).copy()
    for got in (a, b, c):
        assert rel_err(got, pred) <= TOL


# ---- regressions for the round-1 review findings ---------------------------------------------------------------------------
def test_growing_the_workspace_keeps_a_staged_batch(golden_dir):
    """A larger batch i + 1 is reserved and staged while batch i is still staged (the prefetch loop with ragged batches):
    batch i must survive and train exactly as it does without the growth."""
    from gcnn_cut_selector_b200 import GCNN, HostBatch
    path = os.path.join(golden_dir, "state_stream.pkl")
    batches = [batching.concat_samples(synth.make_samples("setcov", k, seed0=30 + k)) for k in (1, 3, 2)]
    a, b = GCNN(device="cuda:0", seed=3), GCNN(device="cuda:0", seed=4)
    a.restore_state(path); b.restore_state(path)
    ha, hb = [HostBatch(x) for x in batches], [HostBatch(x) for x in batches]
    losses_a = [a.train_step_host(h, 1e-3) for h in ha]
    b.stage_host(hb[0], 0)
    losses_b = []
    for i in range(3):
        if i + 1 < 3:
            b.stage_host(hb[i + 1], (i + 1) & 1)  # reserves for the (larger) next batch while batch i is staged
        losses_b.append(b.train_step_staged(i & 1, 1e-3))
    assert losses_a == losses_b and a.adam_step == b.adam_step == 3
    torch.testing.assert_close(a.flat_params.detach(), b.flat_params.detach(), rtol=0, atol=0)


def test_stale_activations_are_rejected(model, golden_dir):
    """A second forward on the same model before backward() overwrites the shared activations: the bridge refuses."""
    z = np.load(os.path.join(golden_dir, "fwd_mini2.npz"))
    y = torch.from_numpy(z["targets"]).to(model.device)
    pred = model(golden_inputs(z), True)
    with torch.no_grad():
        model(golden_inputs(z), False)  # an evaluation forward in between
    with pytest.raises(RuntimeError, match="overwritten"):
        ((y - pred) ** 2).mean().backward()
    model.flat_params.grad = None
    pred = model(golden_inputs(z), True)
    ((y - pred) ** 2).mean().backward()  # undisturbed: fine
    model.flat_params.grad = None


def test_training_calls_report_bad_indices(model, golden_dir):
    from gcnn_cut_selector_b200 import InvalidArgumentError
    z = np.load(os.path.join(golden_dir, "fwd_tiny3.npz"))
    inputs = list(golden_inputs(z))
    bad = inputs[1].copy()
    bad[0, 2] = -1
    inputs[1] = bad
    with pytest.raises(InvalidArgumentError):
        model.loss_and_grads(tuple(inputs), z["targets"])
    model.loss_and_grads(golden_inputs(z), z["targets"])  # the workspace stays usable


def test_seed_none_follows_the_default_generator():
    from gcnn_cut_selector_b200 import GCNN
    torch.manual_seed(123)
    a = GCNN(device="cuda:0").flat_params.detach().clone()
    after = torch.rand(1)
    torch.manual_seed(123)
    b = GCNN(device="cuda:0").flat_params.detach().clone()
    assert torch.equal(a, b) and torch.equal(after, torch.rand(1))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_model_on_a_non_current_device(golden_dir):
    from gcnn_cut_selector_b200 import GCNN
    z = np.load(os.path.join(golden_dir, "fwd_mini2.npz"))
    torch.cuda.set_device(0)
    m = GCNN(device="cuda:1", seed=0)
    m.restore_state(os.path.join(golden_dir, "state_stream.pkl"))
    with torch.no_grad():
        out = m(golden_inputs(z), False)
    assert out.device.index == 1 and rel_err(out.cpu().numpy(), z["scores_f64"]) <= TOL
    assert torch.cuda.current_device() == 0


# ---- ranking accuracy of the training loop (model_trainer.py:279-302) ----------------------------------------------------
def test_ranking_deviation_bit_exact(golden_dir):
    """The device kernel against the oracle's line-by-line restatement (itself pinned to the reference's own `process`
    by tests/test_oracle_golden.py): integer deviations bit-exact, ties, single-cut samples, one large sample."""
    from gcnn_cut_selector_b200 import ranking_accuracy, ranking_deviation
    fr = np.array([0.25, 0.5, 0.75, 1])
    z = np.load(os.path.join(golden_dir, "metric_process.npz"))
    acc_total, n_samples = np.zeros(4), 0
    for b in range(int(z["n_batches"])):
        pred, true, n_cuts = z[f"pred{b}"], z[f"true{b}"], z[f"n_cuts{b}"]
        want_acc, want_dev = orc.ranking_accuracy(pred, true, n_cuts, fr)
        got = ranking_deviation(torch.from_numpy(pred).cuda(), true, n_cuts)
        np.testing.assert_array_equal(got.cpu().numpy(), want_dev)
        acc = ranking_accuracy(torch.from_numpy(pred).cuda(), true, n_cuts, fr)
        np.testing.assert_array_equal(acc, want_acc)
        acc_total += acc
        n_samples += len(n_cuts)
    np.testing.assert_allclose(acc_total / n_samples, z["mean_acc"], rtol=0, atol=1e-12)
    rng = np.random.default_rng(5)
    n_cuts = np.array([1, 64, 1, 7000, 333, 2], dtype=np.int64)  # 7000 cuts: the > 48 KB shared-memory path
    total = int(n_cuts.sum())
    true = np.round(rng.uniform(0, 1, total), 3).astype(np.float32)
    pred = true.copy()
    flip = rng.random(total) < 0.001
    pred[flip] = np.round(rng.uniform(0, 1, int(flip.sum())), 3)
    want_acc, want_dev = orc.ranking_accuracy(pred, true, n_cuts, fr)
    got = ranking_deviation(torch.from_numpy(pred).cuda(), torch.from_numpy(true).cuda(), n_cuts)
    np.testing.assert_array_equal(got.cpu().numpy(), want_dev)


# ---- cut selection after scoring (model_benchmarker.py:108-157) ---------------------------------------------------------
def test_select_cuts_bit_exact(golden_dir):
    """The device kernel against the reference's own ``CustomCutsel.cutselselect`` (golden fixture) and, on larger random
    cases with ties, forced cuts and a 3,000-cut candidate set, against the oracle's line-by-line restatement."""
    from gcnn_cut_selector_b200 import select_cuts
    z = np.load(os.path.join(golden_dir, "selector.npz"))
    for name in z["names"]:
        pf = z[f"{name}_par_forced"]
        order, n_sel = select_cuts(z[f"{name}_quality"], z[f"{name}_par"], pf if len(pf) else None,
                                   max_selected=int(z[f"{name}_max_selected"]))
        np.testing.assert_array_equal(order.cpu().numpy(), z[f"{name}_order"])
        assert int(n_sel.item()) == int(z[f"{name}_n_selected"])
    rng = np.random.default_rng(9)
    for n, nf, sparsity in ((257, 3, 0.9), (1500, 0, 0.97), (3000, 2, 0.995)):
        q = np.round(rng.uniform(0, 1, n), 3).astype(np.float32)
        par = rng.uniform(0, 1, (n, n)).astype(np.float32)
        par = np.maximum(par, par.T)
        mask = rng.random((n, n)) < sparsity
        par[np.maximum(mask, mask.T)] = 0.0
        pf = rng.uniform(0, 1, (nf, n)).astype(np.float32)
        pf[rng.random((nf, n)) < 0.9] = 0.0
        want_order, want_n = orc.select_cuts(q, par, pf if nf else None, max_selected=n // 2)
        order, n_sel = select_cuts(torch.from_numpy(q).cuda(), par, pf if nf else None, max_selected=n // 2)
        np.testing.assert_array_equal(order.cpu().numpy(), want_order)
        assert int(n_sel.item()) == want_n


def test_graph_scoring_matches_eager_scoring(golden_dir):
    """The serving path: host arrays in, host scores out, replayed as one CUDA graph per shape from the third call on
    (first call eager, second captured).  Bit-identical to the eager host path; new shapes, changed inputs of a cached
    shape and a workspace growth in between are handled."""
    from gcnn_cut_selector_b200 import GCNN, HostBatch
    m = GCNN(device="cuda:0", seed=0)
    m.restore_state(os.path.join(golden_dir, "state_stream.pkl"))
    batches = [batching.concat_samples(synth.make_samples(shape, 1, seed0=s)) for shape, s in
               (("combauc", 1), ("combauc", 2), ("indset", 3), ("mini", 4))]
    hbs = [HostBatch(b) for b in batches]
    eager = [m.score_host(h).copy() for h in hbs]
    want = [m.score_host(h, graph=True).copy() for h in hbs]  # first call of a shape: the same launches, not yet captured
    for w, e in zip(want, eager):  # (score_host uses the batch's per-sample counts -> block kernels: another summation order)
        assert rel_err(w, e) <= 2e-6
    for rep in range(4):
        for h, w in zip(hbs, want):
            np.testing.assert_array_equal(m.score_host(h, graph=True), w)
    # (0 would mean the capture fell back to eager launches; the library leaves the reason in its error string)
    assert m._lib.gcnn_serve_graph_count(m._ws) >= 1, m._lib.gcnn_last_error().decode()
    # more shapes than graph slots (8): the least recently used graphs are evicted, results stay exact
    many = [HostBatch(batching.concat_samples(synth.make_samples("tiny", k, seed0=50 + k))) for k in range(1, 12)]
    first = [m.score_host(h, graph=True).copy() for h in many]  # (sizes grow: every call also grows the workspace)
    for h, w in zip(many, first):  # three calls in a row per shape: eager (new slot), captured, replayed
        for rep in range(3):
            np.testing.assert_array_equal(m.score_host(h, graph=True), w)
    assert m._lib.gcnn_serve_graph_count(m._ws) == 8  # eleven shapes went through eight slots
    for h, w in zip(many[::-1], first[::-1]):  # the last eight are cached, the first three come back eagerly
        np.testing.assert_array_equal(m.score_host(h, graph=True), w)
        assert rel_err(w, m.score_host(h)) <= 2e-6
    big = HostBatch(batching.concat_samples(synth.make_samples("setcov", 2, seed0=5)))  # grows the workspace
    w_big = m.score_host(big, graph=True).copy()
    assert rel_err(w_big, m.score_host(big)) <= 2e-6
    for rep in range(3):
        np.testing.assert_array_equal(m.score_host(big, graph=True), w_big)
        np.testing.assert_array_equal(m.score_host(hbs[0], graph=True), want[0])
    bad = list(batches[0])
    ei = bad[5].copy()
    ei[1, 0] = int(bad[8].sum())  # out-of-range variable index on a cached shape
    bad[5] = ei
    from gcnn_cut_selector_b200 import InvalidArgumentError
    with pytest.raises(InvalidArgumentError):
        m.score_host(HostBatch(tuple(bad)), graph=True)
    np.testing.assert_array_equal(m.score_host(hbs[0], graph=True), want[0])


# ---- bf16 MLP path: the 1e-2 accuracy class of BASELINE.json -------------------------------------------------------------
BF16_TOL = 1e-2


BF16_SINGLE_TOL = 3e-2  # one product per MMA: measured 1.2-1.3e-2 -- outside BASELINE's 1e-2 class, tested as what it is


@pytest.mark.parametrize("counts", [False, True], ids=["totals", "per_sample_counts"])
@pytest.mark.parametrize("shape,n", [("setcov", 3), ("combauc", 4), ("indset", 4), ("capfac", 1), ("setcov", 32)])
@pytest.mark.parametrize("precision,tol", [(1, BF16_TOL), (2, BF16_SINGLE_TOL)], ids=["three_products", "one_product"])
def test_bf16_mlp_mode_within_1e2(golden_dir, oracle64, shape, n, counts, precision, tol):
    """option "precision" = 1: the dense layers run as bf16 tensor-core products with fp32 accumulation, three products
    per MMA (hi*hi + hi*lo + lo*hi) instead of the six of the fp32-accurate path; edge kernels, loss and Adam stay fp32.
    Scores, loss and every parameter gradient within 1e-2 of the fp64 oracle (relative to the tensor's max-abs / L2) on
    all four problem classes and at BASELINE config 2's size; the mode is off by default and switches back cleanly.
    "precision" = 2 (one product, operands rounded to bf16) is held to 3e-2: it does NOT meet the 1e-2 class."""
    from gcnn_cut_selector_b200 import GCNN
    m = GCNN(device="cuda:0", seed=0)
    m.restore_state(os.path.join(golden_dir, "state_stream.pkl"))
    batch = batching.concat_samples(synth.make_samples(shape, n, seed0=1000 + n, n_structures=min(n, 8)))
    inputs, targets = batching.model_inputs(batch, per_sample_counts=counts), batch[10]
    if n == 32:
        loss, pred, grads = _oracle_at_size(golden_dir, "bf16_c2", batching.model_inputs(batch), targets)
    else:
        l, p, grads = orc.loss_and_grads(oracle64, batching.model_inputs(batch), targets)
        loss, pred = float(l), p.numpy()
    _, s32 = m.loss_and_grads(inputs, targets)
    g32 = m.flat_grads.clone()
    err32 = rel_err(s32.cpu().numpy(), pred)
    m.set_option("precision", precision)
    loss_sum, scores = m.loss_and_grads(inputs, targets)
    torch.cuda.synchronize()
    err = rel_err(scores.cpu().numpy(), pred)
    print(f"precision={precision} {shape} x{n}: score error {err:.3e} (fp32-accurate path {err32:.3e})")
    assert err <= tol
    assert not torch.equal(scores, s32)  # (it really is another numerics path, not the fp32-accurate one)
    assert abs(float(loss_sum) / scores.numel() - loss) <= 2 * tol * loss
    assert_grads_close(m.flat_grads.cpu().numpy(), grads, tol=tol)
    m.set_option("precision", 0)
    _, s_back = m.loss_and_grads(inputs, targets)
    assert torch.equal(s_back, s32) and torch.equal(m.flat_grads, g32)


def test_bf16_mlp_mode_needs_the_chain_kernels(golden_dir):
    from gcnn_cut_selector_b200 import GCNN, InvalidArgumentError
    m = GCNN(device="cuda:0", seed=0)
    if not m._lib.gcnn_has_alt_paths():  # the product build: the alternates cannot even be selected
        with pytest.raises(InvalidArgumentError):
            m.set_option("tensor_cores", 0)
        m.set_option("tensor_cores", 1)
        m.set_option("precision", 1)
        m.set_option("precision", 0)
        return
    m.set_option("tensor_cores", 0)
    with pytest.raises(InvalidArgumentError):
        m.set_option("precision", 1)
