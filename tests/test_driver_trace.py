"""The reference's own drivers, replayed (tests/driver_replay.py): model_trainer.pretrain, model_trainer.process (training
and evaluation branch) and model_tester.process ran UNMODIFIED over the reference's GCNN in oracle/make_golden.py; the
recorded call sequence is issued here on the CPU oracle (CPU suite) and on gcnn_cut_selector_b200.GCNN (``-m gpu``)."""
import os

import pytest

import driver_replay as dr

STATE = os.path.join(dr.ROOT, "tests", "golden", "state_stream.pkl")


def test_trace_covers_the_three_drivers():
    t = dr.load_trace()
    ops = [e["op"] for e in t["events"]]
    assert [e["name"] for e in t["events"] if e["op"] == "phase"] == ["pretrain", "train", "valid", "test"]
    assert ops.count("pretrain_next") == 12 and ops.count("tape_gradient") == 3 and ops.count("apply_gradients") == 3
    assert ops.count("call") == 9 and "getattr" not in ops  # the drivers touch nothing but the replayed interface
    assert abs(t["lr"] - 1e-4) < 1e-12  # model_trainer.py:53


def test_oracle_follows_the_recorded_drivers():
    """fp64 oracle vs the reference's own model driven by the reference's own loops: agreement to fp64 rounding (this also
    validates the replayer)."""
    rep = dr.replay(dr.load_trace(), dr.OracleSubject(STATE), tol_scores=1e-9, tol_grads=2e-7, tol_trained=2e-7)  # (the fixture holds the gradient vector in fp32)
    assert rep["events"] == 61 and set(rep["phases"]) == {"train", "valid", "test"}
    assert rep["final_param_err"] <= 1e-6  # (the fixture stores the final parameters in fp32)


@pytest.mark.gpu
def test_cuda_model_follows_the_recorded_drivers():
    """gcnn_cut_selector_b200.GCNN in place of the reference's GCNN: the same calls in the same order return the same
    booleans, freeze the same layers, and give predictions / loss / gradients within 1e-5 of the reference's fp64 run
    before the first parameter update and within 1e-4 after three fp32 Adam steps; the drivers' mean loss and ranking
    accuracies, recomputed from the CUDA predictions (accuracy through the device ranking kernel), equal the recorded ones."""
    rep = dr.replay(dr.load_trace(), dr.GcnnSubject(STATE), tol_scores=1e-5, tol_grads=1e-5, tol_trained=1e-4)
    print(rep)
    assert rep["events"] == 61
    assert rep["final_param_err"] <= 1e-5
