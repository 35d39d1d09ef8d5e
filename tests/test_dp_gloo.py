"""CPU, world_size 2 over gloo: the data-parallel reduction rule of gcnn_cut_selector_b200.trainer.

Each rank holds a different shard (different numbers of cuts), seeds UN-normalised MSE gradients 2 (p - y), and one
all-reduce(sum) over the flat bucket [gradients | local cut count | local squared-error sum] followed by a division by
the global cut count must reproduce the single-process gradient of the mean over all cuts (model_trainer.py:271).
Per-rank gradients come from the CPU oracle here (the CUDA kernels need a GPU); the bucket layout and the collective
are the product's (`reduce_bucket`)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import gcnn_oracle as orc
from gcnn_cut_selector_b200 import batching, synth
from gcnn_cut_selector_b200.trainer import reduce_bucket

N = orc.N_TRAINABLE


def _samples():
    # rank 0 gets 1 graph (4 cuts), rank 1 gets 2 graphs of another shape (16 cuts): unequal cut counts on purpose
    return [synth.make_samples("tiny", 1, seed0=5), synth.make_samples("mini", 2, seed0=6)]


def _flat(grads):
    return torch.cat([grads[n].reshape(-1) for n, _ in orc.TRAINABLE])


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    params = orc.init_params(seed=12345, dtype=torch.float64)
    model = orc.OracleGCNN(params, dtype=torch.float64)
    batch = batching.concat_samples(_samples()[rank])
    n_cuts = batch[4].shape[0]
    loss, pred, grads = orc.loss_and_grads(model, batching.model_inputs(batch), batch[10], normaliser=1.0)
    bucket = torch.zeros(N + 2, dtype=torch.float64)
    bucket[:N] = _flat(grads)
    bucket[N] = n_cuts
    bucket[N + 1] = float(loss)  # normaliser 1 -> the squared-error sum
    reduce_bucket(bucket)
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), bucket.numpy())
    dist.destroy_process_group()


def test_dp_bucket_reduction_matches_single_process(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    b0, b1 = (np.load(tmp_path / f"rank{r}.npy") for r in range(2))
    np.testing.assert_array_equal(b0, b1)  # every rank ends with the same bucket

    shards = _samples()
    whole = batching.concat_samples(shards[0] + shards[1])
    model = orc.OracleGCNN(orc.init_params(seed=12345, dtype=torch.float64), dtype=torch.float64)
    loss, pred, grads = orc.loss_and_grads(model, batching.model_inputs(whole), whole[10])
    n_global = whole[4].shape[0]
    assert b0[N] == n_global == 20
    want = _flat(grads).numpy()
    got = b0[:N] / b0[N]
    assert np.abs(got - want).max() <= 1e-12 * np.abs(want).max()
    assert abs(b0[N + 1] / b0[N] - float(loss)) <= 1e-12 * float(loss)

    # a mean of per-rank means is NOT the same thing when ranks hold different numbers of cuts
    per_rank = []
    for shard in shards:
        bt = batching.concat_samples(shard)
        per_rank.append(_flat(orc.loss_and_grads(model, batching.model_inputs(bt), bt[10])[2]).numpy())
    naive = 0.5 * (per_rank[0] + per_rank[1])
    assert np.abs(naive - want).max() > 1e-3 * np.abs(want).max()


def test_reduce_bucket_is_identity_without_process_group():
    b = torch.arange(5, dtype=torch.float32)
    assert reduce_bucket(b.clone()).equal(b)
