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
from gcnn_cut_selector_b200.trainer import gather_prenorm_stats, ordered_bucket_sum, reduce_bucket

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


def _ordered_worker(rank, world, port, out_dir):
    """What csrc/dp.cu does on the GPUs, emulated over gloo: every rank gathers all fp32 buckets and adds them in RANK
    order (not in an order the collective library picks), then applies the same Adam update to its replica."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(40 + rank)
    bucket = torch.from_numpy((rng.standard_normal(N + 2) * 10.0 ** rng.integers(-6, 3, N + 2)).astype(np.float32))
    bucket[N] = 7 + rank
    parts = [torch.empty_like(bucket) for _ in range(world)]
    dist.all_gather(parts, bucket)
    total = ordered_bucket_sum([p.numpy() for p in parts])
    # Keras Adam on the replica (model_trainer.py:131, 273), step 1, from identical parameters
    p0 = np.linspace(-1, 1, N).astype(np.float32)
    g = (total[:N] / total[N]).astype(np.float32)
    m, v = (g * np.float32(0.1)).astype(np.float32), (g * g * np.float32(0.001)).astype(np.float32)
    lr_t = np.float32(1e-3 * np.sqrt(1 - 0.999) / (1 - 0.9))
    p1 = (p0 - lr_t * m / (np.sqrt(v) + np.float32(1e-7))).astype(np.float32)
    np.save(os.path.join(out_dir, f"ordered{rank}.npy"), np.concatenate([total, p1]))
    np.save(os.path.join(out_dir, f"bucket{rank}.npy"), bucket.numpy())
    dist.destroy_process_group()


def test_fixed_order_reduction_is_bit_identical_on_all_ranks(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_ordered_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    o0, o1 = (np.load(tmp_path / f"ordered{r}.npy") for r in range(2))
    np.testing.assert_array_equal(o0, o1)  # same sums, same updated parameters, bit for bit
    b0, b1 = (np.load(tmp_path / f"bucket{r}.npy") for r in range(2))
    np.testing.assert_array_equal(o0[:N + 2], (b0.astype(np.float32) + b1.astype(np.float32)).astype(np.float32))
    # the order is part of the contract: fp32 addition does not associate
    three = [b0, b1, (b0 * np.float32(-1.0000001)).astype(np.float32)]
    assert np.any(ordered_bucket_sum(three) != ordered_bucket_sum(three[::-1]))


def test_reduce_bucket_is_identity_without_process_group():
    b = torch.arange(5, dtype=torch.float32)
    assert reduce_bucket(b.clone()).equal(b)


# ---- pre-norm statistics over sharded data: all-gather of per-rank triples, Chan merge in rank order -------
def _rank_batches(rank):
    rng = np.random.default_rng(100 + rank)
    return [rng.standard_normal((5 + 3 * rank + b, 4)) * (1 + rank) + b for b in range(3)]  # unequal counts per rank


def _stats_worker(rank, world, port, out_dir):
    from gcnn_cut_selector_b200.model import PreNormLayer
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    layer = PreNormLayer(None, 0, "cons_emb/prenorm", 4, None, 0)
    layer.start_updates()
    for x in _rank_batches(rank):
        for stats in gather_prenorm_stats(x.mean(0), x.var(0), x.shape[0]):
            layer.update_params(*stats)
    np.save(os.path.join(out_dir, f"stats{rank}.npy"), np.concatenate([layer.mean, layer.var, [layer.count]]))
    dist.destroy_process_group()


def test_dp_prenorm_statistics_equal_the_single_process_merge(tmp_path):
    from gcnn_cut_selector_b200.model import PreNormLayer
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_stats_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    s0, s1 = (np.load(tmp_path / f"stats{r}.npy") for r in range(2))
    np.testing.assert_array_equal(s0, s1)  # every rank freezes the same values
    # a single process that sees rank 0's and rank 1's batches alternately (model.py:416-423, batch by batch)
    layer = PreNormLayer(None, 0, "cons_emb/prenorm", 4, None, 0)
    layer.start_updates()
    for b in range(3):
        for rank in range(2):
            x = _rank_batches(rank)[b]
            layer.update_params(x.mean(0), x.var(0), x.shape[0])
    np.testing.assert_array_equal(s0, np.concatenate([layer.mean, layer.var, [layer.count]]))
    # ... which is the statistics of all rows together, up to fp32 rounding of the merges
    allx = np.concatenate([x for rank in range(2) for x in _rank_batches(rank)])
    assert np.abs(s0[:4] - allx.mean(0)).max() <= 1e-5 and np.abs(s0[4:8] - allx.var(0)).max() <= 1e-4
    assert s0[8] == allx.shape[0]


def test_cpulist_parsing_and_binding_without_topology():
    from gcnn_cut_selector_b200.trainer import bind_host_to_device, parse_cpulist
    assert parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert parse_cpulist("") == set()
    if not torch.cuda.is_available():
        assert bind_host_to_device(0) is None  # no driver, no sysfs entry: the process is left alone
