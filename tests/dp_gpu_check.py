"""Multi-GPU check of the peer-memory gradient exchange (run under torchrun on a box with >= 2 GPUs):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dp_gpu_check.py

Every rank trains three steps on its own shard (different numbers of graphs / cuts per rank) twice: once with the fused
peer-memory exchange (csrc/dp.cu) and once with NCCL all_reduce + Adam.  Checks: (1) all ranks hold bit-identical
parameters after the peer path; (2) the peer path's parameters equal the fixed-rank-order emulation applied to the gathered
buckets (bit-exact); (3) peer and NCCL paths agree to fp32 rounding; (4) the reported global mean loss is the mean over
all cuts of all ranks; (5) the one-call staged path (gcnn_dp_train_step_staged_async) reports the same losses
to fp32 rounding; (6) a rank whose peers never arrive gets an error after the configured timeout, not a hung GPU.  Prints one JSON line on rank 0.  Test infrastructure (not collected by pytest: needs torchrun)."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gcnn_cut_selector_b200 import GCNN, DataParallelTrainer, batching, synth  # noqa: E402
from gcnn_cut_selector_b200.trainer import ordered_bucket_sum  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    state = os.path.join(ROOT, "tests", "golden", "state_stream.pkl")
    steps = [batching.concat_samples(synth.make_samples("mini", 1 + (rank + s) % 3, seed0=100 * rank + s)) for s in range(3)]

    def run(peer):
        m = GCNN(device=dev, seed=0)
        m.restore_state(state)
        t = DataParallelTrainer(m, lr=1e-3, peer_exchange=peer)
        losses, buckets = [], []
        for b in steps:
            loss = t.step(batching.model_inputs(b, per_sample_counts=True), b[10])
            losses.append(float(loss.item()))
        torch.cuda.synchronize()
        return m, t, losses

    def run_staged():  # the one-call-per-step path: host batches staged in alternating slots, loss read one step later
        from gcnn_cut_selector_b200 import HostBatch
        m = GCNN(device=dev, seed=0)
        m.restore_state(state)
        t = DataParallelTrainer(m, lr=1e-3, peer_exchange=True)
        hosts = [HostBatch(b) for b in steps]
        losses = []
        m.stage_host(hosts[0], 0)
        for i in range(len(steps)):
            if i + 1 < len(steps):
                m.stage_host(hosts[i + 1], (i + 1) & 1)
            t.step_staged_async(i & 1)
            if i > 0:
                losses.append(t.step_result((i - 1) & 1))
        losses.append(t.step_result((len(steps) - 1) & 1))
        torch.cuda.synchronize()
        return m, t, losses

    m_peer, t_peer, loss_peer = run(True)
    m_nccl, t_nccl, loss_nccl = run(False)
    m_stg, t_stg, loss_stg = run_staged()
    # (the staged host batch carries its sortedness hints, so its layouts come from the per-block transpose instead of the
    # radix sort: another summation order in the backward pass, losses equal to fp32 rounding rather than bit for bit)
    staged_same = all(abs(a - c) <= 5e-5 * abs(c) for a, c in zip(loss_stg, loss_peer))
    staged_param_diff = float((m_stg.flat_params.detach() - m_peer.flat_params.detach()).abs().mean())
    ok_peer = bool(t_peer.peer)
    p = m_peer.flat_params.detach().clone()
    gathered = [torch.empty_like(p) for _ in range(world)]
    dist.all_gather(gathered, p)
    identical = all(torch.equal(gathered[0], g) for g in gathered)
    diff = float((m_peer.flat_params.detach() - m_nccl.flat_params.detach()).abs().max())
    # one more step by hand: gather the buckets, emulate the fixed-order sum + Adam on the host, compare bit for bit
    b = steps[0]
    before = m_peer.flat_params.detach().cpu().numpy().copy()
    m_prev, v_prev = m_peer.adam_m.cpu().numpy().copy(), m_peer.adam_v.cpu().numpy().copy()
    m_peer.loss_and_grads(batching.model_inputs(b, per_sample_counts=True), b[10], seed_scale=1.0,
                          loss_out=t_peer.bucket[t_peer.N + 1:t_peer.N + 2])
    mine = t_peer.bucket.detach().clone()
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    t_peer._finish(True)
    torch.cuda.synchronize()
    total = ordered_bucket_sum([x.cpu().numpy() for x in parts])
    N = t_peer.N
    f = np.float32
    g = (total[:N] / total[N]).astype(f)
    step = m_peer.adam_step
    lr_t = f(1e-3 * np.sqrt(1.0 - 0.999 ** step) / (1.0 - 0.9 ** step))
    mm = (m_prev + (g - m_prev) * f(1.0 - f(0.9))).astype(f)
    vv = (v_prev + (g * g - v_prev) * f(1.0 - f(0.999))).astype(f)
    want = (before - lr_t * mm / (np.sqrt(vv) + f(1e-7))).astype(f)
    got = m_peer.flat_params.detach().cpu().numpy()
    emu_max = float(np.abs(got - want).max())
    emu_exact = bool(np.array_equal(got, want))
    if rank == 0:
        print(json.dumps({"world": world, "peer_exchange_active": ok_peer, "ranks_bit_identical": identical,
                          "peer_vs_nccl_max_abs": diff, "loss_peer": loss_peer, "loss_nccl": loss_nccl, "loss_staged_async": loss_stg,
                          "staged_async_matches_step": staged_same, "staged_mean_abs_param_diff": staged_param_diff,
                          "emulation_bit_exact": emu_exact, "emulation_max_abs": emu_max}), flush=True)
    assert ok_peer and identical and diff < 1e-6 and emu_max < 1e-7 and staged_same
    assert all(abs(a - c) <= 1e-5 * abs(c) for a, c in zip(loss_peer, loss_nccl))
    # (6) a peer that never arrives: rank 0 alone enqueues one more exchange with a 300 ms timeout.  Its kernel must give
    # up, leave the parameters untouched and report error bit 16 as InvalidArgumentError instead of hanging the GPU.
    dist.barrier()
    timed_out = None
    if rank == 0:
        from gcnn_cut_selector_b200 import InvalidArgumentError
        m_peer.set_option("dp_timeout_ms", 300)
        before_lone = m_peer.flat_params.detach().clone()
        t_peer._finish(False)
        torch.cuda.synchronize()
        assert torch.equal(before_lone, m_peer.flat_params.detach())
        try:
            m_peer.check_indices = True
            m_peer._check_indices()
            timed_out = False
        except InvalidArgumentError as exc:
            timed_out = "peer" in str(exc) or "16" in str(exc)
        print(json.dumps({"lone_rank_times_out_with_error": timed_out}), flush=True)
        assert timed_out
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
