"""Data-parallel training of the GCNN over the GPUs of one box (BASELINE config 4; no reference counterpart --
the reference trains in a single process, model_trainer.py:52, 128-131).

Samples are independent (block-diagonal batches, utils.py:403-407), so each rank runs the fused forward+backward on
its own shard with an UN-normalised MSE seed 2 (p - y); one all-reduce (sum) over a flat bucket
``[93,121 gradients | local cut count | local squared-error sum]`` followed by a fused Adam that divides by the
global cut count reproduces the single-process mean over all cuts exactly (a mean of per-rank means would not when
ranks hold different numbers of cuts).  The message is 372 KB: latency-bound.  On the GPUs of one box the exchange and
the optimiser are ONE kernel per rank over NVLink peer memory (``peer_exchange``: every rank reads the other ranks'
buckets in rank order through CUDA-IPC mappings and applies Adam, csrc/dp.cu); the NCCL ``all_reduce`` + Adam launch pair
is the fallback (gloo on CPU in the tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import _lib


def reduce_bucket(bucket: torch.Tensor, group=None) -> torch.Tensor:
    """Sum ``bucket`` over all ranks in place (NCCL on GPU tensors, gloo on CPU tensors in the tests)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(bucket, op=dist.ReduceOp.SUM, group=group)
    return bucket


def bind_host_to_device(device) -> dict | None:
    """Restrict this process to the CPUs of the NUMA node the GPU ``device`` is attached to, so that the pinned staging
    buffers it allocates afterwards (first touch) and the threads that fill them sit next to the GPU's PCIe root port.
    With eight ranks streaming ~28 GB/s each, buffers on the far socket turn the inter-socket link into the bottleneck.
    Returns {"numa_node", "cpus"} or None when the topology is not exposed (containers without sysfs PCI entries)."""
    import os
    try:
        p = torch.cuda.get_device_properties(device)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as fh:
            node = int(fh.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as fh:
            cpus = parse_cpulist(fh.read())
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return {"numa_node": node, "cpus": len(cpus)}
    except (OSError, ValueError, AttributeError, RuntimeError):
        return None


def parse_cpulist(text: str) -> set:
    """sysfs cpulist syntax ("0-3,8,10-11") -> set of CPU numbers."""
    cpus = set()
    for part in text.strip().split(","):
        if part:
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def ordered_bucket_sum(buckets):
    """CPU emulation of the peer-memory exchange's reduction: the ranks' buckets are added element-wise IN RANK ORDER,
    in fp32 (csrc/dp.cu) -- every rank forms the same bits.  ``buckets``: list of 1-D float32 arrays / tensors."""
    import numpy as np
    acc = np.zeros_like(np.asarray(buckets[0], dtype=np.float32))
    for b in buckets:
        acc = (acc + np.asarray(b, dtype=np.float32)).astype(np.float32)
    return acc


class _DeviceArray:
    """A library-owned device buffer as a ``__cuda_array_interface__`` object (torch.as_tensor maps it without a copy)."""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 2}


def gather_prenorm_stats(mean, var, count, group=None):
    """Every rank's batch statistics of one pre-norm layer, as a list ``[(mean, var, count), ...]`` in rank order (one
    small all-gather of fp64 triples; gloo on CPU, NCCL through a device copy).  Feeding the list, in order, to
    ``PreNormLayer.update_params`` (Chan's merge, model.py:416-423) leaves every rank with the statistics a single
    process would have after seeing the ranks' batches one after the other."""
    import numpy as np
    mean, var = np.atleast_1d(np.asarray(mean, np.float64)), np.atleast_1d(np.asarray(var, np.float64))
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1):
        return [(mean, var, float(count))]
    local = torch.from_numpy(np.concatenate([[float(count)], mean, var]))
    on_gpu = dist.get_backend(group) == "nccl"
    if on_gpu:
        local = local.cuda()
    parts = [torch.empty_like(local) for _ in range(dist.get_world_size(group))]
    dist.all_gather(parts, local, group=group)
    n = mean.shape[0]
    out = []
    for p in parts:
        p = p.cpu().numpy()
        out.append((p[1:1 + n], p[1 + n:1 + 2 * n], float(p[0])))
    return out


class DataParallelTrainer:
    """Owns the all-reduce bucket; ``model.flat_grads`` becomes a view of it so nothing is copied per step."""

    N = _lib.N_TRAINABLE

    def __init__(self, model, lr: float = 1e-4, group=None, peer_exchange: bool | None = None):
        self.model, self.lr, self.group = model, lr, group
        self.bucket = torch.zeros(self.N + 2, dtype=torch.float32, device=model.device)
        model.flat_grads = self.bucket[:self.N]
        # the library writes the local cut count and squared-error sum straight into the bucket tail (no torch ops on
        # the step's critical path: a scalar assignment from the host alone costs ~60 us of pipeline stall)
        self.tail_on_device = model.device.type == "cuda"
        if self.tail_on_device:
            model.set_option("count_before_loss", 1)
        # peer-memory exchange fused with Adam: default on for NCCL groups (one box); GCNN_DP_PEER=0 keeps NCCL
        self.peer = False
        self._pending_loss = {}
        self._peer_buckets = None
        self._sums = None
        world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        if peer_exchange is None:
            import os
            peer_exchange = (world > 1 and self.tail_on_device and dist.get_backend(group) == "nccl"
                             and os.environ.get("GCNN_DP_PEER", "1") != "0")
        if peer_exchange and world > 1:
            self._connect_peers(world, dist.get_rank(group))

    def _connect_peers(self, world: int, rank: int):
        """Exchange the CUDA-IPC handles of the ranks' communication blocks and map them (all ranks on one box)."""
        import ctypes as C
        m, lib = self.model, self.model._lib
        handle = (C.c_ubyte * 64)()
        ok = torch.ones(1, dtype=torch.int32, device=m.device)
        try:
            with torch.cuda.device(m.device):
                _lib.check(lib.gcnn_dp_create(m._ws, world, rank, handle))
        except Exception:
            ok.zero_()
        mine = torch.tensor(list(bytes(handle)), dtype=torch.uint8, device=m.device)
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine, group=self.group)
        if int(ok.item()):
            blob = bytes(torch.cat(parts).cpu().numpy().tobytes())
            try:
                with torch.cuda.device(m.device):
                    _lib.check(lib.gcnn_dp_connect(m._ws, blob))
            except Exception:
                ok.zero_()
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)  # every rank or none
        if not int(ok.item()):
            return
        self._peer_buckets = [torch.as_tensor(_DeviceArray(int(lib.gcnn_dp_bucket(m._ws, p)), self.N + 2), device=m.device)
                              for p in (0, 1)]
        self._sums = torch.zeros(2, dtype=torch.float32, device=m.device)
        self.peer = True
        self._use_bucket(int(lib.gcnn_dp_next_parity(m._ws)))

    def _use_bucket(self, parity: int):
        self.bucket = self._peer_buckets[parity]
        self.model.flat_grads = self.bucket[:self.N]

    def broadcast_parameters(self, src: int = 0):
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
            dist.broadcast(self.model.flat_params.detach(), src, group=self.group)
            dist.broadcast(self.model.flat_prenorm, src, group=self.group)

    def pretrain_fused(self, batches) -> int:
        """Pre-norm pretraining over sharded data: the 7-pass schedule of ``GCNN.pretrain_fused`` where, for
        every batch, each rank computes the statistics of ITS batch and all ranks merge all of them in rank order.  Every
        rank ends with the same shift / scale values -- those a single process gets from the batch sequence
        (rank 0's first batch, rank 1's first batch, ..., rank 0's second batch, ...).  All ranks must iterate the same
        number of batches.  Returns the number of passes."""
        m = self.model
        m.pretrain_init()
        layers, passes = m._prenorm_layers, 0
        for group in [layers[:5]] + [[layer] for layer in layers[5:]]:
            seen = False
            for b in batches:
                dev_inputs = m.prepare_inputs(b)
                for layer in group:
                    for stats in gather_prenorm_stats(*m._prenorm_batch_stats(layer, dev_inputs), group=self.group):
                        layer.update_params(*stats)
                    layer.received_updates = True
                seen = True
            if not seen:
                break
            for layer in group:
                layer.stop_updates()
            passes += 1
        return passes

    def _finish(self, want_loss: bool):
        m = self.model
        if self.peer:  # one kernel: wait for the peers' buckets, sum in rank order, Adam (csrc/dp.cu)
            with torch.cuda.device(m.device):
                _lib.check(m._lib.gcnn_dp_allreduce_adam(m._ws, m.flat_params.data_ptr(), m.adam_m.data_ptr(),
                                                         m.adam_v.data_ptr(), self.lr, 0.9, 0.999, 1e-7, m.adam_step + 1,
                                                         self._sums.data_ptr(), m._stream()))
            m.adam_step += 1
            self._use_bucket(int(m._lib.gcnn_dp_next_parity(m._ws)))
            return self._sums[1:2] / self._sums[0:1] if want_loss else None
        reduce_bucket(self.bucket, self.group)
        m.apply_gradients(self.lr, grad_divisor=self.bucket[self.N:self.N + 1])
        return self.bucket[self.N + 1:self.N + 2] / self.bucket[self.N:self.N + 1] if want_loss else None

    def step(self, inputs, targets, want_loss: bool = True):
        """One data-parallel optimisation step; returns the global mean loss as a device tensor [1] (``want_loss=False``
        skips that division and returns None)."""
        m = self.model
        if self.tail_on_device:
            m.loss_and_grads(inputs, targets, seed_scale=1.0, loss_out=self.bucket[self.N + 1:self.N + 2])
        else:
            loss_sum, scores = m.loss_and_grads(inputs, targets, seed_scale=1.0)
            self.bucket[self.N] = float(scores.numel())
            self.bucket[self.N + 1:self.N + 2].copy_(loss_sum)
        return self._finish(want_loss)

    def step_staged_async(self, slot: int):
        """Enqueue one data-parallel step on the batch staged in ``slot`` with ONE library call (backward into the
        communication bucket, peer all-reduce + Adam); the global mean loss is read later with ``step_result(slot)``.
        Needs the peer exchange (NCCL ranks of one box); falls back to ``step_staged`` otherwise."""
        m = self.model
        if not self.peer:
            self._pending_loss[slot] = self.step_staged(slot).clone()
            return
        with torch.cuda.device(m.device):
            _lib.check(m._lib.gcnn_dp_train_step_staged_async(m._ws, slot, m.flat_params.data_ptr(),
                                                              m.flat_prenorm.data_ptr(), m.adam_m.data_ptr(),
                                                              m.adam_v.data_ptr(), self.lr, m.adam_step + 1, m._stream()))
        m.adam_step += 1
        self._use_bucket(int(m._lib.gcnn_dp_next_parity(m._ws)))

    def step_result(self, slot: int) -> float:
        """Global mean loss of the step enqueued on ``slot`` by ``step_staged_async`` (waits for that step only)."""
        if not self.peer:
            return float(self._pending_loss.pop(slot).item())
        return self.model.train_step_result(slot)

    def step_staged(self, slot: int):
        """``step`` on the host batch staged in ``slot`` (GCNN.stage_host): its copy overlapped the previous step."""
        self.model.loss_and_grads_staged(slot, seed_scale=1.0, loss_out=self.bucket[self.N + 1:self.N + 2])
        return self._finish(True)
