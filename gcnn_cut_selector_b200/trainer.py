"""Data-parallel training of the GCNN over the GPUs of one box (BASELINE config 4; no reference counterpart --
the reference trains in a single process, model_trainer.py:52, 128-131).

Samples are independent (block-diagonal batches, utils.py:403-407), so each rank runs the fused forward+backward on
its own shard with an UN-normalised MSE seed 2 (p - y); one all-reduce (sum) over a flat bucket
``[93,121 gradients | local cut count | local squared-error sum]`` followed by a fused Adam that divides by the
global cut count reproduces the single-process mean over all cuts exactly (a mean of per-rank means would not when
ranks hold different numbers of cuts).  The message is 372 KB: latency-bound, one NCCL call over NVLink.
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


class DataParallelTrainer:
    """Owns the all-reduce bucket; ``model.flat_grads`` becomes a view of it so nothing is copied per step."""

    N = _lib.N_TRAINABLE

    def __init__(self, model, lr: float = 1e-4, group=None):
        self.model, self.lr, self.group = model, lr, group
        self.bucket = torch.zeros(self.N + 2, dtype=torch.float32, device=model.device)
        model.flat_grads = self.bucket[:self.N]
        # the library writes the local cut count and squared-error sum straight into the bucket tail (no torch ops on
        # the step's critical path: a scalar assignment from the host alone costs ~60 us of pipeline stall)
        self.tail_on_device = model.device.type == "cuda"
        if self.tail_on_device:
            model.set_option("count_before_loss", 1)

    def broadcast_parameters(self, src: int = 0):
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
            dist.broadcast(self.model.flat_params.detach(), src, group=self.group)
            dist.broadcast(self.model.flat_prenorm, src, group=self.group)

    def _finish(self, want_loss: bool):
        reduce_bucket(self.bucket, self.group)
        self.model.apply_gradients(self.lr, grad_divisor=self.bucket[self.N:self.N + 1])
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

    def step_staged(self, slot: int):
        """``step`` on the host batch staged in ``slot`` (GCNN.stage_host): its copy overlapped the previous step."""
        self.model.loss_and_grads_staged(slot, seed_scale=1.0, loss_out=self.bucket[self.N + 1:self.N + 2])
        return self._finish(True)
