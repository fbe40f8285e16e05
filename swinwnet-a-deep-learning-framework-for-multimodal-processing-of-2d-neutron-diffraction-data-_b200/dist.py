"""Data-parallel gradient reduction for the trainers of the reference: one process per GPU,
replicated parameters, gradients summed over ranks with ``torch.distributed`` (NCCL over NVLink on the GPUs, gloo in the
CPU tests) and divided by the world size.

Buckets follow the model's top-level sub-modules (``patch_embed``, ``segmentator_encoder``, ... — the units the
reference's trainers freeze and optimise separately, FullModel_supervised_trainer.py:81-93), so a bucket can be reduced
as soon as its backward has finished.  Parameters whose ``.grad`` is ``None`` (frozen branches of the even / odd steps,
unused cross-attention paths) are tolerated: a presence bitmap is agreed on first, a rank that lacks a gradient another
rank has contributes zeros, and gradients nobody produced stay ``None``.

The forward/backward kernels of this repository are inference-only so far; the reducer is independent of
them and is exercised against plain ``torch.nn`` modules (``tests/test_dist_gloo.py``)."""
import torch
import torch.distributed as dist


class GradReducer:
    def __init__(self, model, process_group=None, average=True):
        self.group = process_group
        self.average = average
        self.buckets = []                      # [(name, [parameters])] in registration order
        top = {}
        for name, p in model.named_parameters():
            if not p.requires_grad:
                continue
            top.setdefault(name.split(".", 1)[0], []).append(p)
        self.buckets = list(top.items())

    def world_size(self):
        return dist.get_world_size(self.group) if dist.is_initialized() else 1

    @torch.no_grad()
    def reduce_bucket(self, params):
        """all-reduce the gradients of one bucket in a single flat buffer; returns the number of reduced elements."""
        if not params:
            return 0
        dev = params[0].device
        have = torch.tensor([0.0 if p.grad is None else 1.0 for p in params], device=dev)
        if self.world_size() > 1:
            dist.all_reduce(have, op=dist.ReduceOp.MAX, group=self.group)      # does ANY rank hold this gradient?
        live = [p for p, h in zip(params, have.tolist()) if h > 0]
        if not live:
            return 0
        flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1).float() for p in live])
        if self.world_size() > 1:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
            if self.average:
                flat /= self.world_size()
        off = 0
        for p in live:
            n = p.numel()
            g = flat[off:off + n].view_as(p).to(p.dtype)
            if p.grad is None:
                p.grad = g.clone()
            else:
                p.grad.copy_(g)
            off += n
        return off

    @torch.no_grad()
    def reduce(self):
        """reduce every bucket (call after ``loss.backward()``, before ``optimizer.step()``); returns elements reduced."""
        return sum(self.reduce_bucket(ps) for _, ps in self.buckets)
