"""Training-side runtime of the drop-in: the optimizer step and the data-parallel gradient
exchange of the reference trainers on hand-written kernels.

* ``FusedAdamW`` — ``torch.optim.Optimizer`` with the semantics of the ``torch.optim.AdamW`` the reference trainers build
  (``FullModel_supervised_trainer.py:85-92``, ``Segmentator_pretrain.py:124-131``): ONE kernel launch updates every
  parameter tensor (``swn_adamw_multi``), instead of ~10 foreach launches over 600 tensors.  Works with
  ``torch.cuda.amp.GradScaler`` (``scaler.step(optimizer)`` unscales, then calls ``step()``) and with parameters whose
  ``.grad`` is ``None`` in a step (frozen sub-modules, the cross-attention branch the even / odd step does not use).
* ``DistributedGradSync`` — at the end of every backward pass the gradients of all ranks are summed with ONE NCCL
  all-reduce of a flat fp32 bucket (packed / unpacked by ``swn_grad_bucket_copy``; 29.16 M elements = 117 MB for SwinWNet) and
  averaged.  A parameter that has no gradient on this rank contributes zeros; a parameter that has no gradient on ANY
  rank keeps ``grad = None`` (a presence bitmap is max-reduced first), so AdamW's moments of unused branches are not
  decayed — the behaviour of single-process training.  One process per GPU, ``torch.distributed`` (NCCL over NVLink).

The trainers themselves are the reference's (unmodified): they accept ``optimizer=`` and call ``model.segment_1`` etc.,
which run the sm_100a kernels forward and the ATen restatement backward (``autograd.py``).
"""
import ctypes

import torch
import torch.distributed as dist

from . import ops

_CHUNK = 4096


class _ParamTable:
    """device-side descriptor table (swn_param_desc) + chunk list for a fixed list of parameters"""

    def __init__(self, params):
        self.params = list(params)
        dev = self.params[0].device
        for p in self.params:
            if p.device != dev or p.dtype != torch.float32 or not p.is_contiguous():
                raise RuntimeError("swinwnet_b200.train: parameters must be contiguous fp32 tensors on one CUDA device")
        self.device = dev
        offs, chunks, off = [], [], 0
        for i, p in enumerate(self.params):
            offs.append(off)
            off += (p.numel() + 3) // 4 * 4            # keep every tensor 16-byte aligned inside the flat bucket
            chunks += [(i, c) for c in range(0, p.numel(), _CHUNK)]
        self.flat_elems, self.offsets = off, offs
        self.chunks = torch.tensor(chunks, dtype=torch.int32).to(dev)
        self.n_chunks = len(chunks)
        self._host = torch.zeros(len(self.params), 7, dtype=torch.int64).pin_memory()     # swn_param_desc: 56 bytes
        self._dev = torch.zeros(len(self.params), 7, dtype=torch.int64, device=dev)
        self._bc = self._host.view(torch.float32).view(len(self.params), 14)[:, 12:14]    # (bc1, bc2_sqrt) of slot 6
        self._copied = None            # event: the last async H2D copy of the table has consumed the pinned buffer

    def refresh(self, m=None, v=None, grads=None, bias_corr=None):
        """(re)write the descriptor table: gradient pointers change from step to step (zero_grad(set_to_none=True))"""
        h = self._host
        if self._copied is not None:
            self._copied.synchronize()
        for i, p in enumerate(self.params):
            g = grads[i] if grads is not None else p.grad
            h[i, 0] = p.data_ptr()
            h[i, 1] = 0 if g is None else g.data_ptr()
            h[i, 2] = 0 if m is None else m[i].data_ptr()
            h[i, 3] = 0 if v is None else v[i].data_ptr()
            h[i, 4] = p.numel()
            h[i, 5] = self.offsets[i]
            if bias_corr is not None:
                self._bc[i, 0], self._bc[i, 1] = bias_corr[i]
        self._dev.copy_(h, non_blocking=True)
        self._copied = torch.cuda.Event()
        self._copied.record(torch.cuda.current_stream(self.device))
        return self._dev


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._tables = {}

    def _table(self, gi, group):
        ps = [p for p in group["params"] if p.requires_grad]
        key = tuple(id(p) for p in ps)
        t = self._tables.get(gi)
        if t is None or t[0] != key:
            tab = _ParamTable(ps)
            m = [torch.zeros_like(p) for p in ps]
            v = [torch.zeros_like(p) for p in ps]
            for p, mi, vi in zip(ps, m, v):           # kept in optimizer.state like torch's AdamW (state_dict / release)
                self.state[p]["exp_avg"], self.state[p]["exp_avg_sq"] = mi, vi
            t = (key, tab, m, v, [0] * len(ps))
            self._tables[gi] = t
        return t

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for gi, group in enumerate(self.param_groups):
            if not any(p.requires_grad for p in group["params"]):
                continue
            key, tab, m, v, steps = self._table(gi, group)
            b1, b2 = group["betas"]
            grads, bc = [], []
            for i, p in enumerate(tab.params):
                g = p.grad
                if g is not None and (g.dtype != torch.float32 or not g.is_contiguous()):
                    g = g.float().contiguous()
                grads.append(g)
                if g is not None:
                    steps[i] += 1          # like torch.optim.AdamW: the step count is per parameter
                t = max(steps[i], 1)
                bc.append((1.0 - b1 ** t, (1.0 - b2 ** t) ** 0.5))
            if all(g is None for g in grads):
                continue
            ops.adamw_multi(tab.refresh(m, v, grads, bc), tab.chunks, tab.n_chunks, float(group["lr"]), float(b1), float(b2),
                            float(group["eps"]), float(group["weight_decay"]), 1.0)
        return loss


class FlatGradReducer:
    """sum (and average) the gradients of `params` over the ranks with one all-reduce of a flat bucket"""

    def __init__(self, params, process_group=None, average=True):
        self.tab = _ParamTable([p for p in params if p.requires_grad])
        self.group, self.average = process_group, average
        self.flat = torch.zeros(self.tab.flat_elems, dtype=torch.float32, device=self.tab.device)

    def world_size(self):
        return dist.get_world_size(self.group) if dist.is_initialized() else 1

    @torch.no_grad()
    def reduce(self):
        ps, ws = self.tab.params, self.world_size()
        have = torch.tensor([0.0 if p.grad is None else 1.0 for p in ps], device=self.tab.device)
        if ws > 1:
            dist.all_reduce(have, op=dist.ReduceOp.MAX, group=self.group)       # does ANY rank hold this gradient?
        have = have.bool().tolist()
        for p, h in zip(ps, have):
            if h and p.grad is None:
                p.grad = torch.zeros_like(p)                                     # this rank contributes zeros
            elif p.grad is not None and (p.grad.dtype != torch.float32 or not p.grad.is_contiguous()):
                p.grad = p.grad.float().contiguous()
        if not any(have):
            return 0
        table = self.tab.refresh()
        ops.grad_bucket_copy(table, self.tab.chunks, self.tab.n_chunks, self.flat, unpack=False)
        if ws > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
        ops.grad_bucket_copy(table, self.tab.chunks, self.tab.n_chunks, self.flat, unpack=True,
                             scale=(1.0 / ws) if self.average else 1.0)
        return sum(p.numel() for p, h in zip(ps, have) if h)


class DistributedGradSync:
    """Data-parallel gradient exchange for UNMODIFIED trainers: construct it once for the model
    (``DistributedGradSync(model)``); from then on every ``loss.backward()`` ends with the all-reduce of the gradients it
    produced (a callback queued on the autograd engine by the first gradient that is accumulated — the mechanism
    ``DistributedDataParallel`` uses), i.e. before ``GradScaler.unscale_`` / ``optimizer.step()``, so non-finite gradients
    reach every rank and all ranks skip or take the step together.  Which parameters receive gradients may differ
    from step to step (even / odd steps of FullModel_supervised_trainer.py:231-288) but must be the same on all ranks
    up to parameters that happen to be unused on one rank (handled by the presence bitmap)."""

    def __init__(self, model, process_group=None, average=True, reducer=None):
        params = [p for p in model.parameters() if p.requires_grad]
        if reducer is None:
            if params[0].is_cuda:
                reducer = FlatGradReducer(params, process_group, average)
            else:                      # host-side logic test (gloo): the bucketed torch reducer of dist.py
                from .dist import GradReducer
                reducer = GradReducer(model, process_group, average)
        self.reducer = reducer
        self._queued = False
        self.reduced_elements = 0
        self._handles = [p.register_post_accumulate_grad_hook(self._on_grad) for p in params]

    def _on_grad(self, _param):
        if not self._queued:
            self._queued = True
            torch.autograd.Variable._execution_engine.queue_callback(self._finish)

    def _finish(self):
        self._queued = False
        self.reduced_elements = self.reducer.reduce()

    def remove(self):
        for h in self._handles:
            h.remove()
        self._handles = []
