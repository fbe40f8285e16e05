"""Autograd boundary of the drop-in modules ("drop-in for the supervised trainers").

Every leaf operator of the hot path (patch embed, Swin block, PatchMerging, PatchExpanding(+crop), decoder linear,
cross-attention block, the two heads) is one ``torch.autograd.Function``:

  forward   the sm_100a kernels, exactly the inference lowering of ``model.py`` (grad mode is off inside ``forward``)
  backward  the operator is re-evaluated by its fp32 torch restatement (``torch_ref.py``) on the SAVED INPUTS with autograd
            enabled, and the incoming gradient is back-propagated through ATen

so only operator inputs are kept alive between forward and backward (activation checkpointing at operator granularity),
the reference trainers' freeze logic (``requires_grad`` flags, ``Segmentator_pretrain.py:74-93``) decides which
gradients are produced, and ``torch.cuda.amp.autocast`` regions hand the kernels fp32 tensors (``custom_fwd``).  The
backward is ATen, not hand-written kernels: correct first; the fused AdamW step and the gradient
bucket kernels are in ``train.py`` / ``csrc/train_ops.cu``.
"""
import torch


class KernelOp(torch.autograd.Function):
    """apply(kernel_fn, torch_fn, *tensors): kernel_fn(*tensors) -> Tensor runs the CUDA path; torch_fn(*tensors) is the
    differentiable restatement of the same operator."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, kernel_fn, torch_fn, *tensors):
        out = kernel_fn(*tensors)
        ctx.torch_fn = torch_fn
        ctx.save_for_backward(*tensors)
        return out

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, gout):
        tensors = ctx.saved_tensors
        need = ctx.needs_input_grad[2:]
        with torch.enable_grad():
            ins = [t.detach().requires_grad_(bool(n)) for t, n in zip(tensors, need)]
            out = ctx.torch_fn(*ins)
            wanted = [t for t, n in zip(ins, need) if n]
            grads = torch.autograd.grad(out, wanted, gout.to(out.dtype), allow_unused=True) if wanted else ()
        it = iter(grads)
        return (None, None) + tuple(next(it) if n else None for n in need)


def op(kernel_fn, torch_fn, *tensors):
    return KernelOp.apply(kernel_fn, torch_fn, *tensors)
