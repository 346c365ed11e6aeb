"""Opt-in fused parameter update.

`FusedAdam` is a drop-in for the optimizer the reference builds in
/root/reference/scene/gaussian_model.py:148-166 (`torch.optim.Adam(l, lr=0.0, eps=1e-15)`) and steps at
/root/reference/train.py:138-140.  It IS a `torch.optim.Adam` (same param_groups, same per-parameter
state keys `step`, `exp_avg`, `exp_avg_sq`, same state_dict), so the reference's densification code that
edits the optimizer state in place (gaussian_model.py:263-345: replace_tensor_to_optimizer,
_prune_optimizer, cat_tensors_to_optimizer) keeps working; only `.step()` is replaced: every parameter
of every group is updated by ONE CUDA launch (csrc/optim.cu) instead of ~12 elementwise passes per group.

`densification_stats(...)` fuses /root/reference/train.py:125-128 and gaussian_model.py:405-407.

No CPU path: parameters must be CUDA float32 tensors.
"""
import ctypes
import math

import torch

from . import _cabi


class FusedAdam(torch.optim.Adam):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        super().__init__(params, lr=lr, betas=betas, eps=eps, weight_decay=0, amsgrad=False,
                         foreach=False, fused=False, capturable=False, differentiable=False, maximize=False)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _cabi.load()
        # one launch per (device, betas, eps) bucket: the reference has exactly one
        buckets = {}
        for group in self.param_groups:
            beta1, beta2 = group["betas"]
            lr = float(group["lr"])
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not p.is_cuda or p.dtype != torch.float32 or p.grad.dtype != torch.float32:
                    raise RuntimeError("FusedAdam: parameters and gradients must be CUDA float32 tensors (no CPU path)")
                if p.grad.is_sparse:
                    raise RuntimeError("FusedAdam does not support sparse gradients")
                if not p.is_contiguous():
                    raise RuntimeError("FusedAdam: parameters must be contiguous")
                state = self.state[p]
                if len(state) == 0:                         # same lazy state as torch.optim.Adam
                    state["step"] = torch.tensor(0.0, dtype=torch.float32)
                    state["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    state["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                state["step"] += 1
                step = float(state["step"])
                grad = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                m, v = state["exp_avg"], state["exp_avg_sq"]
                if not (m.is_contiguous() and v.is_contiguous()):
                    raise RuntimeError("FusedAdam: optimizer state must be contiguous")
                bc1 = 1.0 - beta1 ** step
                bc2 = 1.0 - beta2 ** step
                entry = (p, grad, m, v, lr / bc1, math.sqrt(bc2))
                buckets.setdefault((p.device, float(beta1), float(beta2), float(group["eps"])), []).append(entry)
        for (dev, beta1, beta2, eps), entries in buckets.items():
            with torch.cuda.device(dev):
                stream = torch.cuda.current_stream(dev).cuda_stream
                for i in range(0, len(entries), _cabi.ADAM_MAX_GROUPS):
                    chunk = entries[i:i + _cabi.ADAM_MAX_GROUPS]
                    table = (_cabi.AdamGroup * len(chunk))()
                    for g, (p, grad, m, v, step_size, bc2_sqrt) in zip(table, chunk):
                        g.param, g.grad, g.exp_avg, g.exp_avg_sq = p.data_ptr(), grad.data_ptr(), m.data_ptr(), v.data_ptr()
                        g.n, g.step_size, g.bias2_sqrt = p.numel(), step_size, bc2_sqrt
                    _cabi.check(lib.surfel_adam_step(len(chunk), table, beta1, beta2, eps, stream))
        return loss


@torch.no_grad()
def densification_stats(xyz_gradient_accum, denom, max_radii2D, viewspace_grad, radii):
    """In place, where radii > 0:  max_radii2D = max(max_radii2D, radii);
    xyz_gradient_accum += |viewspace_grad|;  denom += 1   (train.py:125-128, gaussian_model.py:405-407)."""
    lib = _cabi.load()
    P = radii.shape[0]
    for t in (xyz_gradient_accum, denom, viewspace_grad, radii):
        if not t.is_cuda:
            raise RuntimeError("densification_stats: CUDA tensors required (no CPU path)")
    if radii.dtype != torch.int32 or viewspace_grad.shape != (P, 3) or not viewspace_grad.is_contiguous():
        raise RuntimeError("densification_stats: radii must be int32 (P), viewspace_grad float32 (P,3) contiguous")
    dev = radii.device
    with torch.cuda.device(dev):
        _cabi.check(lib.surfel_densify_stats(
            P, radii.data_ptr(), viewspace_grad.data_ptr(), xyz_gradient_accum.data_ptr(), denom.data_ptr(),
            max_radii2D.data_ptr() if max_radii2D is not None else None, torch.cuda.current_stream(dev).cuda_stream))
