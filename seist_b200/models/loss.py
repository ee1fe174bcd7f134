"""Loss modules with the reference's names (models/loss.py) so `config.py` imports unchanged.

Hot-path losses — `BCELoss` (reference :32-56, used by every seist_*_dpk via config.py:138) and
`HuberLoss` (reference :3 re-exports torch's; config.py:158) — run as fused CUDA kernels through the
C-ABI (seist_bce_fwd/bwd, seist_huber_fwd/bwd) when given CUDA tensors.  The remaining classes belong
to other models' tasks (outside the accelerated path) and are thin torch expressions kept
only so the registry surface is complete.
"""
from typing import Tuple

import torch
import torch.nn as nn

from .. import _lib


def _as_weight(weight, name):
    if weight is None:
        return torch.tensor(1.0, dtype=torch.float32)
    print(f"[{name}] Loss Weight:", weight)
    return torch.tensor(weight, dtype=torch.float32)


def _stream():
    return torch.cuda.current_stream().cuda_stream


class _BCEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, preds, targets, wvec, eps):
        preds = preds.contiguous()
        targets = targets.contiguous()
        n, c, l = preds.shape
        acc = torch.empty(1, dtype=torch.float64, device=preds.device)
        out = torch.empty((), dtype=torch.float32, device=preds.device)
        _lib.check(_lib.lib().seist_bce_fwd(preds.data_ptr(), targets.data_ptr(), wvec.data_ptr(), n, c, l,
                                            eps, acc.data_ptr(), out.data_ptr(), _stream()), "seist_bce_fwd")
        ctx.save_for_backward(preds, targets, wvec)
        ctx.eps = eps
        return out

    @staticmethod
    def backward(ctx, gout):
        preds, targets, wvec = ctx.saved_tensors
        n, c, l = preds.shape
        d = torch.empty_like(preds)
        gout = gout.contiguous().float()
        _lib.check(_lib.lib().seist_bce_bwd(preds.data_ptr(), targets.data_ptr(), wvec.data_ptr(),
                                            gout.data_ptr(), n, c, l, ctx.eps, d.data_ptr(), _stream()),
                   "seist_bce_bwd")
        return d, None, None, None


class BCELoss(nn.Module):
    """mean(-w * (t*log(p+eps) + (1-t)*log(1-p+eps))), eps = 1e-6 inside the logs; `weight`
    broadcasts over (N, C, L) — per-channel column [[w0],[w1],...] or a scalar."""

    _epsilon = 1e-6

    def __init__(self, weight=None) -> None:
        super().__init__()
        self.register_buffer("weight", _as_weight(weight, self._get_name()))

    def forward(self, preds, targets):
        if not preds.is_cuda:
            raise RuntimeError("seist_b200.BCELoss has no CPU path")
        c = preds.shape[1]
        w = self.weight.to(preds.device, torch.float32)
        if w.numel() == 1:
            wvec = w.reshape(1).expand(c).contiguous()
        elif w.numel() == c:
            wvec = w.reshape(c).contiguous()
        else:
            raise ValueError(f"BCELoss weight of shape {tuple(w.shape)} does not broadcast over {c} channels")
        if preds.dim() != 3:
            raise ValueError(f"BCELoss expects (N, C, L) predictions, got {tuple(preds.shape)}")
        if targets.shape != preds.shape:          # the reference's torch expression would broadcast (or raise)
            targets = targets.expand_as(preds)
        # the kernels read raw fp32: cast labels of any other dtype / device first (never reinterpret)
        targets = targets.to(device=preds.device, dtype=torch.float32)
        return _BCEFn.apply(preds.float(), targets, wvec, float(self._epsilon))


class _HuberFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, preds, targets, delta):
        preds = preds.contiguous()
        targets = targets.contiguous()
        acc = torch.empty(1, dtype=torch.float64, device=preds.device)
        out = torch.empty((), dtype=torch.float32, device=preds.device)
        _lib.check(_lib.lib().seist_huber_fwd(preds.data_ptr(), targets.data_ptr(), preds.numel(), delta,
                                              acc.data_ptr(), out.data_ptr(), _stream()), "seist_huber_fwd")
        ctx.save_for_backward(preds, targets)
        ctx.delta = delta
        return out

    @staticmethod
    def backward(ctx, gout):
        preds, targets = ctx.saved_tensors
        d = torch.empty_like(preds)
        gout = gout.contiguous().float()
        _lib.check(_lib.lib().seist_huber_bwd(preds.data_ptr(), targets.data_ptr(), gout.data_ptr(),
                                              preds.numel(), ctx.delta, d.data_ptr(), _stream()),
                   "seist_huber_bwd")
        return d, None, None


class HuberLoss(nn.Module):
    """torch.nn.HuberLoss(reduction='mean', delta=1.0) semantics."""

    def __init__(self, reduction: str = "mean", delta: float = 1.0) -> None:
        super().__init__()
        if reduction != "mean":
            raise NotImplementedError("only reduction='mean' (the reference's use) is accelerated")
        self.delta = float(delta)

    def forward(self, preds, targets):
        if not preds.is_cuda:
            raise RuntimeError("seist_b200.HuberLoss has no CPU path")
        if preds.shape != targets.shape:
            targets = targets.expand_as(preds)
        return _HuberFn.apply(preds.float(), targets.to(device=preds.device, dtype=torch.float32), self.delta)


class _CEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, preds, targets, wvec, eps):
        n, c = preds.shape
        acc = torch.empty(1, dtype=torch.float64, device=preds.device)
        out = torch.empty((), dtype=torch.float32, device=preds.device)
        _lib.check(_lib.lib().seist_ce_fwd(preds.data_ptr(), targets.data_ptr(), wvec.data_ptr(), n, c, eps, acc.data_ptr(),
                                           out.data_ptr(), _stream()), "seist_ce_fwd")
        ctx.save_for_backward(preds, targets, wvec)
        ctx.eps = eps
        return out

    @staticmethod
    def backward(ctx, gout):
        preds, targets, wvec = ctx.saved_tensors
        n, c = preds.shape
        d = torch.empty_like(preds)
        gout = gout.contiguous().float()
        _lib.check(_lib.lib().seist_ce_bwd(preds.data_ptr(), targets.data_ptr(), wvec.data_ptr(), gout.data_ptr(), n, c, ctx.eps,
                                           d.data_ptr(), _stream()), "seist_ce_bwd")
        return d, None, None, None


class CELoss(nn.Module):
    """mean_n sum_c -w[c] * t[n, c] * log(p[n, c] + eps) on class probabilities (reference models/loss.py:8-29; the loss of
    the seist_*_pmp variants, config.py:147-155).  Fused CUDA kernels for (N, C) CUDA tensors; other shapes (other
    models' uses) fall back to the reference's torch expression."""
    _epsilon = 1e-6

    def __init__(self, weight=None) -> None:
        super().__init__()
        self.register_buffer("weight", _as_weight(weight, self._get_name()))

    def forward(self, preds, targets):
        if preds.is_cuda and preds.dim() == 2 and targets.shape == preds.shape and self.weight.numel() in (1, preds.shape[1]):
            c = preds.shape[1]
            w = self.weight.to(preds.device, torch.float32)
            wvec = (w.reshape(1).expand(c) if w.numel() == 1 else w.reshape(c)).contiguous()
            return _CEFn.apply(preds.contiguous().float(), targets.to(device=preds.device, dtype=torch.float32).contiguous(),
                               wvec, float(self._epsilon))
        return (-(targets * (preds + self._epsilon).log()) * self.weight).sum(1).mean()


# ---- losses of tasks outside the accelerated path (kept for the registry surface) ---------------


class MSELoss(nn.Module):
    def __init__(self, weight=None) -> None:
        super().__init__()
        self.register_buffer("weight", _as_weight(weight, self._get_name()))

    def forward(self, preds, targets):
        return ((preds - targets).square() * self.weight).mean()


class FocalLoss(nn.Module):
    _epsilon = 1e-6

    def __init__(self, gamma=2, weight=None, has_softmax=True):
        super().__init__()
        self.gamma, self.has_softmax = gamma, has_softmax
        self.register_buffer("weight", _as_weight(weight, self._get_name()))

    def forward(self, preds, targets):
        p = preds.softmax(1) if self.has_softmax else preds
        ce = -targets * (p + self._epsilon).log()
        return (ce * (1 - p).pow(self.gamma) * self.weight).sum(1).mean()


class BinaryFocalLoss(nn.Module):
    _epsilon = 1e-6

    def __init__(self, gamma=2, alpha=1, weight=None):
        super().__init__()
        self.gamma, self.alpha = gamma, alpha
        self.register_buffer("weight", _as_weight(weight, self._get_name()))

    def forward(self, preds, targets):
        pos = self.alpha * (1 - preds).pow(self.gamma) * targets * (preds + self._epsilon).log()
        neg = (1 - self.alpha) * preds.pow(self.gamma) * (1 - targets) * (1 - preds + self._epsilon).log()
        return (-(pos + neg) * self.weight).mean()


class CombinationLoss(nn.Module):
    def __init__(self, losses: list, losses_weights: list = None) -> None:
        super().__init__()
        if len(losses) < 2:
            raise Exception(f"`CombinationLoss` needs at least two loss modules, got {len(losses)}.")
        self.losses_weights = list(losses_weights) if losses_weights is not None else [1.0] * len(losses)
        assert len(self.losses_weights) == len(losses)
        self.losses = nn.ModuleList([make() for make in losses])

    def forward(self, preds: Tuple[torch.Tensor], targets: Tuple[torch.Tensor]):
        return sum(fn(p, t) * w for p, t, fn, w in zip(preds, targets, self.losses, self.losses_weights))


class MousaviLoss(nn.Module):
    def forward(self, preds, targets):
        y_hat, s = preds[:, 0].reshape(-1, 1), preds[:, 1].reshape(-1, 1)
        return (0.5 * (-s).exp() * (targets - y_hat).square() + 0.5 * s).sum()
