"""Input side on the device: per-trace normalisation and the dpk soft labels.

The reference does both per waveform in numpy inside the DataLoader workers (`training/preprocess.py`:
`DataPreprocessor._normalize` :224-242, `_generate_soft_label` :544-683).  At the throughput of the fused training step
eight worker processes cannot feed one GPU; here a raw (N, C, L) batch and the (N, K) phase indices are turned into the
model input and the (N, 3, L) label tensor by two kernel launches on the device the step runs on.  No CPU path.
"""
from __future__ import annotations

import torch

from . import _lib

ABSENT = -10_000_000                       # "no phase" padding of the (N, K) index tensors
_MODES = {"": 0, "std": 1, "max": 2}
_SHAPES = {"gaussian": 0, "triangle": 1, "box": 2}


def _s() -> int:
    return torch.cuda.current_stream().cuda_stream


def normalize_(x: torch.Tensor, mode: str = "std") -> torch.Tensor:
    """In-place `_normalize(data, mode)` of every (n, c) trace of a contiguous CUDA fp32 (N, C, L) batch."""
    if mode not in _MODES:
        raise ValueError(f"Supported mode: 'max','std', got '{mode}'")       # the reference's message
    if not x.is_cuda or x.dtype != torch.float32 or not x.is_contiguous() or x.dim() != 3:
        raise RuntimeError("seist_b200.preprocess.normalize_ needs a contiguous CUDA float32 (N, C, L) tensor (no CPU path)")
    n, c, l = x.shape
    _lib.check(_lib.lib().seist_normalize(x.data_ptr(), n * c, l, _MODES[mode], _s()), "seist_normalize")
    return x


def dpk_soft_labels(ppks: torch.Tensor, spks: torch.Tensor, length: int, soft_label_width: int,
                    soft_label_shape: str = "gaussian", coda_ratio: float = 1.4) -> torch.Tensor:
    """(N, K) int64 P / S sample indices (pad missing phases with `ABSENT`) -> (N, 3, length) float32 labels
    [det, ppk, spk] as `DataPreprocessor._generate_soft_label` builds them (sigmoid-shaped windows: not implemented)."""
    if soft_label_shape not in _SHAPES:
        raise NotImplementedError(f"Unsupported label shape: '{soft_label_shape}'")
    if not ppks.is_cuda or ppks.shape != spks.shape or ppks.dim() != 2:
        raise RuntimeError("dpk_soft_labels needs two CUDA (N, K) index tensors of the same shape (no CPU path)")
    p = ppks.to(torch.int64).contiguous()
    s = spks.to(device=p.device, dtype=torch.int64).contiguous()
    n, k = p.shape
    out = torch.empty(n, 3, length, dtype=torch.float32, device=p.device)
    _lib.check(_lib.lib().seist_dpk_labels(p.data_ptr(), s.data_ptr(), n, k, length, int(soft_label_width), _SHAPES[soft_label_shape],
                                           float(coda_ratio), out.data_ptr(), _s()), "seist_dpk_labels")
    return out
