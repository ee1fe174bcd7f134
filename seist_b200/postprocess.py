"""Device-side post-processing of the dpk head's probability traces.

Mirrors the reference's `training/postprocess.py` (`process_outputs` :196-250, `_pick_phase` :161-193, `_detect_event`
:114-158) and the precision / recall / residual counters of `utils/metrics.py:141-247`, but runs on the GPU the outputs
already live on: the reference copies the full (B, 3, L) output to the host and loops over B x 3 traces in Python on every
training step (`training/train.py:141`).  Integer results equal the numpy oracle (`oracle/postprocess_ref.py`, pinned to
the reference's own `_detect_peaks`) bit for bit.  There is no CPU path.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Union

import torch
import torch.distributed as dist

from . import _lib

PAD_PHASE = int(-1e7)        # padding of `_pick_phase` (postprocess.py:226)


def _s() -> int:
    return torch.cuda.current_stream().cuda_stream


def _check(outputs: torch.Tensor, channel: int):
    if not outputs.is_cuda:
        raise RuntimeError("seist_b200.postprocess has no CPU path: the probabilities must live on a CUDA device")
    if outputs.dim() != 3 or not (0 <= channel < outputs.shape[1]):
        raise ValueError(f"expected (N, C, L) probabilities and a channel < C, got {tuple(outputs.shape)} / {channel}")
    return outputs.contiguous().float()


def pick_phase(outputs: torch.Tensor, channel: int, prob_threshold: float, min_peak_dist: int, topk: int,
               padding_value: int = PAD_PHASE) -> torch.Tensor:
    """`_pick_phase(outputs[:, channel], ...)`: (N, topk) int64 sample indices, padded with `padding_value`."""
    y = _check(outputs, channel)
    n, c, l = y.shape
    out = torch.empty(n, topk, dtype=torch.int64, device=y.device)
    _lib.check(_lib.lib().seist_pick_phase(y.data_ptr(), n, c, channel, l, float(prob_threshold), int(min_peak_dist), int(topk),
                                           int(padding_value), out.data_ptr(), _s()), "seist_pick_phase")
    return out


def detect_event(outputs: torch.Tensor, channel: int, prob_threshold: float, topk: int) -> torch.Tensor:
    """`_detect_event(outputs[:, channel], ...)`: (N, 2 * topk) int64 [on, off] pairs, padded with [1, 0]."""
    y = _check(outputs, channel)
    n, c, l = y.shape
    out = torch.empty(n, 2 * topk, dtype=torch.int64, device=y.device)
    _lib.check(_lib.lib().seist_detect_event(y.data_ptr(), n, c, channel, l, float(prob_threshold), int(topk), out.data_ptr(),
                                             _s()), "seist_detect_event")
    return out


def process_outputs(args, outputs: Union[Sequence[torch.Tensor], torch.Tensor], label_names: List, sampling_rate: int
                    ) -> Dict[str, torch.Tensor]:
    """Drop-in for the reference's `process_outputs(args, outputs, label_names, sampling_rate)` (postprocess.py:196-250):
    same argument meaning (`args.ppk_threshold`, `.spk_threshold`, `.det_threshold`, `.min_peak_dist`,
    `.max_detect_event_num`), same result dictionary, computed on the device without a host round trip."""
    outs = list(outputs) if isinstance(outputs, (tuple, list)) else [outputs]
    results: Dict[str, torch.Tensor] = {}
    for out, group in zip(outs, label_names):
        if isinstance(group, (tuple, list)):
            for i, name in enumerate(group):
                if name in ("ppk", "spk"):
                    thr = args.ppk_threshold if name == "ppk" else args.spk_threshold
                    results[name] = pick_phase(out, i, thr, int(args.min_peak_dist * sampling_rate), args.max_detect_event_num)
                elif name == "det":
                    results[name] = detect_event(out, i, args.det_threshold, args.max_detect_event_num)
                else:
                    tmp = out[:, i]
                    results[name] = tmp.unsqueeze(-1) if tmp.dim() < 2 else tmp
        else:
            results[group] = out
    return results


class StepCounters:
    """The counters behind the reference's `Metrics` for the tasks ppk / spk / det (utils/metrics.py:141-193,205-232) for any
    number of tasks and steps in ONE device vector, so that `synchronize()` is a single all-reduce instead of the
    reference's two barriers + one all-reduce per counter per task per step (utils/metrics.py:83-98)."""
    PICK = ("data_size", "tp", "predp", "possp", "sum_res", "sum_squ_res", "sum_abs_res")
    DET = ("data_size", "tp", "predp", "possp")

    def __init__(self, tasks: Sequence[str], num_samples: int, time_threshold_samples: int, device):
        self.tasks = list(tasks)
        self.num_samples, self.t_thres = int(num_samples), int(time_threshold_samples)
        self.off = {}
        n = 0
        for t in self.tasks:
            self.off[t] = n
            n += 7 if t in ("ppk", "spk") else 4
        self.acc = torch.zeros(max(n, 1), dtype=torch.float64, device=device)

    def update(self, task: str, targets: torch.Tensor, preds: torch.Tensor):
        t = targets.to(self.acc.device, torch.int64).contiguous()
        p = preds.to(self.acc.device, torch.int64).contiguous()
        acc_ptr = self.acc.data_ptr() + 8 * self.off[task]
        lib = _lib.lib()
        if task in ("ppk", "spk"):
            if t.shape != p.shape or (t.dim() == 2 and t.shape[1] != 1):
                raise NotImplementedError("StepCounters: one phase per waveform (max_detect_event_num = 1, the reference default)")
            _lib.check(lib.seist_pick_counters(t.data_ptr(), p.data_ptr(), t.numel(), self.num_samples, self.t_thres, acc_ptr, _s()),
                       "seist_pick_counters")
        else:
            n = t.shape[0]
            _lib.check(lib.seist_det_counters(t.data_ptr(), p.data_ptr(), n, t.numel() // (2 * n), p.numel() // (2 * n),
                                              self.num_samples, acc_ptr, _s()), "seist_det_counters")

    def synchronize(self):
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.acc)

    def result(self) -> Dict[str, Dict[str, float]]:
        """precision / recall / f1 (+ mean / rmse / mae over the true positives' residuals for the picks), metrics.py:300-383."""
        v = self.acc.cpu().tolist()
        out = {}
        for t in self.tasks:
            names = self.PICK if t in ("ppk", "spk") else self.DET
            d = dict(zip(names, v[self.off[t]:self.off[t] + len(names)]))
            eps = 1e-6
            pr = d["tp"] / (d["predp"] + eps)
            rc = d["tp"] / (d["possp"] + eps)
            d.update(precision=pr, recall=rc, f1=2 * pr * rc / (pr + rc + eps))
            if t in ("ppk", "spk") and d["data_size"] > 0:
                d.update(mean=d["sum_res"] / d["data_size"], rmse=(d["sum_squ_res"] / d["data_size"]) ** 0.5,
                         mae=d["sum_abs_res"] / d["data_size"])
            out[t] = d
        return out
