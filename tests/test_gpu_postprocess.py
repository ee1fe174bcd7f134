"""-m gpu: device post-processing (seist_pick_phase / seist_detect_event / counters through the C-ABI) against the numpy
oracle pinned to the reference's `_detect_peaks` (tests/test_cpu_postprocess.py) — integer outputs bit for bit."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from oracle import postprocess_ref as PR
from seist_b200 import postprocess as PP

pytestmark = pytest.mark.gpu


def _traces(rng, n, L):
    t = np.arange(L, dtype=np.float32)
    out = np.zeros((n, 3, L), dtype=np.float32)
    for i in range(n):
        for c in range(3):
            for _ in range(rng.integers(0, 6)):
                ctr, w, a = rng.integers(0, L), rng.uniform(3, 60), rng.uniform(0.1, 1.0)
                out[i, c] += a * np.exp(-((t - ctr) ** 2) / (2 * w * w)).astype(np.float32)
            out[i, c] += 0.02 * rng.standard_normal(L).astype(np.float32)
            out[i, c] = (out[i, c] - out[i, c].min()) / max(1.0, float(out[i, c].max() - out[i, c].min()) * 1.01)
    return out


@pytest.mark.parametrize("L,topk,mpd", [(8192, 1, 50), (8192, 3, 50), (1000, 5, 8), (2048, 2, 200)])
def test_pick_and_detect_match_oracle(L, topk, mpd):
    rng = np.random.default_rng(L + topk)
    x = _traces(rng, 40, L)
    x[0] = 0.0
    x[1, 0, :] = 0.9                       # one run covering the whole trace
    x[2, 0, :] = 0.0
    x[2, 0, L - 1] = 0.9                   # single-sample run at the end
    x[2, 0, 10:14] = 0.6
    x[2, 0, 30:34] = 0.6                   # equal lengths: earlier first
    x[3, 1, 100:110] = 0.8                 # flat-topped peak
    xg = torch.from_numpy(x).cuda()
    for thr in (0.3, 0.05):
        for ch in (1, 2):
            got = PP.pick_phase(xg, ch, thr, mpd, topk).cpu().numpy()
            want = PR.pick_phase(x[:, ch], thr, mpd, topk)
            assert np.array_equal(got, want), (thr, ch, np.argwhere(got != want)[:4])
        got = PP.detect_event(xg, 0, 0.5 if thr > 0.1 else 0.2, topk).cpu().numpy()
        want = PR.detect_event(x[:, 0], 0.5 if thr > 0.1 else 0.2, topk)
        assert np.array_equal(got, want), (thr, np.argwhere(got != want)[:4])


def test_process_outputs_and_counters():
    rng = np.random.default_rng(7)
    L, N = 4096, 64
    x = _traces(rng, N, L)
    xg = torch.from_numpy(x).cuda()
    args = SimpleNamespace(ppk_threshold=0.3, spk_threshold=0.3, det_threshold=0.5, min_peak_dist=1.0, max_detect_event_num=1)
    res = PP.process_outputs(args, xg, [["det", "ppk", "spk"]], sampling_rate=50)
    assert set(res) == {"det", "ppk", "spk"} and res["ppk"].shape == (N, 1) and res["det"].shape == (N, 2)
    tp = torch.from_numpy(PR.pick_phase(x[:, 1], 0.3, 50, 1)) + torch.randint(-8, 9, (N, 1))
    td = torch.from_numpy(PR.detect_event(x[:, 0], 0.5, 1))
    td[:, 0] -= 7
    ctr = PP.StepCounters(["ppk", "det"], L, 5, "cuda")
    for _ in range(2):                                  # accumulates over steps
        ctr.update("ppk", tp.cuda(), res["ppk"])
        ctr.update("det", td.cuda(), res["det"])
    ctr.synchronize()
    got = ctr.result()
    wp = PR.pick_counters(tp.numpy(), res["ppk"].cpu().numpy(), L, 5)
    wd = PR.det_counters(td.numpy(), res["det"].cpu().numpy(), L)
    for k, v in wp.items():
        assert got["ppk"][k] == 2 * v, (k, got["ppk"][k], v)
    for k, v in wd.items():
        assert got["det"][k] == 2 * v, (k, got["det"][k], v)
    assert 0.0 <= got["ppk"]["f1"] <= 1.0


def test_cpu_tensor_fails_loudly():
    with pytest.raises(RuntimeError):
        PP.pick_phase(torch.zeros(1, 3, 64), 1, 0.3, 50, 1)
