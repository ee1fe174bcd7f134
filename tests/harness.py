"""Shared test helpers: build CPU (interpreter) and CUDA plans of the same model and compare them
op by op with teacher forcing (the GPU op always starts from the interpreter's state)."""
import copy
import ctypes

import torch

from oracle.plan_interp import Interp
from seist_b200 import _lib
from seist_b200 import plan as P
from seist_b200.models import create_model

ZERO_DROPS = dict(path_drop_rate=0, attn_drop_rate=0, key_drop_rate=0, mlp_drop_rate=0, other_drop_rate=0)


def randomize(model, seed=0, wstd=0.25):
    """Non-degenerate parameters / running statistics (random-init eval output is a flat 0.5)."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if p.dim() > 1:
                fan = p[0].numel()
                p.copy_(torch.randn(p.shape, generator=g) * (1.0 / fan ** 0.5))
            elif name.endswith("norm.weight") or ".norm" in name and name.endswith("weight") or "norms." in name and name.endswith("weight"):
                p.copy_(1.0 + 0.3 * torch.randn(p.shape, generator=g))
            else:
                p.copy_(0.2 * torch.randn(p.shape, generator=g))
        for name, b in model.named_buffers():
            if name.endswith("running_mean"):
                b.copy_(0.1 * torch.randn(b.shape, generator=g))
            elif name.endswith("running_var"):
                b.copy_(0.5 + torch.rand(b.shape, generator=g))
    return model


def build_pair(name, N, L, training, drops=None, seed=0, hp_overrides=None, state_dict=None):
    """(cpu_plan, gpu_plan, interp) for two identical copies of model `name`."""
    m_cpu = create_model(name, in_channels=3, in_samples=L, **(hp_overrides or {}))
    if state_dict is not None:
        m_cpu.load_state_dict(state_dict, strict=True)
    else:
        randomize(m_cpu, seed)
    m_cpu.set_drop_rates(**(ZERO_DROPS if drops is None else drops))
    m_cpu.train(training)
    m_gpu = copy.deepcopy(m_cpu)
    f_cpu = P.FlatState(m_cpu, torch.device("cpu"))
    p_cpu = P.PlanBuilder(m_cpu, f_cpu, N, L, training).build()
    P.allocate(p_cpu, training)
    dev = torch.device("cuda:0")
    m_gpu.to(dev)
    f_gpu = P.FlatState(m_gpu, dev)
    p_gpu = P.finalize(P.PlanBuilder(m_gpu, f_gpu, N, L, training).build(), training)
    assert p_cpu.arena.numel() == p_gpu.arena.numel()
    return p_cpu, p_gpu, Interp(p_cpu), m_cpu, m_gpu


def push_state(p_cpu, p_gpu):
    p_gpu.arena.copy_(p_cpu.arena)
    p_gpu.stat.copy_(p_cpu.stat)
    p_gpu.gstat.copy_(p_cpu.gstat)
    p_gpu.flat.G.copy_(p_cpu.flat.G)
    p_gpu.flat.RB.copy_(p_cpu.flat.RB)
    p_gpu.step_seed.copy_(p_cpu.step_seed)
    p_gpu.Wx.copy_(p_cpu.Wx)
    p_gpu.dWx.copy_(p_cpu.dWx)


def run_gpu_op(p_gpu, c_ops, i):
    base = ctypes.addressof(c_ops) + i * ctypes.sizeof(_lib.SeistOp)
    _lib.check(_lib.lib().seist_plan_run(base, 1, torch.cuda.current_stream().cuda_stream), f"op {i}")
    torch.cuda.synchronize()


def rel_err(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).abs().max() / (b.abs().max() + 1e-20)).item(), b.abs().max().item()
