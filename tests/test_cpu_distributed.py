"""not-gpu: the data-parallel path (world_size 2, gloo on CPU).  Each rank compiles the plan for its half of
the batch with world=2, executes it with the CPU plan interpreter and all-reduces the SyncBatchNorm (g)stat
slots exactly where the compiler placed the sync points; parameter gradients are then averaged with one
flat all-reduce.  The result must equal ONE process on the concatenated batch — the reference's semantics
for DDP + SyncBatchNorm with equal per-rank batches (training/train.py:369-374)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import seist_ref as R
from oracle.plan_interp import Interp
from seist_b200 import plan as P
from seist_b200.models import create_model

ZERO = dict(path_drop_rate=0, attn_drop_rate=0, key_drop_rate=0, mlp_drop_rate=0, other_drop_rate=0)
NAME, L, NB = "seist_s_dpk", 512, 4


def _model():
    from harness import randomize
    m = randomize(create_model(NAME, in_channels=3, in_samples=L), seed=5)
    m.set_drop_rates(**ZERO)
    return m.train()


def _run(plan, x, dy, world):
    it = Interp(plan)
    plan.x_in.x.copy_(x)
    plan.stat.zero_()
    for op in plan.fwd_ops:
        for b in op.sync_bn:
            e = plan.bns[b]
            dist.all_reduce(plan.stat[e.st_off:e.st_off + 2 * e.C])
        it.run_fwd_op(op)
    y = plan.y_out.x.clone()
    plan.gstat.zero_()
    plan.flat.G.zero_()
    plan.dWx.zero_()
    plan.y_out.dxd.copy_(dy)
    for op in plan.bwd_ops:
        for b in op.sync_bn:
            e = plan.bns[b]
            dist.all_reduce(plan.gstat[e.st_off:e.st_off + 2 * e.C])
        it.run_bwd_op(op)
    if world > 1:
        dist.all_reduce(plan.flat.G)
        plan.flat.G.div_(world)
    return y


def _worker(rank, world, store_path, q):
    # file rendezvous: no TCP port to race for (a port probed free by the parent can be taken before rank 0 binds it)
    os.environ.setdefault("GLOO_SOCKET_IFNAME", "lo")
    dist.init_process_group("gloo", init_method=f"file://{store_path}", rank=rank, world_size=world)
    torch.manual_seed(0)
    x, tgt = R.synth_waveforms(NB, L, seed=3)
    m = _model()
    flat = P.FlatState(m, torch.device("cpu"))
    n_loc = NB // world
    pl = P.PlanBuilder(m, flat, n_loc, L, training=True, world=world).build()
    P.allocate(pl, True)
    xs = x[rank * n_loc:(rank + 1) * n_loc]
    # dL/dy of the LOCAL mean loss (what each DDP rank back-propagates)
    with torch.no_grad():
        y0 = _run_forward_only(pl, xs)
    p = y0.clone().requires_grad_(True)
    R.bce_loss(p, tgt[rank * n_loc:(rank + 1) * n_loc]).backward()
    y = _run(pl, xs, p.grad, world)
    if rank == 0:
        q.put((y, flat.G.clone(), flat.RB.clone()))
    dist.barrier()
    dist.destroy_process_group()


def _run_forward_only(plan, x):
    it = Interp(plan)
    plan.x_in.x.copy_(x)
    plan.stat.zero_()
    rb = plan.flat.RB.clone()
    nbt = plan.flat.NBT.clone()
    for op in plan.fwd_ops:
        for b in op.sync_bn:
            e = plan.bns[b]
            dist.all_reduce(plan.stat[e.st_off:e.st_off + 2 * e.C])
        it.run_fwd_op(op)
    plan.flat.RB.copy_(rb)          # undo the running-stat update of this probe pass
    plan.flat.NBT.copy_(nbt)
    return plan.y_out.x.clone()


def test_two_ranks_equal_one_process_on_concatenated_batch():
    import tempfile
    store_path = os.path.join(tempfile.mkdtemp(prefix="seist_gloo_"), "rendezvous")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, store_path, q)) for r in range(2)]
    for p_ in procs:
        p_.start()
    y2, g2, rb2 = q.get(timeout=600)
    for p_ in procs:
        p_.join(timeout=600)
        assert p_.exitcode == 0

    # single process, whole batch (no process group needed: sync lists are empty for world=1)
    x, tgt = R.synth_waveforms(NB, L, seed=3)
    m = _model()
    flat = P.FlatState(m, torch.device("cpu"))
    pl = P.PlanBuilder(m, flat, NB, L, training=True, world=1).build()
    P.allocate(pl, True)
    it = Interp(pl)
    y1 = it.run_fwd(x).clone()
    p = y1.clone().requires_grad_(True)
    R.bce_loss(p, tgt).backward()
    it.run_bwd(p.grad)
    assert (y2 - y1[:NB // 2]).abs().max().item() < 1e-5
    gmax = flat.G.abs().max().item()
    assert (g2 - flat.G).abs().max().item() < 2e-4 * gmax
    assert (rb2 - flat.RB).abs().max().item() < 1e-4 * (flat.RB.abs().max().item() + 1e-3)
