"""NVLink peer-memory exchange state of one data-parallel rank (C side: csrc/comm.cu, `SeistComm`).

The reference wraps the model in DistributedDataParallel and converts every BatchNorm to SyncBatchNorm
(training/train.py:367-374): 2 x 115 tiny NCCL collectives per seist_m_dpk step plus the gradient all-reduce.
Here every rank allocates ONE symmetric-memory blob (torch.distributed._symmetric_memory: cuMem allocations
exchanged between the processes and mapped into every rank's address space) holding

    [ stat_acc : 2*sumC doubles | gstat_acc : 2*sumC doubles | flat gradient : numel floats | signal pad ]

and the kernels read the peers' parts directly over NVLink/NVSwitch: the BatchNorm statistic sum is fused into the
BN_PREPARE kernel, the gradient all-reduce is one kernel (`seist_comm_allreduce`), and a whole training step
contains no NCCL call - so it is captured into one CUDA graph on every rank.
"""
from __future__ import annotations

import ctypes
import os
from typing import List, Optional

import numpy as np
import torch

from . import _lib


def _align(n: int, a: int = 256) -> int:
    return (n + a - 1) // a * a


class PeerComm:
    def __init__(self, device: torch.device, world: int, rank: int, n_stat: int, n_grad: int,
                 blob: Optional[torch.Tensor] = None, peer_bases: Optional[List[int]] = None, group=None):
        """`blob` / `peer_bases`: pre-made buffers (tests run several virtual ranks inside one process); otherwise the
        blob comes from symmetric memory and `peer_bases` from the rendezvous over `group`."""
        if world > _lib.MAX_WORLD:
            raise ValueError(f"PeerComm supports up to {_lib.MAX_WORLD} ranks per node, got {world}")
        self.device, self.world, self.rank = device, world, rank
        self.n_stat, self.n_grad = n_stat, n_grad
        self.off_stat = 0
        self.off_gstat = _align(8 * n_stat)
        self.off_grad = self.off_gstat + _align(8 * n_stat)
        self.off_sig = self.off_grad + _align(4 * n_grad)
        self.nbytes = self.off_sig + _align(4 * _lib.SIG_LANES * _lib.MAX_WORLD)
        if blob is None:
            import torch.distributed as dist
            import torch.distributed._symmetric_memory as symm_mem
            blob = symm_mem.empty(self.nbytes, dtype=torch.uint8, device=device)
            blob.zero_()
            torch.cuda.synchronize(device)
            hdl = symm_mem.rendezvous(blob, group if group is not None else dist.group.WORLD)
            peer_bases = [int(p) for p in hdl.buffer_ptrs]
            self._hdl = hdl
            if len(peer_bases) != world or hdl.rank != rank:
                raise RuntimeError("symmetric-memory rendezvous does not match the process group")
            dist.barrier()          # every rank has zeroed its blob before anybody signals
        self.blob = blob
        self.peer_bases = list(peer_bases)
        assert self.peer_bases[rank] == blob.data_ptr()
        self.stat_acc = blob[self.off_stat:self.off_stat + 8 * n_stat].view(torch.float64)
        self.gstat_acc = blob[self.off_gstat:self.off_gstat + 8 * n_stat].view(torch.float64)
        self.grad = blob[self.off_grad:self.off_grad + 4 * n_grad].view(torch.float32)
        self.epoch = torch.zeros(_lib.SIG_LANES, dtype=torch.int32, device=device)
        self.err = torch.zeros(1, dtype=torch.int32, device=device)
        self.grad_red = torch.zeros(n_grad, dtype=torch.float32, device=device)    # all-reduced gradients (local)
        c = _lib.SeistComm()
        c.world, c.rank = world, rank
        for p in range(world):
            c.stat_peer[p] = self.peer_bases[p] + self.off_stat
            c.gstat_peer[p] = self.peer_bases[p] + self.off_gstat
            c.grad_peer[p] = self.peer_bases[p] + self.off_grad
            c.sig_peer[p] = self.peer_bases[p] + self.off_sig
        c.epoch = self.epoch.data_ptr()
        c.err = self.err.data_ptr()
        self.host = c
        self.dev = torch.from_numpy(np.frombuffer(bytes(c), dtype=np.uint8).copy()).to(device)

    @property
    def dev_ptr(self) -> int:
        return self.dev.data_ptr()

    def barrier(self, lane: int = 3, stream: Optional[int] = None):
        s = torch.cuda.current_stream().cuda_stream if stream is None else stream
        _lib.check(_lib.lib().seist_comm_barrier(self.dev_ptr, lane, s), "seist_comm_barrier")

    def allreduce_grads(self, stream: Optional[int] = None) -> torch.Tensor:
        """sum over ranks of the symmetric flat gradient buffers -> `grad_red` (local); includes both barriers."""
        s = torch.cuda.current_stream().cuda_stream if stream is None else stream
        _lib.check(_lib.lib().seist_comm_allreduce(self.dev_ptr, self.world, self.grad_red.data_ptr(), self.n_grad, s),
                   "seist_comm_allreduce")
        return self.grad_red

    def timed_out(self) -> bool:
        return bool(self.err.item())


def symmetric_memory_enabled() -> bool:
    """SEIST_SYMM=0 falls back to NCCL calls at the plan's sync points (the round-1 path)."""
    return os.environ.get("SEIST_SYMM", "1") != "0"
