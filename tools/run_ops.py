"""Run selected plan ops of the bench configuration a few times (for ncu captures).
usage: python tools/run_ops.py <model> <batch> <reps> <name-substring> [<name-substring> ...]"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from seist_b200 import _lib  # noqa: E402
from seist_b200.models import create_model  # noqa: E402

model, batch, reps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
pats = sys.argv[4:]
torch.manual_seed(0)
m = create_model(model, in_channels=3, in_samples=8192).cuda().train()
x = torch.randn(batch, 3, 8192, device="cuda")
y = m(x)
y.square().mean().backward()          # populate every buffer once
torch.cuda.synchronize()
plan = m.engine().last_plan
lib = _lib.lib()
torch.cuda.profiler.start()      # ncu --profile-from-start off: only the selected launches are captured
size = ctypes.sizeof(_lib.SeistOp)
for ops, c_ops in ((plan.fwd_ops, plan.c_fwd), (plan.bwd_ops, plan.c_bwd)):
    base = ctypes.addressof(c_ops)
    for i, op in enumerate(ops):
        if any(op.name == p or (p.endswith("*") and op.name.startswith(p[:-1])) for p in pats):
            for _ in range(reps):
                _lib.check(lib.seist_plan_run(base + i * size, 1, torch.cuda.current_stream().cuda_stream))
            torch.cuda.synchronize()
            print("ran", op.name)
torch.cuda.profiler.stop()
