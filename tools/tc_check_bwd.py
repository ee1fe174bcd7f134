"""GPU check of the tensor-core weight-gradient kernel: CONV_BWD_W ops vs the CPU interpreter (teacher forcing)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch  # noqa: E402

from harness import build_pair, push_state, rel_err, run_gpu_op  # noqa: E402
from seist_b200 import _lib  # noqa: E402

name, N, L = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
drops = dict(path_drop_rate=0.2, attn_drop_rate=0.1, key_drop_rate=0.1, mlp_drop_rate=0.2, other_drop_rate=0.1) if len(sys.argv) > 4 else None
p_cpu, p_gpu, it, _, _ = build_pair(name, N, L, True, drops)
torch.manual_seed(1)
p_cpu.step_seed.fill_(77)
it.run_fwd(torch.randn(N, 3, L))
p_cpu.gstat.zero_(); p_cpu.flat.G.zero_(); p_cpu.dWx.zero_()
g = torch.Generator().manual_seed(2)
p_cpu.y_out.dxd.copy_(torch.randn(p_cpu.y_out.dxd.shape, generator=g) / p_cpu.y_out.dxd[0].numel() ** 0.5)
worst = 0.0
for i, (bc, bg) in enumerate(zip(p_cpu.bwd_ops, p_gpu.bwd_ops)):
    push_state(p_cpu, p_gpu)
    if bc.kind == _lib.ATT_BWD_KV:
        run_gpu_op(p_gpu, p_gpu.c_bwd, i - 1)
    g_before = p_cpu.flat.G.clone()
    it.run_bwd_op(bc)
    run_gpu_op(p_gpu, p_gpu.c_bwd, i)
    if bc.kind == _lib.CONV_BWD_W and bc.fwd.k == 1:
        f = bc.fwd
        sl = slice(f.W.off, f.W.off + f.W.numel)
        err, ref = rel_err(p_gpu.flat.G[sl] - g_before[sl].cuda(), p_cpu.flat.G[sl] - g_before[sl])
        berr = 0.0
        if f.bias is not None:
            bs = slice(f.bias.off, f.bias.off + f.bias.numel)
            berr, _ = rel_err(p_gpu.flat.G[bs], p_cpu.flat.G[bs])
        worst = max(worst, err, berr)
        if err > 1e-4 or berr > 1e-4 or i % 40 == 0:
            print(f"bwd[{i}] {bc.name:50s} Cin={f.Cin:3d} Cout={f.Cout:3d} L={f.L_out:5d} dW rel_err={err:.2e} dbias={berr:.2e}")
print("worst", worst, "tc_error_flag", _lib.lib().seist_tc_error_flag())
