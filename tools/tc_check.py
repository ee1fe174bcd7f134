"""Quick GPU check of the tcgen05 pointwise kernel: forward ops of a small plan vs the CPU interpreter."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch  # noqa: E402

from harness import build_pair, push_state, rel_err, run_gpu_op  # noqa: E402
from seist_b200 import _lib  # noqa: E402

name, N, L, training = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), bool(int(sys.argv[4]))
p_cpu, p_gpu, it, _, _ = build_pair(name, N, L, training)
torch.manual_seed(1)
p_cpu.x_in.x.copy_(torch.randn(N, 3, L))
p_cpu.stat.zero_()
worst = 0.0
for i, (fc, fg) in enumerate(zip(p_cpu.fwd_ops, p_gpu.fwd_ops)):
    push_state(p_cpu, p_gpu)
    it.run_fwd_op(fc)
    run_gpu_op(p_gpu, p_gpu.c_fwd, i)
    if fc.kind == _lib.CONV_FWD and fc.k == 1 and fc.pool == 1:
        sl = slice(fc.out.c0, fc.out.c0 + fc.out.C)
        err, ref = rel_err(fg.out.buf.x[:, sl], fc.out.buf.x[:, sl])
        tag = "TC?" if (8 <= fc.Cout <= 128 and fc.L_out % 4 == 0) else "simt"
        worst = max(worst, err)
        if err > 1e-4 or i < 12:
            print(f"fwd[{i}] {fc.name:45s} Cin={fc.Cin:3d} Cout={fc.Cout:3d} L={fc.L_out:5d} {tag} rel_err={err:.2e}")
        if training and fc.out.bn >= 0:
            e = p_cpu.bns[fc.out.bn]
            serr, _ = rel_err(p_gpu.stat[e.st_off:e.st_off + 2 * e.C], p_cpu.stat[e.st_off:e.st_off + 2 * e.C])
            if serr > 1e-4:
                print(f"      stat rel_err={serr:.2e}")
print("worst rel err", worst, "tc_error_flag", _lib.lib().seist_tc_error_flag())
