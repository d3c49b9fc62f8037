"""One kernel family on one BASELINE operand, for ncu:  python tools/prof_one.py WHAT
  c3_tc_f16 | c3_tc_tf32 | c3_sddmm_tc | c5_n16_gather | c4_sddmm | c5_tc_256"""
import sys
from pathlib import Path
import numpy as np
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from matrix_multiplication_b200 import capi, synthetic
from matrix_multiplication_b200._capi import OP_FORCE_TC, OP_NO_TC
what = sys.argv[1]
dev = torch.device("cuda")


def up(*a):
    return [torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in a]


if what.startswith("c3"):
    rp, ci, va = synthetic.bernoulli_csr(3072, 768, 0.9, seed=0)
    M, K, N = 3072, 768, 4096
elif what.startswith("c5"):
    rp, ci, va = synthetic.bernoulli_csr(16384, 16384, 0.9, seed=0)
    M = K = 16384
    N = 16 if "n16" in what else 256
else:
    rp, ci, va, (M, K) = synthetic.rmat_csr(20, 16, seed=0)
    N = 128
crow, col, val = up(rp, ci, va)
B = torch.rand(K, N, device=dev)
G = torch.rand(M, N, device=dev)
C = torch.empty(M, N, device=dev)
dv = torch.empty(len(va), device=dev)
plan = capi.Plan(crow, col, val, (M, K))
if "tf32" in what:
    capi.set_option("tc_kind", 0)
for _ in range(3):
    if what in ("c3_tc_f16", "c3_tc_tf32", "c5_tc_256"):
        plan.forward_ex(B, C, flags=OP_FORCE_TC)
    elif what == "c3_sddmm_tc":
        plan.backward_A_ex(G, B, dv, flags=OP_FORCE_TC)
    elif what == "c5_n16_gather":
        capi.set_option("slab_mode", 0)
        plan.forward_ex(B, C, flags=OP_NO_TC)
    elif what == "c5_n16_slab":
        plan.forward_ex(B, C, flags=OP_NO_TC)
    elif what == "c4_sddmm":
        plan.backward_A_ex(G, B, dv, flags=OP_NO_TC)
torch.cuda.synchronize()
print(what, "ok", plan.tc_timeout())
