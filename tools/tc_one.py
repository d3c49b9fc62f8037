"""One block-dense forward on a BASELINE config-3 shaped operand (for ncu): python tools/tc_one.py KIND M K N [iters]"""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from matrix_multiplication_b200 import capi
from matrix_multiplication_b200._capi import OP_FORCE_TC
kind, M, K, N = (int(x) for x in sys.argv[1:5])
iters = int(sys.argv[5]) if len(sys.argv) > 5 else 3
density = float(sys.argv[6]) if len(sys.argv) > 6 else 0.1
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(0)
A = (torch.rand(M, K, device=dev, generator=g) * (torch.rand(M, K, device=dev, generator=g) < density)).float()
B = torch.rand(K, N, device=dev, generator=g)
sp = A.to_sparse_csr()
crow, col, val = sp.crow_indices().int().contiguous(), sp.col_indices().int().contiguous(), sp.values().float().contiguous()
plan = capi.Plan(crow, col, val, (M, K))
capi.set_option("tc_kind", kind)
C = torch.empty(M, N, device=dev)
for _ in range(iters):
    plan.forward_ex(B, C, flags=OP_FORCE_TC)
torch.cuda.synchronize()
ref = A.double() @ B.double()
if len(sys.argv) > 7 and sys.argv[7] == "sddmm":
    G = torch.rand(M, N, device=dev, generator=g)
    dv = torch.empty(val.numel(), device=dev)
    for _ in range(iters):
        plan.backward_A_ex(G, B, dv, flags=OP_FORCE_TC)
    torch.cuda.synchronize()
print("tc_last_kernel", capi.get_option("tc_last_kernel"), "pair capacity", capi.get_option("tc_pair_capacity"))
print("rel_err", float((C.double() - ref).abs().max() / ref.abs().max()), "timeout", plan.tc_timeout())
