"""Timeline of CTA 0 of the tensor-core kernel (producer lane, MMA lane, first epilogue warp) on a config-3 shaped operand:
python tools/tc_trace.py [M K N density] [name=value ...]   (tc_terms bit 9 switches the recording on)"""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from matrix_multiplication_b200 import capi
from matrix_multiplication_b200._capi import OP_FORCE_TC

nums = [a for a in sys.argv[1:] if "=" not in a]
M, K, N = (int(x) for x in nums[:3]) if len(nums) >= 3 else (3072, 768, 4096)
density = float(nums[3]) if len(nums) > 3 else 0.1
terms = 7
for a in sys.argv[1:]:
    if "=" in a:
        k, v = a.split("=")
        if k == "tc_terms":
            terms = int(v)
        else:
            capi.set_option(k, int(v))
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(0)
A = (torch.randn(M, K, device=dev, generator=g) * (torch.rand(M, K, device=dev, generator=g) < density)).float()
sp = A.to_sparse_csr()
crow, col, val = sp.crow_indices().int().contiguous(), sp.col_indices().int().contiguous(), sp.values().float().contiguous()
B = torch.randn(K, N, device=dev, generator=g)
C = torch.empty(M, N, device=dev)
plan = capi.Plan(crow, col, val, (M, K))
capi.set_option("tc_terms", terms)
for _ in range(3):
    plan.forward_ex(B, C, flags=OP_FORCE_TC)
torch.cuda.synchronize()
capi.set_option("tc_trace_sync", 0)   # clear
capi.set_option("tc_terms", terms | 512)
plan.forward_ex(B, C, flags=OP_FORCE_TC)
torch.cuda.synchronize()
capi.set_option("tc_terms", terms)
capi.set_option("tc_trace_sync", 1)   # copy back
REGION = 2048
names = {1: "P wait slot", 2: "P slot free -> copies", 10: "M wait acc", 11: "M acc free", 12: "M wait operands",
         13: "M operands landed", 20: "E wait sums", 21: "E sums ready", 22: "E drained", 23: "E store begins", 24: "E stored", 25: "E scales fetched",
         30: "E g0 filled", 31: "E g1 filled", 32: "E g2 filled", 33: "E g3 filled",
         40: "E g0 written", 41: "E g1 written", 42: "E g2 written", 43: "E g3 written"}
ev = []
for role in range(3):
    capi.set_option("tc_trace_idx", role * REGION)
    words = [capi.get_option("tc_trace_word") for _ in range(REGION)]
    for i in range(0, REGION, 2):
        if words[i] == 0:
            break
        ev.append((words[i + 1], role, words[i]))
ev.sort()
t0 = ev[0][0]
print(f"{len(ev)} events; times in us since the first event (globaltimer)")
last = {0: None, 1: None, 2: None}
for t, role, code in ev:
    d = "" if last[role] is None else f"(+{(t - last[role]) / 1e3:.2f})"
    last[role] = t
    print(f"{(t - t0) / 1e3:9.2f}  {'   ' * role * 9}{names.get(code, code)} {d}")
