"""Run one warm GEMM of each operand-major kind at a Llama shape (for an ncu --set full capture)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from llark_b200 import ops
M, N, K = 4096, 11008, 4096
a = (torch.randn(M, K, device="cuda") * 0.1).to(torch.bfloat16)
b = (torch.randn(N, K, device="cuda") * 0.1).to(torch.bfloat16)
at, bt = a.t().contiguous(), b.t().contiguous()
out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
for _ in range(2):
    ops.gemm(a, b, out=out)
    ops.gemm(a, bt, b_mn=True, out=out)
    ops.gemm(at, bt, a_mn=True, b_mn=True, out=out)
torch.cuda.synchronize()
