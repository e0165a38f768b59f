"""Time one GEMM shape in the three operand-major kinds (no correctness check)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from llark_b200 import ops
M, N, K = 4096, 11008, 4096
a = (torch.randn(M, K, device="cuda") * 0.1).to(torch.bfloat16)
b = (torch.randn(N, K, device="cuda") * 0.1).to(torch.bfloat16)
at, bt = a.t().contiguous(), b.t().contiguous()
out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
def t(fn):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): fn()
    e1.record(); torch.cuda.synchronize()
    return 2.0 * M * N * K / (e0.elapsed_time(e1) / 20) / 1e9
ref = a.float() @ b.float().t()
def err(fn):
    fn(); torch.cuda.synchronize()
    return ((out.float() - ref).norm() / ref.norm()).item()
print("rel err KK %.3f K,MN %.3f MN,MN %.3f" % (err(lambda: ops.gemm(a, b, out=out)), err(lambda: ops.gemm(a, bt, b_mn=True, out=out)),
      err(lambda: ops.gemm(at, bt, a_mn=True, b_mn=True, out=out))))
print(os.environ.get("LLARK_GEMM_DEBUG", "0"), "KK %.0f  K,MN %.0f  MN,MN %.0f TFLOP/s" % (
    t(lambda: ops.gemm(a, b, out=out)), t(lambda: ops.gemm(a, bt, b_mn=True, out=out)),
    t(lambda: ops.gemm(at, bt, a_mn=True, b_mn=True, out=out))), flush=True)
