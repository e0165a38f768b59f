"""MN-major B operand: throughput vs leading dimension (is the slow-down a page/TLB effect of the k-row stride?)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from llark_b200 import ops
def t(fn, flops):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): fn()
    e1.record(); torch.cuda.synchronize()
    return flops / (e0.elapsed_time(e1) / 20) / 1e9
for (M, N, K) in ((16384, 256, 8192), (16384, 512, 8192), (8192, 1024, 8192), (8192, 2048, 8192), (4096, 4096, 8192),
                  (4096, 8192, 8192), (4096, 16384, 8192), (4096, 4096, 1024), (4096, 4096, 2048)):
    a = (torch.randn(M, K, device="cuda") * 0.1).to(torch.bfloat16)
    b = (torch.randn(N, K, device="cuda") * 0.1).to(torch.bfloat16)
    bt = b.t().contiguous()
    at = a.t().contiguous()
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    fl = 2.0 * M * N * K
    print(f"M{M} N{N} K{K}: KK {t(lambda: ops.gemm(a, b, out=out), fl):.0f}  K,MN {t(lambda: ops.gemm(a, bt, b_mn=True, out=out), fl):.0f}"
          f"  MN,K {t(lambda: ops.gemm(at, b, a_mn=True, out=out), fl):.0f}  MN,MN {t(lambda: ops.gemm(at, bt, a_mn=True, b_mn=True, out=out), fl):.0f}", flush=True)
    del a, b, bt, at, out
