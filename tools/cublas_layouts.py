"""How does cuBLAS (torch.matmul) fare on the MN-major operand layouts at a Llama shape?"""
import torch
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
print("cublas  A[M,K]@B[N,K]^T %.0f | A[M,K]@Bt[K,N] %.0f | At[K,M]^T@Bt[K,N] %.0f TFLOP/s" % (
    t(lambda: torch.matmul(a, b.t(), out=out)), t(lambda: torch.matmul(a, bt, out=out)),
    t(lambda: torch.matmul(at.t(), bt, out=out))), flush=True)
