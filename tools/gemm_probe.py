"""GEMM probe: prints a relative-error table for a ladder of shapes (one operand-major combination per process,
so a trap in one combination cannot poison the others). Usage: python tools/gemm_probe.py <a_mn> <b_mn> <fp32>"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from llark_b200 import ops

a_mn, b_mn, fp32 = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
from llark_b200._lib import lib
print("gemm mode (1 = CTA pair):", lib().llark_gemm_get_mode(), flush=True)
dt = torch.float32 if fp32 else torch.bfloat16
shapes = [(128, 256, 16), (128, 256, 32), (128, 256, 64), (128, 256, 128), (128, 256, 256), (128, 256, 512),
          (256, 256, 64), (128, 512, 64), (256, 512, 320), (4096, 4096, 4096), (64, 64, 64), (240, 4096, 4800)]
if fp32:
    shapes = [(128, 256, 8), (128, 256, 32), (128, 256, 64), (256, 512, 256), (2048, 3600, 4800)]
for (M, N, K) in shapes:
    torch.manual_seed(0)
    a = (torch.randn(M, K, device="cuda") * 0.5).to(dt)
    b = (torch.randn(N, K, device="cuda") * 0.5).to(dt)
    ref = a.double() @ b.double().t()
    ai = a.t().contiguous() if a_mn else a
    bi = b.t().contiguous() if b_mn else b
    try:
        out = ops.gemm(ai, bi, a_mn=bool(a_mn), b_mn=bool(b_mn))
        torch.cuda.synchronize()
        rel = ((out.double() - ref).norm() / ref.norm()).item()
        # quadrant diagnostics for the single-tile cases
        msg = ""
        if rel > 1e-2 and M <= 256:
            d = (out.double() - ref).abs()
            msg = " rowerr[0:32,32:64,64:96,96:128]=%s colerr[by64]=%s" % (
                [round(d[i:i + 32].mean().item(), 3) for i in range(0, min(M, 128), 32)],
                [round(d[:, j:j + 64].mean().item(), 3) for j in range(0, min(N, 256), 64)])
        print(f"probe a_mn={a_mn} b_mn={b_mn} fp32={fp32} M={M} N={N} K={K} rel={rel:.3e}{msg}", flush=True)
    except Exception as e:  # noqa
        print(f"probe a_mn={a_mn} b_mn={b_mn} fp32={fp32} M={M} N={N} K={K} EXC {e}", flush=True)
        break
# quick timing of the big one
if not fp32:
  for (M, N, K) in ((8192, 8192, 8192), (4096, 12288, 4096), (4096, 4096, 11008), (4096, 22016, 4096)):
      a = torch.randn(M, K, device="cuda").to(dt); b = torch.randn(N, K, device="cuda").to(dt)
      ai = a.t().contiguous() if a_mn else a
      bi = b.t().contiguous() if b_mn else b
      out = torch.empty(M, N, device="cuda", dtype=dt)
      for _ in range(3): ops.gemm(ai, bi, a_mn=bool(a_mn), b_mn=bool(b_mn), out=out)
      torch.cuda.synchronize()
      e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      e0.record()
      for _ in range(10): ops.gemm(ai, bi, a_mn=bool(a_mn), b_mn=bool(b_mn), out=out)
      e1.record(); torch.cuda.synchronize()
      ms = e0.elapsed_time(e1) / 10
      print(f"timing a_mn={a_mn} b_mn={b_mn} {M}x{N}x{K}: {ms:.3f} ms  {2*M*N*K/ms/1e9:.1f} TFLOP/s", flush=True)
      ref = torch.matmul(a, b.t())
      t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
      for _ in range(3): torch.matmul(a, b.t(), out=ref)
      t0.record()
      for _ in range(10): torch.matmul(a, b.t(), out=ref)
      t1.record(); torch.cuda.synchronize()
      ms2 = t0.elapsed_time(t1) / 10
      print(f"timing cublas {M}x{N}x{K}: {ms2:.3f} ms  {2*M*N*K/ms2/1e9:.1f} TFLOP/s", flush=True)
