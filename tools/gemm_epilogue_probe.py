"""Where does the GEMM epilogue time go? Times C[M,N] = A[M,K] @ B[N,K]^T for the prior's c_proj shape (K = 1200, the
epilogue-bound case) and a K = 4800 control with every output variant: bf16 / fp32 output, with / without a residual.
Run once per epilogue mode (LLARK_GEMM_EPILOGUE=staged|direct is read once per process):

    python tools/gemm_epilogue_probe.py            # auto
    LLARK_GEMM_EPILOGUE=direct python tools/gemm_epilogue_probe.py
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from llark_b200 import ops  # noqa: E402


def time_gemm(M, N, K, out_dtype, residual, reps=8):
    a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    b = torch.randn(N, K, device="cuda").to(torch.bfloat16)
    out = torch.empty(M, N, device="cuda", dtype=out_dtype)
    r = torch.randn(M, N, device="cuda").to(out_dtype) if residual else None
    for _ in range(2):
        ops.gemm(a, b, out=out, residual=r)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        ops.gemm(a, b, out=out, residual=r)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    tiles = ((M + 255) // 256) * ((N + 255) // 256)
    return {"ms": round(ms, 4), "tflops": round(2.0 * M * N * K / ms / 1e9, 1),
            "us_per_tile_per_cluster": round(ms * 1e3 / (tiles / 74.0), 2)}


def main():
    res = {"mode": os.environ.get("LLARK_GEMM_EPILOGUE", "auto")}
    for K in (1200, 4800):
        for dt, name in ((torch.bfloat16, "bf16"), (torch.float32, "fp32")):
            for resid in (False, True):
                res[f"K{K} {name} {'R' if resid else 'noR'}"] = time_gemm(65536, 4800, K, dt, resid)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
