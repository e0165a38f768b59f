"""One training step of the bench workload (Llama-2-7B dims, `--layers` decoder layers, per-GPU batch B, S = 512) that
records every GEMM launch's (kind, M, N, K) IN LAUNCH ORDER to a JSON file. Run once plainly and once under
`ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:gemm` : the n-th GEMM row of
the ncu CSV is the n-th entry of the JSON, which is how profiles/gemm_traffic.json gets per-shape DRAM traffic for the
bench's OWN shapes (tools/gpu/r2_evidence.sh, tools/summarize_ncu.py gemm_traffic).

    python tools/gemm_shape_trace.py --out gpurun_out/gemm_shapes.json [--batch 32] [--layers 2]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", required=True)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--layers", type=int, default=2)
    a = ap.parse_args()
    from llark_b200 import ops
    from llark_b200.optim import LlarkAdamW

    torch.manual_seed(1234)
    model = bench.build_model(a.layers)
    opt = LlarkAdamW(model, lr=5e-5, max_grad_norm=1.0)
    batch = {k: v.cuda() for k, v in bench.synth_batch(a.batch, 1000).items()}
    ops.GEMM_PROFILE = []
    out = model(**batch)
    out.loss.backward()
    opt.step()
    torch.cuda.synchronize()
    rows = []
    for _, _, flops, shp in ops.GEMM_PROFILE:
        kind = {(0, 0): "fwd(K,K)", (0, 1): "dgrad(K,MN)", (1, 1): "wgrad(MN,MN)", (1, 0): "(MN,K)"}[shp[3:]]
        M, N, K = shp[:3]
        rows.append({"key": f"{kind} M{M} N{N} K{K}", "flops": flops,
                     "algorithmic_bytes": 2 * (M * K + N * K + M * N)})  # bf16 A + B read once, C written once
    with open(a.out, "w") as f:
        json.dump(rows, f)
    print(len(rows), "gemm launches traced")


if __name__ == "__main__":
    main()
