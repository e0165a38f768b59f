"""Time the attention kernels at the bench shape (B 16, S 512, 32 heads, d 128) and check them against fp32 torch.

    python tools/attn_time.py            # tcgen05 forward (default)
    LLARK_ATTN_TC=0 python tools/attn_time.py   # mma.sync forward
Prints one JSON line: forward / backward microseconds and effective causal TFLOP/s (4*S^2*d*NH*B/2 forward)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from llark_b200 import ops

B, S, NH, D = int(os.environ.get("ATTN_B", 16)), int(os.environ.get("ATTN_S", 512)), 32, 128
torch.manual_seed(0)
qkv = (torch.randn(B * S, 3 * NH * D, device="cuda") * 0.8).to(torch.bfloat16)
out, lse = ops.attn_fwd(qkv, None, B, S, NH, D)
# fp32 reference on two (batch, head) pairs
err = 0.0
for (b, h) in ((0, 0), (B - 1, NH - 1)):
    x = qkv.view(B, S, 3, NH, D)[b, :, :, h].float()
    q, k, v = x[:, 0], x[:, 1], x[:, 2]
    sc = (q @ k.t()) / D ** 0.5
    sc = sc.masked_fill(torch.triu(torch.ones(S, S, device="cuda", dtype=torch.bool), 1), float("-inf"))
    ref = torch.softmax(sc, -1) @ v
    got = out.view(B, S, NH, D)[b, :, h].float()
    err = max(err, ((got - ref).norm() / ref.norm()).item())
    lref = torch.logsumexp(sc, -1)
    err_l = (lse[b, h] - lref).abs().max().item()
dout = torch.randn_like(out)
ws = torch.empty(ops.lib().llark_attn_bwd_workspace(B, S, NH), device="cuda", dtype=torch.uint8)
dqkv = torch.empty_like(qkv)
# backward against fp32 autograd on one (batch, head)
ops.attn_bwd(qkv, None, out, dout, lse, B, S, NH, D, dqkv=dqkv, workspace=ws)
b, h = B - 1, 1
x = qkv.view(B, S, 3, NH, D)[b, :, :, h].float().clone().requires_grad_(True)
sc = (x[:, 0] @ x[:, 1].t()) / D ** 0.5
sc = sc.masked_fill(torch.triu(torch.ones(S, S, device="cuda", dtype=torch.bool), 1), float("-inf"))
(torch.softmax(sc, -1) @ x[:, 2]).backward(dout.view(B, S, NH, D)[b, :, h].float())
gref = x.grad  # [S, 3, D]
ggot = dqkv.view(B, S, 3, NH, D)[b, :, :, h].float()
err_b = [((ggot[:, i] - gref[:, i]).norm() / gref[:, i].norm()).item() for i in range(3)]

def t(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1000.0

us_f = t(lambda: ops.attn_fwd(qkv, None, B, S, NH, D, out=out, lse=lse))
us_b = t(lambda: ops.attn_bwd(qkv, None, out, dout, lse, B, S, NH, D, dqkv=dqkv, workspace=ws))
fl = 4.0 * S * S * D * NH * B / 2
print(json.dumps({"attn_tc": os.environ.get("LLARK_ATTN_TC", "1"), "B": B, "S": S, "fwd_us": round(us_f, 1),
                  "bwd_us": round(us_b, 1), "fwd_tflops_causal": round(fl / us_f / 1e6, 1),
                  "bwd_tflops_causal": round(2.5 * fl / us_b / 1e6, 1), "fwd_rel_err": err, "lse_abs_err": err_l, "bwd_rel_err_qkv": err_b}))
