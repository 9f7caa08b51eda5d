"""In-graph time of the 2048x768x768 GEMM under the epilogue variants the step uses.  Diagnostic only."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools"))
from gemm_bench import time_ours
BF = torch.bfloat16; dev = "cuda"
M, N, K = 2048, 768, 768
A = torch.randn(M, K, device=dev).to(BF); B = torch.randn(N, K, device=dev).to(BF); Bt = B.t().contiguous()
ob = torch.empty(M, N, device=dev, dtype=BF); of = torch.empty(M, N, device=dev)
res = torch.randn(M, N, device=dev); bias = torch.randn(N, device=dev); mask = torch.randn(M, N, device=dev).to(BF)
rng = torch.tensor([1234, 7], dtype=torch.int64, device=dev)
cases = [
    ("bf16 plain", lambda r: r.gemm(M, N, K, A, K, 0, B, K, 0, ob, N, 0)),
    ("bf16 + bias + relu + dropout", lambda r: r.gemm(M, N, K, A, K, 0, B, K, 0, ob, N, 0, bias=bias, relu=1, drop_p=0.1, sid=3, rng=rng)),
    ("fp32 + fp32 residual", lambda r: r.gemm(M, N, K, A, K, 0, B, K, 0, of, N, 1, residual=res, ldr=N, res_fp32=1)),
    ("fp32 + fp32 residual + dropout", lambda r: r.gemm(M, N, K, A, K, 0, B, K, 0, of, N, 1, residual=res, ldr=N, res_fp32=1, drop_p=0.1, sid=3, rng=rng)),
    ("fp32 + bias + residual + dropout", lambda r: r.gemm(M, N, K, A, K, 0, B, K, 0, of, N, 1, bias=bias, residual=res, ldr=N, res_fp32=1, drop_p=0.1, sid=3, rng=rng)),
    ("dgrad bf16 + relu mask + dropout", lambda r: r.gemm(M, N, K, A, K, 0, Bt, N, 1, ob, N, 0, relu_mask=mask, ldm=N, drop_p=0.1, sid=3, rng=rng)),
    ("dgrad fp32 + residual", lambda r: r.gemm(M, N, K, A, K, 0, Bt, N, 1, of, N, 1, residual=res, ldr=N, res_fp32=1)),
]
for name, fn in cases:
    print("%-36s %5.1f us" % (name, time_ours(fn)))
