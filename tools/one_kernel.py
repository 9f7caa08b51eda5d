"""Run one GEMM / conv shape a few times (for `ncu --set full -k regex:gemm_tcgen05`)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import t5_resnet_vqa_b200 as pkg
from util import Caller
BF = torch.bfloat16
C = Caller(pkg)
kind = sys.argv[1]
if kind == "conv":
    H, Cin, Cout, R, res, bn = [int(x) for x in sys.argv[2:8]]
    N = 64
    x = torch.randn(N, H, H, Cin, device="cuda").to(BF); w = torch.randn(Cout, R, R, Cin, device="cuda").to(BF)
    b = torch.randn(Cout, device="cuda"); out = torch.empty(N, H, H, Cout, device="cuda", dtype=BF)
    rs = torch.randn(N, H, H, Cout, device="cuda").to(BF) if res else None
    for _ in range(4):
        C.conv(N, H, H, Cin, Cout, R, 1, R // 2, x, w, out, bias=b, residual=rs, relu=1, bn=bn)
else:
    M, N, K, bn = [int(x) for x in sys.argv[2:6]]
    A = torch.randn(M, K, device="cuda").to(BF); B = torch.randn(N, K, device="cuda").to(BF)
    out = torch.empty(M, N, device="cuda", dtype=BF)
    for _ in range(4):
        C.gemm(M, N, K, A, K, 0, B, K, 0, out, N, 0, bn=bn)
torch.cuda.synchronize()
print("ok")
