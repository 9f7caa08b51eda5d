"""Mainloop diagnosis: time GEMM shapes with VQA_B200_GEMM_DBG = 0 (normal) / 1 (no MMAs) / 2 (no loads), replayed as a
CUDA graph (device time per launch).  Diagnostic only."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# needs the diagnostic build (python t5-resnet-vqa_b200/build.py --debug): the default library has no instrumentation
_DBG = os.path.join(ROOT, "t5-resnet-vqa_b200", "libvqa_b200_dbg.so")
if not os.path.exists(_DBG):
    raise SystemExit("build the diagnostic library first: python t5-resnet-vqa_b200/build.py --debug")
os.environ.setdefault("VQA_B200_LIB", _DBG)
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools"))
from gemm_bench import time_ours
BF = torch.bfloat16; dev = "cuda"
shapes = [(2048, 768, 3072, 64), (2048, 768, 3072, 128), (2048, 768, 3072, 256), (2048, 3072, 768, 256),
          (8192, 3072, 768, 128), (8192, 3072, 768, 256), (2048, 768, 768, 128), (2048, 768, 768, 64)]
if len(sys.argv) > 1:
    shapes = [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]]
for M, N, K, bn in shapes:
    A = torch.randn(M, K, device=dev).to(BF); B = torch.randn(N, K, device=dev).to(BF)
    out = torch.empty(M, N, device=dev, dtype=BF)
    us = time_ours(lambda r: r.gemm(M, N, K, A, K, 0, B, K, 0, out, N, 0, bn=bn))
    tiles = ((M + 127) // 128) * ((N + bn - 1) // bn)
    waves = (tiles + 147) // 148
    print("dbg=%s M%d N%d K%d bn%d: %.1f us  (%d tiles, %d waves, %.0f ns per k-block-wave)" % (
        os.environ.get("VQA_B200_GEMM_DBG", "0"), M, N, K, bn, us, tiles, waves, us * 1e3 / (waves * K / 64)), flush=True)
