"""Phase timing (clock64 of CTA 0) of the tcgen05 attention kernels.  Diagnostic only."""
import math, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# needs the diagnostic build (python t5-resnet-vqa_b200/build.py --debug): the default library has no instrumentation
_DBG = os.path.join(ROOT, "t5-resnet-vqa_b200", "libvqa_b200_dbg.so")
if not os.path.exists(_DBG):
    raise SystemExit("build the diagnostic library first: python t5-resnet-vqa_b200/build.py --debug")
os.environ.setdefault("VQA_B200_LIB", _DBG)
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import t5_resnet_vqa_b200 as pkg
from util import Caller
C = Caller(pkg); lib = pkg.lib.load()
dev = "cuda"; B, L = 64, 32
rng = torch.tensor([1234, 7], dtype=torch.int64, device=dev)
names = ["start", "setup done", "pdl_wait done", "tma landed (t0)", "pre mma wait", "S ready", "P staged", "synced", "O ready", "stored"]
for (H, hd, t5) in ((12, 64, True), (8, 96, False)):
    Dm = H * hd
    qkv = torch.randn(B * L, 3 * Dm, device=dev).to(torch.bfloat16) * 0.3
    out = torch.empty(B * L, Dm, device=dev, dtype=torch.bfloat16); dO = torch.randn_like(out); dqkv = torch.empty_like(qkv)
    stats = torch.zeros(B * H * L, 2, device=dev)
    bias = torch.randn(H, L, L, device=dev) if t5 else None
    mask = torch.ones(B, L, dtype=torch.int64, device=dev) if t5 else None
    dbias = torch.zeros(H, L, L, device=dev) if t5 else None
    sc = 1.0 if t5 else 1 / math.sqrt(hd)
    q, k, v = qkv, qkv.data_ptr() + 2 * Dm, qkv.data_ptr() + 4 * Dm
    dq, dk, dv = dqkv, dqkv.data_ptr() + 2 * Dm, dqkv.data_ptr() + 4 * Dm
    for which in ("fwd", "bwd"):
        buf = torch.zeros(64, dtype=torch.int64, device=dev)
        lib.vqa_debug_attn_timing(buf.data_ptr())
        for _ in range(3):
            if which == "fwd":
                C.attn_fwd(B, H, L, L, hd, q, 3 * Dm, k, 3 * Dm, v, 3 * Dm, out, Dm, None, bias, mask, sc, 0.1, 3, rng, stats=stats)
            else:
                C.attn_bwd(B, H, L, L, hd, q, 3 * Dm, k, 3 * Dm, v, 3 * Dm, None, dO, Dm, dq, 3 * Dm, dk, 3 * Dm, dv, 3 * Dm, dbias, sc, 0.1, 3, rng, stats=stats, bias=bias, key_mask=mask)
            torch.cuda.synchronize()
        lib.vqa_debug_attn_timing(None)
        t = buf.cpu().view(4, 16)
        print("hd%d %s  (cycles since start, warp0 | warp3)" % (hd, which))
        for i, n in enumerate(names):
            print("   %-18s %8d %8d" % (n, int(t[0, i] - t[0, 0]) if t[0, i] else -1, int(t[3, i] - t[0, 0]) if t[3, i] else -1))
