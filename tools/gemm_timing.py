"""Phase timing (clock64 of CTA 0) of the tcgen05 GEMM kernel on the step's shapes.  Diagnostic only."""
import os, sys
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
BF = torch.bfloat16; dev = "cuda"
names = ["start", "setup done", "pdl_wait done", "first stage full (mma)", "tile0 mma committed", "tile0 acc ready (epi)", "epi loop done", "stores drained", "teardown sync", "ksplit: partials dumped", "ksplit: fence + barrier", "ksplit: peers arrived"]
shapes = [("fwd", 2048, 768, 3072, 128, 0, 0), ("fwd", 2048, 768, 3072, 256, 0, 0), ("fwd", 2048, 768, 768, 128, 0, 0), ("fwd", 2048, 768, 768, 128, 1, 0), ("fwd", 2048, 2304, 768, 256, 0, 0),
          ("fwd", 2048, 768, 3072, 128, 1, 0), ("wgrad", 768, 768, 2048, 64, 1, 0), ("wgrad", 3072, 768, 2048, 128, 1, 0),
          ("fwd", 2048, 2304, 768, 256, 0, 1), ("fwd", 8192, 3072, 768, 256, 0, 1), ("fwd", 8192, 3072, 768, 256, 0, 0)]
KS = 1
if len(sys.argv) > 1 and sys.argv[1] == "ksplit":
    shapes = [("fwd", 2048, 768, 3072, 256, 0, 0)]
    KS = int(sys.argv[2])
elif len(sys.argv) > 1 and sys.argv[1] == "conv":
    shapes = [("conv", 14, 256, 1024, 256, 1, 1), ("conv", 14, 256, 1024, 256, 0, 1), ("conv", 56, 64, 256, 256, 1, 1), ("conv", 56, 64, 256, 64, 1, 1),
              ("conv", 14, 256, 256, 256, 0, 3), ("conv", 28, 128, 512, 256, 1, 1)]
elif len(sys.argv) > 1:
    shapes = shapes[-3:]
for kind, M, N, K, bn, fp32, pair in shapes:
    if kind == "conv":
        H, Cin, Cout, bn, has_res, R = M, N, K, bn, fp32, pair
        NB = 64
        x = torch.randn(NB, H, H, Cin, device=dev).to(BF); w = torch.randn(Cout, R, R, Cin, device=dev).to(BF)
        b = torch.randn(Cout, device=dev); out = torch.empty(NB, H, H, Cout, device=dev, dtype=BF)
        rs = torch.randn(NB, H, H, Cout, device=dev).to(BF) if has_res else None
        buf = torch.zeros(22 * 16, dtype=torch.int64, device=dev)
        for it in range(3):
            if it == 2:
                lib.vqa_debug_gemm_timing(buf.data_ptr())
            C.conv(NB, H, H, Cin, Cout, R, 1, R // 2, x, w, out, bias=b, residual=rs, relu=1, bn=bn)
            torch.cuda.synchronize()
        lib.vqa_debug_gemm_timing(None)
        t = buf.cpu().view(22, 16)
        t0 = int(t[0, 0])
        tiles = ((NB * H * H + 127) // 128) * ((Cout + bn - 1) // bn)
        print("conv %dx%d %d->%d k%d res=%d bn%d (%d tiles, %.2f per SM)  cycles since start: producer(w0) mma(w1) epi(w2) epi(w9)" % (H, H, Cin, Cout, R, has_res, bn, tiles, tiles / 148))
        for i, n in enumerate(names):
            print("   %-26s" % n + "".join("%9s" % (str(int(t[w, i]) - t0) if int(t[w, i]) else "-") for w in (0, 1, 2, 9)))
        print("   cta0: producer cycles waiting for empty slots %d of %d total | MMA warp cycles waiting for data %d, issuing %d" % (int(t[0, 12]), int(t[0, 13]), int(t[1, 12]), int(t[1, 13])))
        continue
    A = torch.randn(M, K, device=dev).to(BF); B = torch.randn(N, K, device=dev).to(BF)
    At, Bt = A.t().contiguous(), B.t().contiguous()
    out = torch.empty(M, N, device=dev, dtype=torch.float32 if fp32 else BF)
    res = torch.randn(M, N, device=dev) if (fp32 and kind == "fwd") else None
    buf = torch.zeros(22 * 16, dtype=torch.int64, device=dev)
    for it in range(3):
        if it == 2:
            lib.vqa_debug_gemm_timing(buf.data_ptr())
        if kind == "fwd":
            C.gemm(M, N, K, A, K, 0, B, K, 0, out, N, fp32, bn=bn, residual=res, ldr=N, res_fp32=1, pair=bool(pair), ksplit=KS)
        else:
            C.gemm(M, N, K, At, M, 1, Bt, N, 1, out, N, fp32, bn=bn, pair=bool(pair))
        torch.cuda.synchronize()
    lib.vqa_debug_gemm_timing(None)
    t = buf.cpu().view(22, 16)
    t0 = int(t[0, 0])
    print("%s M%d N%d K%d bn%d fp32=%d pair=%d   cycles since start: producer(w0) mma(w1) epi(w2) epi(w9)" % (kind, M, N, K, bn, fp32, pair))
    for i, n in enumerate(names):
        print("   %-26s" % n + "".join("%9s" % (str(int(t[w, i]) - t0) if int(t[w, i]) else "-") for w in (0, 1, 2, 9)))
    print("   cta0: producer cycles waiting for empty slots %d of %d total | MMA warp cycles waiting for data %d, issuing %d" % (int(t[0, 12]), int(t[0, 13]), int(t[1, 12]), int(t[1, 13])))
    print("   producer cta0 empty-pass i=0..6:", [int(t[0, 9 + i]) - t0 for i in range(7)])
    print("   producer cta1 empty-pass i=0..6:", [int(t[11, 9 + i]) - int(t[11, 0]) for i in range(7)], "cta1 start-cta0 start", int(t[11, 0]) - t0)
    print("   mma cta0 full-pass i=1..6      :", [int(t[1, 9 + i]) - t0 for i in range(6)])
