"""GEMM time with L2-cold weights (as inside the training step, where the AdamW pass has streamed 5 GB through the L2
since the weights were last read) vs L2-warm weights (what tools/gemm_bench.py measures).  Diagnostic only."""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools"))
import t5_resnet_vqa_b200 as pkg
from t5_resnet_vqa_b200.engine import _Rec
BF = torch.bfloat16; dev = "cuda"
lib = pkg.lib.load()


def run(M, N, K, ncopies, reps=48):
    A = torch.randn(M, K, device=dev).to(BF)
    Bs = [torch.randn(N, K, device=dev).to(BF) for _ in range(ncopies)]
    out = torch.empty(M, N, device=dev, dtype=BF)
    plan = lib.vqa_plan_create()
    rec = _Rec(lib, plan, None)
    for i in range(reps):
        rec.gemm(M, N, K, A, K, 0, Bs[i % ncopies], K, 0, out, N, 0)
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        sp = ctypes.c_void_p(side.cuda_stream)
        pkg.lib.check(lib.vqa_plan_run(plan, sp)); side.synchronize()
        pkg.lib.check(lib.vqa_plan_capture_graph(plan, sp))
        for _ in range(2):
            pkg.lib.check(lib.vqa_plan_run(plan, sp))
        side.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(side)
        for _ in range(3):
            pkg.lib.check(lib.vqa_plan_run(plan, sp))
        e1.record(side); side.synchronize()
    lib.vqa_plan_destroy(plan)
    return e0.elapsed_time(e1) / (3 * reps) * 1e3


for M, N, K in [(2048, 2304, 768), (2048, 768, 768), (2048, 3072, 768), (2048, 768, 3072)]:
    wbytes = N * K * 2
    ncold = max(2, int(400e6 // wbytes))     # > 3x the L2: every launch reads weights that left the cache long ago
    ncold = min(ncold, 256)
    warm, cold = run(M, N, K, 1, reps=max(48, ncold)), run(M, N, K, ncold, reps=max(48, ncold))
    print("M%d N%d K%d: warm weights %.1f us, cold weights (%d copies, %.0f MB) %.1f us" % (M, N, K, warm, ncold, ncold * wbytes / 1e6, cold))
