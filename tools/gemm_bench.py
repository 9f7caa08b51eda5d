"""Micro-benchmark of the tcgen05 GEMM / conv kernel on the step's shapes, next to cuBLAS / cuDNN (torch) on the
same shapes.  Diagnostic only (not a bench.py number): back-to-back launches, CUDA events, inputs L2-warm.

    python tools/gemm_bench.py
"""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import t5_resnet_vqa_b200 as pkg  # noqa: E402
from util import Caller  # noqa: E402

BF = torch.bfloat16


REPS = 20


def _time_graph(launch, n_inner):
    for _ in range(3):
        launch()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        launch()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (5 * n_inner) * 1e3  # us per op


def time_ours(record):
    """record(rec) appends ONE op to a plan; the plan holds REPS copies and is replayed as a CUDA graph, so the
    number is device time per launch (back-to-back), free of Python / ctypes overhead."""
    import ctypes
    from t5_resnet_vqa_b200.engine import _Rec
    lib = pkg.lib.load()
    plan = lib.vqa_plan_create()
    rec = _Rec(lib, plan, None)
    for _ in range(REPS):
        record(rec)
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        sp = ctypes.c_void_p(side.cuda_stream)
        pkg.lib.check(lib.vqa_plan_run(plan, sp))
        side.synchronize()
        pkg.lib.check(lib.vqa_plan_capture_graph(plan, sp))
        t = _time_graph(lambda: pkg.lib.check(lib.vqa_plan_run(plan, sp)), REPS)
    lib.vqa_plan_destroy(plan)
    return t


def time_torch(fn):
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn()
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            for _ in range(REPS):
                fn()
    return _time_graph(g.replay, REPS)


def main():
    C = Caller(pkg)
    dev = "cuda"
    print("%-34s %7s %7s %7s %7s %7s %8s | %8s" % ("shape", "bn64", "bn128", "bn256", "pair128", "pair256", "best TF", "cuBLAS"))
    # forward-style (K-major x K-major), dgrad-style (B MN-major), wgrad-style (both MN-major)
    shapes = [("fwd", 2048, 768, 768), ("fwd", 2048, 2304, 768), ("fwd", 2048, 3072, 768), ("fwd", 2048, 768, 3072),
              ("dgrad", 2048, 768, 2304), ("dgrad", 2048, 3072, 768), ("dgrad", 2048, 768, 3072),
              ("wgrad", 768, 768, 2048), ("wgrad", 2304, 768, 2048), ("wgrad", 3072, 768, 2048),
              ("wgrad", 768, 3072, 2048), ("fwd", 8192, 768, 768), ("fwd", 8192, 3072, 768)]
    for kind, M, N, K in shapes:
        A = torch.randn(M, K, device=dev).to(BF)
        B = torch.randn(N, K, device=dev).to(BF)
        At, Bt = A.t().contiguous(), B.t().contiguous()
        fp32 = kind == "wgrad"
        out = torch.empty(M, N, device=dev, dtype=torch.float32 if fp32 else BF)
        res = []
        for bn, pair in ((64, False), (128, False), (256, False), (128, True), (256, True)):
            if kind == "fwd":
                fn = lambda r: r.gemm(M, N, K, A, K, 0, B, K, 0, out, N, int(fp32), bn=bn, pair=pair)
            elif kind == "dgrad":
                fn = lambda r: r.gemm(M, N, K, A, K, 0, Bt, N, 1, out, N, int(fp32), bn=bn, pair=pair)
            else:
                fn = lambda r: r.gemm(M, N, K, At, M, 1, Bt, N, 1, out, N, int(fp32), bn=bn, pair=pair)
            res.append(time_ours(fn))
        # cluster split-K variants that fit (tiles x ks <= 148, >= 4 k-blocks per slice)
        ks_best = None
        for bn in (128, 256):
            tiles = ((M + 127) // 128) * ((N + bn - 1) // bn)
            for ks in (2, 3, 4):
                if tiles * ks > 148 or (K // 64) < 4 * ks:
                    continue
                if kind == "fwd":
                    fn = lambda r: r.gemm(M, N, K, A, K, 0, B, K, 0, out, N, int(fp32), bn=bn, ksplit=ks)
                elif kind == "dgrad":
                    fn = lambda r: r.gemm(M, N, K, A, K, 0, Bt, N, 1, out, N, int(fp32), bn=bn, ksplit=ks)
                else:
                    fn = lambda r: r.gemm(M, N, K, At, M, 1, Bt, N, 1, out, N, int(fp32), bn=bn, ksplit=ks)
                t = time_ours(fn)
                if ks_best is None or t < ks_best[0]:
                    ks_best = (t, bn, ks)
        ks_txt = "ksplit best %5.1f us (bn%d x%d)" % ks_best if ks_best else "ksplit n/a"
        ref = time_torch(lambda: torch.matmul(A, B.t()))
        fl = 2.0 * M * N * K
        best = min(res + ([ks_best[0]] if ks_best else []))
        print("%-6s M%-6d N%-5d K%-5d       %7.1f %7.1f %7.1f %7.1f %7.1f %8.0f | %8.1f us (%4.0f TF) | %s" % (
            kind, M, N, K, res[0], res[1], res[2], res[3], res[4], fl / best / 1e6, ref, fl / ref / 1e6, ks_txt))
    # convolutions of ResNet-50 at batch 64 (NHWC) vs cuDNN channels_last bf16
    convs = [(56, 64, 64, 1, 1), (56, 64, 64, 3, 1), (56, 64, 256, 1, 1), (56, 256, 64, 1, 1), (28, 128, 128, 3, 1),
             (28, 128, 512, 1, 1), (28, 512, 128, 1, 1), (14, 256, 256, 3, 1), (14, 256, 1024, 1, 1),
             (14, 1024, 256, 1, 1), (7, 512, 512, 3, 1), (7, 512, 2048, 1, 1), (7, 2048, 512, 1, 1), (7, 2048, 768, 3, 1)]
    print("\n%-36s %7s %7s %7s %7s %7s | %8s" % ("conv (B=64)", "bn64", "bn128", "bn256", "pair128", "pair256", "cuDNN"))
    if os.environ.get("GEMM_BENCH_SKIP_CONV"):
        return
    for H, Cin, Cout, R, s in convs:
        N = 64
        x = torch.randn(N, H, H, Cin, device=dev).to(BF)
        w = torch.randn(Cout, R, R, Cin, device=dev).to(BF)
        b = torch.randn(Cout, device=dev)
        out = torch.empty(N, H, H, Cout, device=dev, dtype=BF)
        resd = torch.randn(N, H, H, Cout, device=dev).to(BF)
        res = []
        for bn, pair in ((64, False), (128, False), (256, False), (128, True), (256, True)):
            if bn > 64 and Cout < bn:
                res.append(float("nan"))
                continue
            res.append(time_ours(lambda r: r.conv(N, H, H, Cin, Cout, R, s, R // 2, x, w, out, bias=b, residual=resd,
                                                  relu=1, bn=bn, pair=pair)))
        xc = x.permute(0, 3, 1, 2)
        wc = w.permute(0, 3, 1, 2).contiguous(memory_format=torch.channels_last)
        ref = time_torch(lambda: F.conv2d(xc, wc, None, s, R // 2))
        fl = 2.0 * N * H * H * Cout * R * R * Cin
        best = min(r for r in res if r == r)
        print("%2dx%-2d %4d->%-4d k%d +res          %7.1f %7.1f %7.1f %7.1f %7.1f | %8.1f us  ours %4.0f TF" % (
            H, H, Cin, Cout, R, res[0], res[1], res[2], res[3], res[4], ref, fl / best / 1e6))


if __name__ == "__main__":
    main()
