"""Device time per launch of the non-GEMM kernels of the step (norms, attention, elementwise) at the step's shapes:
REPS copies of one op in a plan, replayed as a CUDA graph (back-to-back, PDL on).  Diagnostic only.

    python tools/op_bench.py
"""
import ctypes
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import t5_resnet_vqa_b200 as pkg  # noqa: E402
from gemm_bench import time_ours  # noqa: E402

BF, F32 = torch.bfloat16, torch.float32


def main():
    dev = "cuda"
    M, D, B, L = 2048, 768, 64, 32
    x = torch.randn(M, D, device=dev)
    w = torch.randn(D, device=dev)
    yb = torch.empty(M, D, device=dev, dtype=BF)
    yf = torch.empty(M, D, device=dev)
    rstd = torch.rand(M, device=dev) + 0.5
    mean = torch.zeros(M, device=dev)
    dyb = torch.randn(M, D, device=dev).to(BF)
    dres = torch.randn(M, D, device=dev)
    dx = torch.empty(M, D, device=dev)
    dw = torch.zeros(D, device=dev)
    db = torch.zeros(D, device=dev)
    gcs = torch.zeros(D, device=dev)
    gb = torch.empty(M, D, device=dev, dtype=BF)
    rng = torch.tensor([1234, 7], dtype=torch.int64, device=dev)
    rows = []

    def add(name, fn, nbytes):
        t = time_ours(fn)
        rows.append((name, t, nbytes / t / 1e3 if nbytes else 0.0))

    add("rmsnorm_fwd", lambda r: r.rmsnorm_fwd(x, w, yb, None, rstd, M, D, 1e-6, 0.0, 0, None), M * D * 6)
    add("rmsnorm_bwd (+g_out, dropout)", lambda r: r.rmsnorm_bwd(dyb, 0, x, w, rstd, dres, dx, dw, M, D, 0.0, 0, rng,
                                                                  gb, 0.1, 5), M * D * 17)
    add("rmsnorm_bwd (no dw)", lambda r: r.rmsnorm_bwd(dyb, 0, x, w, rstd, dres, dx, None, M, D, 0.0, 0, rng,
                                                        gb, 0.1, 5), M * D * 17)
    add("rmsnorm_bwd (plain)", lambda r: r.rmsnorm_bwd(dyb, 0, x, w, rstd, dres, dx, dw, M, D, 0.0, 0, None,
                                                        None, 0.0, 0), M * D * 14)
    add("layernorm_fwd", lambda r: r.layernorm_fwd(x, w, w, yb, yf, mean, rstd, M, D, 1e-5), M * D * 10)
    add("layernorm_bwd (+g_out, colsum)", lambda r: r.layernorm_bwd(dres, x, w, mean, rstd, dx, dw, db, M, D, gb, 0.1, 9,
                                                                     rng, gcs), M * D * 14)
    add("colsum_bf16 [2048,768]", lambda r: r.colsum_bf16(dyb, D, dw, M, D), M * D * 2)
    big = torch.randn(M, 3 * D, device=dev).to(BF)
    dw3 = torch.zeros(3 * D, device=dev)
    add("colsum_bf16 [2048,2304]", lambda r: r.colsum_bf16(big, 3 * D, dw3, M, 3 * D), M * D * 6)
    # attention
    for (H, hd, t5) in ((12, 64, True), (8, 96, False)):
        Dm = H * hd
        qkv = torch.randn(B * L, 3 * Dm, device=dev).to(BF) * 0.3
        out = torch.empty(B * L, Dm, device=dev, dtype=BF)
        dO = torch.randn(B * L, Dm, device=dev).to(BF)
        dqkv = torch.empty_like(qkv)
        stats = torch.zeros(B * H * L, 2, device=dev)
        probs = torch.zeros(B * H * L * L, device=dev)
        bias = torch.randn(H, L, L, device=dev) if t5 else None
        mask = torch.ones(B, L, dtype=torch.int64, device=dev) if t5 else None
        dbias = torch.zeros(H, L, L, device=dev) if t5 else None
        sc = 1.0 if t5 else 1 / math.sqrt(hd)
        q, k, v = qkv, qkv.data_ptr() + 2 * Dm, qkv.data_ptr() + 4 * Dm
        dq, dk, dv = dqkv, dqkv.data_ptr() + 2 * Dm, dqkv.data_ptr() + 4 * Dm
        for tc in (True, False):
            st = stats if tc else None
            tag = "%s hd%d %s" % ("T5" if t5 else "SGA", hd, "tcgen05" if tc else "simt")
            add("attn_fwd " + tag, lambda r: r.attn_fwd(B, H, L, L, hd, q, 3 * Dm, k, 3 * Dm, v, 3 * Dm, out, Dm, probs,
                                                        bias, mask, sc, 0.1, 3, rng, stats=st), 0)
            add("attn_bwd " + tag, lambda r: r.attn_bwd(B, H, L, L, hd, q, 3 * Dm, k, 3 * Dm, v, 3 * Dm, probs, dO, Dm, dq,
                                                        3 * Dm, dk, 3 * Dm, dv, 3 * Dm, dbias, sc, 0.1, 3, rng, stats=st,
                                                        bias=bias, key_mask=mask), 0)
            if t5:
                add("attn_bwd " + tag + " (no dbias)",
                    lambda r: r.attn_bwd(B, H, L, L, hd, q, 3 * Dm, k, 3 * Dm, v, 3 * Dm, probs, dO, Dm, dq, 3 * Dm, dk,
                                         3 * Dm, dv, 3 * Dm, None, sc, 0.1, 3, rng, stats=st, bias=bias, key_mask=mask), 0)
    print("%-40s %8s %10s" % ("op", "us", "GB/s"))
    for name, t, bw in rows:
        print("%-40s %8.1f %10.0f" % (name, t, bw))


if __name__ == "__main__":
    main()
