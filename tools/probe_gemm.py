"""Bring-up probe for the tcgen05 GEMM / implicit-GEMM conv kernels.  Each case runs in its own
subprocess (a device trap poisons the CUDA context), with a timeout, and reports error patterns so a
wrong descriptor / swizzle hypothesis can be told apart from an indexing bug.

    python tools/probe_gemm.py            # run every case, write gpurun_out/probe_gemm.json
    python tools/probe_gemm.py --case X   # one case, in-process
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _setup():
    import torch
    import t5_resnet_vqa_b200 as pkg
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    return torch, pkg.lib


def report(name, got, ref, extra=None):
    import torch
    got = got.float()
    ref = ref.float()
    err = (got - ref).abs()
    denom = ref.abs().max().item() + 1e-12
    rel_fro = (got - ref).norm().item() / (ref.norm().item() + 1e-12)
    bad = err > (2e-2 * denom)
    out = {"case": name, "max_abs_err": err.max().item(), "ref_absmax": denom, "rel_fro": rel_fro,
           "frac_bad": bad.float().mean().item(), "nan": bool(torch.isnan(got).any().item()),
           "ok": bool(rel_fro < 1e-2 and not torch.isnan(got).any().item())}
    if not out["ok"] and got.dim() == 2:
        # error pattern: which rows / cols (mod 8, mod 64, mod 128) are wrong
        rows_bad = bad.any(dim=1).nonzero().flatten()[:16].tolist()
        cols_bad = bad.any(dim=0).nonzero().flatten()[:16].tolist()
        out["first_bad_rows"] = rows_bad
        out["first_bad_cols"] = cols_bad
        out["bad_by_row_mod8"] = [bad[i::8].float().mean().item() for i in range(8)]
        out["bad_by_col_mod8"] = [bad[:, i::8].float().mean().item() for i in range(8)]
        out["bad_by_col_blk64"] = [bad[:, i:i + 64].float().mean().item() for i in range(0, min(bad.shape[1], 512), 64)]
        out["bad_by_row_blk32"] = [bad[i:i + 32].float().mean().item() for i in range(0, min(bad.shape[0], 256), 32)]
        out["sample_got"] = got[:2, :8].tolist()
        out["sample_ref"] = ref[:2, :8].tolist()
    if extra:
        out.update(extra)
    print("PROBE " + json.dumps(out))
    return out


def run_gemm(torch, lib, M, N, K, a_mn, b_mn, bn, split_k=1, out_fp32=0, bias=False, relu=False,
             residual=None, lda_pad=0, name=""):
    L = lib.load()
    dev = "cuda"
    g = torch.Generator(device="cpu").manual_seed(1234)
    A = torch.randn(M, K, generator=g).to(dev).bfloat16()
    B = torch.randn(N, K, generator=g).to(dev).bfloat16()
    ref = A.float() @ B.float().t()
    bias_t = None
    if bias:
        bias_t = torch.randn(N, generator=g).to(dev)
        ref = ref + bias_t
    if relu:
        ref = ref.relu()
    res_t = None
    if residual == "fp32":
        res_t = torch.randn(M, N, generator=g).to(dev)
        ref = ref + res_t
    elif residual == "bf16":
        res_t = torch.randn(M, N, generator=g).to(dev).bfloat16()
        ref = ref + res_t.float()
    def pad8(t):  # row stride must be a multiple of 8 elements (16 B) for TMA
        r, c = t.shape
        c8 = (c + 7) // 8 * 8
        buf = torch.zeros(r, c8, device=dev, dtype=t.dtype)
        buf[:, :c] = t
        return buf[:, :c]
    A_st = pad8(A.t()) if a_mn else pad8(A)
    B_st = pad8(B.t()) if b_mn else pad8(B)
    ldo = N if (N % 8 == 0) else ((N + 7) // 8) * 8
    out = torch.zeros(M, ldo, device=dev, dtype=torch.float32 if out_fp32 else torch.bfloat16)
    a = lib.GemmArgs()
    a.M, a.N, a.K = M, N, K
    a.A, a.lda, a.a_mn = lib.ptr(A_st), A_st.stride(0), a_mn
    a.B, a.ldb, a.b_mn = lib.ptr(B_st), B_st.stride(0), b_mn
    a.out, a.ldo, a.out_fp32 = lib.ptr(out), ldo, out_fp32
    a.bias = lib.ptr(bias_t)
    a.relu = int(relu)
    a.relu_mask, a.ldm = None, 0
    a.drop_p, a.drop_sid, a.rng = 0.0, 0, None
    a.residual, a.ldr, a.res_fp32 = lib.ptr(res_t), N, int(residual == "fp32")
    a.alpha = 1.0
    a.bn, a.split_k = bn, split_k
    lib.check(L.vqa_gemm_bf16(ctypes.byref(a), lib.stream_ptr()), "gemm")
    torch.cuda.synchronize()
    return report(name, out[:, :N], ref)


def nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


def run_conv(torch, lib, Nimg, H, W, Cin, Cout, R, stride, pad, bn, relu=True, residual=False, name=""):
    import torch.nn.functional as F
    L = lib.load()
    dev = "cuda"
    g = torch.Generator(device="cpu").manual_seed(4321)
    x = torch.randn(Nimg, Cin, H, W, generator=g).to(dev).bfloat16()
    w = (torch.randn(Cout, Cin, R, R, generator=g) / (Cin * R * R) ** 0.5).to(dev).bfloat16()
    b = torch.randn(Cout, generator=g).to(dev)
    ref = F.conv2d(x.float(), w.float(), b, stride=stride, padding=pad)
    Ho, Wo = ref.shape[2], ref.shape[3]
    res = None
    if residual:
        res = torch.randn(Nimg, Cout, Ho, Wo, generator=g).to(dev).bfloat16()
        ref = ref + res.float()
    if relu:
        ref = ref.relu()
    x_n = nhwc(x)
    w_k = w.permute(0, 2, 3, 1).contiguous().view(Cout, R * R * Cin)
    res_n = nhwc(res) if residual else None
    out = torch.zeros(Nimg, Ho, Wo, Cout, device=dev, dtype=torch.bfloat16)
    a = lib.ConvArgs()
    a.N, a.H, a.W, a.Cin, a.Cout, a.R, a.S, a.stride, a.pad, a.Ho, a.Wo, a.stem7 = \
        Nimg, H, W, Cin, Cout, R, R, stride, pad, Ho, Wo, 0
    a.x, a.w, a.out, a.out_fp32 = lib.ptr(x_n), lib.ptr(w_k), lib.ptr(out), 0
    a.bias, a.residual, a.relu, a.bn = lib.ptr(b), lib.ptr(res_n), int(relu), bn
    lib.check(L.vqa_conv2d_bf16(ctypes.byref(a), lib.stream_ptr()), "conv")
    torch.cuda.synchronize()
    return report(name, out.view(-1, Cout), nhwc(ref).view(-1, Cout))


def run_stem(torch, lib, Nimg, H, W, bn, name=""):
    import torch.nn.functional as F
    L = lib.load()
    dev = "cuda"
    g = torch.Generator(device="cpu").manual_seed(99)
    x = torch.rand(Nimg, 3, H, W, generator=g).to(dev).bfloat16()
    w = (torch.randn(64, 3, 7, 7, generator=g) / 147 ** 0.5).to(dev).bfloat16()
    b = torch.randn(64, generator=g).to(dev)
    ref = F.conv2d(x.float(), w.float(), b, stride=2, padding=3).relu()
    Ho, Wo = ref.shape[2], ref.shape[3]
    xp = torch.zeros(Nimg, H, W + 8, 8, device=dev, dtype=torch.bfloat16)
    xp[:, :, 3:3 + W, :3] = x.permute(0, 2, 3, 1)
    wk = torch.zeros(64, 7, 8, 8, device=dev, dtype=torch.bfloat16)
    wk[:, :, :7, :3] = w.permute(0, 2, 3, 1)
    wk = wk.view(64, 448).contiguous()
    out = torch.zeros(Nimg, Ho, Wo, 64, device=dev, dtype=torch.bfloat16)
    a = lib.ConvArgs()
    a.N, a.H, a.W, a.Cin, a.Cout, a.R, a.S, a.stride, a.pad, a.Ho, a.Wo, a.stem7 = \
        Nimg, H, W, 8, 64, 7, 7, 2, 3, Ho, Wo, 1
    a.x, a.w, a.out, a.out_fp32 = lib.ptr(xp), lib.ptr(wk), lib.ptr(out), 0
    a.bias, a.residual, a.relu, a.bn = lib.ptr(b), None, 1, bn
    lib.check(L.vqa_conv2d_bf16(ctypes.byref(a), lib.stream_ptr()), "stem")
    torch.cuda.synchronize()
    return report(name, out.view(-1, 64), nhwc(ref).view(-1, 64))


def run_wgrad(torch, lib, Nimg, H, W, Cin, Cout, bn, split_k, name=""):
    import torch.nn.functional as F
    L = lib.load()
    dev = "cuda"
    g = torch.Generator(device="cpu").manual_seed(77)
    x = torch.randn(Nimg, Cin, H, W, generator=g).to(dev).bfloat16()
    dy = torch.randn(Nimg, Cout, H, W, generator=g).to(dev).bfloat16()
    w = torch.zeros(Cout, Cin, 3, 3, device=dev, requires_grad=True)
    y = F.conv2d(x.float(), w, None, stride=1, padding=1)
    y.backward(dy.float())
    ref = w.grad.permute(0, 2, 3, 1).contiguous().view(Cout, 9 * Cin)
    dw = torch.zeros(Cout, 9 * Cin, device=dev)
    a = lib.ConvWgradArgs()
    a.N, a.H, a.W, a.Cin, a.Cout, a.R, a.S, a.pad = Nimg, H, W, Cin, Cout, 3, 3, 1
    a.dy, a.x, a.dw, a.bn, a.split_k = lib.ptr(nhwc(dy)), lib.ptr(nhwc(x)), lib.ptr(dw), bn, split_k
    lib.check(L.vqa_conv2d_wgrad_bf16(ctypes.byref(a), lib.stream_ptr()), "wgrad")
    torch.cuda.synchronize()
    return report(name, dw, ref)


def cases():
    c = {}
    # K-major x K-major (forward)
    c["tn_256_bn128"] = lambda t, l, n: run_gemm(t, l, 256, 256, 256, 0, 0, 128, name=n)
    c["tn_256_bn64"] = lambda t, l, n: run_gemm(t, l, 256, 256, 256, 0, 0, 64, name=n)
    c["tn_256_bn256"] = lambda t, l, n: run_gemm(t, l, 256, 256, 256, 0, 0, 256, name=n)
    c["tn_t5_qkv"] = lambda t, l, n: run_gemm(t, l, 2048, 2304, 768, 0, 0, 256, name=n)
    c["tn_bias_relu"] = lambda t, l, n: run_gemm(t, l, 2048, 768, 768, 0, 0, 128, bias=True, relu=True, name=n)
    c["tn_res_fp32"] = lambda t, l, n: run_gemm(t, l, 2048, 768, 3072, 0, 0, 64, out_fp32=1, residual="fp32", name=n)
    c["tn_res_bf16"] = lambda t, l, n: run_gemm(t, l, 1024, 512, 512, 0, 0, 128, residual="bf16", name=n)
    c["tn_tails"] = lambda t, l, n: run_gemm(t, l, 200, 170, 776, 0, 0, 64, out_fp32=1, bias=True, name=n)
    c["tn_small_m"] = lambda t, l, n: run_gemm(t, l, 64, 170, 768, 0, 0, 64, out_fp32=1, bias=True, name=n)
    c["tn_splitk"] = lambda t, l, n: run_gemm(t, l, 256, 256, 2048, 0, 0, 128, split_k=4, out_fp32=1, name=n)
    # dgrad: A K-major, B MN-major
    c["nn_256"] = lambda t, l, n: run_gemm(t, l, 256, 256, 256, 0, 1, 128, name=n)
    c["nn_dgrad"] = lambda t, l, n: run_gemm(t, l, 2048, 768, 3072, 0, 1, 256, name=n)
    c["nn_k170"] = lambda t, l, n: run_gemm(t, l, 64, 768, 176, 0, 1, 128, out_fp32=1, name=n)
    # wgrad: both MN-major
    c["tt_256"] = lambda t, l, n: run_gemm(t, l, 256, 256, 256, 1, 1, 128, out_fp32=1, name=n)
    c["tt_wgrad"] = lambda t, l, n: run_gemm(t, l, 768, 768, 2048, 1, 1, 128, split_k=4, out_fp32=1, name=n)
    c["tt_wgrad_bn256"] = lambda t, l, n: run_gemm(t, l, 3072, 768, 2048, 1, 1, 256, out_fp32=1, name=n)
    c["tt_tails"] = lambda t, l, n: run_gemm(t, l, 170, 768, 64, 1, 1, 64, out_fp32=1, name=n)
    # A MN-major only
    c["tk_256"] = lambda t, l, n: run_gemm(t, l, 256, 256, 256, 1, 0, 128, out_fp32=1, name=n)
    # convolutions
    c["conv3x3_56"] = lambda t, l, n: run_conv(t, l, 4, 56, 56, 64, 64, 3, 1, 1, 64, name=n)
    c["conv1x1_56"] = lambda t, l, n: run_conv(t, l, 4, 56, 56, 64, 256, 1, 1, 0, 128, relu=False, name=n)
    c["conv1x1_res"] = lambda t, l, n: run_conv(t, l, 4, 28, 28, 128, 512, 1, 1, 0, 256, residual=True, name=n)
    c["conv3x3_s2"] = lambda t, l, n: run_conv(t, l, 4, 56, 56, 128, 128, 3, 2, 1, 128, name=n)
    c["conv1x1_s2"] = lambda t, l, n: run_conv(t, l, 4, 56, 56, 256, 512, 1, 2, 0, 128, relu=False, name=n)
    c["conv3x3_14"] = lambda t, l, n: run_conv(t, l, 8, 14, 14, 256, 256, 3, 1, 1, 128, name=n)
    c["conv3x3_7"] = lambda t, l, n: run_conv(t, l, 8, 7, 7, 512, 512, 3, 1, 1, 128, name=n)
    c["conv3x3_8"] = lambda t, l, n: run_conv(t, l, 4, 8, 8, 512, 768, 3, 1, 1, 128, relu=False, name=n)
    c["stem7"] = lambda t, l, n: run_stem(t, l, 2, 224, 224, 64, name=n)
    c["stem7_64"] = lambda t, l, n: run_stem(t, l, 2, 64, 64, 64, name=n)
    c["wgrad7"] = lambda t, l, n: run_wgrad(t, l, 8, 7, 7, 256, 128, 128, 1, name=n)
    c["wgrad7_split"] = lambda t, l, n: run_wgrad(t, l, 16, 7, 7, 512, 768, 256, 4, name=n)
    c["wgrad8"] = lambda t, l, n: run_wgrad(t, l, 4, 8, 8, 128, 64, 64, 1, name=n)
    c["wgrad14"] = lambda t, l, n: run_wgrad(t, l, 2, 14, 14, 128, 128, 128, 1, name=n)
    return c


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", default=None)
    ap.add_argument("--only", default=None, help="comma separated prefixes")
    ap.add_argument("--dbg", default=None, help="a_lbo,a_sbo,b_lbo,b_sbo override")
    args = ap.parse_args()
    cs = cases()
    if args.case:
        torch, lib = _setup()
        if args.dbg:
            v = [int(x) for x in args.dbg.split(",")]
            lib.load().vqa_debug_set_umma(*v)
        cs[args.case](torch, lib, args.case)
        return
    names = list(cs)
    if args.only:
        pref = args.only.split(",")
        names = [n for n in names if any(n.startswith(p) for p in pref)]
    results = []
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    for n in names:
        t0 = time.time()
        cmd = [sys.executable, os.path.abspath(__file__), "--case", n]
        try:
            p = subprocess.run(cmd, capture_output=True, text=True, timeout=180)
            lines = [ln for ln in p.stdout.splitlines() if ln.startswith("PROBE ")]
            if lines:
                r = json.loads(lines[-1][6:])
            else:
                r = {"case": n, "ok": False, "rc": p.returncode, "stderr": p.stderr[-1500:], "stdout": p.stdout[-500:]}
        except subprocess.TimeoutExpired:
            r = {"case": n, "ok": False, "timeout": True}
        r["secs"] = round(time.time() - t0, 1)
        results.append(r)
        print(("OK   " if r.get("ok") else "FAIL ") + n + " " + json.dumps({k: v for k, v in r.items() if k in ("rel_fro", "max_abs_err", "frac_bad", "rc", "timeout")}), flush=True)
        # MN-major alternative hypothesis if the default fails
        if not r.get("ok") and (n.startswith("nn_256") or n.startswith("tt_256") or n.startswith("tk_256")):
            for dbg in ("1024,8192,1024,8192",):
                try:
                    p = subprocess.run(cmd + ["--dbg", dbg], capture_output=True, text=True, timeout=180)
                    lines = [ln for ln in p.stdout.splitlines() if ln.startswith("PROBE ")]
                    r2 = json.loads(lines[-1][6:]) if lines else {"ok": False, "rc": p.returncode, "stderr": p.stderr[-800:]}
                except subprocess.TimeoutExpired:
                    r2 = {"ok": False, "timeout": True}
                r2["case"] = n + "@dbg=" + dbg
                results.append(r2)
                print(("OK   " if r2.get("ok") else "FAIL ") + r2["case"], flush=True)
        with open(os.path.join(ROOT, "gpurun_out", "probe_gemm.json"), "w") as f:
            json.dump(results, f, indent=1)
    nfail = sum(1 for r in results if not r.get("ok"))
    print("probe_gemm: %d cases, %d failed" % (len(results), nfail))


if __name__ == "__main__":
    main()
