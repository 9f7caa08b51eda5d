"""Kernel-level parity (GPU): every C-ABI entry point against the matching torch fp32 op on the same inputs.
Tolerances: bf16-operand contractions rel-Frobenius <= 6e-3 (bf16 output rounding 2^-9 dominates);
fp32-output contractions <= 2e-5 relative to an fp32 reference fed the same bf16-rounded operands;
bandwidth kernels <= 1e-5 (fp32) / 4e-3 (bf16 outputs)."""
import ctypes
import math

import pytest
import torch
import torch.nn.functional as F

from util import Caller, cosine, rel_fro

pytestmark = pytest.mark.gpu
BF, F32 = torch.bfloat16, torch.float32


@pytest.fixture(scope="module")
def C(pkg, cuda):
    return Caller(pkg)


def rnd(*shape, seed=0, scale=1.0, dtype=F32):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).cuda().to(dtype)


def rng_state(seed=1234, offset=7):
    return torch.tensor([seed, offset], dtype=torch.int64, device="cuda")


# ------------------------------------------------------------------------------------------------
# GEMM
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,N,K,bn", [(2048, 768, 768, 64), (2048, 2304, 768, 256), (2048, 3072, 768, 256),
                                      (2048, 768, 3072, 128), (64, 170, 768, 64), (200, 170, 776, 64),
                                      (128, 768, 768, 64), (196, 1536, 768, 128)])
def test_linear_forward_bias_relu(C, M, N, K, bn):
    X, W, b = rnd(M, K, seed=1, dtype=BF), rnd(N, K, seed=2, scale=K ** -0.5, dtype=BF), rnd(N, seed=3)
    ldo = (N + 7) // 8 * 8
    out = torch.zeros(M, ldo, dtype=BF, device="cuda")
    C.linear(X, M, K, K, W, N, out, ldo, bias=b, relu=1, bn=bn)
    ref = F.relu(X.float() @ W.float().t() + b)
    assert rel_fro(out[:, :N], ref) < 6e-3
    if ldo > N:
        assert float(out[:, N:].abs().max()) == 0.0


def test_linear_fp32_residual_and_accumulate(C):
    M, N, K = 2048, 768, 3072
    X, W, R = rnd(M, K, seed=1, dtype=BF), rnd(N, K, seed=2, scale=K ** -0.5, dtype=BF), rnd(M, N, seed=3)
    out = torch.zeros(M, N, device="cuda")
    C.linear(X, M, K, K, W, N, out, N, out_fp32=1, residual=R, ldr=N, res_fp32=1, bn=64)
    ref = X.float() @ W.float().t() + R
    assert rel_fro(out, ref) < 2e-5
    C.linear(X, M, K, K, W, N, out, N, out_fp32=1, residual=R, ldr=N, res_fp32=1, accumulate=1, bn=128)
    assert rel_fro(out, 2 * ref) < 2e-5


def test_split_k(C):
    M, N, K = 256, 256, 2048
    X, W = rnd(M, K, seed=1, dtype=BF), rnd(N, K, seed=2, dtype=BF)
    out = torch.zeros(M, N, device="cuda")
    C.linear(X, M, K, K, W, N, out, N, out_fp32=1, bn=128, split_k=4)
    assert rel_fro(out, X.float() @ W.float().t()) < 2e-5


@pytest.mark.parametrize("M,N,K,bn", [(2048, 768, 3072, 256), (2048, 3072, 768, 128), (64, 170, 768, 64),
                                      (2048, 2304, 768, 128)])
def test_dgrad_with_relu_mask(C, M, N, K, bn):
    """dX[M,K] = (dY[M,N] @ W[N,K]) masked by saved activation > 0."""
    ld = (N + 7) // 8 * 8
    dY = torch.zeros(M, ld, dtype=BF, device="cuda")
    dY[:, :N] = rnd(M, N, seed=1, dtype=BF)
    W, act = rnd(N, K, seed=2, scale=N ** -0.5, dtype=BF), rnd(M, K, seed=3, dtype=BF)
    out = torch.zeros(M, K, dtype=BF, device="cuda")
    C.dgrad(dY, M, N, ld, W, K, out, K, relu_mask=act, ldm=K, bn=bn)
    ref = (dY[:, :N].float() @ W.float()) * (act.float() > 0)
    assert rel_fro(out, ref) < 6e-3


@pytest.mark.parametrize("M,N,K,bn", [(2048, 768, 768, 64), (2048, 3072, 768, 128), (2048, 768, 3072, 256),
                                      (64, 170, 768, 64), (196, 1536, 768, 128), (2048, 2304, 768, 128)])
def test_wgrad(C, M, N, K, bn):
    """dW[N,K] = dY[M,N]^T @ X[M,K] (both operands MN-major)."""
    ld = (N + 7) // 8 * 8
    dY = torch.zeros(M, ld, dtype=BF, device="cuda")
    dY[:, :N] = rnd(M, N, seed=1, dtype=BF)
    X = rnd(M, K, seed=2, dtype=BF)
    dW = torch.zeros(N, K, device="cuda")
    C.wgrad(dY, M, N, ld, X, K, K, dW, bn=bn)
    assert rel_fro(dW, dY[:, :N].float().t() @ X.float()) < 2e-5


@pytest.mark.parametrize("cap", [1, 40, 148, 500])
def test_capped_persistent_grid(C, cap):
    """vqa_gemm_args.max_ctas: the persistent grid walks the same tiles with fewer CTAs (side-lane weight gradients)."""
    M, N, K = 2048, 768, 768
    dY, X = rnd(M, N, seed=1, dtype=BF), rnd(M, K, seed=2, dtype=BF)
    dW = torch.zeros(N, K, device="cuda")
    C.wgrad(dY, M, N, N, X, K, K, dW, bn=128, max_ctas=cap)
    assert rel_fro(dW, dY.float().t() @ X.float()) < 2e-5


def test_memcpy_d2d_plan_step(C):
    src = rnd(1000, 768, seed=1, dtype=BF)
    dst = torch.zeros_like(src)
    C.memcpy_d2d(dst, src, src.numel() * 2)
    assert torch.equal(dst, src)


# ------------------------------------------------------------------------------------------------
# two-term operand split (hi + lo bf16): the early T5 blocks' forward GEMMs (vqa_gemm_args.B_lo / a_lo_col)
# ------------------------------------------------------------------------------------------------
def _split(x):
    hi = x.to(BF)
    lo = (x - hi.float()).to(BF)
    return hi, lo


@pytest.mark.parametrize("M,N,K,bn,a_split,pair", [(2048, 2304, 768, 256, True, None), (2048, 3072, 768, 128, True, None),
                                                   (2048, 768, 768, 64, False, None), (2048, 768, 3072, 128, False, None),
                                                   (200, 192, 128, 64, True, None), (2048, 3072, 768, 256, True, 1)])
def test_linear_two_term_split(C, M, N, K, bn, a_split, pair):
    """x_hi W_hi + [x_lo W_hi +] x_hi W_lo in one k-loop: against the fp32 product of the fp32 operands the error must be far
    below a plain bf16 GEMM's (2^-9 per operand), and it must equal the fp32 sum of the same bf16 terms."""
    Xf, Wf = rnd(M, K, seed=1), rnd(N, K, seed=2, scale=K ** -0.5)
    Xh, Xl = _split(Xf)
    Wh, Wl = _split(Wf)
    X2 = torch.cat([Xh, Xl], dim=1).contiguous() if a_split else Xh.contiguous()
    out = torch.zeros(M, N, device="cuda")
    C.linear(X2, M, K, X2.shape[1], Wh, N, out, N, out_fp32=1, bn=bn, b_lo=Wl, a_lo_col=K if a_split else 0, pair=pair)
    terms = Xh.float() @ Wh.float().t() + Xh.float() @ Wl.float().t()
    if a_split:
        terms = terms + Xl.float() @ Wh.float().t()
    assert rel_fro(out, terms) < 2e-5
    exact = (Xf.double() @ Wf.double().t()).float()
    plain = torch.zeros(M, N, device="cuda")
    C.linear(Xh.contiguous(), M, K, K, Wh, N, plain, N, out_fp32=1, bn=bn)
    err_split, err_plain = rel_fro(out, exact), rel_fro(plain, exact)
    assert err_plain > 1e-3                                   # the plain bf16 product really is at bf16 precision
    assert err_split < (3e-5 if a_split else 0.75 * err_plain)   # both operands split: ~2^-16; weights only: x's rounding remains


def test_linear_two_term_split_epilogues(C):
    """The split k-loop under the epilogues the T5 blocks use: bf16 out + ReLU + dropout (wi), fp32 out + fp32 residual (o/wo)."""
    M, N, K = 2048, 3072, 768
    Xf, Wf = rnd(M, K, seed=1), rnd(N, K, seed=2, scale=K ** -0.5)
    Xh, Xl = _split(Xf)
    Wh, Wl = _split(Wf)
    X2 = torch.cat([Xh, Xl], dim=1).contiguous()
    out = torch.zeros(M, N, dtype=BF, device="cuda")
    C.linear(X2, M, K, 2 * K, Wh, N, out, N, relu=1, b_lo=Wl, a_lo_col=K, bn=128)
    ref = F.relu((Xh.float() + Xl.float()) @ (Wh.float() + Wl.float()).t())
    assert rel_fro(out, ref) < 4e-3
    R = rnd(M, 768, seed=3)
    W2f = rnd(768, N, seed=4, scale=N ** -0.5)
    W2h, W2l = _split(W2f)
    o2 = torch.zeros(M, 768, device="cuda")
    C.linear(out, M, N, N, W2h, 768, o2, 768, out_fp32=1, residual=R, ldr=768, res_fp32=1, b_lo=W2l, bn=64)
    ref2 = out.float() @ (W2h.float() + W2l.float()).t() + R
    assert rel_fro(o2, ref2) < 2e-5


def test_rmsnorm_split_and_weight_lo(C):
    M, D = 2048, 768
    x, w = rnd(M, D, seed=1, scale=3.0), 0.75 + 0.5 * torch.rand(D, device="cuda")
    y2, rstd = torch.zeros(M, 2 * D, dtype=BF, device="cuda"), torch.zeros(M, device="cuda")
    C.rmsnorm_fwd_split(x, w, y2, rstd, M, D, 1e-6)
    ref = w * (x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + 1e-6))
    hi, lo = y2[:, :D].float(), y2[:, D:].float()
    assert rel_fro(hi, ref) < 4e-3
    assert rel_fro(hi + lo, ref) < 3e-5
    y1 = torch.zeros(M, D, dtype=BF, device="cuda")
    C.rmsnorm_fwd(x, w, y1, None, rstd, M, D, 1e-6, 0.0, 0, None)
    assert torch.equal(y1, y2[:, :D])
    W = rnd(3000, 769, seed=5)
    Wl = torch.zeros_like(W, dtype=BF)
    C.split_lo_bf16(W, Wl, W.numel())
    assert torch.equal(Wl, (W - W.to(BF).float()).to(BF))


# ------------------------------------------------------------------------------------------------
# cluster split-K: a thread-block cluster per output tile, k-slices per CTA, partial sums exchanged through a
# workspace, every CTA finishing (bias / ReLU / dropout / residual / store) a column range of the tile
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,N,K,bn,ks", [(2048, 768, 3072, 256, 3), (2048, 768, 3072, 256, 2), (2048, 768, 2304, 256, 3),
                                         (2048, 768, 3072, 128, 1), (1000, 700, 1544, 128, 2), (256, 170, 2048, 64, 4),
                                         (2048, 768, 768, 256, 3)])
def test_ksplit_linear_forward_bias_relu(C, M, N, K, bn, ks):
    X, W, b = rnd(M, K, seed=1, dtype=BF), rnd(N, K, seed=2, scale=K ** -0.5, dtype=BF), rnd(N, seed=3)
    ldo = (N + 7) // 8 * 8
    out = torch.zeros(M, ldo, dtype=BF, device="cuda")
    C.linear(X, M, K, K, W, N, out, ldo, bias=b, relu=1, bn=bn, ksplit=ks)
    ref = F.relu(X.float() @ W.float().t() + b)
    assert rel_fro(out[:, :N], ref) < 6e-3
    if ldo > N:
        assert float(out[:, N:].abs().max()) == 0.0
    # the exchange must not depend on what an earlier launch left in the workspace
    C.linear(X, M, K, K, W, N, out, ldo, bias=b, relu=1, bn=bn, ksplit=ks)
    assert rel_fro(out[:, :N], ref) < 6e-3


def test_ksplit_fp32_residual_dropout(C):
    M, N, K = 2048, 768, 3072
    X, W, R = rnd(M, K, seed=1, dtype=BF), rnd(N, K, seed=2, scale=K ** -0.5, dtype=BF), rnd(M, N, seed=3)
    out = torch.zeros(M, N, device="cuda")
    C.linear(X, M, K, K, W, N, out, N, out_fp32=1, residual=R, ldr=N, res_fp32=1, bn=256, ksplit=3)
    ref = X.float() @ W.float().t() + R
    assert rel_fro(out, ref) < 2e-5
    # dropout: the mask depends on the element index only, so it must equal the unsplit launch's exactly
    a, b = torch.zeros(M, N, device="cuda"), torch.zeros(M, N, device="cuda")
    rs = rng_state()
    C.linear(X, M, K, K, W, N, a, N, out_fp32=1, residual=R, ldr=N, res_fp32=1, drop_p=0.1, sid=5, rng=rs, bn=256, ksplit=3)
    C.linear(X, M, K, K, W, N, b, N, out_fp32=1, residual=R, ldr=N, res_fp32=1, drop_p=0.1, sid=5, rng=rs, bn=128)
    assert torch.equal(a == R, b == R)
    assert rel_fro(a, b) < 2e-5


@pytest.mark.parametrize("M,N,K,bn,ks", [(2048, 3072, 768, 256, 3), (2048, 2304, 768, 256, 3)])
def test_ksplit_dgrad_with_relu_mask(C, M, N, K, bn, ks):
    dY = rnd(M, N, seed=1, dtype=BF)
    W, act = rnd(N, K, seed=2, scale=N ** -0.5, dtype=BF), rnd(M, K, seed=3, dtype=BF)
    out = torch.zeros(M, K, dtype=BF, device="cuda")
    C.dgrad(dY, M, N, N, W, K, out, K, relu_mask=act, ldm=K, bn=bn, ksplit=ks)
    ref = (dY.float() @ W.float()) * (act.float() > 0)
    assert rel_fro(out, ref) < 6e-3


@pytest.mark.parametrize("M,N,K,bn,ks", [(2048, 768, 768, 128, 4), (2048, 768, 3072, 256, 2), (2048, 2304, 768, 256, 2)])
def test_ksplit_wgrad(C, M, N, K, bn, ks):
    dY, X = rnd(M, N, seed=1, dtype=BF), rnd(M, K, seed=2, dtype=BF)
    dW = torch.full((N, K), 7.0, device="cuda")     # a plain store: stale contents must not matter
    C.wgrad(dY, M, N, N, X, K, K, dW, bn=bn, ksplit=ks)
    assert rel_fro(dW, dY.float().t() @ X.float()) < 2e-5


@pytest.mark.parametrize("N,H,Cin,Cout,R,bn,ks,res", [(64, 7, 512, 512, 3, 256, 2, True), (16, 7, 512, 512, 3, 128, 4, False),
                                                      (8, 14, 256, 256, 3, 256, 3, True)])
def test_ksplit_conv_bias_residual_relu(C, N, H, Cin, Cout, R, bn, ks, res):
    x = rnd(N, Cin, H, H, seed=1, dtype=BF)
    w = rnd(Cout, Cin, R, R, seed=2, scale=(Cin * R * R) ** -0.5, dtype=BF)
    b = rnd(Cout, seed=3)
    ref = F.conv2d(x.float(), w.float(), b, stride=1, padding=R // 2)
    r = None
    if res:
        r = rnd(*ref.shape, seed=4, dtype=BF)
        ref = ref + r.float()
    ref = F.relu(ref)
    out = torch.zeros(N, H, H, Cout, dtype=BF, device="cuda")
    wk = w.permute(0, 2, 3, 1).contiguous()
    C.conv(N, H, H, Cin, Cout, R, 1, R // 2, nhwc(x), wk, out, bias=b, residual=nhwc(r) if res else None, relu=1, bn=bn,
           ksplit=ks)
    assert rel_fro(out, nhwc(ref)) < 6e-3


def test_ksplit_rejects_bad_configurations(C):
    X, W = rnd(2048, 768, seed=1, dtype=BF), rnd(3072, 768, seed=2, dtype=BF)
    out = torch.zeros(2048, 3072, dtype=BF, device="cuda")
    with pytest.raises(RuntimeError, match="148"):
        C.linear(X, 2048, 768, 768, W, 3072, out, 3072, bn=256, ksplit=2)   # 192 tiles x 2 CTAs do not fit the SMs
    with pytest.raises(RuntimeError, match="k-blocks"):
        C.linear(X[:128], 128, 768, 768, W[:256], 256, out, 3072, bn=256, ksplit=8)


# ------------------------------------------------------------------------------------------------
# CTA pairs (cta_group::2): a cluster of two CTAs runs one 256-row MMA, each staging half of the B tile
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,N,K,bn", [(2048, 2304, 768, 256), (2048, 768, 768, 128), (2048, 3072, 768, 256),
                                      (300, 256, 512, 128), (8192, 3072, 768, 256), (196, 1536, 768, 256)])
def test_pair_linear_forward(C, M, N, K, bn):
    X, W, b = rnd(M, K, seed=1, dtype=BF), rnd(N, K, seed=2, scale=K ** -0.5, dtype=BF), rnd(N, seed=3)
    out = torch.zeros(M, N, dtype=BF, device="cuda")
    C.linear(X, M, K, K, W, N, out, N, bias=b, relu=1, bn=bn, pair=True)
    assert rel_fro(out, F.relu(X.float() @ W.float().t() + b)) < 6e-3


def test_pair_fp32_residual_dropout_and_accumulate(C):
    M, N, K = 2048, 768, 3072
    X, W, R = rnd(M, K, seed=1, dtype=BF), rnd(N, K, seed=2, scale=K ** -0.5, dtype=BF), rnd(M, N, seed=3)
    out = torch.zeros(M, N, device="cuda")
    C.linear(X, M, K, K, W, N, out, N, out_fp32=1, residual=R, ldr=N, res_fp32=1, bn=128, pair=True)
    ref = X.float() @ W.float().t() + R
    assert rel_fro(out, ref) < 2e-5
    C.linear(X, M, K, K, W, N, out, N, out_fp32=1, residual=R, ldr=N, res_fp32=1, accumulate=1, bn=128, pair=True)
    assert rel_fro(out, 2 * ref) < 2e-5
    rng = rng_state()
    a, b = torch.zeros(M, N, device="cuda"), torch.zeros(M, N, device="cuda")
    C.linear(X, M, K, K, W, N, a, N, out_fp32=1, residual=R, ldr=N, res_fp32=1, drop_p=0.1, sid=4, rng=rng, bn=128, pair=True)
    C.linear(X, M, K, K, W, N, b, N, out_fp32=1, residual=R, ldr=N, res_fp32=1, drop_p=0.1, sid=4, rng=rng, bn=128)
    assert rel_fro(a, b) < 1e-6       # same dropout mask, same sums as the single-CTA kernel


@pytest.mark.parametrize("M,N,K,bn", [(2048, 768, 3072, 256), (2048, 3072, 768, 128), (2048, 2304, 768, 128)])
def test_pair_dgrad_with_relu_mask(C, M, N, K, bn):
    dY = rnd(M, N, seed=1, dtype=BF)
    W, act = rnd(N, K, seed=2, scale=N ** -0.5, dtype=BF), rnd(M, K, seed=3, dtype=BF)
    out = torch.zeros(M, K, dtype=BF, device="cuda")
    C.dgrad(dY, M, N, N, W, K, out, K, relu_mask=act, ldm=K, bn=bn, pair=True)
    assert rel_fro(out, (dY.float() @ W.float()) * (act.float() > 0)) < 6e-3


def test_pair_split_k(C):
    M, N, K = 512, 512, 4096
    X, W = rnd(M, K, seed=1, dtype=BF), rnd(N, K, seed=2, dtype=BF)
    out = torch.zeros(M, N, device="cuda")
    C.linear(X, M, K, K, W, N, out, N, out_fp32=1, bn=128, split_k=4, pair=True)
    assert rel_fro(out, X.float() @ W.float().t()) < 2e-5


@pytest.mark.parametrize("M,N,K,bn", [(2048, 3072, 768, 128), (2048, 768, 3072, 256), (196, 1536, 768, 128),
                                      (2048, 2304, 768, 128), (1024, 3072, 3072, 128)])
def test_pair_wgrad(C, M, N, K, bn):
    dY, X = rnd(M, N, seed=1, dtype=BF), rnd(M, K, seed=2, dtype=BF)
    dW = torch.zeros(N, K, device="cuda")
    C.wgrad(dY, M, N, N, X, K, K, dW, bn=bn, pair=True)
    assert rel_fro(dW, dY.float().t() @ X.float()) < 2e-5


@pytest.mark.parametrize("N,H,Cin,Cout,R,s,p,bn,res", [
    (4, 56, 64, 256, 1, 1, 0, 128, True), (4, 28, 128, 512, 1, 1, 0, 256, True),
    (8, 14, 256, 256, 3, 1, 1, 128, True), (5, 7, 512, 2048, 1, 1, 0, 256, True),
    (4, 56, 256, 512, 1, 2, 0, 128, False), (3, 8, 512, 768, 3, 1, 1, 256, False)])
def test_pair_conv_bias_residual_relu(C, N, H, Cin, Cout, R, s, p, bn, res):
    x = rnd(N, Cin, H, H, seed=1, dtype=BF)
    w = rnd(Cout, Cin, R, R, seed=2, scale=(Cin * R * R) ** -0.5, dtype=BF)
    b = rnd(Cout, seed=3)
    ref = F.conv2d(x.float(), w.float(), b, stride=s, padding=p)
    r = None
    if res:
        r = rnd(*ref.shape, seed=4, dtype=BF)
        ref = ref + r.float()
    ref = F.relu(ref)
    out = torch.zeros(N, ref.shape[2], ref.shape[3], Cout, dtype=BF, device="cuda")
    wk = w.permute(0, 2, 3, 1).contiguous()
    C.conv(N, H, H, Cin, Cout, R, s, p, x.permute(0, 2, 3, 1).contiguous(), wk, out, bias=b,
           residual=r.permute(0, 2, 3, 1).contiguous() if res else None, relu=1, bn=bn, pair=True)
    assert rel_fro(out, ref.permute(0, 2, 3, 1).contiguous()) < 6e-3


# ------------------------------------------------------------------------------------------------
# convolutions
# ------------------------------------------------------------------------------------------------
def nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


@pytest.mark.parametrize("N,H,Cin,Cout,R,s,p,bn,res", [
    (4, 56, 64, 64, 3, 1, 1, 64, False), (4, 56, 64, 256, 1, 1, 0, 128, False),
    (4, 28, 128, 512, 1, 1, 0, 256, True), (4, 56, 128, 128, 3, 2, 1, 128, False),
    (4, 56, 256, 512, 1, 2, 0, 128, False), (8, 14, 256, 256, 3, 1, 1, 128, True),
    (8, 7, 512, 512, 3, 1, 1, 64, True), (4, 8, 512, 768, 3, 1, 1, 128, False)])
def test_conv_bias_residual_relu(C, N, H, Cin, Cout, R, s, p, bn, res):
    x = rnd(N, Cin, H, H, seed=1, dtype=BF)
    w = rnd(Cout, Cin, R, R, seed=2, scale=(Cin * R * R) ** -0.5, dtype=BF)
    b = rnd(Cout, seed=3)
    ref = F.conv2d(x.float(), w.float(), b, stride=s, padding=p)
    r = None
    if res:
        r = rnd(*ref.shape, seed=4, dtype=BF)
        ref = ref + r.float()
    ref = F.relu(ref)
    out = torch.zeros(N, ref.shape[2], ref.shape[3], Cout, dtype=BF, device="cuda")
    wk = w.permute(0, 2, 3, 1).contiguous()
    C.conv(N, H, H, Cin, Cout, R, s, p, nhwc(x), wk, out, bias=b, residual=nhwc(r) if res else None, relu=1, bn=bn)
    assert rel_fro(out, nhwc(ref)) < 6e-3


@pytest.mark.parametrize("N,H", [(2, 224), (2, 64), (3, 256)])
def test_stem_fold_pack_maxpool(C, N, H):
    """image_to_stem + fold_conv_bn + stem conv + maxpool == conv7x7/2 -> BN(eval) -> ReLU -> MaxPool(3,2,1)."""
    img = torch.rand(N, 3, H, H, generator=torch.Generator().manual_seed(5)).cuda()
    w = rnd(64, 3, 7, 7, seed=6, scale=147 ** -0.5)
    gamma, beta = 0.5 + torch.rand(64).cuda(), rnd(64, seed=7, scale=0.1)
    mean, var = rnd(64, seed=8, scale=0.1), 0.5 + torch.rand(64).cuda()
    ref = F.relu(F.batch_norm(F.conv2d(img, w, None, 2, 3), mean, var, gamma, beta, False, 0.0, 1e-5))
    ref = F.max_pool2d(ref, 3, 2, 1)
    xp = torch.empty(N, H, H + 8, 8, dtype=BF, device="cuda")
    C.image_to_stem(img, xp, N, H, H)
    wk = torch.empty(64 * 7 * 8 * 8, dtype=BF, device="cuda")
    bias = torch.empty(64, device="cuda")
    C.fold_conv_bn(w, gamma, beta, mean, var, 1e-5, wk, bias, 64, 3, 7, 7, 8, 8)
    Ho = (H + 6 - 7) // 2 + 1
    c1 = torch.zeros(N, Ho, Ho, 64, dtype=BF, device="cuda")
    C.conv(N, H, H, 8, 64, 7, 2, 3, xp, wk, c1, bias=bias, relu=1, stem7=1, bn=64)
    Hp = (Ho + 2 - 3) // 2 + 1
    out = torch.zeros(N, Hp, Hp, 64, dtype=BF, device="cuda")
    C.maxpool3x3s2(c1, out, N, Ho, Ho, 64)
    assert rel_fro(out, nhwc(ref)) < 8e-3


@pytest.mark.parametrize("N,H,Cin,bn,split,pair", [(8, 7, 512, 256, 1, False), (16, 7, 2048, 256, 2, False),
                                                   (4, 8, 512, 128, 1, False), (16, 7, 2048, 256, 2, True),
                                                   (8, 7, 512, 256, 1, True)])
def test_convT_projection_forward_and_wgrad(C, N, H, Cin, bn, split, pair):
    """ConvTranspose2d(k3,s1,p1) == 3x3 same conv on the prepared weight; wgrad maps back to [Cin,Cout,3,3]."""
    Cout = 768
    x = rnd(N, Cin, H, H, seed=1, dtype=BF)
    w = rnd(Cin, Cout, 3, 3, seed=2, scale=(Cout * 9) ** -0.5).requires_grad_(True)
    b = rnd(Cout, seed=3)
    ref = F.conv_transpose2d(x.float(), w, b, 1, 1)
    dy = rnd(N, Cout, H, H, seed=4, dtype=BF)
    ref.backward(dy.float())
    wk = torch.empty(Cout * 9 * Cin, dtype=BF, device="cuda")
    C.convT_weight_prep(w.detach(), wk, Cin, Cout)
    out = torch.zeros(N * H * H, Cout, dtype=BF, device="cuda")
    C.conv(N, H, H, Cin, Cout, 3, 1, 1, nhwc(x), wk, out, bias=b, relu=0, pair=pair)
    assert rel_fro(out, nhwc(ref.detach()).view(-1, Cout)) < 8e-3
    dwc = torch.zeros(Cout, 9 * Cin, device="cuda")
    C.conv_wgrad(N, H, H, Cin, Cout, nhwc(dy), nhwc(x), dwc, bn, split, pair=pair)
    dw = torch.empty_like(w)
    C.convT_wgrad_unprep(dwc, dw, Cin, Cout)
    assert rel_fro(dw, w.grad) < 2e-5
    bsum = torch.zeros(Cout, device="cuda")
    C.colsum_bf16(nhwc(dy), Cout, bsum, N * H * H, Cout)
    assert rel_fro(bsum, dy.float().sum((0, 2, 3))) < 1e-5


def test_nhwc_to_nchw(C):
    x = rnd(3, 7, 7, 512, seed=1, dtype=BF)
    out = torch.empty(3, 512, 7, 7, device="cuda")
    C.nhwc_to_nchw_f32(x, out, 3, 7, 7, 512)
    assert torch.equal(out, x.float().permute(0, 3, 1, 2))


# ------------------------------------------------------------------------------------------------
# norms
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M", [128, 2048, 77])
def test_rmsnorm_fwd_bwd(C, M):
    D = 768
    x = rnd(M, D, seed=1, scale=3.0).requires_grad_(True)
    w = (0.75 + 0.5 * torch.rand(D)).cuda().requires_grad_(True)
    var = x.pow(2).mean(-1, keepdim=True)
    y = w * (x * torch.rsqrt(var + 1e-6))
    dy, dres = rnd(M, D, seed=2), rnd(M, D, seed=3)
    y.backward(dy)
    yb, yf, rstd = torch.empty(M, D, dtype=BF, device="cuda"), torch.empty(M, D, device="cuda"), torch.empty(M, device="cuda")
    C.rmsnorm_fwd(x.detach(), w.detach(), yb, yf, rstd, M, D, 1e-6, 0.0, 0, None)
    assert rel_fro(yf, y.detach()) < 1e-5 and rel_fro(yb, y.detach()) < 4e-3
    dx, dw = torch.empty(M, D, device="cuda"), torch.zeros(D, device="cuda")
    C.rmsnorm_bwd(dy, 1, x.detach(), w.detach(), rstd, dres, dx, dw, M, D, 0.0, 0, None, None, 0.0, 0)
    assert rel_fro(dx, x.grad + dres) < 1e-5 and rel_fro(dw, w.grad) < 1e-5
    # bf16 upstream gradient, in-place residual accumulation (dx aliases dres)
    dyb = dy.to(BF)
    acc = dres.clone()
    dw.zero_()
    gb = torch.empty(M, D, dtype=BF, device="cuda")
    rng = rng_state()
    C.rmsnorm_bwd(dyb, 0, x.detach(), w.detach(), rstd, acc, acc, dw, M, D, 0.0, 0, rng, gb, 0.1, 5)
    gb2 = torch.empty_like(gb)
    C.dropout_cast(acc, gb2, M, D, 0.1, 5, rng)          # fused masked copy == separate dropout_cast of dx
    assert torch.equal(gb, gb2)
    x.grad = None
    w.grad = None
    y2 = w * (x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + 1e-6))
    y2.backward(dyb.float())
    assert rel_fro(acc, x.grad + dres) < 1e-5 and rel_fro(dw, w.grad) < 1e-5


@pytest.mark.parametrize("M", [128, 2048, 50])
def test_layernorm_fwd_bwd(C, M):
    D = 768
    z = rnd(M, D, seed=1, scale=2.0).requires_grad_(True)
    g = (0.75 + 0.5 * torch.rand(D)).cuda().requires_grad_(True)
    b = rnd(D, seed=2, scale=0.1).requires_grad_(True)
    y = F.layer_norm(z, (D,), g, b, 1e-5)
    dy = rnd(M, D, seed=3)
    y.backward(dy)
    yb, yf = torch.empty(M, D, dtype=BF, device="cuda"), torch.empty(M, D, device="cuda")
    mean, rstd = torch.empty(M, device="cuda"), torch.empty(M, device="cuda")
    C.layernorm_fwd(z.detach(), g.detach(), b.detach(), yb, yf, mean, rstd, M, D, 1e-5)
    assert rel_fro(yf, y.detach()) < 1e-5 and rel_fro(yb, y.detach()) < 4e-3
    dz, dg, db = torch.empty(M, D, device="cuda"), torch.zeros(D, device="cuda"), torch.zeros(D, device="cuda")
    gb, gcs = torch.empty(M, D, dtype=BF, device="cuda"), torch.zeros(D, device="cuda")
    rng = rng_state()
    C.layernorm_bwd(dy, z.detach(), g.detach(), mean, rstd, dz, dg, db, M, D, gb, 0.1, 9, rng, gcs)
    gb2 = torch.empty_like(gb)
    C.dropout_cast(dz, gb2, M, D, 0.1, 9, rng)
    assert torch.equal(gb, gb2) and rel_fro(gcs, gb2.float().sum(0)) < 2e-3
    assert rel_fro(dz, z.grad) < 1e-5 and rel_fro(dg, g.grad) < 1e-5 and rel_fro(db, b.grad) < 1e-5


# ------------------------------------------------------------------------------------------------
# T5 embedding / bias
# ------------------------------------------------------------------------------------------------
def test_embedding_and_bias(C, pkg):
    from t5_resnet_vqa_b200.engine import t5_relative_buckets
    M, D, V = 256, 768, 1000
    ids = torch.randint(0, V, (M,), generator=torch.Generator().manual_seed(0)).cuda()
    ids[:8] = 3  # duplicates exercise the scatter-add
    table = rnd(V, D, seed=1).requires_grad_(True)
    out = torch.empty(M, D, device="cuda")
    C.embedding_fwd(ids, table.detach(), out, M, D, V, 0.0, 0, None)
    assert torch.equal(out, table.detach()[ids])
    dout = rnd(M, D, seed=2)
    F.embedding(ids, table).backward(dout)
    dt = torch.zeros(V, D, device="cuda")
    C.embedding_bwd(ids, dout, dt, M, D, V, 0.0, 0, None)
    assert rel_fro(dt, table.grad) < 1e-6
    H, Lq = 12, 32
    bucket = t5_relative_buckets(Lq, Lq).cuda()
    tb = rnd(32, H, seed=3).requires_grad_(True)
    ref = tb[bucket.long()].permute(2, 0, 1)
    bias = torch.empty(H, Lq, Lq, device="cuda")
    C.t5_bias_build(tb.detach(), bucket, bias, H, Lq, 32)
    assert torch.equal(bias, ref.detach())
    db = rnd(H, Lq, Lq, seed=4)
    ref.backward(db)
    dtb = torch.zeros(32, H, device="cuda")
    C.t5_bias_grad(db, bucket, dtb, H, Lq, 32)
    assert rel_fro(dtb, tb.grad) < 1e-5


# ------------------------------------------------------------------------------------------------
# attention
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,H,Lq,Lk,hd,t5,tc", [
    (4, 12, 32, 32, 64, True, False), (3, 12, 16, 16, 64, True, False), (4, 8, 32, 32, 96, False, False),
    (4, 8, 32, 49, 96, False, False), (2, 8, 16, 64, 96, False, False), (2, 8, 32, 196, 96, False, False),
    # tcgen05 flash kernels, 32-row slots (Lq, Lk <= 32): full groups, ragged last group, short sequences, ragged Lk
    (4, 12, 32, 32, 64, True, True), (3, 12, 16, 16, 64, True, True), (64, 12, 32, 32, 64, True, True),
    (4, 8, 32, 32, 96, False, True), (3, 7, 32, 32, 96, False, True), (5, 6, 20, 27, 64, True, True),
    (2, 8, 16, 16, 96, False, True),
    # 64-row slots (33 <= max(Lq, Lk) <= 64): the guided attention over 49 (224x224) / 64 (256x256) vision tokens, the
    # reference's native 16-token questions, odd pair counts, long questions through the T5 path (bias + key mask)
    (64, 8, 32, 49, 96, False, True), (4, 8, 32, 64, 96, False, True), (2, 8, 16, 64, 96, False, True),
    (3, 7, 32, 49, 96, False, True), (3, 12, 64, 64, 64, True, True), (5, 12, 40, 40, 64, True, True),
    (2, 12, 33, 50, 64, True, True)])
def test_attention_fwd_bwd(C, B, H, Lq, Lk, hd, t5, tc):
    D = H * hd
    q = rnd(B * Lq, D, seed=1, scale=1.0 if not t5 else 0.4, dtype=BF)
    k = rnd(B * Lk, D, seed=2, scale=1.0 if not t5 else 0.4, dtype=BF)
    v = rnd(B * Lk, D, seed=3, dtype=BF)
    dO = rnd(B * Lq, D, seed=4, dtype=BF)
    bias = mask = None
    scale = 1.0 if t5 else 1.0 / math.sqrt(hd)
    qf, kf, vf = [t.float().requires_grad_(True) for t in (q, k, v)]
    qh = qf.view(B, Lq, H, hd).transpose(1, 2)
    kh = kf.view(B, Lk, H, hd).transpose(1, 2)
    vh = vf.view(B, Lk, H, hd).transpose(1, 2)
    s = torch.matmul(qh, kh.transpose(-1, -2)) * scale
    if t5:
        bias = rnd(H, Lq, Lk, seed=5).requires_grad_(True)
        mask = torch.ones(B, Lk, dtype=torch.long, device="cuda")
        mask[:, Lk - 5:] = 0
        mask[0] = 1
        s = s + bias[None] + (1.0 - mask[:, None, None, :].float()) * torch.finfo(torch.float32).min
    p = F.softmax(s, dim=-1)
    o = torch.matmul(p, vh).transpose(1, 2).reshape(B * Lq, D)
    o.backward(dO.float())
    out = torch.zeros(B * Lq, D, dtype=BF, device="cuda")
    probs = None if tc else torch.zeros(B * H * Lq * Lk, dtype=F32, device="cuda")
    stats = torch.zeros(B * H * Lq, 2, dtype=F32, device="cuda") if tc else None
    bdet = bias.detach() if t5 else None
    C.attn_fwd(B, H, Lq, Lk, hd, q, D, k, D, v, D, out, D, probs, bdet, mask, scale, 0.0, 0, None, stats=stats)
    assert rel_fro(out, o.detach()) < 6e-3
    if tc:
        sv = stats.view(B, H, Lq, 2)
        assert rel_fro(sv[..., 0], s.detach().max(-1).values) < 1e-5
        assert rel_fro(1.0 / sv[..., 1], (s.detach() - s.detach().max(-1, keepdim=True).values).exp().sum(-1)) < 1e-4
    else:
        assert rel_fro(probs.view(B, H, Lq, Lk), p.detach()) < 2e-3
    dq, dk, dv = [torch.zeros_like(t) for t in (q, k, v)]
    dbias = torch.zeros(H, Lq, Lk, device="cuda") if t5 else None
    C.attn_bwd(B, H, Lq, Lk, hd, q, D, k, D, v, D, probs, dO, D, dq, D, dk, D, dv, D, dbias, scale, 0.0, 0, None,
               stats=stats, bias=bdet, key_mask=mask)
    assert rel_fro(dq, qf.grad) < 1.5e-2 and rel_fro(dk, kf.grad) < 1.5e-2 and rel_fro(dv, vf.grad) < 1.5e-2
    assert cosine(dq, qf.grad) > 0.9999 and cosine(dk, kf.grad) > 0.9999
    if t5:
        assert rel_fro(dbias, bias.grad) < 1.5e-2


@pytest.mark.parametrize("H,hd,Lk", [(12, 64, 32), (8, 96, 32), (6, 64, 27), (8, 96, 49), (8, 96, 64), (12, 64, 40)])
def test_attention_tcgen05_dropout_matches_simt(C, H, hd, Lk):
    """Both kernels draw the dropout mask of probability (b,h,i,j) from the same Philox stream, so with dropout on
    the tcgen05 path must reproduce the SIMT path (forward and every gradient) up to bf16 rounding of P."""
    B, Lq, D = 5, 32, H * hd
    q, k = rnd(B * Lq, D, seed=1, scale=0.5, dtype=BF), rnd(B * Lk, D, seed=2, scale=0.5, dtype=BF)
    v, dO = rnd(B * Lk, D, seed=3, dtype=BF), rnd(B * Lq, D, seed=4, dtype=BF)
    rng = rng_state()
    res = []
    for tc in (False, True):
        out = torch.zeros(B * Lq, D, dtype=BF, device="cuda")
        probs = None if tc else torch.zeros(B * H * Lq * Lk, dtype=F32, device="cuda")
        stats = torch.zeros(B * H * Lq, 2, dtype=F32, device="cuda") if tc else None
        C.attn_fwd(B, H, Lq, Lk, hd, q, D, k, D, v, D, out, D, probs, None, None, 0.125, 0.1, 3, rng, stats=stats)
        dq, dk, dv = [torch.zeros_like(t) for t in (q, k, v)]
        C.attn_bwd(B, H, Lq, Lk, hd, q, D, k, D, v, D, probs, dO, D, dq, D, dk, D, dv, D, None, 0.125, 0.1, 3, rng,
                   stats=stats)
        res.append((out, dq, dk, dv))
    for a, b in zip(*res):
        assert rel_fro(b, a) < 1.2e-2 and cosine(a, b) > 0.9999


# ------------------------------------------------------------------------------------------------
# dropout: keep-rate, scale, and forward/backward mask identity
# ------------------------------------------------------------------------------------------------
def test_dropout_mask_consistency(C):
    M, N, K, p = 2048, 768, 768, 0.1
    rng = rng_state()
    ones = torch.ones(M, N, device="cuda")
    m1 = torch.empty(M, N, dtype=BF, device="cuda")
    C.dropout_cast(ones, m1, M, N, p, 11, rng)
    keep = float((m1 > 0).float().mean())
    assert abs(keep - 0.9) < 2e-3
    assert torch.allclose(m1[m1 > 0].float(), torch.tensor(1 / 0.9).cuda(), rtol=4e-3)
    # the GEMM epilogue regenerates the same mask for the same (sid, rng)
    X = torch.eye(K, dtype=BF, device="cuda").repeat(M // K + 1, 1)[:M].contiguous()
    W = torch.ones(N, K, dtype=BF, device="cuda")
    out = torch.empty(M, N, dtype=BF, device="cuda")
    C.linear(X, M, K, K, W, N, out, N, drop_p=p, sid=11, rng=rng, bn=128)
    assert torch.equal(out > 0, m1 > 0)
    m2 = torch.empty_like(m1)
    C.dropout_cast(ones, m2, M, N, p, 12, rng)      # another stream id -> another mask
    assert not torch.equal(m2 > 0, m1 > 0)
    C.rng_advance(rng)
    C.dropout_cast(ones, m2, M, N, p, 11, rng)      # next step -> another mask
    assert not torch.equal(m2 > 0, m1 > 0)
    assert int(rng[1]) == 8


# ------------------------------------------------------------------------------------------------
# head
# ------------------------------------------------------------------------------------------------
def test_pooler_and_loss(C):
    B, Lq, D, A = 64, 32, 768, 170
    x = rnd(B, Lq, D, seed=1).requires_grad_(True)
    a = rnd(1, D, seed=2, scale=D ** -0.5).requires_grad_(True)
    b = rnd(1, seed=3).requires_grad_(True)
    w = F.softmax(F.linear(x, a, b), dim=1).transpose(1, 2)
    pooled = torch.bmm(w, x).squeeze(1)
    dp = rnd(B, D, seed=4)
    pooled.backward(dp)
    wo, pf, pb = torch.empty(B, Lq, device="cuda"), torch.empty(B, D, device="cuda"), torch.empty(B, D, dtype=BF, device="cuda")
    C.pooler_fwd(x.detach(), a.detach(), b.detach(), wo, pf, pb, B, Lq, D)
    assert rel_fro(pf, pooled.detach()) < 1e-5 and rel_fro(pb, pooled.detach()) < 4e-3
    dx, da, db = torch.empty(B, Lq, D, device="cuda"), torch.zeros(D, device="cuda"), torch.zeros(1, device="cuda")
    C.pooler_bwd(x.detach(), a.detach(), wo, dp, dx, da, db, B, Lq, D)
    assert rel_fro(dx, x.grad) < 1e-5 and rel_fro(da, a.grad.flatten()) < 1e-4
    assert abs(float(db) - float(b.grad)) < 1e-4
    # log_softmax + NLL
    ld = 176
    logits = torch.zeros(B, ld, device="cuda")
    logits[:, :A] = rnd(B, A, seed=5, scale=2.0)
    lg = logits[:, :A].clone().requires_grad_(True)
    labels = torch.randint(0, A, (B,), generator=torch.Generator().manual_seed(6)).cuda()
    ref_lp = F.log_softmax(lg, -1)
    ref_loss = F.nll_loss(ref_lp, labels)
    gl = torch.tensor([0.7], device="cuda")
    glp = rnd(B, A, seed=7, scale=0.01)
    (ref_loss * gl[0] + (ref_lp * glp).sum()).backward()
    lp, loss = torch.empty(B, A, device="cuda"), torch.zeros(1, device="cuda")
    C.logsoftmax_nll_fwd(logits, ld, labels, lp, loss, B, A)
    assert rel_fro(lp, ref_lp.detach()) < 1e-6 and abs(float(loss) - float(ref_loss)) < 1e-5
    dl = torch.empty(B, ld, dtype=BF, device="cuda")
    C.logsoftmax_nll_bwd(lp, labels, gl, glp, dl, ld, B, A)
    assert rel_fro(dl[:, :A], lg.grad) < 4e-3 and float(dl[:, A:].abs().max()) == 0.0


# ------------------------------------------------------------------------------------------------
# optimizer
# ------------------------------------------------------------------------------------------------
def test_adamw_amsgrad_matches_torch(C, pkg):
    n = 1_000_003
    p0 = rnd(n, seed=1)
    ref = p0.clone().requires_grad_(True)
    opt = torch.optim.AdamW([ref], lr=5e-3, weight_decay=0.1, amsgrad=True)
    mine = p0.clone()
    m, v, vm = torch.zeros_like(mine), torch.zeros_like(mine), torch.zeros_like(mine)
    shadow = torch.empty(n, dtype=BF, device="cuda")
    lib = pkg.lib.load()
    s = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    for t in range(1, 6):
        g = rnd(n, seed=10 + t, scale=1.0 / t)
        ref.grad = g.clone()
        opt.step()
        pkg.lib.check(lib.vqa_adamw_amsgrad(None, mine.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(),
                                            vm.data_ptr(), shadow.data_ptr(), n, 5e-3, 0.9, 0.999, 1e-8, 0.1,
                                            1 - 0.9 ** t, 1 - 0.999 ** t, None, 0.0, 1, s))
    assert float((mine - ref.detach()).abs().max()) < 2e-6
    st = opt.state[ref]
    assert rel_fro(m, st["exp_avg"]) < 1e-6 and rel_fro(v, st["exp_avg_sq"]) < 1e-6
    assert rel_fro(vm, st["max_exp_avg_sq"]) < 1e-6
    assert torch.equal(shadow, mine.to(BF))
    ss = torch.zeros(1, device="cuda")
    C.sumsq_f32(mine, n, ss)
    assert abs(float(ss) / float(mine.double().pow(2).sum()) - 1) < 1e-5


def test_fused_optimizer_clip_matches_clip_grad_norm(C, pkg):
    torch.manual_seed(0)
    ps = [rnd(1000, 64, seed=i) for i in range(4)]
    ref = [p.clone().requires_grad_(True) for p in ps]
    mine = [p.clone().requires_grad_(True) for p in ps]
    o_ref = torch.optim.AdamW([{"params": ref[:2], "lr": 1e-3}, {"params": ref[2:], "lr": 5e-4}],
                              weight_decay=0.1, amsgrad=True)
    o_mine = torch.optim.VQAFusedAdamW([{"params": mine[:2], "lr": 1e-3}, {"params": mine[2:], "lr": 5e-4}],
                                       weight_decay=0.1, amsgrad=True, max_grad_norm=1.0)
    for step in range(3):
        for a, b in zip(ref, mine):
            g = rnd(*a.shape, seed=100 + step)
            a.grad, b.grad = g.clone(), g.clone()
        torch.nn.utils.clip_grad_norm_(ref, 1.0)
        o_ref.step()
        o_mine.step()
    for a, b in zip(ref, mine):
        assert float((a - b).abs().max()) < 2e-6
    sd = o_mine.state_dict()
    assert set(sd["state"][0].keys()) == {"step", "exp_avg", "exp_avg_sq", "max_exp_avg_sq"}


# ------------------------------------------------------------------------------------------------
# plans
# ------------------------------------------------------------------------------------------------
def test_plan_replay_and_graph(C, pkg):
    from t5_resnet_vqa_b200.engine import _Rec
    lib = pkg.lib.load()
    plan = lib.vqa_plan_create()
    r = _Rec(lib, plan, None)
    M, N, K = 256, 256, 256
    X, W = rnd(M, K, seed=1, dtype=BF), rnd(N, K, seed=2, dtype=BF)
    y = torch.zeros(M, N, device="cuda")
    acc = torch.zeros(M, N, device="cuda")
    r.linear(X, M, K, K, W, N, y, N, out_fp32=1, bn=128)
    r.axpy_f32(acc, y, 1.0, M * N)
    assert lib.vqa_plan_size(plan) == 2
    assert float(acc.abs().max()) == 0.0            # nothing ran while recording
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        sp = ctypes.c_void_p(side.cuda_stream)
        pkg.lib.check(lib.vqa_plan_run(plan, sp))
        pkg.lib.check(lib.vqa_plan_capture_graph(plan, sp))
        pkg.lib.check(lib.vqa_plan_run(plan, sp))
        side.synchronize()
    assert rel_fro(acc, 2 * (X.float() @ W.float().t())) < 2e-5
    lib.vqa_plan_destroy(plan)


def test_plan_time_ops_selects_a_kernel_family(C, pkg):
    """bench.py's roofline timing: only the named ops are replayed (back to back), flops and launch counts come from
    the recorded notes, and the per-launch profile of the same plan still covers every op."""
    from t5_resnet_vqa_b200.engine import _Rec
    lib = pkg.lib.load()
    plan = lib.vqa_plan_create()
    r = _Rec(lib, plan, None)
    M, N, K = 512, 256, 384
    X, W = rnd(M, K, seed=1, dtype=BF), rnd(N, K, seed=2, dtype=BF)
    y = torch.zeros(M, N, device="cuda")
    acc = torch.zeros(M, N, device="cuda")
    for _ in range(3):
        r.linear(X, M, K, K, W, N, y, N, out_fp32=1, bn=128)
        r.axpy_f32(acc, y, 1.0, M * N)
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        sp = ctypes.c_void_p(side.cuda_stream)
        ms, fl, n = ctypes.c_float(), ctypes.c_double(), ctypes.c_int()
        pkg.lib.check(lib.vqa_plan_time_ops(plan, sp, b"gemm", 4, ctypes.byref(ms), ctypes.byref(fl), ctypes.byref(n)))
        assert n.value == 3 and fl.value == 3 * 2.0 * M * N * K and 0.0 < ms.value < 50.0
        assert float(acc.abs().max()) == 0.0            # the axpy ops were not part of the replay
        pkg.lib.check(lib.vqa_plan_time_ops(plan, sp, b"conv,nothing", 2, ctypes.byref(ms), ctypes.byref(fl), ctypes.byref(n)))
        assert n.value == 0 and ms.value == 0.0
        per = (ctypes.c_float * 6)()
        pkg.lib.check(lib.vqa_plan_profile(plan, sp, per, 100))
        side.synchronize()
        assert all(p > 0.0 for p in per)
    assert rel_fro(acc, 3 * (X.float() @ W.float().t())) < 2e-5
    lib.vqa_plan_destroy(plan)
