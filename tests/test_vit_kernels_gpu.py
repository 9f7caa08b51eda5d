"""Kernel-level parity (GPU) of the VitVQAModel step's own entry points (include/vqa_b200.h, "VitVQAModel step") against
the matching torch fp32 ops on the same inputs.  Tolerances as in test_kernels_gpu.py: bf16 outputs <= 6e-3 rel-Frobenius,
fp32 plumbing exact or <= 1e-6."""
import math

import pytest
import torch
import torch.nn.functional as F

from util import Caller, rel_fro

pytestmark = pytest.mark.gpu
BF, F32 = torch.bfloat16, torch.float32


@pytest.fixture(scope="module")
def C(pkg, cuda):
    return Caller(pkg)


def rnd(*shape, seed=0, scale=1.0, dtype=F32):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).cuda().to(dtype)


@pytest.mark.parametrize("B,H,L", [(3, 12, 197), (2, 4, 128), (2, 3, 256), (5, 2, 50), (1, 12, 129)])
def test_attention_long_fwd(C, B, H, L):
    """softmax(Q K^T / 8) V over L tokens on both sides, fused q|k|v rows as the ViT's QKV GEMM writes them."""
    hd, D = 64, H * 64
    qkv = rnd(B * L, 3 * D, seed=1, scale=1.5, dtype=BF)
    out = torch.full((B * L, D), float("nan"), dtype=BF, device="cuda")
    e = qkv.element_size()
    probs = torch.full((B, H, L, L), float("nan"), device="cuda") if B <= 3 else None
    C.attention_long_fwd(qkv, 3 * D, qkv.data_ptr() + D * e, 3 * D, qkv.data_ptr() + 2 * D * e, 3 * D, out, D, B, H, L, hd,
                         1.0 / math.sqrt(hd), probs)
    q, k, v = [t.float().reshape(B, L, H, hd).transpose(1, 2) for t in qkv.split(D, dim=1)]
    p_ref = F.softmax(torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(hd), dim=-1)
    ref = torch.matmul(p_ref, v).transpose(1, 2).reshape(B * L, D)
    assert bool(torch.isfinite(out.float()).all())
    assert rel_fro(out, ref) < 6e-3
    if probs is not None:          # the attention maps generate_answers hands to the heat-map script
        assert rel_fro(probs, p_ref) < 1e-4


def test_patchify_assemble_match_conv_embeddings(C):
    """vit: ViTEmbeddings = Conv2d(3, 768, 16, 16) patches + cls token + position embeddings."""
    B, Himg, P, D = 3, 224, 16, 768
    NP = (Himg // P) ** 2
    img = rnd(B, 3, Himg, Himg, seed=1)
    w, bias = rnd(D, 3, P, P, seed=2, scale=0.03), rnd(D, seed=3, scale=0.1)
    cls, pos = rnd(1, 1, D, seed=4), rnd(1, NP + 1, D, seed=5)
    patches = torch.zeros(B * NP, 3 * P * P, dtype=BF, device="cuda")
    C.vit_patchify(img, patches, B, Himg, Himg, P)
    ref_p = F.unfold(img, P, stride=P).transpose(1, 2).reshape(B * NP, 3 * P * P)
    assert torch.equal(patches, ref_p.to(BF))
    proj = torch.zeros(B * NP, D, device="cuda")
    C.linear(patches, B * NP, 3 * P * P, 3 * P * P, w.reshape(D, -1).to(BF).contiguous(), D, proj, D, out_fp32=1, bias=bias)
    hidden = torch.zeros(B * (NP + 1), D, device="cuda")
    C.vit_assemble(proj, cls, pos, hidden, B, NP, D)
    x = F.conv2d(img.to(BF).float(), w.to(BF).float(), bias, stride=P).flatten(2).transpose(1, 2)
    ref = torch.cat([cls.expand(B, -1, -1), x], dim=1) + pos
    assert rel_fro(hidden, ref.reshape(B * (NP + 1), D)) < 2e-5


@pytest.mark.parametrize("M,N,K,bn", [(12608, 3072, 768, 256), (394, 3072, 768, 128), (200, 192, 128, 64)])
def test_linear_bias_gelu_epilogue(C, M, N, K, bn):
    """fc1 of the ViT layer: bf16 out = gelu(x W^T + b) in the GEMM epilogue (vqa_gemm_args.relu = 2)."""
    X, W, b = rnd(M, K, seed=1, dtype=BF), rnd(N, K, seed=2, scale=2.0 * K ** -0.5, dtype=BF), rnd(N, seed=3)
    out = torch.zeros(M, N, dtype=BF, device="cuda")
    C.linear(X, M, K, K, W, N, out, N, bias=b, relu=2, bn=bn)
    assert rel_fro(out, F.gelu(X.float() @ W.float().t() + b)) < 6e-3


def test_gelu_and_fuse_concat(C):
    x = rnd(1000, 3072, seed=1, scale=2.0, dtype=BF)
    y = x.clone()
    C.gelu_bf16(y, y.numel())
    assert rel_fro(y, F.gelu(x.float())) < 4e-3
    B, L, D = 7, 16, 768
    pre, enc = rnd(B, D, seed=2), rnd(B * L, D, seed=3)
    out = torch.zeros(B, 2 * D, dtype=BF, device="cuda")
    pooled = torch.zeros(B, D, device="cuda")
    C.vit_fuse_concat(pre, enc, L, out, pooled, B, D)
    ref = torch.cat([torch.tanh(pre), enc.view(B, L, D)[:, 0]], dim=1)
    assert rel_fro(pooled, torch.tanh(pre)) < 1e-6
    assert torch.equal(out, ref.to(BF)) or rel_fro(out, ref) < 4e-3


@pytest.mark.parametrize("p", [0.0, 0.1])
def test_one_key_cross_attention(C, p):
    """ctx = dropout(softmax over ONE key = 1) * v, and its backward: the same mask in both directions, mean keep-rate ~ 1-p."""
    B, H, Lq, hd = 6, 12, 20, 64
    D = H * hd
    rng = torch.tensor([99, 3], dtype=torch.int64, device="cuda")
    v = rnd(B, D, seed=1, dtype=BF)
    ctx = torch.zeros(B * Lq, D, dtype=BF, device="cuda")
    C.xattn1_fwd(v, ctx, B, H, Lq, hd, p, 5, rng)
    ratio = ctx.float().view(B, Lq, H, hd) / v.float().view(B, 1, H, hd)
    mask = ratio[..., 0]                                   # one decision per (b, q, h)
    assert rel_fro(ratio, mask[..., None].expand_as(ratio)) < 1e-2
    if p == 0.0:
        assert torch.equal(ctx.view(B, Lq, D), v[:, None, :].expand(B, Lq, D))
    else:
        keep = (mask > 0).float()
        assert abs(float(keep.mean()) - (1 - p)) < 0.05
        assert rel_fro(mask[mask > 0], torch.full_like(mask[mask > 0], 1 / (1 - p))) < 1e-2
    dctx = rnd(B * Lq, D, seed=2, dtype=BF)
    dv = torch.zeros(B, D, dtype=BF, device="cuda")
    C.xattn1_bwd(dctx, dv, B, H, Lq, hd, p, 5, rng)
    m = (mask > 0).float() / (1 - p)
    ref = (dctx.float().view(B, Lq, H, hd) * m[..., None]).sum(1).reshape(B, D)
    assert rel_fro(dv, ref) < 6e-3


def test_gather_scatter_rows_and_causal_bias(C):
    B, L, D = 5, 20, 768
    src = rnd(B * L, D, seed=1)
    mask = torch.zeros(B, L, dtype=torch.long, device="cuda")
    lens = [20, 1, 7, 0, 13]
    for b, n in enumerate(lens):
        mask[b, :n] = 1
    last = [max(n - 1, 0) for n in lens]
    out_b, out_f = torch.zeros(B, D, dtype=BF, device="cuda"), torch.zeros(B, D, device="cuda")
    C.gather_rows(src, mask, out_b, out_f, B, L, D)
    ref = torch.stack([src.view(B, L, D)[b, last[b]] for b in range(B)])
    assert torch.equal(out_f, ref) and torch.equal(out_b, ref.to(BF))
    C.gather_rows(src, None, None, out_f, B, L, D)
    assert torch.equal(out_f, src.view(B, L, D)[:, 0])
    g = rnd(B, D, seed=2)
    dst = torch.full((B * L, D), 7.0, device="cuda")
    C.scatter_rows(g, mask, dst, B, L, D)
    ref = torch.zeros(B, L, D, device="cuda")
    for b in range(B):
        ref[b, last[b]] = g[b]
    assert torch.equal(dst.view(B, L, D), ref)
    H = 12
    bias = rnd(H, L, L, seed=3)
    keep = bias.clone()
    C.t5_bias_causal(bias, H, L)
    tri = torch.tril(torch.ones(L, L, dtype=torch.bool, device="cuda"))
    assert torch.equal(bias[:, tri], keep[:, tri])
    assert bool((bias[:, ~tri] == torch.finfo(torch.float32).min).all())


def test_relu_dropout_bwd(C):
    n = 64 * 768
    dy, y = rnd(n, seed=1), F.relu(rnd(n, seed=2)).to(BF)
    out = torch.zeros(n, dtype=BF, device="cuda")
    C.relu_dropout_bwd(dy, y, out, 2.0, n)
    assert torch.equal(out, torch.where(y.float() > 0, dy * 2.0, torch.zeros_like(dy)).to(BF))
