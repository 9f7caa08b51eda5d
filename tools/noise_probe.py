"""Where does the bf16 gradient noise come from?  CPU experiment on the oracle: inject bf16 rounding at named
points of the SGA stack's forward (straight-through) and backward, and report per-tensor gradient cosines against
the unperturbed fp32 oracle.  Test/diagnostic infrastructure only.

    python tools/noise_probe.py r50 fwd:kv,bwd:dvk
"""
import math
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import vqa_oracle as O  # noqa: E402

ON = set()


class _RoundF(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.bfloat16().float()

    @staticmethod
    def backward(ctx, g):
        return g


class _RoundB(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return g.bfloat16().float()


def rf(x, name):
    return _RoundF.apply(x) if ("fwd:" + name) in ON or "fwd:all" in ON else x


def rb(x, name):
    return _RoundB.apply(x) if ("bwd:" + name) in ON or "bwd:all" in ON else x


def lin(x, w, b, tag):
    x = rb(rf(x, tag + "_in"), tag + "_din")
    w = rf(w, "w")
    return F.linear(x, w, b)


def mhatt(sd, p, v, k, q, tag):
    B, H, hd = q.shape[0], 8, 96
    v = rb(rf(lin(v, sd[p + "linear_v.weight"], sd[p + "linear_v.bias"], tag + "v"), "kv"), "dvk").view(B, -1, H, hd).transpose(1, 2)
    k = rb(rf(lin(k, sd[p + "linear_k.weight"], sd[p + "linear_k.bias"], tag + "k"), "kv"), "dvk").view(B, -1, H, hd).transpose(1, 2)
    q = rb(rf(lin(q, sd[p + "linear_q.weight"], sd[p + "linear_q.bias"], tag + "q"), "q"), "dq").view(B, -1, H, hd).transpose(1, 2)
    s = torch.matmul(q, k.transpose(-2, -1)) / math.sqrt(hd)
    a = torch.matmul(F.softmax(s, dim=-1), v).transpose(1, 2).contiguous().view(B, -1, H * hd)
    a = rb(rf(a, "ctx"), "dctx")
    return lin(a, sd[p + "linear_merge.weight"], sd[p + "linear_merge.bias"], tag + "m")


def sga(sd, p, x, y):
    def ln(k, t):
        return F.layer_norm(t, (768,), sd[p + k + ".norm.weight"], sd[p + k + ".norm.bias"], 1e-5)
    x = ln("norm1", x + rb(mhatt(sd, p + "mhatt1.", x, x, x, "m1"), "g"))
    x = ln("norm2", x + rb(mhatt(sd, p + "mhatt2.", y, y, x, "m2"), "g"))
    h = rb(rf(F.relu(lin(x, sd[p + "ffn.mlp.fc1.weight"], sd[p + "ffn.mlp.fc1.bias"], "fc1")), "hm"), "dpre")
    f = lin(h, sd[p + "ffn.mlp.fc2.weight"], sd[p + "ffn.mlp.fc2.bias"], "fc2")
    return ln("norm3", x + rb(f, "g"))


def run(sd, vision, batch, feat, text):
    keys = [k for k in O.trainable_keys(sd, vision) if not k.startswith("lang_model.")]
    work = {k: (v.clone().requires_grad_(True) if k in set(keys) else v) for k, v in sd.items()}
    proj = "downscale_layer" if vision == "resnet50" else "upscale_layer"
    ve = F.conv_transpose2d(rf(feat, "feat"), rf(work[proj + ".weight"], "w"), work[proj + ".bias"], 1, 1)
    y = rb(rf(ve.view(ve.shape[0], ve.shape[1], -1).permute(0, 2, 1), "y0"), "dy0")
    fused = None
    for l in range(3):
        fused = sga(work, "sga_modules.%d." % l, text, y)
        y = rf(fused, "yl")
    w = F.softmax(F.linear(fused, work["attention_pooler.attention.0.weight"], work["attention_pooler.attention.0.bias"]), dim=1).transpose(1, 2)
    pooled = torch.bmm(w, fused).squeeze(1)
    logp = F.log_softmax(lin(pooled, work["classification_layer.weight"], work["classification_layer.bias"], "cls"), -1)
    loss = F.nll_loss(logp, batch["annotation_ids"])
    loss.backward()
    return {k: work[k].grad for k in keys}


def main():
    vision = {"r50": "resnet50", "r34": "resnet34"}[sys.argv[1]]
    sd = O.random_state_dict(vision, 170, seed=0)
    batch = O.synthetic_batch(int(os.environ.get("PB", 2)), 32, 224, 224, 170, seed=1, masked_tail=int(os.environ.get("PMASK", 10)))
    with torch.no_grad():
        feat = O.resnet_body(sd, batch["image_tensors"], vision)
        text = O.t5_encoder(sd, batch["question_input_ids"], batch["question_attention_masks"])
    ON.clear()
    base = run(sd, vision, batch, feat, text)
    for spec in sys.argv[2:]:
        ON.clear()
        ON.update(spec.split(","))
        g = run(sd, vision, batch, feat, text)
        cos = []
        for k in base:
            a, b = g[k].flatten().double(), base[k].flatten().double()
            if float(b.norm()) > 1e-7:
                cos.append((float(a @ b / (a.norm() * b.norm())), k))
        cos.sort()
        print(spec, "| worst:", ["%.5f %s" % c for c in cos[:4]], "| median %.6f" % cos[len(cos) // 2][0])


if __name__ == "__main__":
    main()
