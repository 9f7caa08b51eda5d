"""Which operand precision does the north-star gradient-cosine bar (>= 0.999 on EVERY tensor) need?

CPU experiment on the oracle (test / diagnostic infrastructure only): every F.linear / matmul / conv_transpose2d operand of
the oracle forward is rounded (straight-through, so autograd still flows) to one of
    1 = bf16                      (what a plain bf16 tcgen05 GEMM sees)
    2 = bf16 hi + bf16 lo         (two-term split: ~16 mantissa bits; x_hi*W_hi + x_hi*W_lo [+ x_lo*W_hi] on the GPU)
    0 = fp32 (untouched)
chosen per operand class by rules on the weight's state_dict key, and the per-tensor gradient cosines against the
unperturbed fp32 oracle are reported.

    python tools/precision_probe.py r34 4 "w=1,x=1" "t5.w=2,w=1,x=1" ...
rule syntax: `;`-separated `[prefix.]{w|x}=mode` or `re:<regex>/{w|x}=mode`, first match wins; operand classes: t5.b<block>.<q|k|v|o|wi|wo>,
sga, proj, cls, mm (attention matmuls).
"""
import os
import re
import sys
import types

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import vqa_oracle as O  # noqa: E402


class _Round(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, mode):
        hi = x.bfloat16().float()
        if mode == 1:
            return hi
        return hi + (x - hi).bfloat16().float()

    @staticmethod
    def backward(ctx, g):
        return g, None


def rnd(x, mode):
    return x if mode == 0 else _Round.apply(x, mode)


RULES = []
KEYS = {}


def mode_for(cls, operand):
    for prefix, op, mode in RULES:
        if op != operand:
            continue
        if prefix.startswith("re:"):
            if re.fullmatch(prefix[3:], cls):
                return mode
        elif prefix == "" or cls == prefix or cls.startswith(prefix + "."):
            return mode
    return 0


def classify(w):
    k = KEYS.get(id(w), "")
    if k.startswith("lang_model.block."):
        return "t5.b%s.%s" % (k.split(".")[2], k.split(".")[-2])
    if k.startswith("lang_model."):
        return "t5"
    if k.startswith("sga_modules."):
        return "sga"
    if k.startswith("classification_layer") or k.startswith("attention_pooler"):
        return "cls"
    if "scale_layer" in k:
        return "proj"
    return "other"


def install():
    f = types.SimpleNamespace(**{n: getattr(F, n) for n in dir(F) if not n.startswith("__")})

    def linear(x, w, b=None):
        c = classify(w)
        return F.linear(rnd(x, mode_for(c, "x")), rnd(w, mode_for(c, "w")), b)

    def conv_transpose2d(x, w, b, s, p):
        return F.conv_transpose2d(rnd(x, mode_for("proj", "x")), rnd(w, mode_for("proj", "w")), b, s, p)
    f.linear, f.conv_transpose2d = linear, conv_transpose2d
    O.F = f
    t = types.SimpleNamespace(**{n: getattr(torch, n) for n in dir(torch) if not n.startswith("__")})

    def matmul(a, b):
        m = mode_for("mm", "x")
        return torch.matmul(rnd(a, m), rnd(b, m))
    t.matmul = matmul
    O.torch = t


def run(sd, vision, batch):
    keys = O.trainable_keys(sd, vision)
    ks = set(keys)
    work = {k: (v.clone().requires_grad_(True) if k in ks else v) for k, v in sd.items()}
    KEYS.clear()
    KEYS.update({id(v): k for k, v in work.items()})
    logp, loss = O.forward(work, vision, batch["question_input_ids"], batch["question_attention_masks"],
                           batch["annotation_ids"], batch["image_tensors"])
    loss.backward()
    return logp.detach(), float(loss), {k: work[k].grad for k in keys}


def main():
    vision = {"r50": "resnet50", "r34": "resnet34", "r18": "resnet18"}[sys.argv[1]]
    B = int(sys.argv[2])
    torch.set_num_threads(os.cpu_count())
    sd = O.random_state_dict(vision, 170, seed=0)
    batch = O.synthetic_batch(B, int(os.environ.get("PL", 32)), int(os.environ.get("PH", 224)), int(os.environ.get("PH", 224)), 170,
                              seed=int(os.environ.get("PSEED", 1)),
                              masked_tail=int(os.environ.get("PMASK", 0)))
    install()
    RULES.clear()
    logp0, loss0, base = run(sd, vision, batch)
    for spec in sys.argv[3:]:
        RULES.clear()
        for item in spec.split(";"):
            lhs, mode = item.split("=")
            prefix, _, op = lhs.rpartition("/" if "/" in lhs else ".")
            RULES.append((prefix, op, int(mode)))
        logp, loss, g = run(sd, vision, batch)
        cos = []
        for k in base:
            a, b = g[k].flatten().double(), base[k].flatten().double()
            if float(b.norm()) > 1e-7:
                cos.append((float(a @ b / (a.norm() * b.norm())), k.replace("lang_model.", "")))
        cos.sort()
        top1 = float((logp.argmax(1) == logp0.argmax(1)).float().mean())
        below = sum(1 for c, _ in cos if c < 0.999)
        print("%-40s | below 0.999: %d/%d | worst: %s | median %.6f | logp rel %.2e loss rel %.2e top1 %.4f" % (
            spec, below, len(cos), ["%.5f %s" % c for c in cos[:3]], cos[len(cos) // 2][0],
            float((logp - logp0).norm() / logp0.norm()), abs(loss - loss0) / abs(loss0), top1), flush=True)


def top1_main():
    """python tools/precision_probe.py top1 r34 512 <spec>...: forward-only top-1 agreement over many samples."""
    vision = {"r50": "resnet50", "r34": "resnet34", "r18": "resnet18"}[sys.argv[2]]
    n = int(sys.argv[3])
    torch.set_num_threads(os.cpu_count())
    sd = O.random_state_dict(vision, 170, seed=0)
    install()
    KEYS.clear()
    KEYS.update({id(v): k for k, v in sd.items()})
    res = {}
    for chunk in range(n // 64):
        batch = O.synthetic_batch(64, 32, 224, 224, 170, seed=100 + chunk)
        outs = {}
        for spec in ["base"] + sys.argv[4:]:
            RULES.clear()
            if spec != "base":
                for item in spec.split(";"):
                    lhs, mode = item.split("=")
                    prefix, _, op = lhs.rpartition("/" if "/" in lhs else ".")
                    RULES.append((prefix, op, int(mode)))
            with torch.no_grad():
                logp, _ = O.forward(sd, vision, batch["question_input_ids"], batch["question_attention_masks"],
                                    batch["annotation_ids"], batch["image_tensors"])
            outs[spec] = logp
        for spec in sys.argv[4:]:
            same = int((outs[spec].argmax(1) == outs["base"].argmax(1)).sum())
            res[spec] = res.get(spec, 0) + same
        top2 = outs["base"].topk(2, dim=1).values
        print(chunk, {k[:30]: v for k, v in res.items()}, "min margin %.2e" % float((top2[:, 0] - top2[:, 1]).min()), flush=True)
    for spec, v in res.items():
        print("%s: top-1 agreement %d/%d = %.4f" % (spec, v, n, v / n))


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "top1":
    top1_main()
    sys.exit(0)


if __name__ == "__main__":
    main()
