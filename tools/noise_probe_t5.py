"""Same experiment as noise_probe.py for the T5 encoder: which bf16 rounding points cost gradient cosine."""
import os, sys, math
import torch, torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import vqa_oracle as O
from tools.noise_probe import rf, rb, ON

def t5(sd, ids, mask, p="lang_model."):
    T5 = O.T5
    B, L = ids.shape
    nh, dk, eps = T5["num_heads"], T5["d_kv"], T5["eps"]
    h = F.embedding(ids, sd[p + "embed_tokens.weight"])
    table = sd[p + "block.0.layer.0.SelfAttention.relative_attention_bias.weight"]
    bias = table[O.t5_buckets(L, L)].permute(2, 0, 1).unsqueeze(0)
    ext = (1.0 - mask[:, None, None, :].float()) * torch.finfo(torch.float32).min
    bias = bias + ext
    for b in range(T5["num_layers"]):
        a = "%sblock.%d.layer.0." % (p, b)
        n = rb(rf(O._rms(h, sd[a + "layer_norm.weight"], eps), "y"), "dy")
        q = rb(rf(F.linear(n, rf(sd[a + "SelfAttention.q.weight"], "w")), "qkv"), "dqkv").view(B, L, nh, dk).transpose(1, 2)
        k = rb(rf(F.linear(n, rf(sd[a + "SelfAttention.k.weight"], "w")), "qkv"), "dqkv").view(B, L, nh, dk).transpose(1, 2)
        v = rb(rf(F.linear(n, rf(sd[a + "SelfAttention.v.weight"], "w")), "qkv"), "dqkv").view(B, L, nh, dk).transpose(1, 2)
        s = torch.matmul(q, k.transpose(3, 2)) + bias
        w = F.softmax(s.float(), dim=-1)
        ctx = rb(rf(torch.matmul(w, v).transpose(1, 2).contiguous().view(B, L, nh * dk), "ctx"), "dctx")
        h = h + rb(F.linear(ctx, rf(sd[a + "SelfAttention.o.weight"], "w")), "g")
        f = "%sblock.%d.layer.1." % (p, b)
        n = rb(rf(O._rms(h, sd[f + "layer_norm.weight"], eps), "y"), "dy")
        hh = rb(rf(F.relu(F.linear(n, rf(sd[f + "DenseReluDense.wi.weight"], "w"))), "h"), "dpre")
        h = h + rb(F.linear(hh, rf(sd[f + "DenseReluDense.wo.weight"], "w")), "g")
    return O._rms(h, sd[p + "final_layer_norm.weight"], eps), h

def run(sd, batch, proj):
    keys = [k for k in sd if k.startswith("lang_model.")]
    work = {k: (v.clone().requires_grad_(True) if k in set(keys) else v) for k, v in sd.items()}
    out, h = t5(work, batch["question_input_ids"], batch["question_attention_masks"])
    (out * proj).sum().backward()
    return {k: work[k].grad for k in keys}, float(h.std())

def main():
    sd = O.random_state_dict("resnet18", 170, seed=0)
    batch = O.synthetic_batch(int(os.environ.get("PB", 2)), 16, 64, 64, 170, seed=1, masked_tail=3)
    g = torch.Generator().manual_seed(5)
    proj = torch.randn(batch["question_input_ids"].shape[0], 16, 768, generator=g) * 1e-3
    ON.clear()
    base, hstd = run(sd, batch, proj)
    print("final hidden std", hstd)
    for spec in sys.argv[1:]:
        ON.clear(); ON.update(spec.split(","))
        gr, _ = run(sd, batch, proj)
        cos = []
        for k in base:
            a, b = gr[k].flatten().double(), base[k].flatten().double()
            if float(b.norm()) > 1e-9:
                cos.append((float(a @ b / (a.norm() * b.norm())), k.replace("lang_model.", "")))
        cos.sort()
        print(spec, "| worst:", ["%.5f %s" % c for c in cos[:4]], "| median %.6f" % cos[len(cos) // 2][0])
main()
