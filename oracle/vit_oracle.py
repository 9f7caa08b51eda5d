"""ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.

CPU fp32 restatement of `VitVQAModel.forward` (model/vit_vqa_model.py:127-227, SURVEY.md 8f-4 / BASELINE config 5) as
plain functional PyTorch over a reference-layout state_dict, so that autograd gives the reference gradients:

    frozen ViT-B/16 (transformers ViTModel, `vit:` = models/vit/modeling_vit.py) -> pooler_output        (:184-186, no_grad)
    T5 encoder over the question, token 0                                                                  (:189-195)
    Linear(1536 -> 768) + ReLU + Dropout(0.5) over [pooled | token 0]                                      (:198-203)
    T5 decoder over decoder_question_input_ids, cross-attending to that ONE fused token                    (:207-212)
    gather of the last un-padded decoder position, Linear(768 -> answers), log_softmax, NLLLoss            (:215-227)

Pinned like oracle/vqa_oracle.py: oracle/make_golden.py imports the UNMODIFIED reference class, loads the deterministic
state_dict of `random_state_dict` below and freezes its outputs under tests/golden/vit_*.pt; tests/test_oracle.py checks
this restatement against those files on every run.  `hf:` = transformers/models/t5/modeling_t5.py.
"""
import math

import torch
import torch.nn.functional as F

from . import vqa_oracle as O

T5 = O.T5
VIT = dict(hidden=768, heads=12, layers=12, inter=3072, patch=16, image=224, eps=1e-12)


# --------------------------------------------------------------------------------------------------
# deterministic random state_dict in the reference's key layout
# --------------------------------------------------------------------------------------------------
def state_dict_spec(answer_spaces=170):
    """Ordered (key, shape, kind) of every entry of VitVQAModel.state_dict().  The four names of T5's tied token table
    (shared / encoder.embed_tokens / decoder.embed_tokens / lm_head) are one tensor: kind 'tied'."""
    spec = []
    d, dff = VIT["hidden"], VIT["inter"]
    v = "vision_model."
    spec.append((v + "embeddings.cls_token", (1, 1, d), "vit_tok"))
    spec.append((v + "embeddings.position_embeddings", (1, (VIT["image"] // VIT["patch"]) ** 2 + 1, d), "vit_tok"))
    spec.append((v + "embeddings.patch_embeddings.projection.weight", (d, 3, VIT["patch"], VIT["patch"]), "vit_w"))
    spec.append((v + "embeddings.patch_embeddings.projection.bias", (d,), "vit_b"))
    for i in range(VIT["layers"]):
        p = "%sencoder.layer.%d." % (v, i)
        for nm in ("query", "key", "value"):
            spec.append((p + "attention.attention.%s.weight" % nm, (d, d), "vit_w"))
            spec.append((p + "attention.attention.%s.bias" % nm, (d,), "vit_b"))
        spec.append((p + "attention.output.dense.weight", (d, d), "vit_w"))
        spec.append((p + "attention.output.dense.bias", (d,), "vit_b"))
        spec.append((p + "intermediate.dense.weight", (dff, d), "vit_w"))
        spec.append((p + "intermediate.dense.bias", (dff,), "vit_b"))
        spec.append((p + "output.dense.weight", (d, dff), "vit_w"))
        spec.append((p + "output.dense.bias", (d,), "vit_b"))
        for nm in ("layernorm_before", "layernorm_after"):
            spec.append((p + nm + ".weight", (d,), "norm_w"))
            spec.append((p + nm + ".bias", (d,), "norm_b"))
    spec.append((v + "layernorm.weight", (d,), "norm_w"))
    spec.append((v + "layernorm.bias", (d,), "norm_b"))
    spec.append((v + "pooler.dense.weight", (d, d), "vit_w"))
    spec.append((v + "pooler.dense.bias", (d,), "vit_b"))

    t = "lang_model."
    dm, inner, tff = T5["d_model"], T5["d_kv"] * T5["num_heads"], T5["d_ff"]
    spec.append((t + "shared.weight", (T5["vocab"], dm), "embed"))

    def attn(p, bias_table):
        spec.append((p + "q.weight", (inner, dm), "t5_q"))
        spec.append((p + "k.weight", (inner, dm), "t5_kv"))
        spec.append((p + "v.weight", (inner, dm), "t5_kv"))
        spec.append((p + "o.weight", (dm, inner), "t5_o"))
        if bias_table:
            spec.append((p + "relative_attention_bias.weight", (T5["num_buckets"], T5["num_heads"]), "t5_bias"))

    for stack in ("encoder", "decoder"):
        s = t + stack + "."
        spec.append((s + "embed_tokens.weight", (T5["vocab"], dm), "tied"))
        for b in range(T5["num_layers"]):
            p = "%sblock.%d.layer." % (s, b)
            attn(p + "0.SelfAttention.", b == 0)
            spec.append((p + "0.layer_norm.weight", (dm,), "norm_w"))
            ff = 1
            if stack == "decoder":
                attn(p + "1.EncDecAttention.", False)
                spec.append((p + "1.layer_norm.weight", (dm,), "norm_w"))
                ff = 2
            spec.append(("%s%d.DenseReluDense.wi.weight" % (p, ff), (tff, dm), "t5_wi"))
            spec.append(("%s%d.DenseReluDense.wo.weight" % (p, ff), (dm, tff), "t5_wo"))
            spec.append(("%s%d.layer_norm.weight" % (p, ff), (dm,), "norm_w"))
        spec.append((s + "final_layer_norm.weight", (dm,), "norm_w"))
    spec.append((t + "lm_head.weight", (T5["vocab"], dm), "tied"))
    spec.append(("fusing_layer.0.weight", (768, 1536), "linear"))
    spec.append(("fusing_layer.0.bias", (768,), "bias1536"))
    spec.append(("classification_layer.weight", (answer_spaces, 768), "linear"))
    spec.append(("classification_layer.bias", (answer_spaces,), "bias"))
    return spec


def random_state_dict(answer_spaces=170, seed=0):
    g = torch.Generator().manual_seed(seed)
    sd = {}
    d = T5["d_model"]
    for key, shape, kind in state_dict_spec(answer_spaces):
        if kind == "tied":
            t = sd["lang_model.shared.weight"]
        elif kind == "vit_w":
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            t = torch.randn(shape, generator=g) * (1.0 / math.sqrt(fan_in))
        elif kind == "vit_b":
            t = 0.1 * torch.randn(shape, generator=g)
        elif kind == "vit_tok":
            t = 0.5 * torch.randn(shape, generator=g)
        elif kind == "norm_w":
            t = 0.75 + 0.5 * torch.rand(shape, generator=g)
        elif kind == "norm_b":
            t = 0.1 * torch.randn(shape, generator=g)
        elif kind in ("linear", "bias", "bias1536"):
            fan_in = shape[1] if len(shape) > 1 else (1536 if kind == "bias1536" else 768)
            t = (torch.rand(shape, generator=g) * 2 - 1) / math.sqrt(fan_in)
        elif kind == "embed":
            t = torch.randn(shape, generator=g)
        elif kind == "t5_q":
            t = torch.randn(shape, generator=g) * (d * T5["d_kv"]) ** -0.5
        elif kind in ("t5_kv", "t5_wi", "t5_o"):
            t = torch.randn(shape, generator=g) * d ** -0.5
        elif kind == "t5_wo":
            t = torch.randn(shape, generator=g) * T5["d_ff"] ** -0.5
        elif kind == "t5_bias":
            t = 0.5 * torch.randn(shape, generator=g)
        else:
            raise KeyError(kind)
        sd[key] = t
    return sd


def synthetic_batch(B, L=32, Ld=20, answers=170, seed=1, masked_tail=0, dec_lengths=True):
    """pixel_values in [-1, 1) (ViTImageProcessor: rescale 1/255, mean 0.5, std 0.5); question ids; decoder question ids
    padded to Enums.MAX_LEN = 20 (dataset_utils/vit_vqa_daquar_dataset.py:165-166) with per-sample lengths."""
    g = torch.Generator().manual_seed(seed)
    batch = dict(
        pixel_values=torch.rand(B, 3, VIT["image"], VIT["image"], generator=g) * 2 - 1,
        question_input_ids=torch.randint(2, 32100, (B, L), generator=g),
        question_attention_masks=torch.ones(B, L, dtype=torch.long),
        decoder_question_input_ids=torch.randint(2, 32100, (B, Ld), generator=g),
        decoder_question_attention_masks=torch.ones(B, Ld, dtype=torch.long),
        annotation_ids=torch.randint(0, answers, (B,), generator=g),
    )
    if masked_tail:
        batch["question_attention_masks"][:, L - masked_tail:] = 0
    if dec_lengths:
        lens = torch.randint(3, Ld + 1, (B,), generator=g)
        for b in range(B):
            batch["decoder_question_attention_masks"][b, int(lens[b]):] = 0
            batch["decoder_question_input_ids"][b, int(lens[b]):] = 0      # the tokenizer's pad id
    return batch


# --------------------------------------------------------------------------------------------------
# forward restatement
# --------------------------------------------------------------------------------------------------
def vit_pooled(sd, pixel_values, p="vision_model.", return_hidden=False, return_attn=False):
    """ViTModel(pixel_values).pooler_output in eval mode (vit: ViTEmbeddings, ViTLayer (pre-LN, exact GELU), ViTPooler)."""
    d, nh = VIT["hidden"], VIT["heads"]
    hd = d // nh
    x = F.conv2d(pixel_values, sd[p + "embeddings.patch_embeddings.projection.weight"],
                 sd[p + "embeddings.patch_embeddings.projection.bias"], stride=VIT["patch"])
    x = x.flatten(2).transpose(1, 2)
    B = x.shape[0]
    x = torch.cat([sd[p + "embeddings.cls_token"].expand(B, -1, -1), x], dim=1) + sd[p + "embeddings.position_embeddings"]
    T = x.shape[1]
    attn = []
    for i in range(VIT["layers"]):
        l = "%sencoder.layer.%d." % (p, i)
        n = F.layer_norm(x, (d,), sd[l + "layernorm_before.weight"], sd[l + "layernorm_before.bias"], VIT["eps"])
        a = l + "attention.attention."
        q = F.linear(n, sd[a + "query.weight"], sd[a + "query.bias"]).view(B, T, nh, hd).transpose(1, 2)
        k = F.linear(n, sd[a + "key.weight"], sd[a + "key.bias"]).view(B, T, nh, hd).transpose(1, 2)
        v = F.linear(n, sd[a + "value.weight"], sd[a + "value.bias"]).view(B, T, nh, hd).transpose(1, 2)
        w = F.softmax(torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(hd), dim=-1)
        attn.append(w)
        ctx = torch.matmul(w, v).transpose(1, 2).reshape(B, T, d)
        x = x + F.linear(ctx, sd[l + "attention.output.dense.weight"], sd[l + "attention.output.dense.bias"])
        n = F.layer_norm(x, (d,), sd[l + "layernorm_after.weight"], sd[l + "layernorm_after.bias"], VIT["eps"])
        h = F.gelu(F.linear(n, sd[l + "intermediate.dense.weight"], sd[l + "intermediate.dense.bias"]))
        x = x + F.linear(h, sd[l + "output.dense.weight"], sd[l + "output.dense.bias"])
    x = F.layer_norm(x, (d,), sd[p + "layernorm.weight"], sd[p + "layernorm.bias"], VIT["eps"])
    pooled = torch.tanh(F.linear(x[:, 0], sd[p + "pooler.dense.weight"], sd[p + "pooler.dense.bias"]))
    if return_attn:          # ViTModel(..., output_attentions=True).attentions (model/vit_vqa_model.py:240-243)
        return pooled, tuple(attn)
    return (pooled, x) if return_hidden else pooled


def t5_buckets_causal(Lq, Lk, num_buckets=32, max_distance=128):
    """hf:189-234 with bidirectional=False (decoder self-attention): only the distance into the past counts."""
    rel = torch.arange(Lk)[None, :] - torch.arange(Lq)[:, None]
    rel = -torch.min(rel, torch.zeros_like(rel))
    max_exact = num_buckets // 2
    large = max_exact + (torch.log(rel.float() / max_exact) / math.log(max_distance / max_exact)
                         * (num_buckets - max_exact)).long()
    large = torch.min(large, torch.full_like(large, num_buckets - 1))
    return torch.where(rel < max_exact, rel, large)


def t5_decoder(sd, ids, mask, enc, p="lang_model.decoder."):
    """T5Stack decoder in eval mode (hf:637-792 with is_decoder): causal self-attention with its own relative-position table
    (block 0, unidirectional buckets) and the padding mask as additive finfo.min, cross-attention onto `enc` [B, S, 768]
    with a zero position bias, ReLU FFN.  `embed_tokens` is the shared table."""
    B, L = ids.shape
    nh, dk, eps = T5["num_heads"], T5["d_kv"], T5["eps"]
    h = F.embedding(ids, sd[p + "embed_tokens.weight"])
    table = sd[p + "block.0.layer.0.SelfAttention.relative_attention_bias.weight"]
    bias = table[t5_buckets_causal(L, L)].permute(2, 0, 1).unsqueeze(0)            # [1, nh, L, L]
    allowed = torch.tril(torch.ones(L, L, dtype=torch.bool))[None, None]
    if mask is not None:
        allowed = allowed & mask[:, None, None, :].bool()
    bias = bias + (~allowed).float() * torch.finfo(torch.float32).min

    def heads(t):
        return t.view(B, -1, nh, dk).transpose(1, 2)

    for b in range(T5["num_layers"]):
        a = "%sblock.%d.layer.0." % (p, b)
        n = O._rms(h, sd[a + "layer_norm.weight"], eps)
        q, k, v = (heads(F.linear(n, sd[a + "SelfAttention.%s.weight" % nm])) for nm in "qkv")
        w = F.softmax((torch.matmul(q, k.transpose(3, 2)) + bias).float(), dim=-1)
        ctx = torch.matmul(w, v).transpose(1, 2).contiguous().view(B, L, nh * dk)
        h = h + F.linear(ctx, sd[a + "SelfAttention.o.weight"])
        c = "%sblock.%d.layer.1." % (p, b)
        n = O._rms(h, sd[c + "layer_norm.weight"], eps)
        q = heads(F.linear(n, sd[c + "EncDecAttention.q.weight"]))
        k = heads(F.linear(enc, sd[c + "EncDecAttention.k.weight"]))
        v = heads(F.linear(enc, sd[c + "EncDecAttention.v.weight"]))
        w = F.softmax(torch.matmul(q, k.transpose(3, 2)).float(), dim=-1)
        ctx = torch.matmul(w, v).transpose(1, 2).contiguous().view(B, L, nh * dk)
        h = h + F.linear(ctx, sd[c + "EncDecAttention.o.weight"])
        f = "%sblock.%d.layer.2." % (p, b)
        n = O._rms(h, sd[f + "layer_norm.weight"], eps)
        h = h + F.linear(F.relu(F.linear(n, sd[f + "DenseReluDense.wi.weight"])), sd[f + "DenseReluDense.wo.weight"])
    return O._rms(h, sd[p + "final_layer_norm.weight"], eps)


def forward(sd, question_input_ids, decoder_question_input_ids, question_attention_masks,
            decoder_question_attention_masks, annotation_ids, pixel_values):
    """VitVQAModel.forward in eval mode with grad enabled (model/vit_vqa_model.py:168-227)."""
    with torch.no_grad():
        pooled = vit_pooled(sd, pixel_values.float())
    enc = O.t5_encoder(sd, question_input_ids, question_attention_masks, p="lang_model.encoder.")
    cat = torch.cat([pooled, enc[:, 0, :]], dim=1)
    fused = F.relu(F.linear(cat, sd["fusing_layer.0.weight"], sd["fusing_layer.0.bias"]))
    dec = t5_decoder(sd, decoder_question_input_ids, decoder_question_attention_masks, fused.unsqueeze(1))
    m = decoder_question_attention_masks
    last = torch.where(m == 1, torch.arange(m.shape[1])[None, :].expand_as(m), torch.zeros_like(m)).max(dim=1).values
    ans = dec[torch.arange(dec.shape[0]), last]
    logp = F.log_softmax(F.linear(ans, sd["classification_layer.weight"], sd["classification_layer.bias"]), -1)
    loss = F.nll_loss(logp, annotation_ids) if annotation_ids is not None else None
    return logp, loss


TIED = ("lang_model.encoder.embed_tokens.weight", "lang_model.decoder.embed_tokens.weight", "lang_model.lm_head.weight")


def trainable_keys(sd):
    """named_parameters() of the reference that receive a gradient: everything outside the frozen ViT; the tied token table
    appears once, as lang_model.shared.weight."""
    return [k for k, v in sd.items() if not k.startswith("vision_model.") and k not in TIED and v.is_floating_point()]


def forward_backward(sd, batch):
    keys = trainable_keys(sd)
    ks = set(keys)
    work = {k: (v.clone().requires_grad_(True) if k in ks else v) for k, v in sd.items()}
    for k in TIED:
        work[k] = work["lang_model.shared.weight"]
    logp, loss = forward(work, batch["question_input_ids"], batch["decoder_question_input_ids"],
                         batch["question_attention_masks"], batch["decoder_question_attention_masks"],
                         batch["annotation_ids"], batch["pixel_values"])
    loss.backward()
    grads = {k: (work[k].grad if work[k].grad is not None else torch.zeros_like(work[k])) for k in keys}
    return logp.detach(), loss.detach(), grads
