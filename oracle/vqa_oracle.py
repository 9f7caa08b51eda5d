"""ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.

A CPU fp32 restatement of the reference's hot path (ResnetVQAModel forward, model/resnet_vqa_model.py:101-165)
written as plain functional PyTorch over a reference-layout state_dict, so that autograd gives the reference
gradients as well.  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import it; the
package under t5-resnet-vqa_b200/ never does.

Pinning: the reference ships no tests, fixtures or golden vectors, so this file is
pinned against the UNMODIFIED reference itself: oracle/make_golden.py imports /root/reference/model/
resnet_vqa_model.py in the build container, loads the same deterministic state_dict into it, and freezes its
outputs (log-probs, loss, per-tensor gradient norms, a few full gradients) under tests/golden/.
tests/test_oracle.py then checks this restatement against those files on every run.

`tv:` = torchvision/models/resnet.py, `hf:` = transformers/models/t5/modeling_t5.py (third-party files the
reference assembles its model from; pinned torchvision==0.16.0 / transformers==4.34.1 in requirements.txt:46,48,
installed 0.26.0 / 5.5.0 — same arithmetic).
"""
import math

import torch
import torch.nn.functional as F

RESNET_CFG = {"resnet18": ("basic", [2, 2, 2, 2]), "resnet34": ("basic", [3, 4, 6, 3]),
              "resnet50": ("bottleneck", [3, 4, 6, 3])}
T5 = dict(vocab=32128, d_model=768, d_kv=64, d_ff=3072, num_layers=12, num_heads=12, num_buckets=32,
          max_distance=128, eps=1e-6)


# --------------------------------------------------------------------------------------------------
# deterministic random state_dict in the reference's key layout
# --------------------------------------------------------------------------------------------------
def state_dict_spec(vision_name, answer_spaces=170, num_attention_blocks=3):
    """Ordered list of (key, shape, kind) for every entry of ResnetVQAModel.state_dict()."""
    spec = []

    def conv(k, o, i, r):
        spec.append((k + ".weight", (o, i, r, r), "conv"))

    frcnn = vision_name == "faster-rcnn"

    def bn(k, c, last=False):
        spec.extend([(k + ".weight", (c,), "bn_w_last" if last else "bn_w"), (k + ".bias", (c,), "bn_b"),
                     (k + ".running_mean", (c,), "bn_m"), (k + ".running_var", (c,), "bn_v")])
        if not frcnn:      # torchvision FrozenBatchNorm2d (the detector's backbone) keeps no batch counter
            spec.append((k + ".num_batches_tracked", (), "count"))

    def linear(k, o, i, bias=True, kind="linear"):
        spec.append((k + ".weight", (o, i), kind))
        if bias:
            spec.append((k + ".bias", (o,), "bias"))

    kind, layers = RESNET_CFG["resnet50" if frcnn else vision_name]
    exp = 4 if kind == "bottleneck" else 1
    v = "vision_model.body." if frcnn else "vision_model."
    conv(v + "conv1", 64, 3, 7); bn(v + "bn1", 64)
    inpl = 64
    for li, (planes, n) in enumerate(zip([64, 128, 256, 512], layers)):
        for bi in range(n):
            stride = 2 if (bi == 0 and li > 0) else 1
            p = "%slayer%d.%d." % (v, li + 1, bi)
            if kind == "bottleneck":
                conv(p + "conv1", planes, inpl, 1); bn(p + "bn1", planes)
                conv(p + "conv2", planes, planes, 3); bn(p + "bn2", planes)
                conv(p + "conv3", planes * 4, planes, 1); bn(p + "bn3", planes * 4, True)
            else:
                conv(p + "conv1", planes, inpl, 3); bn(p + "bn1", planes)
                conv(p + "conv2", planes, planes, 3); bn(p + "bn2", planes, True)
            if bi == 0 and (stride != 1 or inpl != planes * exp):
                conv(p + "downsample.0", planes * exp, inpl, 1); bn(p + "downsample.1", planes * exp)
            inpl = planes * exp
    if frcnn:
        # FeaturePyramidNetwork([256, 512, 1024, 2048], 256): lateral 1x1 and output 3x3 convs with bias (key '.0.':
        # Conv2dNormActivation without norm / activation)
        for i, cin in enumerate([256, 512, 1024, 2048]):
            spec.append(("vision_model.fpn.inner_blocks.%d.0.weight" % i, (256, cin, 1, 1), "fpn_conv"))
            spec.append(("vision_model.fpn.inner_blocks.%d.0.bias" % i, (256,), "bn_b"))
        for i in range(4):
            spec.append(("vision_model.fpn.layer_blocks.%d.0.weight" % i, (256, 256, 3, 3), "fpn_conv"))
            spec.append(("vision_model.fpn.layer_blocks.%d.0.bias" % i, (256,), "bn_b"))
    else:
        linear(v + "fc", 1000, 512 * exp)
    t = "lang_model."
    d, inner, dff = T5["d_model"], T5["d_kv"] * T5["num_heads"], T5["d_ff"]
    spec.append((t + "embed_tokens.weight", (T5["vocab"], d), "embed"))
    for b in range(T5["num_layers"]):
        p = "%sblock.%d.layer.0.SelfAttention." % (t, b)
        spec.append((p + "q.weight", (inner, d), "t5_q"))
        spec.append((p + "k.weight", (inner, d), "t5_kv"))
        spec.append((p + "v.weight", (inner, d), "t5_kv"))
        spec.append((p + "o.weight", (d, inner), "t5_o"))
        if b == 0:
            spec.append((p + "relative_attention_bias.weight", (T5["num_buckets"], T5["num_heads"]), "t5_bias"))
        spec.append(("%sblock.%d.layer.0.layer_norm.weight" % (t, b), (d,), "norm_w"))
        spec.append(("%sblock.%d.layer.1.DenseReluDense.wi.weight" % (t, b), (dff, d), "t5_wi"))
        spec.append(("%sblock.%d.layer.1.DenseReluDense.wo.weight" % (t, b), (d, dff), "t5_wo"))
        spec.append(("%sblock.%d.layer.1.layer_norm.weight" % (t, b), (d,), "norm_w"))
    spec.append((t + "final_layer_norm.weight", (d,), "norm_w"))
    for name, cin in ((("upscale_layer", 256),) if frcnn else (("upscale_layer", 512), ("downscale_layer", 2048))):
        spec.append((name + ".weight", (cin, 768, 3, 3), "convT"))
        spec.append((name + ".bias", (768,), "bias"))
    for l in range(num_attention_blocks):
        for mh in ("mhatt1", "mhatt2"):
            for ln in ("linear_v", "linear_k", "linear_q", "linear_merge"):
                linear("sga_modules.%d.%s.%s" % (l, mh, ln), 768, 768)
        linear("sga_modules.%d.ffn.mlp.fc1" % l, 768, 768)
        linear("sga_modules.%d.ffn.mlp.fc2" % l, 768, 768)
        for nm in ("norm1", "norm2", "norm3"):
            spec.append(("sga_modules.%d.%s.norm.weight" % (l, nm), (768,), "norm_w"))
            spec.append(("sga_modules.%d.%s.norm.bias" % (l, nm), (768,), "norm_b"))
    linear("classification_layer", answer_spaces, 768)
    linear("attention_pooler.attention.0", 1, 768)
    return spec


def random_state_dict(vision_name, answer_spaces=170, seed=0, num_attention_blocks=3):
    """Deterministic (CPU generator) weights with every component switched 'on': BatchNorm statistics, norm
    gains and the relative-position table are non-trivial so folding / bias / mask bugs are observable."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    d = T5["d_model"]
    for key, shape, kind in state_dict_spec(vision_name, answer_spaces, num_attention_blocks):
        if kind == "count":
            sd[key] = torch.tensor(0, dtype=torch.long)
            continue
        if kind == "conv":
            std = math.sqrt(2.0 / (shape[0] * shape[2] * shape[3]))
            t = torch.randn(shape, generator=g) * std
        elif kind == "fpn_conv":
            bound = math.sqrt(3.0 / (shape[1] * shape[2] * shape[3]))     # kaiming_uniform_(a=1)
            t = (torch.rand(shape, generator=g) * 2 - 1) * bound
        elif kind in ("bn_w", "bn_v"):
            t = 0.5 + torch.rand(shape, generator=g)
        elif kind == "bn_w_last":
            # gain of the last BatchNorm of every residual block: small, as in trained networks, so the 16
            # residual additions keep the feature map O(1) (with O(1) gains it reaches ~1e2-1e3, the guided
            # attention over the vision tokens saturates to one-hot and its gradients become pure rounding noise)
            t = 0.4 * (0.5 + torch.rand(shape, generator=g))
        elif kind in ("bn_b", "bn_m", "norm_b"):
            t = 0.1 * torch.randn(shape, generator=g)
        elif kind == "norm_w":
            t = 0.75 + 0.5 * torch.rand(shape, generator=g)
        elif kind in ("linear", "bias", "convT"):
            fan_in = shape[1] * 9 if kind == "convT" else (shape[1] if len(shape) > 1 else 768)
            bound = 1.0 / math.sqrt(fan_in)
            t = (torch.rand(shape, generator=g) * 2 - 1) * bound
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


def synthetic_batch(B, L=32, H=224, W=224, answers=170, seed=1, masked_tail=0):
    """SURVEY.md section 8d inputs: images U[0,1), ids U{2..32099}, labels U{0..answers-1}."""
    g = torch.Generator().manual_seed(seed)
    batch = dict(
        image_tensors=torch.rand(B, 3, H, W, generator=g),
        question_input_ids=torch.randint(2, 32100, (B, L), generator=g),
        question_attention_masks=torch.ones(B, L, dtype=torch.long),
        annotation_ids=torch.randint(0, answers, (B,), generator=g),
    )
    if masked_tail:
        batch["question_attention_masks"][:, L - masked_tail:] = 0
    return batch


# --------------------------------------------------------------------------------------------------
# forward restatement
# --------------------------------------------------------------------------------------------------
def _bn(sd, k, x):
    # eval-mode BatchNorm2d: the reference calls vision_model.eval() (model/resnet_vqa_model.py:116,127)
    return F.batch_norm(x, sd[k + ".running_mean"], sd[k + ".running_var"], sd[k + ".weight"], sd[k + ".bias"],
                        False, 0.0, 1e-5)


def resnet_body(sd, x, vision_name, p="vision_model.", return_levels=False):
    """conv1/bn1/relu/maxpool/layer1..4 without avgpool/fc (model/resnet_vqa_model.py:119-121; tv:266-277)."""
    kind, layers = RESNET_CFG[vision_name]
    x = F.relu(_bn(sd, p + "bn1", F.conv2d(x, sd[p + "conv1.weight"], None, 2, 3)))
    x = F.max_pool2d(x, 3, 2, 1)
    levels = []
    for li, n in enumerate(layers):
        for bi in range(n):
            b = "%slayer%d.%d." % (p, li + 1, bi)
            stride = 2 if (bi == 0 and li > 0) else 1
            idn = x
            if kind == "bottleneck":   # tv:143-163 (v1.5: stride on conv2)
                o = F.relu(_bn(sd, b + "bn1", F.conv2d(x, sd[b + "conv1.weight"])))
                o = F.relu(_bn(sd, b + "bn2", F.conv2d(o, sd[b + "conv2.weight"], None, stride, 1)))
                o = _bn(sd, b + "bn3", F.conv2d(o, sd[b + "conv3.weight"]))
            else:                      # tv:89-105
                o = F.relu(_bn(sd, b + "bn1", F.conv2d(x, sd[b + "conv1.weight"], None, stride, 1)))
                o = _bn(sd, b + "bn2", F.conv2d(o, sd[b + "conv2.weight"], None, 1, 1))
            if (b + "downsample.0.weight") in sd:
                idn = _bn(sd, b + "downsample.1", F.conv2d(x, sd[b + "downsample.0.weight"], None, stride, 0))
            x = F.relu(o + idn)
        levels.append(x)
    return levels if return_levels else x


def fpn_backbone(sd, x, p="vision_model."):
    """`fasterrcnn_resnet50_fpn(pretrained=True).backbone` in eval mode (model/faster_rcnn_vqa_model.py:51-53,102-108):
    ResNet-50 body with FrozenBatchNorm2d (same arithmetic as eval BatchNorm, eps 1e-5) returning layer1..4, then
    torchvision's FeaturePyramidNetwork (lateral 1x1, nearest top-down, 3x3 output convs) and LastLevelMaxPool
    (max_pool2d(kernel 1, stride 2) of the top level) -> {'0','1','2','3','pool'}."""
    c = resnet_body(sd, x, "resnet50", p + "body.", return_levels=True)

    def inner(i, t):
        return F.conv2d(t, sd[p + "fpn.inner_blocks.%d.0.weight" % i], sd[p + "fpn.inner_blocks.%d.0.bias" % i])

    def layer(i, t):
        return F.conv2d(t, sd[p + "fpn.layer_blocks.%d.0.weight" % i], sd[p + "fpn.layer_blocks.%d.0.bias" % i], 1, 1)
    last = inner(3, c[3])
    out = {"3": layer(3, last)}
    for i in (2, 1, 0):
        lat = inner(i, c[i])
        last = lat + F.interpolate(last, size=lat.shape[-2:], mode="nearest")
        out[str(i)] = layer(i, last)
    out["pool"] = F.max_pool2d(out["3"], kernel_size=1, stride=2, padding=0)
    return out


def t5_buckets(Lq, Lk, num_buckets=32, max_distance=128):
    """hf:189-234, bidirectional."""
    rel = torch.arange(Lk)[None, :] - torch.arange(Lq)[:, None]
    nb = num_buckets // 2
    out = (rel > 0).long() * nb
    rel = rel.abs()
    max_exact = nb // 2
    large = max_exact + (torch.log(rel.float() / max_exact) / math.log(max_distance / max_exact)
                         * (nb - max_exact)).long()
    large = torch.min(large, torch.full_like(large, nb - 1))
    return out + torch.where(rel < max_exact, rel, large)


def _rms(x, w, eps):
    # hf:55-68
    var = x.float().pow(2).mean(-1, keepdim=True)
    return w * (x * torch.rsqrt(var + eps))


def t5_encoder(sd, ids, mask, p="lang_model."):
    """T5Stack encoder in eval mode (dropout off): hf:637-792, block hf:424-498, attention hf:253-344
    (no 1/sqrt(d) scaling; position bias computed once from block 0's table and shared, hf:755-758;
    additive key mask = (1-mask)*finfo.min)."""
    B, L = ids.shape
    nh, dk, eps = T5["num_heads"], T5["d_kv"], T5["eps"]
    h = F.embedding(ids, sd[p + "embed_tokens.weight"])
    table = sd[p + "block.0.layer.0.SelfAttention.relative_attention_bias.weight"]
    bias = table[t5_buckets(L, L)].permute(2, 0, 1).unsqueeze(0)            # [1, nh, L, L]
    if mask is not None:
        ext = (1.0 - mask[:, None, None, :].float()) * torch.finfo(torch.float32).min
        bias = bias + ext
    for b in range(T5["num_layers"]):
        a = "%sblock.%d.layer.0." % (p, b)
        n = _rms(h, sd[a + "layer_norm.weight"], eps)
        q = F.linear(n, sd[a + "SelfAttention.q.weight"]).view(B, L, nh, dk).transpose(1, 2)
        k = F.linear(n, sd[a + "SelfAttention.k.weight"]).view(B, L, nh, dk).transpose(1, 2)
        v = F.linear(n, sd[a + "SelfAttention.v.weight"]).view(B, L, nh, dk).transpose(1, 2)
        s = torch.matmul(q, k.transpose(3, 2)) + bias
        w = F.softmax(s.float(), dim=-1)
        ctx = torch.matmul(w, v).transpose(1, 2).contiguous().view(B, L, nh * dk)
        h = h + F.linear(ctx, sd[a + "SelfAttention.o.weight"])
        f = "%sblock.%d.layer.1." % (p, b)
        n = _rms(h, sd[f + "layer_norm.weight"], eps)
        h = h + F.linear(F.relu(F.linear(n, sd[f + "DenseReluDense.wi.weight"])), sd[f + "DenseReluDense.wo.weight"])
    return _rms(h, sd[p + "final_layer_norm.weight"], eps)


def _mhatt(sd, p, v, k, q):
    """model/multi_head_vision_text_attn.py:38-86 (mask is always None on this path)."""
    B, H, hd = q.shape[0], 8, 96
    v = F.linear(v, sd[p + "linear_v.weight"], sd[p + "linear_v.bias"]).view(B, -1, H, hd).transpose(1, 2)
    k = F.linear(k, sd[p + "linear_k.weight"], sd[p + "linear_k.bias"]).view(B, -1, H, hd).transpose(1, 2)
    q = F.linear(q, sd[p + "linear_q.weight"], sd[p + "linear_q.bias"]).view(B, -1, H, hd).transpose(1, 2)
    s = torch.matmul(q, k.transpose(-2, -1)) / math.sqrt(hd)
    a = torch.matmul(F.softmax(s, dim=-1), v).transpose(1, 2).contiguous().view(B, -1, H * hd)
    return F.linear(a, sd[p + "linear_merge.weight"], sd[p + "linear_merge.bias"])


def sga(sd, p, x, y):
    """model/multi_head_vision_text_attn.py:145-158 (eval mode)."""
    def ln(k, t):
        return F.layer_norm(t, (768,), sd[p + k + ".norm.weight"], sd[p + k + ".norm.bias"], 1e-5)
    x = ln("norm1", x + _mhatt(sd, p + "mhatt1.", x, x, x))
    x = ln("norm2", x + _mhatt(sd, p + "mhatt2.", y, y, x))
    f = F.linear(F.relu(F.linear(x, sd[p + "ffn.mlp.fc1.weight"], sd[p + "ffn.mlp.fc1.bias"])),
                 sd[p + "ffn.mlp.fc2.weight"], sd[p + "ffn.mlp.fc2.bias"])
    return ln("norm3", x + f)


def forward(sd, vision_name, question_input_ids, question_attention_masks, annotation_ids, image_tensors,
            num_attention_blocks=3, return_features=False):
    """ResnetVQAModel.forward in eval mode with grad enabled (model/resnet_vqa_model.py:114-165)."""
    with torch.no_grad():
        if vision_name == "faster-rcnn":      # model/faster_rcnn_vqa_model.py:102-108: the FPN's 'pool' level
            feat_all = fpn_backbone(sd, image_tensors.float())
            feat = feat_all["pool"]
        else:
            feat = resnet_body(sd, image_tensors.float(), vision_name)
            feat_all = feat
    proj = "downscale_layer" if vision_name == "resnet50" else "upscale_layer"
    ve = F.conv_transpose2d(feat, sd[proj + ".weight"], sd[proj + ".bias"], 1, 1)
    text = t5_encoder(sd, question_input_ids, question_attention_masks)
    y = ve.view(ve.shape[0], ve.shape[1], -1).permute(0, 2, 1)
    fused = None
    for l in range(num_attention_blocks):
        fused = sga(sd, "sga_modules.%d." % l, text, y)   # x is ALWAYS the T5 output (:147-149)
        y = fused
    w = F.softmax(F.linear(fused, sd["attention_pooler.attention.0.weight"],
                           sd["attention_pooler.attention.0.bias"]), dim=1).transpose(1, 2)
    pooled = torch.bmm(w, fused).squeeze(1)
    logp = F.log_softmax(F.linear(pooled, sd["classification_layer.weight"], sd["classification_layer.bias"]), -1)
    loss = F.nll_loss(logp, annotation_ids) if annotation_ids is not None else None
    if return_features:
        return logp, loss, feat_all
    return logp, loss


def trainable_keys(sd, vision_name):
    """Keys that receive a gradient in the reference (everything but the frozen backbone and the unused
    scaling layer)."""
    unused = "upscale_layer." if vision_name == "resnet50" else "downscale_layer."     # (faster-rcnn has no downscale_layer)
    return [k for k, v in sd.items()
            if not k.startswith("vision_model.") and not k.startswith(unused) and v.is_floating_point()]


def forward_backward(sd, vision_name, batch):
    """Returns (logp, loss, {key: grad}) with fp32 autograd."""
    keys = trainable_keys(sd, vision_name)
    work = {k: (v.clone().requires_grad_(True) if k in set(keys) else v) for k, v in sd.items()}
    logp, loss = forward(work, vision_name, batch["question_input_ids"], batch["question_attention_masks"],
                         batch["annotation_ids"], batch["image_tensors"])
    loss.backward()
    grads = {k: work[k].grad for k in keys}
    return logp.detach(), loss.detach(), grads
