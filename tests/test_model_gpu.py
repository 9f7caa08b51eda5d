"""Model-level parity (GPU): the B200 ResnetVQAModel against (a) the golden vectors frozen from the UNMODIFIED
reference (tests/golden/*.pt, made by oracle/make_golden.py) and (b) the CPU oracle run live on the same
weights and inputs.  Bars from BASELINE.json north_star (bf16 compute / fp32 accumulate vs fp32 reference,
dropout off): log-prob rel-err <= 2e-2, loss rel-err <= 1e-3, top-1 agreement >= 99 %, per-tensor gradient
cosine >= 0.999, vision_model.*.grad is None."""
import json
import os

import pytest
import torch

from util import cosine

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

LOGP_REL, LOSS_REL, TOP1, GRAD_COS = 2e-2, 1e-3, 0.99, 0.999
# All four bars are asserted as stated.  The gradient-cosine bar holds on EVERY tensor because the first T5 blocks run
# their forward GEMMs with two-term (hi + lo bf16) operands (engine.t5_split_blocks, tools/precision_probe.py: with plain
# bf16 operands the deepest tensors - block 0/1 wi, q, k, layer_norm - sit at 0.9982..0.9989 from ReLU kinks and softmax
# rows moved by the forward rounding).  Top-1 is resolved over 512 samples (one flip = 0.2 %); on the small golden batches
# every sample whose reference margin exceeds the log-prob tolerance must agree.


def build(pkg, vision, sd, device, train=False):
    os.environ["VQA_B200_PRETRAINED"] = "0"
    cls = pkg.FasterRcnnVQAModel if vision == "faster-rcnn" else pkg.ResnetVQAModel
    m = cls(vision, "t5-base", answer_spaces=170)
    m.load_state_dict(sd, strict=True)
    m.to(device)
    m.train(train)
    return m


def run(m, batch, device):
    kw = {k: v.to(device) for k, v in batch.items()}
    return m(question_input_ids=kw["question_input_ids"], decoder_question_input_ids=None,
             question_attention_masks=kw["question_attention_masks"], decoder_question_attention_masks=None,
             annotation_ids=kw["annotation_ids"], image_tensors=kw["image_tensors"])


def report(name, **kv):
    out = os.path.join(os.path.dirname(GOLD), "..", "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "parity.jsonl"), "a") as f:
        f.write(json.dumps(dict(name=name, **kv)) + "\n")


def check_outputs(logp, loss, ref_logp, ref_loss, name=""):
    logp, loss = logp.detach().float().cpu(), float(loss.detach())
    rel = float((logp - ref_logp).norm() / ref_logp.norm())
    # element-wise too: log-probs are O(5); bound the worst element relative to its magnitude
    worst = float(((logp - ref_logp).abs() / ref_logp.abs().clamp_min(1.0)).max())
    lrel = abs(loss - float(ref_loss)) / abs(float(ref_loss))
    same = logp.argmax(1) == ref_logp.argmax(1)
    # a random-init model has near-flat answers: on a batch too small to resolve 99 % a top-1 flip only counts when the
    # reference's own margin between its best two answers is larger than the log-prob tolerance (otherwise the
    # reference itself would flip under an fp32 re-ordering); from 64 samples on the >= 99 % bar is enforced as stated
    # on batches of >= 100 samples, and test_top1_agreement_512 resolves it properly
    top2 = ref_logp.topk(2, dim=1).values
    decisive = (top2[:, 0] - top2[:, 1]) > LOGP_REL * top2[:, 0].abs()
    agree = float(same.float().mean())
    report(name, logp_rel=rel, logp_worst=worst, loss_rel=lrel, top1=agree, n=int(logp.shape[0]),
           decisive=int(decisive.sum()))
    assert rel <= LOGP_REL, "log-prob rel err %.3e" % rel
    assert worst <= LOGP_REL, "worst log-prob element rel err %.3e" % worst
    assert lrel <= LOSS_REL, "loss rel err %.3e" % lrel
    assert bool(same[decisive].all())
    if logp.shape[0] >= 100:
        assert agree >= TOP1, agree
    return agree


@pytest.mark.parametrize("case", ["r34_b4", "r50_b2_masked", "r18_b2_256_l16", "r50_b64", "r50_b64_masked",
                                  "r34_b2_448", "frcnn_b2_256_l16", "frcnn_b2_448"])
def test_parity_with_reference_golden_and_oracle(pkg, cuda, case):
    from oracle import vqa_oracle as O
    gold = torch.load(os.path.join(GOLD, case + ".pt"), weights_only=False)
    c = gold["case"]
    sd = O.random_state_dict(c["vision"], 170, seed=0)
    batch = O.synthetic_batch(c["B"], c["L"], c["H"], c["W"], 170, seed=1, masked_tail=c["masked_tail"])
    m = build(pkg, c["vision"], sd, cuda)
    assert list(m.state_dict().keys()) == gold["state_dict_keys"]
    logp, loss = run(m, batch, cuda)
    assert logp.shape == (c["B"], 170) and logp.dtype == torch.float32 and loss.dim() == 0
    check_outputs(logp, loss, gold["logp"], gold["loss"], case + ":golden")
    loss.backward()
    grads = {k: p.grad for k, p in m.named_parameters()}
    assert sorted(k for k, g in grads.items() if g is None) == gold["grad_none"]
    assert all(g is None for k, g in grads.items() if k.startswith("vision_model."))
    # (b) live oracle: full per-tensor cosine (computed first so the report is complete even when (a) fails)
    noise = 1e-5 * max(gold["grad_norm"].values())
    o_logp, o_loss, o_grads = O.forward_backward(sd, c["vision"], batch)
    cos = sorted((cosine(grads[k].float().cpu(), g), k) for k, g in o_grads.items() if float(g.norm()) > 100 * noise)
    report(case + ":grad_cosine", worst=cos[:8], median=cos[len(cos) // 2][0], n=len(cos))
    # (a) golden: norms of every tensor, cosine on the frozen samples
    # (linear_k.bias gradients are mathematically zero - softmax is shift invariant - so tensors whose reference
    # norm is pure rounding noise are only required to stay at noise level)
    bad = []
    for k, n_ref in gold["grad_norm"].items():
        g = grads[k].float().cpu()
        if abs(float(g.norm()) - n_ref) > 2e-2 * n_ref + noise:
            bad.append((k, float(g.norm()), n_ref))
        f = g.flatten()
        s = f.clone() if f.numel() <= 2304 else f[::f.numel() // 128][:128]
        if n_ref > 100 * noise and float(gold["grad_sample"][k].norm()) > 1e-3 * n_ref:
            if cosine(s, gold["grad_sample"][k]) < 0.99:   # 128-element sample: looser than the full tensor
                bad.append((k, "sample cosine", cosine(s, gold["grad_sample"][k])))
    report(case + ":golden_grad_violations", bad=bad[:10], n_bad=len(bad))
    check_outputs(logp, loss, o_logp, o_loss, case + ":oracle")
    assert not bad, bad[:10]
    frac = sum(1 for c_, _ in cos if c_ >= GRAD_COS) / len(cos)
    report(case + ":grad_cosine_summary", frac_ge_0p999=frac, worst=cos[0][0], median=cos[len(cos) // 2][0])
    assert cos[0][0] >= GRAD_COS, cos[:5]            # the north-star bar, on every tensor
    assert cos[len(cos) // 2][0] >= 0.9998, cos[len(cos) // 2]


def test_faster_rcnn_generate_answers_returns_the_fpn_maps(pkg, cuda):
    """FasterRcnnVQAModel.generate_answers (model/faster_rcnn_vqa_model.py:128-197): log-probs as forward, plus the FPN's five
    feature maps ('0'..'3', 'pool') as fp32 NCHW, against the oracle's restatement of torchvision's BackboneWithFPN."""
    from oracle import vqa_oracle as O
    sd = O.random_state_dict("faster-rcnn", 170, seed=0)
    batch = O.synthetic_batch(2, 16, 256, 256, 170, seed=3)
    m = build(pkg, "faster-rcnn", sd, cuda)
    kw = {k: v.to(cuda) for k, v in batch.items()}
    with torch.no_grad():
        logp, loss, d = m.generate_answers(kw["question_input_ids"], None, kw["question_attention_masks"], None,
                                           kw["image_tensors"], annotation_ids=kw["annotation_ids"])
        logp_f, loss_f = run(m, batch, cuda)
    assert torch.equal(logp, logp_f) and torch.equal(loss, loss_f)
    o_logp, o_loss, feats = O.forward(sd, "faster-rcnn", batch["question_input_ids"], batch["question_attention_masks"],
                                      batch["annotation_ids"], batch["image_tensors"], return_features=True)
    assert sorted(d.keys()) == ["0", "1", "2", "3", "pool"]
    for k, ref in feats.items():
        assert d[k].shape == ref.shape and d[k].dtype == torch.float32, k
        assert float((d[k].cpu() - ref).norm() / ref.norm()) < 3e-2, k
    check_outputs(logp, loss, o_logp.detach(), o_loss.detach(), "frcnn_generate_answers")


def test_generate_answers_features_and_eval_determinism(pkg, cuda):
    from oracle import vqa_oracle as O
    sd = O.random_state_dict("resnet34", 170, seed=0)
    batch = O.synthetic_batch(2, 32, 224, 224, 170, seed=3)
    m = build(pkg, "resnet34", sd, cuda)
    kw = {k: v.to(cuda) for k, v in batch.items()}
    with torch.no_grad():
        logp, loss, d = m.generate_answers(kw["question_input_ids"], None, kw["question_attention_masks"], None,
                                           kw["image_tensors"])
        logp2, loss2, _ = m.generate_answers(kw["question_input_ids"], None, kw["question_attention_masks"], None,
                                             kw["image_tensors"], annotation_ids=kw["annotation_ids"])
    assert loss is None and loss2 is not None and torch.equal(logp, logp2)
    o_logp, o_loss, feat = O.forward(sd, "resnet34", batch["question_input_ids"], batch["question_attention_masks"],
                                     batch["annotation_ids"], batch["image_tensors"], return_features=True)
    assert d["features"].shape == feat.shape and d["features"].dtype == torch.float32
    rel = float((d["features"].cpu() - feat).norm() / feat.norm())
    assert rel < 3e-2, rel
    check_outputs(logp2, loss2, o_logp.detach(), o_loss.detach(), "generate_answers")


def test_top1_agreement_512(pkg, cuda):
    """Top-1 answer agreement >= 99 % (north_star) on BASELINE configs[1] (ResNet50, 64 per batch, 224x224, 32 tokens), resolved
    over 512 samples = 8 batches.  The fp32 oracle runs on the GPU here (torch fp32, TF32 off: the reference's own eager
    path) so that 512 ResNet50 samples take seconds; the first batch is cross-checked against the CPU oracle."""
    from oracle import vqa_oracle as O
    sd = O.random_state_dict("resnet50", 170, seed=0)
    sd_gpu = {k: v.to(cuda) for k, v in sd.items()}
    m = build(pkg, "resnet50", sd, cuda)
    same, total, rels = 0, 0, []
    for chunk in range(8):
        batch = O.synthetic_batch(64, 32, 224, 224, 170, seed=100 + chunk, masked_tail=10 if chunk % 2 else 0)
        kw = {k: v.to(cuda) for k, v in batch.items()}
        with torch.no_grad():
            logp, loss = run(m, batch, cuda)
            o_logp, o_loss = O.forward(sd_gpu, "resnet50", kw["question_input_ids"], kw["question_attention_masks"],
                                       kw["annotation_ids"], kw["image_tensors"])
            if chunk == 0:
                c_logp, _ = O.forward(sd, "resnet50", batch["question_input_ids"], batch["question_attention_masks"],
                                      batch["annotation_ids"], batch["image_tensors"])
                assert float((o_logp.cpu() - c_logp).abs().max()) < 1e-4
        o_logp = o_logp.cpu()
        check_outputs(logp, loss, o_logp, o_loss.cpu(), "top1_512:%d" % chunk)
        same += int((logp.cpu().argmax(1) == o_logp.argmax(1)).sum())
        total += 64
    report("top1_512", agree=same / total, same=same, n=total)
    assert same / total >= TOP1, (same, total)


def test_train_mode_dropout_and_fused_optimizer_step(pkg, cuda):
    """train(): dropout on -> different losses step to step; AdamW via the fused kernel moves every trainable
    tensor; torch.optim.AdamW on the same gradients gives the same update (trainer runs with either)."""
    from oracle import vqa_oracle as O
    sd = O.random_state_dict("resnet34", 170, seed=0)
    batch = O.synthetic_batch(4, 32, 224, 224, 170, seed=1)
    m = build(pkg, "resnet34", sd, cuda, train=True)
    groups = [{"params": m.lang_model.parameters(), "lr": 5e-3}, {"params": m.upscale_layer.parameters(), "lr": 5e-4},
              {"params": m.sga_modules.parameters(), "lr": 5e-4}, {"params": m.attention_pooler.parameters(), "lr": 5e-4},
              {"params": m.classification_layer.parameters(), "lr": 1e-5},
              {"params": m.vision_model.parameters(), "lr": 8e-3}]
    opt = torch.optim.VQAFusedAdamW(groups, weight_decay=0.1, amsgrad=True)
    before = {k: p.detach().clone() for k, p in m.named_parameters()}
    losses = []
    for _ in range(3):
        opt.zero_grad()
        logp, loss = run(m, batch, cuda)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(m.parameters(), 1.0)
        if len(losses) == 0:
            g0 = {k: p.grad.detach().clone() for k, p in m.named_parameters() if p.grad is not None}
        opt.step()
        losses.append(float(loss))
    assert all(l == l and l < 20 for l in losses)
    assert len(set(losses)) == 3                       # dropout masks change every step
    moved = [k for k, p in m.named_parameters() if not torch.equal(p.detach(), before[k])]
    frozen = [k for k in before if k not in moved]
    assert all(k.startswith("vision_model.") or k.startswith("downscale_layer.") for k in frozen), frozen[:5]
    # one reference AdamW step from the same start on the first step's (clipped) gradients
    ref = {k: before[k].clone().requires_grad_(True) for k in g0}
    ropt = torch.optim.AdamW([{"params": [ref[k] for k in g0 if k.startswith("lang_model.")], "lr": 5e-3},
                              {"params": [ref[k] for k in g0 if not k.startswith("lang_model.")
                                          and not k.startswith("classification_layer.")], "lr": 5e-4},
                              {"params": [ref[k] for k in g0 if k.startswith("classification_layer.")], "lr": 1e-5}],
                             weight_decay=0.1, amsgrad=True)
    for k in g0:
        ref[k].grad = g0[k]
    ropt.step()
    m2 = build(pkg, "resnet34", sd, cuda, train=True)
    opt2 = torch.optim.VQAFusedAdamW([{"params": m2.lang_model.parameters(), "lr": 5e-3},
                                      {"params": list(m2.upscale_layer.parameters()) + list(m2.sga_modules.parameters())
                                       + list(m2.attention_pooler.parameters()), "lr": 5e-4},
                                      {"params": m2.classification_layer.parameters(), "lr": 1e-5}],
                                     weight_decay=0.1, amsgrad=True)
    m2._engine._ensure(cuda)
    for k, p in m2.named_parameters():
        if k in g0:
            p.grad = g0[k].clone()
    opt2.step()
    for k, p in m2.named_parameters():
        if k in g0:
            assert float((p.detach() - ref[k].detach()).abs().max()) < 2e-6, k


def test_clip_grad_norm_fast_path(pkg, cuda):
    """torch.nn.utils.clip_grad_norm_ (the trainer's call, trainer/faster_rcnn_vqa_trainer.py:399-400) is routed
    through the flat-buffer kernels: same norm and same clipped gradients as torch's implementation; once the fused
    optimizer is attached the scaling is folded into its AdamW pass, which must equal torch.optim.AdamW stepping
    from the same state on the clipped gradients."""
    import copy
    from oracle import vqa_oracle as O
    from t5_resnet_vqa_b200 import optim as vo
    assert torch.nn.utils.clip_grad_norm_ is vo.clip_grad_norm_
    sd = O.random_state_dict("resnet18", 170, seed=0)
    batch = O.synthetic_batch(2, 16, 64, 64, 170, seed=1)
    m = build(pkg, "resnet18", sd, cuda)          # eval(): dropout off
    opt = torch.optim.VQAFusedAdamW(m.parameters(), lr=1e-3, weight_decay=0.1, amsgrad=True)
    MAXN = 0.05

    def fwd_bwd():
        opt.zero_grad()
        _, loss = run(m, batch, cuda)
        loss.backward()
        return {k: p.grad.detach().clone() for k, p in m.named_parameters() if p.grad is not None}

    def torch_clip(grads):
        holders = [torch.zeros_like(g).requires_grad_(True) for g in grads.values()]
        for h, g in zip(holders, grads.values()):
            h.grad = g.clone()
        total = vo._torch_clip_grad_norm_(holders, MAXN)
        return total, {k: h.grad for k, h in zip(grads, holders)}

    # step 1: no fused optimizer attached yet -> in-place scaling, .grad holds the clipped values like torch
    g = fwd_bwd()
    want, clipped = torch_clip(g)
    got = torch.nn.utils.clip_grad_norm_(m.parameters(), MAXN)
    assert got.dim() == 0 and abs(float(got) / float(want) - 1) < 1e-5 and float(want) > MAXN
    assert m._engine.pending_clip is None
    for k, p in m.named_parameters():
        if p.grad is not None:
            assert float((p.grad - clipped[k]).abs().max()) <= 1e-6 * max(1.0, float(clipped[k].abs().max())), k
    opt.step()

    # step 2: the optimizer now covers the whole engine -> deferred; gradients stay unscaled until step()
    g = fwd_bwd()
    before = {k: p.detach().clone() for k, p in m.named_parameters()}
    state = copy.deepcopy(opt.state_dict())
    want, clipped = torch_clip(g)
    got = torch.nn.utils.clip_grad_norm_(m.parameters(), MAXN)
    assert abs(float(got) / float(want) - 1) < 1e-5
    assert m._engine.pending_clip == MAXN
    for k, p in m.named_parameters():
        if p.grad is not None:
            assert torch.equal(p.grad, g[k])
    opt.step()
    assert m._engine.pending_clip is None
    ref = [before[k].clone().requires_grad_(True) for k, _ in m.named_parameters()]
    ropt = torch.optim.AdamW(ref, lr=1e-3, weight_decay=0.1, amsgrad=True)
    ropt.load_state_dict(state)
    for r, (k, _) in zip(ref, m.named_parameters()):
        if k in clipped:
            r.grad = clipped[k]
    ropt.step()
    for r, (k, p) in zip(ref, m.named_parameters()):
        assert float((p.detach() - r.detach()).abs().max()) < 2e-6, k
    # step 3: a caller that bound torch's own clip_grad_norm_ before this package was imported (`from torch.nn.utils import
    # clip_grad_norm_`) keeps calling torch's implementation: it scales the flat-buffer views in place, nothing is deferred,
    # and the fused optimizer then updates from the clipped gradients like torch.optim.AdamW
    g = fwd_bwd()
    before = {k: p.detach().clone() for k, p in m.named_parameters()}
    state = copy.deepcopy(opt.state_dict())
    want, clipped = torch_clip(g)
    got = vo._torch_clip_grad_norm_(m.parameters(), MAXN)
    assert abs(float(got) / float(want) - 1) < 1e-5 and m._engine.pending_clip is None
    for k, p in m.named_parameters():
        if p.grad is not None:
            assert float((p.grad - clipped[k]).abs().max()) <= 1e-6 * max(1.0, float(clipped[k].abs().max())), k
    opt.step()
    ref = [before[k].clone().requires_grad_(True) for k, _ in m.named_parameters()]
    ropt = torch.optim.AdamW(ref, lr=1e-3, weight_decay=0.1, amsgrad=True)
    ropt.load_state_dict(state)
    for r, (k, _) in zip(ref, m.named_parameters()):
        if k in clipped:
            r.grad = clipped[k]
    ropt.step()
    for r, (k, p) in zip(ref, m.named_parameters()):
        assert float((p.detach() - r.detach()).abs().max()) < 2e-6, k
    # tensors outside an engine still go to torch
    w = torch.randn(10, 10, device=cuda, requires_grad=True)
    w.grad = torch.ones_like(w)
    assert abs(float(torch.nn.utils.clip_grad_norm_([w], 1.0)) - 10.0) < 1e-4


def test_gradient_accumulation_and_hooks_keep_autograd_semantics(pkg, cuda):
    """Two backward passes without zero_grad add up (micro-batch accumulation), the first pass's gradients are views of the
    engine's flat buffer, and a parameter hook (which forces the autograd AccumulateGrad path) sees the same gradient."""
    from oracle import vqa_oracle as O
    sd = O.random_state_dict("resnet18", 170, seed=0)
    batch = O.synthetic_batch(2, 16, 64, 64, 170, seed=1)
    m = build(pkg, "resnet18", sd, cuda, train=False)
    _, loss = run(m, batch, cuda)
    loss.backward()
    eng = m._engine
    w = m.classification_layer.weight
    assert w.grad.data_ptr() == eng.gp(w)                      # a view of the flat gradient buffer, no copy
    assert all(p.grad is None for p in m.vision_model.parameters())
    g1 = {k: p.grad.detach().clone() for k, p in m.named_parameters() if p.grad is not None}
    _, loss = run(m, batch, cuda)
    loss.backward()                                            # no zero_grad in between
    for k, p in m.named_parameters():
        if p.grad is not None:
            assert torch.allclose(p.grad, 2 * g1[k], rtol=1e-5, atol=1e-7), k
    # hooks: the step falls back to returning gradients through autograd, so the hook fires with the same values
    seen = {}
    h = w.register_hook(lambda g: seen.setdefault("g", g.detach().clone()))
    for p in m.parameters():
        p.grad = None
    _, loss = run(m, batch, cuda)
    loss.backward()
    h.remove()
    assert "g" in seen and torch.allclose(seen["g"], g1["classification_layer.weight"], rtol=1e-5, atol=1e-7)
    assert torch.allclose(w.grad, g1["classification_layer.weight"], rtol=1e-5, atol=1e-7)


def test_weight_caches_follow_non_fused_updates_after_idle_forwards(pkg, cuda):
    """torch.optim.AdamW (the reference config's default "type") and load_state_dict change the parameters behind the engine's
    back; after two forwards without any change (validation epoch) the next update must still reach the bf16 shadow, the
    projection's conv-layout weights and the split-precision low halves: same result as a fresh model on the same weights."""
    from oracle import vqa_oracle as O
    sd = O.random_state_dict("resnet18", 170, seed=0)
    batch = O.synthetic_batch(2, 16, 64, 64, 170, seed=1)
    m = build(pkg, "resnet18", sd, cuda)
    opt = torch.optim.AdamW(m.parameters(), lr=1e-2, weight_decay=0.1, amsgrad=True)
    _, loss = run(m, batch, cuda)
    loss.backward()
    with torch.no_grad():
        a, _ = run(m, batch, cuda)
        b, _ = run(m, batch, cuda)
    assert torch.equal(a, b)
    opt.step()                                   # in-place update of every trainable tensor, projection included
    with torch.no_grad():
        after, _ = run(m, batch, cuda)
    assert not torch.equal(after, a)
    fresh = build(pkg, "resnet18", {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}, cuda)
    with torch.no_grad():
        want, _ = run(fresh, batch, cuda)
    assert torch.equal(after, want)
    # load_state_dict after idle forwards, and an edit through .data followed by invalidate()
    with torch.no_grad():
        run(m, batch, cuda)
    m.load_state_dict(sd, strict=True)
    with torch.no_grad():
        back, _ = run(m, batch, cuda)
    assert torch.equal(back, a)
    m.upscale_layer.weight.data.mul_(1.5)
    m._engine.invalidate()
    with torch.no_grad():
        scaled, _ = run(m, batch, cuda)
    assert not torch.equal(scaled, a)


def test_dropout_stream_and_fused_optimizer_state_reload(pkg, cuda):
    """Train mode: every forward draws new masks (also without a backward in between), a second backward on the same forward
    (retain_graph) regenerates the forward's masks, a new torch.manual_seed restarts the stream; VQAFusedAdamW.load_state_dict
    after a step adopts the loaded moments."""
    from oracle import vqa_oracle as O
    sd = O.random_state_dict("resnet18", 170, seed=0)
    batch = O.synthetic_batch(2, 16, 64, 64, 170, seed=1)
    torch.manual_seed(11)
    m = build(pkg, "resnet18", sd, cuda, train=True)
    with torch.no_grad():
        l1, _ = run(m, batch, cuda)
        l2, _ = run(m, batch, cuda)
    assert not torch.equal(l1, l2)
    _, loss = run(m, batch, cuda)
    loss.backward(retain_graph=True)
    g1 = {k: p.grad.detach().clone() for k, p in m.named_parameters() if p.grad is not None}
    loss.backward()
    for k, p in m.named_parameters():
        if p.grad is not None:
            assert torch.allclose(p.grad, 2 * g1[k], rtol=1e-5, atol=1e-7), k
    assert int(m._engine.rng[0]) == 11 and int(m._engine.rng[1]) == 3      # three training forwards so far
    torch.manual_seed(12)                                                    # a later re-seed is honoured
    with torch.no_grad():
        run(m, batch, cuda)
    assert int(m._engine.rng[0]) == 12 and int(m._engine.rng[1]) == 1
    # optimizer state reload after a step
    opt = torch.optim.VQAFusedAdamW(m.parameters(), lr=1e-3, weight_decay=0.1, amsgrad=True)
    opt.step()
    import copy
    saved = copy.deepcopy(opt.state_dict())
    opt.step()
    opt.load_state_dict(saved)
    opt.step()
    w = m.classification_layer.weight
    assert float(opt.state[w]["step"]) == 2.0          # 1 (loaded) + 1, not 3: the loaded state was adopted


def test_uint8_nhwc_images_equal_the_float_path(pkg, cuda):
    """Input edge: uint8 RGB [B,H,W,3] images (cv2's output before ToTensor) give bit-identical log-probs,
    loss and gradients to the reference format float [B,3,H,W] = uint8 / 255, from device or pinned host memory."""
    from oracle import vqa_oracle as O
    sd = O.random_state_dict("resnet34", 170, seed=0)
    batch = O.synthetic_batch(4, 32, 224, 224, 170, seed=1)
    g = torch.Generator().manual_seed(3)
    u8 = torch.randint(0, 256, (4, 224, 224, 3), dtype=torch.uint8, generator=g)
    batch["image_tensors"] = u8.permute(0, 3, 1, 2).float().div(255.0).contiguous()      # transforms.ToTensor()
    m = build(pkg, "resnet34", sd, cuda)
    logp_f, loss_f = run(m, batch, cuda)
    loss_f.backward()
    gf = m.upscale_layer.weight.grad.detach().clone()
    o_logp, o_loss, _ = O.forward_backward(sd, "resnet34", batch)
    check_outputs(logp_f, loss_f, o_logp, o_loss, "uint8_edge:float")
    for p in m.parameters():
        p.grad = None
    for images in (u8.to(cuda), u8.pin_memory()):
        b2 = dict(batch, image_tensors=images)
        kw = {k: (v.to(cuda) if k != "image_tensors" else v) for k, v in b2.items()}
        logp_u, loss_u = m(question_input_ids=kw["question_input_ids"], question_attention_masks=kw["question_attention_masks"],
                           annotation_ids=kw["annotation_ids"], image_tensors=kw["image_tensors"])
        assert torch.equal(logp_u, logp_f) and torch.equal(loss_u, loss_f)
    loss_u.backward()
    assert torch.equal(m.upscale_layer.weight.grad, gf)
    with pytest.raises(ValueError):
        m(question_input_ids=kw["question_input_ids"], question_attention_masks=kw["question_attention_masks"],
          annotation_ids=kw["annotation_ids"], image_tensors=u8.permute(0, 3, 1, 2).contiguous().to(cuda))
