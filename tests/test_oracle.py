"""CPU: the oracle restatement (oracle/vqa_oracle.py) against the golden vectors frozen from the UNMODIFIED
reference by oracle/make_golden.py (log-probs, loss, per-tensor gradient norms and gradient samples)."""
import os

import pytest
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("case", ["r34_b4", "r18_b2_256_l16", "r50_b2_masked", "r50_b64", "r34_b2_448", "frcnn_b2_256_l16", "frcnn_b2_448"])
def test_oracle_matches_reference_golden(case):
    from oracle import vqa_oracle as O
    gold = torch.load(os.path.join(GOLD, case + ".pt"), weights_only=False)
    c = gold["case"]
    sd = O.random_state_dict(c["vision"], 170, seed=0)
    assert list(sd.keys()) == gold["state_dict_keys"]          # reference state_dict layout, key for key
    batch = O.synthetic_batch(c["B"], c["L"], c["H"], c["W"], 170, seed=1, masked_tail=c["masked_tail"])
    logp, loss, grads = O.forward_backward(sd, c["vision"], batch)
    # fp32 on the same machine type: equal up to thread-count dependent summation order
    assert torch.allclose(logp, gold["logp"], rtol=0, atol=2e-5)
    assert abs(float(loss) - float(gold["loss"])) < 2e-6 * abs(float(gold["loss"])) + 1e-6
    assert sorted(grads.keys()) == sorted(gold["grad_norm"].keys())
    trainable = set(grads)
    assert all(k.startswith("vision_model.") or k.startswith(("upscale_layer.", "downscale_layer."))
               for k in gold["grad_none"])
    noise = 1e-6 * max(gold["grad_norm"].values())
    assert not (trainable & set(gold["grad_none"]))
    scale = max(gold["grad_norm"].values())
    for k, g in grads.items():
        n_ref = gold["grad_norm"][k]
        assert abs(float(g.norm()) - n_ref) <= 1e-4 * n_ref + 1e-6 * scale, k
        f = g.flatten()
        s = f if f.numel() <= 2304 else f[::f.numel() // 128][:128]
        # (the detector backbone's FrozenBatchNorm2d and eval BatchNorm differ in rounding: tensors whose gradient is
        # mathematically zero - linear_k.bias, softmax shift invariance - hold different rounding noise)
        assert torch.allclose(s, gold["grad_sample"][k], rtol=1e-3, atol=1e-6 * scale), k


def test_masked_question_changes_output():
    """Key-padding mask handling is observable: masking the tail of the question moves the log-probs."""
    from oracle import vqa_oracle as O
    sd = O.random_state_dict("resnet18", 170, seed=0)
    a = O.synthetic_batch(2, 16, 64, 64, 170, seed=1)
    b = O.synthetic_batch(2, 16, 64, 64, 170, seed=1, masked_tail=5)
    with torch.no_grad():
        la, _ = O.forward(sd, "resnet18", a["question_input_ids"], a["question_attention_masks"], a["annotation_ids"],
                          a["image_tensors"])
        lb, _ = O.forward(sd, "resnet18", b["question_input_ids"], b["question_attention_masks"], b["annotation_ids"],
                          b["image_tensors"])
    assert float((la - lb).abs().max()) > 1e-3


def test_t5_buckets_against_transformers():
    """Bidirectional T5 buckets: the oracle's restatement against the third-party implementation the reference
    actually executes (transformers T5Attention._relative_position_bucket, hf:189-234), plus hand-checked
    values (exact for |d| < 8, logarithmic beyond, +16 for keys to the right of the query)."""
    from oracle import vqa_oracle as O
    from transformers.models.t5.modeling_t5 import T5Attention
    for L in (16, 20, 32, 64):
        rel = torch.arange(L)[None, :] - torch.arange(L)[:, None]
        want = T5Attention._relative_position_bucket(rel, bidirectional=True, num_buckets=32, max_distance=128)
        assert torch.equal(O.t5_buckets(L, L), want)
    b = O.t5_buckets(32, 32)
    assert b[0, 0] == 0 and b[5, 4] == 1 and b[4, 5] == 17 and b[10, 3] == 7 and b[3, 10] == 23
    assert b[31, 0] == 11 and b[0, 31] == 27          # 8 + floor(log(31/8) / log(16) * 8) = 11
    assert int(b.max()) < 32 and int(b.min()) >= 0


@pytest.mark.parametrize("case", ["vit_b2_l16", "vit_b4_l32"])
def test_vit_oracle_matches_reference_golden(case):
    """oracle/vit_oracle.py (VitVQAModel restatement) against the goldens frozen from the unmodified reference
    class model/vit_vqa_model.py:127-227: ViT pooler_output, log-probs, loss, every gradient norm and the gradient samples."""
    from oracle import vit_oracle as V
    gold = torch.load(os.path.join(GOLD, case + ".pt"), weights_only=False)
    c = gold["case"]
    sd = V.random_state_dict(170, seed=0)
    assert list(sd.keys()) == gold["state_dict_keys"]
    assert V.trainable_keys(sd) == [k for k in gold["param_keys"] if not k.startswith("vision_model.")]
    batch = V.synthetic_batch(c["B"], c["L"], c["Ld"], 170, seed=1, masked_tail=c["masked_tail"])
    with torch.no_grad():
        assert torch.allclose(V.vit_pooled(sd, batch["pixel_values"]), gold["vit_pooled"], rtol=0, atol=2e-5)
    logp, loss, grads = V.forward_backward(sd, batch)
    assert torch.allclose(logp, gold["logp"], rtol=0, atol=2e-5)
    assert abs(float(loss) - float(gold["loss"])) < 2e-6 * abs(float(gold["loss"])) + 1e-6
    assert all(k.startswith("vision_model.") for k in gold["grad_none"])
    scale = max(gold["grad_norm"].values())
    zero = 0
    for k, g in grads.items():
        n_ref = gold["grad_norm"][k]
        zero += n_ref == 0.0
        assert abs(float(g.norm()) - n_ref) <= 1e-3 * n_ref + 1e-6 * scale, k
        f = g.flatten()
        s = f if f.numel() <= 2304 else f[::f.numel() // 128][:128]
        assert torch.allclose(s, gold["grad_sample"][k], rtol=2e-3, atol=1e-5 * scale), k
    assert zero == 36        # one-key cross-attention: q, k and their RMSNorm get exactly zero gradient in the reference too


def test_t5_causal_buckets_against_transformers():
    """Unidirectional (decoder) buckets: oracle and host restatements against transformers' own function (hf:189-234)."""
    from oracle import vit_oracle as V
    from transformers.models.t5.modeling_t5 import T5Attention
    import t5_resnet_vqa_b200.vit_step as S
    for L in (16, 20, 32, 64):
        rel = torch.arange(L)[None, :] - torch.arange(L)[:, None]
        want = T5Attention._relative_position_bucket(rel, bidirectional=False, num_buckets=32, max_distance=128)
        assert torch.equal(V.t5_buckets_causal(L, L), want)
        assert torch.equal(S.t5_causal_buckets(L, L).long(), want)
