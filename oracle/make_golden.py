"""Generates tests/golden/*.pt from the UNMODIFIED reference (run in the build container only).

    python oracle/make_golden.py            # needs /root/reference; writes tests/golden/<case>.pt

The reference model class is imported from /root/reference untouched.  Two monkeypatches replace its weight
DOWNLOADS (there is no network) with random-init constructors of the same architectures; the
deterministic state_dict of oracle.vqa_oracle.random_state_dict is then loaded with strict=True, and one
eval-mode forward + backward on the synthetic batch is frozen: log-probs, loss, per-tensor gradient L2 norms,
full gradients of the small tensors and a strided 128-element sample of the large ones.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import vqa_oracle as O  # noqa: E402

REF = "/root/reference"

CASES = {
    # BASELINE.json configs[0]: ResNet34, batch 4, 224x224, 32-token question, 170 answers
    "r34_b4": dict(vision="resnet34", B=4, L=32, H=224, W=224, masked_tail=0),
    # ResNet50 path (2048->768 projection) with a padded question (key-mask handling is observable)
    "r50_b2_masked": dict(vision="resnet50", B=2, L=32, H=224, W=224, masked_tail=10),
    # the reference's native collate shapes: 256x256 images (64 vision tokens), 16-token questions
    "r18_b2_256_l16": dict(vision="resnet18", B=2, L=16, H=256, W=256, masked_tail=3),
    # BASELINE.json configs[1] (the config every bench number is quoted on): ResNet50, batch 64, 224x224, 32 tokens,
    # all-ones mask, and the same with a padded tail of 10 tokens
    "r50_b64": dict(vision="resnet50", B=64, L=32, H=224, W=224, masked_tail=0),
    "r50_b64_masked": dict(vision="resnet50", B=64, L=32, H=224, W=224, masked_tail=10),
    # BASELINE.json configs[3] regime: 448x448 images -> 14x14 = 196 vision tokens through the 512 -> 768 upscaling path
    "r34_b2_448": dict(vision="resnet34", B=2, L=32, H=448, W=448, masked_tail=4),
    # FasterRcnnVQAModel (model/faster_rcnn_vqa_model.py): FPN 'pool' level, 256 -> 768; the reference's collate shapes
    # (256x256 -> 4x4 = 16 tokens, 16-token questions) and 448x448 (7x7 = 49 tokens)
    "frcnn_b2_256_l16": dict(vision="faster-rcnn", B=2, L=16, H=256, W=256, masked_tail=3),
    "frcnn_b2_448": dict(vision="faster-rcnn", B=2, L=32, H=448, W=448, masked_tail=0),
}


def reference_model(vision):
    sys.path.insert(0, REF)
    import torchvision
    from transformers import T5Config, T5ForQuestionAnswering
    import model.resnet_vqa_model as ref
    for n in ("resnet18", "resnet34", "resnet50"):
        setattr(ref, n, (lambda name: (lambda pretrained=True: getattr(torchvision.models, name)(weights=None)))(n))
    cfg = T5Config(vocab_size=32128, d_model=768, d_kv=64, d_ff=3072, num_layers=12, num_decoder_layers=12,
                   num_heads=12, relative_attention_num_buckets=32, relative_attention_max_distance=128,
                   dropout_rate=0.1, layer_norm_epsilon=1e-6, feed_forward_proj="relu")
    T5ForQuestionAnswering.from_pretrained = staticmethod(lambda name, *a, **k: T5ForQuestionAnswering(cfg))
    if vision == "faster-rcnn":
        # fasterrcnn_resnet50_fpn(pretrained=True) builds its backbone as resnet50(norm_layer=FrozenBatchNorm2d) behind
        # _resnet_fpn_extractor(.., trainable_layers=3) and then downloads weights; the same construction without the download
        import types
        import model.faster_rcnn_vqa_model as fref
        from torchvision.models.detection.backbone_utils import _resnet_fpn_extractor
        from torchvision.ops import misc as misc_nn_ops

        def detector(pretrained=True):
            body = torchvision.models.resnet50(weights=None, norm_layer=misc_nn_ops.FrozenBatchNorm2d)
            return types.SimpleNamespace(backbone=_resnet_fpn_extractor(body, 3))
        fref.fasterrcnn_resnet50_fpn = detector
        return fref.FasterRcnnVQAModel("faster-rcnn", "t5-base", answer_spaces=170)
    return ref.ResnetVQAModel(vision, "t5-base", answer_spaces=170)


VIT_CASES = {
    # VitVQAModel (model/vit_vqa_model.py:127-227, BASELINE.json configs[4]): ViT-B/16 224x224 + T5-base encoder-decoder;
    # decoder questions padded to Enums.MAX_LEN = 20 with per-sample lengths, the reference collate's 'longest' questions
    "vit_b2_l16": dict(B=2, L=16, Ld=20, masked_tail=3),
    "vit_b4_l32": dict(B=4, L=32, Ld=20, masked_tail=0),
}


def reference_vit_model():
    sys.path.insert(0, REF)
    from transformers import T5Config, T5ForConditionalGeneration, ViTConfig, ViTModel
    cfg = T5Config(vocab_size=32128, d_model=768, d_kv=64, d_ff=3072, num_layers=12, num_decoder_layers=12,
                   num_heads=12, relative_attention_num_buckets=32, relative_attention_max_distance=128,
                   dropout_rate=0.1, layer_norm_epsilon=1e-6, feed_forward_proj="relu")
    T5ForConditionalGeneration.from_pretrained = staticmethod(lambda name, *a, **k: T5ForConditionalGeneration(cfg))
    ViTModel.from_pretrained = staticmethod(lambda name, *a, **k: ViTModel(ViTConfig()))
    import model.vit_vqa_model as vref
    return vref.VitVQAModel("google/vit-base-patch16-224-in21k", "t5-base", answer_spaces=170)


def main_vit(out_dir, only):
    from oracle import vit_oracle as V
    for name, c in VIT_CASES.items():
        if only and name not in only:
            continue
        sd = V.random_state_dict(170, seed=0)
        batch = V.synthetic_batch(c["B"], c["L"], c["Ld"], 170, seed=1, masked_tail=c["masked_tail"])
        m = reference_vit_model()
        missing = m.load_state_dict(sd, strict=True)
        m.eval()
        logp, loss = m(question_input_ids=batch["question_input_ids"],
                       decoder_question_input_ids=batch["decoder_question_input_ids"],
                       question_attention_masks=batch["question_attention_masks"],
                       decoder_question_attention_masks=batch["decoder_question_attention_masks"],
                       annotation_ids=batch["annotation_ids"], pixel_values=batch["pixel_values"], image_tensors=None,
                       answer_input_ids=None, answer_attention_masks=None)
        loss.backward()
        with torch.no_grad():
            pooled = m.vision_model(batch["pixel_values"]).pooler_output
        grads = {k: p.grad for k, p in m.named_parameters()}
        none_keys = sorted(k for k, g in grads.items() if g is None)
        gold = dict(case=c, logp=logp.detach().clone(), loss=loss.detach().clone(), vit_pooled=pooled.clone(),
                    grad_norm={k: float(g.norm()) for k, g in grads.items() if g is not None},
                    grad_sample={k: sample(g) for k, g in grads.items() if g is not None},
                    grad_none=none_keys, state_dict_keys=list(m.state_dict().keys()),
                    param_keys=[k for k, _ in m.named_parameters()], versions=dict(torch=torch.__version__))
        o_logp, o_loss, o_grads = V.forward_backward(sd, batch)
        print(name, "loss ref %.6f oracle %.6f | max|dlogp| %.3e | pooled %.3e | n_grad %d n_none %d | %s" % (
            float(loss), float(o_loss), float((logp - o_logp).abs().max()),
            float((pooled - V.vit_pooled(sd, batch["pixel_values"])).abs().max()), len(gold["grad_norm"]), len(none_keys),
            missing))
        worst = max(float((grads[k] - o_grads[k]).norm() / (grads[k].norm() + 1e-30)) for k in o_grads)
        print(name, "worst per-tensor grad rel diff oracle vs reference: %.3e" % worst)
        torch.save(gold, os.path.join(out_dir, name + ".pt"))


def sample(g):
    f = g.flatten()
    if f.numel() <= 2304:
        return f.clone()
    stride = f.numel() // 128
    return f[::stride][:128].clone()


def main():
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    only = sys.argv[1:]
    main_vit(out_dir, only)
    for name, c in CASES.items():
        if only and name not in only:
            continue
        sd = O.random_state_dict(c["vision"], 170, seed=0)
        batch = O.synthetic_batch(c["B"], c["L"], c["H"], c["W"], 170, seed=1, masked_tail=c["masked_tail"])
        m = reference_model(c["vision"])
        missing = m.load_state_dict(sd, strict=True)
        m.eval()  # dropout off; gradients still flow (the backbone is eval/no_grad anyway)
        logp, loss = m(question_input_ids=batch["question_input_ids"], decoder_question_input_ids=None,
                       question_attention_masks=batch["question_attention_masks"],
                       decoder_question_attention_masks=None, annotation_ids=batch["annotation_ids"],
                       image_tensors=batch["image_tensors"])
        loss.backward()
        grads = {k: p.grad for k, p in m.named_parameters()}
        none_keys = sorted(k for k, g in grads.items() if g is None)
        gold = dict(case=c, logp=logp.detach().clone(), loss=loss.detach().clone(),
                    grad_norm={k: float(g.norm()) for k, g in grads.items() if g is not None},
                    grad_sample={k: sample(g) for k, g in grads.items() if g is not None},
                    grad_none=none_keys, state_dict_keys=list(m.state_dict().keys()),
                    versions=dict(torch=torch.__version__))
        # cross-check the restatement right here as well
        o_logp, o_loss, o_grads = O.forward_backward(sd, c["vision"], batch)
        print(name, "loss ref %.6f oracle %.6f | max|dlogp| %.3e | n_grad %d n_none %d | %s" % (
            float(loss), float(o_loss), float((logp - o_logp).abs().max()), len(gold["grad_norm"]),
            len(none_keys), missing))
        worst = max(float((grads[k] - o_grads[k]).norm() / (grads[k].norm() + 1e-30)) for k in o_grads)
        print(name, "worst per-tensor grad rel diff oracle vs reference: %.3e" % worst)
        torch.save(gold, os.path.join(out_dir, name + ".pt"))


if __name__ == "__main__":
    main()
