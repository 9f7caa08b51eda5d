"""Device time of the VitVQAModel training step (BASELINE.json configs[4]: ViT-B/16 + T5-base encoder-decoder) on one B200:
train_one_step of trainer/vit_vqa_trainer.py:450-464 (zero_grad, forward, backward, clip_grad_norm_, AdamW-amsgrad over the
trainer's four parameter groups :300-318) in train() mode on synthetic inputs, CUDA events over K steps after W warm-up steps.
Prints one JSON line (also the eval-mode forward alone).  Not the driver's bench (bench.py measures the north-star
ResnetVQAModel step); this is the measurement record of SURVEY.md 8f-4."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def flops_per_sample(L, Ld):
    d, dff, T = 768, 3072, 197
    vit = 196 * 768 * 768 + 12 * T * (4 * d * d + 2 * d * dff) + 12 * 12 * 2 * T * T * 64 + d * d
    enc = 12 * L * (4 * d * d + 2 * d * dff) + 12 * 12 * 2 * L * L * 64
    dec = 12 * Ld * (4 * d * d + 2 * d * dff) + 12 * 12 * 2 * Ld * Ld * 64 + 12 * (d * d + Ld * d * d)
    head = 1536 * 768 + 768 * 170
    return 2.0 * (vit + 3 * (enc + dec + head))          # MACs -> FLOPs; the frozen ViT runs forward only


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--L", type=int, default=32)
    ap.add_argument("--Ld", type=int, default=20)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--ncu-step", action="store_true",
                    help="after warm-up run ONE step between cudaProfilerStart/Stop and exit (ncu --profile-from-start off)")
    a = ap.parse_args()
    os.environ.setdefault("VQA_B200_PRETRAINED", "0")
    import t5_resnet_vqa_b200 as pkg
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    m = pkg.VitVQAModel("google/vit-base-patch16-224-in21k", "t5-base", 170).to(dev)
    g = torch.Generator().manual_seed(1)
    B = a.batch
    batch = dict(question_input_ids=torch.randint(2, 32100, (B, a.L), generator=g),
                 decoder_question_input_ids=torch.randint(2, 32100, (B, a.Ld), generator=g),
                 question_attention_masks=torch.ones(B, a.L, dtype=torch.long),
                 decoder_question_attention_masks=torch.ones(B, a.Ld, dtype=torch.long),
                 annotation_ids=torch.randint(0, 170, (B,), generator=g),
                 pixel_values=torch.rand(B, 3, 224, 224, generator=g) * 2 - 1)
    batch = {k: v.to(dev) for k, v in batch.items()}
    groups = [dict(params=m.vision_model.parameters(), lr=1e-5, model_name="Vision Model"),
              dict(params=m.lang_model.parameters(), lr=1e-5, model_name="Language Model"),
              dict(params=m.fusing_layer.parameters(), lr=1e-4, model_name="Fusion Layer"),
              dict(params=m.classification_layer.parameters(), lr=1e-4, model_name="Classifier Layer")]
    opt = torch.optim.VQAFusedAdamW(groups, weight_decay=0.1, amsgrad=True)

    def step():
        opt.zero_grad()
        logp, loss = m(**batch)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(m.parameters(), 1.0)
        opt.step()
        return loss
    m.train()
    for _ in range(a.warmup):
        loss = step()
    torch.cuda.synchronize()
    if a.ncu_step:
        torch.cuda.profiler.start()
        step()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        print(json.dumps({"ncu_step": "done"}))
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    st = m._engine.last_state
    m.eval()
    with torch.no_grad():
        for _ in range(3):
            m(**batch)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(a.steps):
            m(**batch)
        e1.record()
        torch.cuda.synchronize()
    ms_fwd = e0.elapsed_time(e1) / a.steps
    fl = flops_per_sample(a.L, a.Ld) * B
    n_params = sum(p.numel() for p in m._engine.params)
    print(json.dumps({"metric": "train samples/s (VitVQAModel step)", "value": B / ms * 1e3, "unit": "samples/s",
                      "ms_per_step": ms, "ms_eval_forward": ms_fwd, "batch": B, "L": a.L, "Ld": a.Ld, "steps": a.steps,
                      "warmup": a.warmup, "loss": float(loss), "trainable_params": n_params,
                      "launches_per_step": st.n_fwd_launches + st.n_bwd_launches,
                      "algorithmic_tflop_per_step": fl / 1e12, "step_tflops": fl / ms / 1e9,
                      "config": "ViT-B/16 (frozen, 197 tokens) + T5-base encoder (L) + decoder (Ld) + fusing layer + classifier, "
                                "dropout on, clip + VQAFusedAdamW(amsgrad)"}))


if __name__ == "__main__":
    main()
