"""The same-box library bar: the reference step executed by PyTorch eager on the B200 - cuDNN convolutions,
cuBLAS GEMMs, ATen softmax / layer_norm / elementwise, torch.optim.AdamW(amsgrad) - through the oracle's functional
restatement of the reference module (dropout off: slightly LESS work than the train-mode step bench.py times), in fp32
(TF32 off and on) and under torch.autocast(bfloat16).  Diagnostic / evidence tool; writes one JSON line.

    python tools/eager_bar.py > gpurun_out/eager_bar.json
"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import vqa_oracle as O  # noqa: E402

VISION, B, L, IMG, A = "resnet50", 64, 32, 224, 170
LRS = dict(lang=0.005, scaler=0.0005, sga=0.0005, pooler=0.0005, classifier=1e-5)


def step_time(mode, steps=20, warmup=5):
    dev = torch.device("cuda:0")
    torch.backends.cuda.matmul.allow_tf32 = mode == "tf32"
    torch.backends.cudnn.allow_tf32 = mode == "tf32"
    sd = {k: v.to(dev) for k, v in O.random_state_dict(VISION, A, seed=0).items()}
    keys = O.trainable_keys(sd, VISION)
    params = {k: sd[k].clone().requires_grad_(True) for k in keys}
    work = dict(sd)
    work.update(params)
    groups = [{"params": [params[k] for k in keys if k.startswith(p)], "lr": lr} for p, lr in
              (("lang_model.", LRS["lang"]), ("downscale_layer.", LRS["scaler"]), ("sga_modules.", LRS["sga"]),
               ("attention_pooler.", LRS["pooler"]), ("classification_layer.", LRS["classifier"]))]
    opt = torch.optim.AdamW(groups, weight_decay=0.1, amsgrad=True)
    batch = {k: v.to(dev) for k, v in O.synthetic_batch(B, L, IMG, IMG, A, seed=1).items()}

    def one():
        opt.zero_grad()
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=(mode == "autocast_bf16")):
            logp, loss = O.forward(work, VISION, batch["question_input_ids"], batch["question_attention_masks"],
                                   batch["annotation_ids"], batch["image_tensors"])
        loss.backward()
        torch.nn.utils.clip_grad_norm_(list(params.values()), 1.0)
        opt.step()
        return loss
    for _ in range(warmup):
        one()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(steps):
        loss = one()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"ms_per_step": ms, "samples_per_s": B / ms * 1e3, "wall_ms_per_step": (time.perf_counter() - t0) / steps * 1e3,
            "loss": float(loss)}


def main():
    out = {"what": "reference step (oracle restatement) under PyTorch eager on this GPU: cuDNN / cuBLAS / ATen + "
                   "torch.optim.AdamW(amsgrad), ResNet50 + T5-base + 3xSGA, batch 64, 224x224, 32 tokens, dropout off",
           "torch": torch.__version__, "gpu": torch.cuda.get_device_name(0)}
    for mode in ("fp32", "tf32", "autocast_bf16"):
        out[mode] = step_time(mode)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
