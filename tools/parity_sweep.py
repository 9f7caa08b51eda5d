"""Parity bars as a function of the split-precision settings (GPU; diagnostic): worst per-tensor gradient cosine on two golden
cases and top-1 agreement over 512 samples, for VQA_B200_T5_SPLIT_BLOCKS x VQA_B200_SPLIT_HEAD.

    python tools/parity_sweep.py "4,1" "4,0" "3,1" "2,1" "0,0"
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["VQA_B200_PRETRAINED"] = "0"
import t5_resnet_vqa_b200 as pkg   # noqa: E402
from oracle import vqa_oracle as O  # noqa: E402  (diagnostic tool: the oracle is the checker)

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
dev = torch.device("cuda:0")


def cos(a, b):
    a, b = a.flatten().double(), b.flatten().double()
    return float(a @ b / (a.norm() * b.norm() + 1e-300))


def build(vision, sd):
    m = pkg.ResnetVQAModel(vision, "t5-base", answer_spaces=170)
    m.load_state_dict(sd, strict=True)
    return m.to(dev).eval()


def run(m, batch):
    kw = {k: v.to(dev) for k, v in batch.items()}
    return m(question_input_ids=kw["question_input_ids"], question_attention_masks=kw["question_attention_masks"],
             annotation_ids=kw["annotation_ids"], image_tensors=kw["image_tensors"])


def main():
    cases = {"r18_b2_256_l16": ("resnet18", O.synthetic_batch(2, 16, 256, 256, 170, seed=1, masked_tail=3)),
             "r50_b64": ("resnet50", O.synthetic_batch(64, 32, 224, 224, 170, seed=1))}
    sds = {v: O.random_state_dict(v, 170, seed=0) for v in ("resnet18", "resnet50")}
    ref = {n: O.forward_backward(sds[v], v, b) for n, (v, b) in cases.items()}
    sd_gpu = {k: v.to(dev) for k, v in sds["resnet50"].items()}
    chunks = [O.synthetic_batch(64, 32, 224, 224, 170, seed=100 + c, masked_tail=10 if c % 2 else 0) for c in range(8)]
    with torch.no_grad():
        ref_top = [O.forward(sd_gpu, "resnet50", *(c[k].to(dev) for k in ("question_input_ids", "question_attention_masks",
                                                                          "annotation_ids", "image_tensors")))[0].cpu()
                   for c in chunks]
    for spec in sys.argv[1:]:
        blocks, head = spec.split(",")
        os.environ["VQA_B200_T5_SPLIT_BLOCKS"], os.environ["VQA_B200_SPLIT_HEAD"] = blocks, head
        res = {"t5_split_blocks": int(blocks), "split_head": int(head)}
        for n, (v, b) in cases.items():
            m = build(v, sds[v])
            logp, loss = run(m, b)
            loss.backward()
            g = dict(m.named_parameters())
            scale = max(float(x.norm()) for x in ref[n][2].values())
            cs = sorted((cos(g[k].grad.float().cpu(), x), k) for k, x in ref[n][2].items() if float(x.norm()) > 1e-3 * scale * 1e-2)
            res[n] = {"worst": cs[0], "below_0p999": sum(1 for c, _ in cs if c < 0.999),
                      "logp_rel": float((logp.detach().cpu() - ref[n][0]).norm() / ref[n][0].norm())}
            del m
        m = build("resnet50", sds["resnet50"])
        same = 0
        with torch.no_grad():
            for c, r in zip(chunks, ref_top):
                same += int((run(m, c)[0].cpu().argmax(1) == r.argmax(1)).sum())
        res["top1_512"] = same
        del m
        print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
