"""N-rank data-parallel step == 1-rank step on the concatenated batch (run with torchrun on >= 2 GPUs).

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/ddp_check.py
"""
import os, sys
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("VQA_B200_PRETRAINED", "0")
import t5_resnet_vqa_b200 as pkg
from oracle import vqa_oracle as O

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
per = 4
sd = O.random_state_dict("resnet18", 170, seed=0)
full = O.synthetic_batch(per * world, 16, 64, 64, 170, seed=1, masked_tail=3)

def run(model, batch):
    kw = {k: v.to(dev) for k, v in batch.items()}
    logp, loss = model(kw["question_input_ids"], None, kw["question_attention_masks"], None, kw["annotation_ids"],
                       kw["image_tensors"])
    loss.backward()
    return float(loss), {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}

m = pkg.ResnetVQAModel("resnet18", "t5-base", 170); m.load_state_dict(sd); m.to(dev).eval()
shard = {k: v[rank * per:(rank + 1) * per] for k, v in full.items()}
loss_r, g_ddp = run(m, shard)
assert m._engine._ddp is not None, "gradient sync was not enabled"
t = torch.tensor([loss_r], device=dev); dist.all_reduce(t); loss_mean = float(t) / world
if rank == 0:
    os.environ["VQA_B200_DDP"] = "0"      # the single-process reference must not join any collective
    m1 = pkg.ResnetVQAModel("resnet18", "t5-base", 170); m1.load_state_dict(sd); m1.to(dev).eval()
    loss_1, g_1 = run(m1, full)
    worst = 0.0
    for k in g_1:
        d = float((g_ddp[k] - g_1[k]).norm() / (g_1[k].norm() + 1e-12))
        if float(g_1[k].norm()) > 1e-6 * max(float(v.norm()) for v in g_1.values()):
            worst = max(worst, d)
    print("ddp_check: world %d loss mean-of-ranks %.6f vs single %.6f; worst per-tensor grad rel diff %.3e" % (
        world, loss_mean, loss_1, worst))
    assert abs(loss_mean - loss_1) < 2e-3 * abs(loss_1) and worst < 3e-2, "DDP mismatch"
    print("ddp_check OK")
dist.barrier(); dist.destroy_process_group()
