"""Bring-up aid: sharded vs replica data-parallel step side by side (torchrun, >= 2 GPUs); prints where they part."""
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
sd = O.random_state_dict("resnet18", 170, seed=0)
full = O.synthetic_batch(4 * world, 16, 64, 64, 170, seed=1, masked_tail=3)
shard = {k: v[rank * 4:(rank + 1) * 4].to(dev) for k, v in full.items()}
models = {}
for mode in ("zero1", "allreduce"):
    os.environ["VQA_B200_DDP_MODE"] = mode
    m = pkg.ResnetVQAModel("resnet18", "t5-base", 170); m.load_state_dict(sd); m.to(dev).eval()
    opt = torch.optim.VQAFusedAdamW(m.parameters(), lr=1e-4, weight_decay=0.1, amsgrad=True)
    with torch.no_grad():     # the engine (and its GradSync, which reads the mode) is set up by the first forward
        m(shard["question_input_ids"], None, shard["question_attention_masks"], None, shard["annotation_ids"], shard["image_tensors"])
    assert m._engine._ddp.mode == mode
    models[mode] = (m, opt)

def fwd_bwd(m, opt):
    opt.zero_grad()
    _, loss = m(shard["question_input_ids"], None, shard["question_attention_masks"], None, shard["annotation_ids"], shard["image_tensors"])
    loss.backward()
    n = torch.nn.utils.clip_grad_norm_(m.parameters(), 1.0)
    torch.cuda.synchronize()
    return float(loss), float(n)

for step in range(3):
    out = {}
    for mode, (m, opt) in models.items():
        loss, n = fwd_bwd(m, opt)
        e = m._engine
        out[mode] = dict(loss=loss, norm=n, shards=e.ddp_shards is not None, gsmall=e.grad[e.n_big:].clone(), gbig=e.grad[:e.n_big].clone())
    a, b = out["zero1"], out["allreduce"]
    ez = models["zero1"][0]._engine
    if rank == 0:
        print("step", step, "loss", a["loss"], b["loss"], "norm", a["norm"], b["norm"], "sharded", a["shards"], b["shards"])
        ga, gb = a["gsmall"].double(), b["gsmall"].double()
        print("   small grad max|diff| %.3e (max %.3e) cosine %.6f" % (float((ga - gb).abs().max()), float(gb.abs().max()),
                                                                       float(ga @ gb / (ga.norm() * gb.norm()))))
        emb_lo = ez.offs[id(ez.model.lang_model.embed_tokens.weight)] - ez.n_big
        print("   small grad without the embedding: cosine %.6f" % float(ga[:emb_lo] @ gb[:emb_lo] / (ga[:emb_lo].norm() * gb[:emb_lo].norm())))
        if ez.ddp_shards:
            for lo, bhi, olo, ohi in ez.ddp_shards:
                print("   own slice [%d,%d) grad max|diff| %.3e" % (olo, ohi, float((a["gbig"][olo:ohi] - b["gbig"][olo:ohi]).abs().max())))
    before = {mode: (m._engine.master.clone(), m._engine.shadow.clone()) for mode, (m, opt) in models.items()}
    shards_now = models["zero1"][0]._engine.ddp_shards
    for mode, (m, opt) in models.items():
        opt.step()
    torch.cuda.synchronize()
    if rank == 0:
        ez, er = models["zero1"][0]._engine, models["allreduce"][0]._engine
        nb = ez.n_big
        dz, dr = (ez.master - before["zero1"][0]).double(), (er.master - before["allreduce"][0]).double()
        print("   small update: norm sharded %.4e replica %.4e cosine %.6f" % (float(dz[nb:].norm()), float(dr[nb:].norm()),
              float(dz[nb:] @ dr[nb:] / (dz[nb:].norm() * dr[nb:].norm() + 1e-300))))
        if shards_now:
            for lo, bhi, olo, ohi in shards_now:
                print("   shard [%d,%d) own [%d,%d): own update norm sharded %.4e replica %.4e cos %.6f | non-own master moved %.3e | shadow range vs replica max|diff| %.3e" % (
                    lo, bhi, olo, ohi, float(dz[olo:ohi].norm()), float(dr[olo:ohi].norm()),
                    float(dz[olo:ohi] @ dr[olo:ohi] / (dz[olo:ohi].norm() * dr[olo:ohi].norm() + 1e-300)),
                    float(dz[lo:bhi].norm() ** 2 - dz[olo:ohi].norm() ** 2) ** 0.5 if True else 0,
                    float((ez.shadow[lo:bhi].float() - er.shadow[lo:bhi].float()).abs().max())))
    ea, eb = models["zero1"][0]._engine, models["allreduce"][0]._engine
    ea.wait_optimizer(); eb.wait_optimizer(); torch.cuda.synchronize()
    prev = globals().setdefault("PREV", {})
    rows = []
    pa, pb = dict(models["zero1"][0].named_parameters()), dict(models["allreduce"][0].named_parameters())
    for k in pa:
        if k.startswith("vision_model.") or pa[k].numel() > 4096 or k.endswith("linear_k.bias"):
            continue
        a0 = prev.get(("a", k), sd[k].to(dev)); b0 = prev.get(("b", k), sd[k].to(dev))
        da, db = (pa[k].detach() - a0).flatten().double(), (pb[k].detach() - b0).flatten().double()
        prev[("a", k)], prev[("b", k)] = pa[k].detach().clone(), pb[k].detach().clone()
        if float(db.norm()) > 0:
            rows.append((float(da @ db / (da.norm() * db.norm() + 1e-300)), k))
    rows.sort()
    if rank == 0:
        print("   per-step update cosine, small tensors, worst:", rows[:3])
        nb = ea.n_big
        print("   after step: small master max|diff| %.3e, shadow max|diff| %.3e, lo max|diff| %.3e, master_stale %s" % (
            float((ea.master[nb:] - eb.master[nb:]).abs().max()), float((ea.shadow[:nb].float() - eb.shadow[:nb].float()).abs().max()),
            float((ea.shadow_lo.float() - eb.shadow_lo.float()).abs().max()), ea.master_stale))
dist.barrier(); dist.destroy_process_group()
