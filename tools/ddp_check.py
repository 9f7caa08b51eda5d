"""N-rank data-parallel step == 1-rank step on the concatenated batch (run with torchrun on >= 2 GPUs).

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/ddp_check.py

Every rank runs forward + backward on its shard with the gradient exchange on; rank 0 then runs the whole batch on one GPU
with the exchange off and compares the loss (mean of the rank losses) and every gradient tensor.  The check is also run
by tests/test_ddp_gpu.py (skipped below 2 GPUs) and by bench.py --gpus N (key "ddp_equivalence" of the bench line).
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# per-tensor relative gradient difference: bf16 wire format (2^-9 per element before averaging) on top of bf16 compute
LOSS_REL, GRAD_REL = 2e-3, 3e-2


def _state_dict(pkg, kind):
    """Deterministic weights from the PACKAGE's own initialisation under a fixed seed (identical on every rank; nothing under
    oracle/ is touched: bench.py --gpus N runs this check).  The last BatchNorm gain of every residual block is scaled down as
    in trained networks, so that sixteen residual additions keep the feature map O(1) and the guided attention stays soft."""
    torch.manual_seed(0)
    if kind == "vit":
        m = pkg.VitVQAModel("google/vit-base-patch16-224-in21k", "t5-base", 170)
    else:
        m = pkg.ResnetVQAModel(kind, "t5-base", 170)
        with torch.no_grad():
            for layer in (m.vision_model.layer1, m.vision_model.layer2, m.vision_model.layer3, m.vision_model.layer4):
                for blk in layer:
                    (blk.bn3 if hasattr(blk, "bn3") else blk.bn2).weight.mul_(0.3)
    return {k: v.detach().clone() for k, v in m.state_dict().items()}


def _batch(n, L, H, W, masked_tail, seed=1, vit=False, Ld=20):
    g = torch.Generator().manual_seed(seed)
    b = dict(question_input_ids=torch.randint(2, 32100, (n, L), generator=g),
             question_attention_masks=torch.ones(n, L, dtype=torch.long),
             annotation_ids=torch.randint(0, 170, (n,), generator=g))
    b["question_attention_masks"][:, L - masked_tail:] = 0
    if vit:
        b["pixel_values"] = torch.rand(n, 3, H, W, generator=g) * 2 - 1
        b["decoder_question_input_ids"] = torch.randint(2, 32100, (n, Ld), generator=g)
        b["decoder_question_attention_masks"] = torch.ones(n, Ld, dtype=torch.long)
        for i in range(n):
            b["decoder_question_attention_masks"][i, 3 + (5 * i) % (Ld - 2):] = 0
    else:
        b["image_tensors"] = torch.rand(n, 3, H, W, generator=g)
    return b


def check(dev, rank, world, per=4, vision="resnet18"):
    """Returns (on rank 0) dict(world, loss_mean_of_ranks, loss_single, worst_grad_rel_diff, identical_across_ranks, ok)."""
    os.environ.setdefault("VQA_B200_PRETRAINED", "0")
    import t5_resnet_vqa_b200 as pkg
    sd = _state_dict(pkg, vision)
    full = _batch(per * world, 16, 64, 64, 3)

    def run(model, batch):
        kw = {k: v.to(dev) for k, v in batch.items()}
        logp, loss = model(kw["question_input_ids"], None, kw["question_attention_masks"], None, kw["annotation_ids"],
                           kw["image_tensors"])
        loss.backward()
        torch.cuda.synchronize()
        return float(loss), {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}

    m = pkg.ResnetVQAModel(vision, "t5-base", 170)
    m.load_state_dict(sd)
    m.to(dev).eval()
    shard = {k: v[rank * per:(rank + 1) * per] for k, v in full.items()}
    loss_r, g_ddp = run(m, shard)
    if m._engine._ddp is None:
        raise RuntimeError("gradient sync was not enabled")
    t = torch.tensor([loss_r], device=dev)
    dist.all_reduce(t)
    loss_mean = float(t) / world
    # replicas must hold bit-identical reduced gradients (else they drift apart step by step)
    flat = m._engine.grad
    hi, lo = flat.clone(), flat.clone()
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    identical = bool(torch.equal(hi, lo))
    res = None
    if rank == 0:
        old = os.environ.get("VQA_B200_DDP")
        os.environ["VQA_B200_DDP"] = "0"      # the single-process run must not join any collective
        try:
            m1 = pkg.ResnetVQAModel(vision, "t5-base", 170)
            m1.load_state_dict(sd)
            m1.to(dev).eval()
            loss_1, g_1 = run(m1, full)
        finally:
            if old is None:
                del os.environ["VQA_B200_DDP"]
            else:
                os.environ["VQA_B200_DDP"] = old
        scale = max(float(v.norm()) for v in g_1.values())
        worst = 0.0
        for k in g_1:
            if float(g_1[k].norm()) > 1e-6 * scale:
                worst = max(worst, float((g_ddp[k] - g_1[k]).norm() / (g_1[k].norm() + 1e-12)))
        res = dict(world=world, wire=m._engine._ddp.wire, loss_mean_of_ranks=loss_mean, loss_single=loss_1,
                   worst_grad_rel_diff=worst, identical_across_ranks=identical,
                   ok=bool(abs(loss_mean - loss_1) < LOSS_REL * abs(loss_1) and worst < GRAD_REL and identical))
    dist.barrier()
    return res


def check_vit(dev, rank, world, per=2):
    """The same gradient equivalence for VitVQAModel (replicas, one all-reduce of the flat gradient behind its backward)."""
    os.environ.setdefault("VQA_B200_PRETRAINED", "0")
    import t5_resnet_vqa_b200 as pkg
    sd = _state_dict(pkg, "vit")
    full = _batch(per * world, 16, 224, 224, 3, vit=True)

    def run(model, batch):
        kw = {k: v.to(dev) for k, v in batch.items()}
        logp, loss = model(**kw)
        loss.backward()
        torch.cuda.synchronize()
        return float(loss), {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}

    def build():
        m = pkg.VitVQAModel("google/vit-base-patch16-224-in21k", "t5-base", 170)
        m.load_state_dict(sd)
        return m.to(dev).eval()
    m = build()
    loss_r, g_ddp = run(m, {k: v[rank * per:(rank + 1) * per] for k, v in full.items()})
    if m._engine._ddp is None:
        raise RuntimeError("gradient sync was not enabled")
    t = torch.tensor([loss_r], device=dev)
    dist.all_reduce(t)
    loss_mean = float(t) / world
    flat = m._engine.grad
    hi, lo = flat.clone(), flat.clone()
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    identical = bool(torch.equal(hi, lo))
    res = None
    if rank == 0:
        old = os.environ.get("VQA_B200_DDP")
        os.environ["VQA_B200_DDP"] = "0"
        try:
            loss_1, g_1 = run(build(), full)
        finally:
            if old is None:
                del os.environ["VQA_B200_DDP"]
            else:
                os.environ["VQA_B200_DDP"] = old
        scale = max(float(v.norm()) for v in g_1.values())
        worst = 0.0
        for k in g_1:
            if float(g_1[k].norm()) > 1e-6 * scale:
                worst = max(worst, float((g_ddp[k] - g_1[k]).norm() / (g_1[k].norm() + 1e-12)))
        res = dict(model="VitVQAModel", world=world, wire=m._engine._ddp.wire, loss_mean_of_ranks=loss_mean, loss_single=loss_1,
                   worst_grad_rel_diff=worst, identical_across_ranks=identical,
                   ok=bool(abs(loss_mean - loss_1) < LOSS_REL * abs(loss_1) and worst < GRAD_REL and identical))
    dist.barrier()
    return res


def check_training(dev, rank, world, per=4, vision="resnet18", steps=3):
    """`steps` training steps (dropout off, VQAFusedAdamW + clip, lr 1e-4) three ways: `world` ranks with the sharded optimizer,
    `world` ranks with full replicas (all-reduce), and one GPU over the concatenated batch.  The two data-parallel runs see
    the same gradient arithmetic, so their parameter updates must agree tightly; against the single-GPU run the bar is the
    bf16 wire format's.  Returns (rank 0) a dict with the worst tensors of both comparisons."""
    os.environ.setdefault("VQA_B200_PRETRAINED", "0")
    import t5_resnet_vqa_b200 as pkg
    sd = _state_dict(pkg, vision)
    full = _batch(per * world, 16, 64, 64, 3)

    def train(batch, env):
        old = {k: os.environ.get(k) for k in env}
        os.environ.update(env)
        try:
            m = pkg.ResnetVQAModel(vision, "t5-base", 170)
            m.load_state_dict(sd)
            m.to(dev).eval()
            opt = torch.optim.VQAFusedAdamW(m.parameters(), lr=1e-4, weight_decay=0.1, amsgrad=True)
            kw = {k: v.to(dev) for k, v in batch.items()}
            losses = []
            for _ in range(steps):
                opt.zero_grad()
                _, loss = m(kw["question_input_ids"], None, kw["question_attention_masks"], None, kw["annotation_ids"],
                            kw["image_tensors"])
                loss.backward()
                torch.nn.utils.clip_grad_norm_(m.parameters(), 1.0)
                opt.step()
                losses.append(float(loss))
            sharded = bool(m._engine.master_stale)
            state = None
            if env.get("VQA_B200_DDP") != "0":
                state = {k: v.detach().float().cpu().clone() for k, v in m.state_dict().items()}   # collective (gathers)
            else:
                state = {k: v.detach().float().cpu().clone() for k, v in m.state_dict().items()}
            return losses, state, sharded
        finally:
            for k, v in old.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v

    def mean_losses(losses):
        t = torch.tensor(losses, device=dev)
        dist.all_reduce(t)
        return (t / world).tolist()

    shard = {k: v[rank * per:(rank + 1) * per] for k, v in full.items()}
    l_z, s_z, sharded = train(shard, {"VQA_B200_DDP_MODE": "zero1"})
    l_z = mean_losses(l_z)
    l_a, s_a, sharded_a = train(shard, {"VQA_B200_DDP_MODE": "allreduce"})
    l_a = mean_losses(l_a)

    def compare(a, b):
        rows = []
        for k, v in b.items():
            if not v.is_floating_point() or k.startswith("vision_model."):
                continue
            if k.endswith("linear_k.bias") or k == "attention_pooler.attention.0.bias":
                continue        # mathematically zero gradients: Adam normalises rounding noise
            d0, d1 = v - sd[k].float(), a[k] - sd[k].float()
            if float(d0.norm()) == 0.0:
                continue
            c = float((d1.flatten().double() @ d0.flatten().double()) / (d1.norm().double() * d0.norm().double() + 1e-300))
            rows.append((c, float((d1 - d0).norm() / d0.norm()), k))
        rows.sort()
        return rows
    res = None
    if rank == 0:
        l_1, s_1, _ = train(full, {"VQA_B200_DDP": "0"})
        za, z1 = compare(s_z, s_a), compare(s_z, s_1)
        lrel = max(abs(a - b) / abs(b) for a, b in zip(l_z, l_1))
        res = dict(world=world, sharded_optimizer=sharded, replicas_run_sharded=sharded_a, losses_sharded=l_z,
                   losses_replicas=l_a, losses_single=l_1, sharded_vs_replicas_worst=za[:4], sharded_vs_single_worst=z1[:4],
                   # (the sharded optimizer needs the bf16 wire format: with fp32 on the wire both runs are replicas)
                   ok=bool(sharded == (os.environ.get("VQA_B200_DDP_GRAD_DTYPE", "bf16") == "bf16") and not sharded_a
                           and za[0][0] > 0.999 and lrel < 2e-2 and z1[0][0] > 0.97))
    dist.barrier()
    return res


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    if os.environ.get("VQA_DDP_CHECK_MODEL") == "vit":
        res = check_vit(dev, rank, world)
        if rank == 0:
            print("ddp_check vit: %s" % res)
            print("ddp_check OK" if res["ok"] else "ddp_check FAILED")
        dist.destroy_process_group()
        if rank == 0 and not res["ok"]:
            sys.exit(1)
        return
    res = check(dev, rank, world)
    res2 = check_training(dev, rank, world)
    if rank == 0:
        print("ddp_check: %s" % res)
        print("ddp_check training: %s" % res2)
        try:
            os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
            with open(os.path.join(ROOT, "gpurun_out", "ddp_check_%s_%s.log" % (
                    os.environ.get("VQA_B200_DDP_MODE", "zero1"), os.environ.get("VQA_B200_DDP_GRAD_DTYPE", "bf16"))), "w") as f:
                f.write("%s\n%s\n" % (res, res2))
        except OSError:
            pass
        print("ddp_check OK" if res["ok"] and res2["ok"] else "ddp_check FAILED")
    dist.destroy_process_group()
    if rank == 0 and not (res["ok"] and res2["ok"]):
        sys.exit(1)


if __name__ == "__main__":
    main()
