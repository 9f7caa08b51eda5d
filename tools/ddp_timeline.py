"""Where does the data-parallel step lose time?  CUDA-event timeline of the backward pass under torchrun (>= 2 GPUs): when
each backward segment ends on the compute stream, when each gradient exchange starts / ends on the communication stream,
and when the compute stream has both (VQA_B200_DDP_TRACE=1 in ddp.GradSync), next to the same step without any exchange.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/ddp_timeline.py
"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["VQA_B200_DDP_TRACE"] = "1"
os.environ.setdefault("VQA_B200_PRETRAINED", "0")
import bench  # noqa: E402


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import t5_resnet_vqa_b200 as pkg
    torch.manual_seed(0)
    model = pkg.ResnetVQAModel("resnet50", "t5-base", answer_spaces=170).to(dev).train()
    B = 64
    g = torch.Generator().manual_seed(1 + rank)
    batch = dict(question_input_ids=torch.randint(2, 32100, (B, 32), generator=g).to(dev),
                 question_attention_masks=torch.ones(B, 32, dtype=torch.long, device=dev),
                 annotation_ids=torch.randint(0, 170, (B,), generator=g).to(dev),
                 image_tensors=torch.rand(B, 3, 224, 224, generator=g).to(dev))
    opt, sched = bench.build_trainer_objects(model, 200)
    marks = []
    for i in range(18):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        ev[0].record()
        opt.zero_grad()
        logp, loss = model(**batch)
        ev[1].record()
        loss.backward()
        ev[2].record()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
        ev[3].record()
        opt.step()
        sched.step()
        ev[4].record()
        marks.append(ev)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    if rank == 0:
        sync = model._engine._ddp
        steps = list(range(8, 18))
        n = len(steps)

        def avg(f):
            return sum(f(i) for i in steps) / n
        out = {"world": world, "wire": sync.wire if sync is not None else None,
               "step_ms": avg(lambda i: marks[i][0].elapsed_time(marks[i + 1][0]) if i + 1 < len(marks) else marks[i][0].elapsed_time(marks[i][4])),
               "forward_ms": avg(lambda i: marks[i][0].elapsed_time(marks[i][1])),
               "backward_incl_exchange_ms": avg(lambda i: marks[i][1].elapsed_time(marks[i][2])),
               "clip_ms": avg(lambda i: marks[i][2].elapsed_time(marks[i][3])),
               "optimizer_enqueue_ms": avg(lambda i: marks[i][3].elapsed_time(marks[i][4]))}
        nseg = len(sync.trace[-1]["seg"]) if sync is not None else 0
        out["segments"] = []
        for k in range(nseg):
            out["segments"].append({
                "MB_fp32": sync.trace[-1]["comm"][k][2] / 1e6,
                "compute_end_ms": avg(lambda i: sync.trace[i]["start"].elapsed_time(sync.trace[i]["seg"][k])),
                "exchange_start_ms": avg(lambda i: sync.trace[i]["start"].elapsed_time(sync.trace[i]["comm"][k][0])),
                "exchange_end_ms": avg(lambda i: sync.trace[i]["start"].elapsed_time(sync.trace[i]["comm"][k][1]))})
        if sync is not None:
            out["backward_done_ms"] = avg(lambda i: sync.trace[i]["start"].elapsed_time(sync.trace[i]["done"]))
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
