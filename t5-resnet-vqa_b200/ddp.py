"""Data-parallel gradient exchange for the training step (new work: the reference is single-device,
trainer/faster_rcnn_vqa_trainer.py:56,147 only logs the device count).

One process per GPU (torchrun environment), full replica per rank, batch sharded by the caller.  The backward
plan is split into segments whose parameter gradients are contiguous in the engine's flat fp32 gradient buffer
(classifier + SGA + projection, then the T5 blocks three at a time, then embeddings + small tensors); as soon as
a segment has been launched its gradient range is all-reduced (NCCL over NVLink 5 / NVSwitch, average) on NCCL's
own stream while the next segment computes.  Every rank then runs the identical clip + AdamW update, so no
parameter broadcast is needed after the first step.
"""
import os

import torch
import torch.distributed as dist


def average_range(flat, lo, hi, group=None, async_op=False):
    """All-reduce-average flat[lo:hi] in place across the group (NCCL: fused AVG; gloo: SUM then scale)."""
    view = flat[lo:hi]
    if dist.get_backend(group) == "nccl":
        return dist.all_reduce(view, op=dist.ReduceOp.AVG, group=group, async_op=async_op)
    w = dist.all_reduce(view, op=dist.ReduceOp.SUM, group=group, async_op=False)
    view.div_(dist.get_world_size(group))
    return w


def segment_ranges(segments, total):
    """Gradient ranges of the backward segments; they must tile [0, total) exactly once."""
    ranges = [(s.grad_lo, s.grad_hi) for s in segments if s.grad_hi > s.grad_lo]
    pos = 0
    for lo, hi in sorted(ranges):
        if lo != pos:
            raise RuntimeError("backward segments do not tile the gradient buffer: gap/overlap at %d" % pos)
        pos = hi
    if pos != total:
        raise RuntimeError("backward segments cover %d of %d gradient elements" % (pos, total))
    return ranges


class GradSync:
    def __init__(self, group=None):
        self.group = group
        self.world = dist.get_world_size(group)
        self.comm_stream = None

    def backward(self, eng, st):
        dev = eng.device
        cur = torch.cuda.current_stream(dev)
        if self.comm_stream is None:
            self.comm_stream = torch.cuda.Stream(device=dev)
        works = []
        for seg in st.bwd_segments:
            eng.run_plan(seg.plan)
            if seg.grad_hi <= seg.grad_lo:
                continue
            ev = torch.cuda.Event()
            ev.record(cur)
            self.comm_stream.wait_event(ev)
            with torch.cuda.stream(self.comm_stream):
                works.append(average_range(eng.grad, seg.grad_lo, seg.grad_hi, self.group, async_op=True))
        for w in works:
            w.wait()           # stream-level: the compute stream waits for NCCL, the host does not block
        cur.wait_stream(self.comm_stream)


def maybe_enable(eng):
    """Turn gradient averaging on when a process group with more than one rank exists (VQA_B200_DDP=0 opts out)."""
    if os.environ.get("VQA_B200_DDP", "1") == "0":
        return
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        if eng._ddp is None:
            eng._ddp = GradSync()
            # identical replicas: broadcast rank 0's parameters once
            with torch.no_grad():
                dist.broadcast(eng.master, src=0)
                for p in eng.model.vision_model.parameters():
                    dist.broadcast(p.data, src=0)
                for b in eng.model.vision_model.buffers():
                    dist.broadcast(b.data, src=0)
            eng.shadow_fresh = False
            eng.vision_sig = None
