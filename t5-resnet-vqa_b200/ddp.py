"""Data-parallel gradient exchange for the training step (new work: the reference is single-device,
trainer/faster_rcnn_vqa_trainer.py:56,147 only logs the device count).

One process per GPU (torchrun environment), full replica per rank, batch sharded by the caller.  The backward
plan is split into segments whose parameter gradients are contiguous in the engine's flat fp32 gradient buffer
(classifier + SGA + projection, then the T5 blocks three at a time, then embeddings + small tensors); as soon as
a segment has been launched its gradient range is all-reduced (NCCL over NVLink 5 / NVSwitch, average) on NCCL's
own stream while the next segment computes.  Every rank then runs the identical clip + AdamW update, so no
parameter broadcast is needed after the first step.

Sharded optimizer (VQA_B200_DDP_MODE=zero1, the default when a VQAFusedAdamW updates the whole engine): each backward
segment's GEMM-weight gradients are REDUCE-SCATTERED (rank r keeps the averaged r-th slice of every segment), the global
gradient norm is the sum of the ranks' partial sums, every rank runs AdamW on its slices only (1 / world of the
HBM-bound optimizer pass) and the updated bf16 weights are all-gathered behind the optimizer, on its stream, while the
next step's frozen backbone already runs.  Small tensors (biases, norms, embeddings) stay replicated and all-reduced.
Same NVLink volume as the all-reduce (reduce-scatter + all-gather), an eighth of the optimizer time at 8 GPUs.  The fp32
master copy of a GEMM weight is then current only on its owner: `state_dict()` gathers it (collective: call it on all
ranks), `.grad` of such a weight holds the averaged gradient only inside the owner's slice.  VQA_B200_DDP_MODE=allreduce
keeps full replicas.

Wire format: bf16 by default (VQA_B200_DDP_GRAD_DTYPE=fp32 restores fp32): a segment's fp32 gradients are cast into a
bf16 staging range, averaged there, and cast back into the fp32 gradient buffer, all on the communication stream - half
the NVLink and HBM traffic of the exchange (141.6 M gradients: 283 MB instead of 567 MB per step), the quantity that
bounds the step at N = 8, where the exchange must fit under ~2 ms of backward.  `.grad`, clip_grad_norm_ and AdamW see
fp32 gradients as before; every rank holds bit-identical reduced values, so replicas do not drift.
"""
import os

import torch
import torch.distributed as dist


def average_range(flat, lo, hi, group=None, async_op=False):
    """All-reduce-average flat[lo:hi] in place across the group (NCCL: fused AVG; gloo: SUM then scale)."""
    if hi <= lo:
        return None
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


def average_range_via(flat, stage, lo, hi, cast_down, cast_up, group=None):
    """flat[lo:hi] (fp32) <- average over ranks, exchanged as stage[lo:hi] (bf16): cast_down(flat_view, stage_view), all-reduce
    on the staging range, cast_up(stage_view, flat_view).  The casts are the caller's (device kernels on the GPU path)."""
    if hi <= lo:
        return
    cast_down(flat[lo:hi], stage[lo:hi])
    average_range(stage, lo, hi, group)
    cast_up(stage[lo:hi], flat[lo:hi])


class GradSync:
    def __init__(self, group=None):
        self.group = group
        self.world = dist.get_world_size(group)
        self.comm_stream = None
        self.wire = os.environ.get("VQA_B200_DDP_GRAD_DTYPE", "bf16")
        if self.wire not in ("bf16", "fp32"):
            raise ValueError("VQA_B200_DDP_GRAD_DTYPE must be bf16 or fp32")
        self.stage = None
        # The token-embedding gradient (24.7 M parameters, 99 MB fp32, 2048 non-zero rows per rank) is exchanged as rows:
        # all-gather (ids, rows) = 6 MB per rank, then an order-independent local scatter (csrc/t5misc.cu).  It is the last
        # gradient of the backward pass, so its exchange cannot overlap anything: the dense form was a third of the tail.
        self.sparse_embedding = os.environ.get("VQA_B200_DDP_SPARSE_EMBEDDING", "1") != "0"
        self._emb = None
        self.mode = os.environ.get("VQA_B200_DDP_MODE", "zero1")
        if self.mode not in ("zero1", "allreduce"):
            raise ValueError("VQA_B200_DDP_MODE must be zero1 or allreduce")
        self.rank = dist.get_rank(group)
        self.own_stage = None      # bf16 landing buffer of the reduce-scatters (this rank's slices, back to back)
        self._tmp = None
        # VQA_B200_DDP_TRACE=1: CUDA-event timeline of every backward (tools/ddp_timeline.py reads self.trace)
        self.tracing = os.environ.get("VQA_B200_DDP_TRACE", "0") == "1"
        self.trace = []

    def shards_for(self, eng, st):
        """[(lo, big_hi, own_lo, own_hi)] per backward segment with GEMM weights, or None when this step cannot be sharded
        (mode, wire format, no fused optimizer over the whole engine, or a segment that does not split evenly)."""
        opt = eng.fused_opt() if eng.fused_opt is not None else None
        if not getattr(eng, "ddp_shardable", True):
            return None
        if self.mode != "zero1" or self.wire != "bf16" or opt is None or opt.max_grad_norm is not None:
            return None
        out = []
        for seg in st.bwd_segments:
            lo, hi = seg.grad_lo, min(seg.grad_hi, eng.n_big)
            if hi <= lo:
                continue
            n = hi - lo
            if n % self.world or (n // self.world) % 8:
                return None
            s = n // self.world
            out.append((lo, hi, lo + self.rank * s, lo + (self.rank + 1) * s))
        return out

    def backward(self, eng, st):
        dev = eng.device
        cur = torch.cuda.current_stream(dev)
        if self.comm_stream is None:
            self.comm_stream = torch.cuda.Stream(device=dev)
        if self.wire == "bf16" and (self.stage is None or self.stage.numel() != eng.total or self.stage.device != dev):
            self.stage = torch.empty(eng.total, dtype=torch.bfloat16, device=dev)
        rec = eng.rec(None)
        works = []
        tr = None
        if self.tracing:
            tr = dict(start=torch.cuda.Event(enable_timing=True), seg=[], comm=[])
            tr["start"].record(cur)
            self.trace.append(tr)
        emb = eng.embedding_param()
        emb_lo = eng.offs[id(emb)]
        shards = self.shards_for(eng, st)
        eng.ddp_shards = shards              # read by clip_grad_norm_ and VQAFusedAdamW.step for THIS step
        if shards is not None:
            need = sum(s[3] - s[2] for s in shards)
            if self.own_stage is None or self.own_stage.numel() != need or self.own_stage.device != dev:
                self.own_stage = torch.empty(need, dtype=torch.bfloat16, device=dev)
        shard_of = {s[0]: (i, s) for i, s in enumerate(shards or [])}
        own_off = 0
        for seg in st.bwd_segments:
            eng.run_plan(seg.plan)
            lo, hi = seg.grad_lo, seg.grad_hi
            if hi <= lo:
                continue
            rows = getattr(st, "emb_rows", None) if (lo <= emb_lo < hi) else None
            if rows is not None:
                hi = emb_lo          # the table itself is not exchanged (it is the last tensor of the flat buffer)
            sh = shard_of.get(lo)
            ev = torch.cuda.Event(enable_timing=self.tracing)
            ev.record(cur)
            self.comm_stream.wait_event(ev)
            if tr is not None:
                tr["seg"].append(ev)
                c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                tr["comm"].append((c0, c1, 4 * (hi - lo)))
                c0.record(self.comm_stream)
            with torch.cuda.stream(self.comm_stream):
                if sh is not None:
                    # sharded: reduce-scatter the GEMM-weight gradients (this rank receives the average of its slice), the
                    # replicated small tensors behind them (last segment only) are all-reduced
                    _, (_, bhi, olo, ohi) = sh
                    n_own = ohi - olo
                    rec.cast_f32_bf16(eng.grad.data_ptr() + 4 * lo, self.stage.data_ptr() + 2 * lo, hi - lo)
                    dist.reduce_scatter_tensor(self.own_stage[own_off:own_off + n_own], self.stage[lo:bhi],
                                               op=dist.ReduceOp.AVG, group=self.group)
                    rec.cast_bf16_f32(self.own_stage.data_ptr() + 2 * own_off, eng.grad.data_ptr() + 4 * olo, n_own)
                    own_off += n_own
                    if hi > bhi:
                        average_range(self.stage, bhi, hi, self.group)
                        rec.cast_bf16_f32(self.stage.data_ptr() + 2 * bhi, eng.grad.data_ptr() + 4 * bhi, hi - bhi)
                elif self.wire == "bf16":
                    # cast -> all-reduce -> cast back, in order on the communication stream (NCCL's own stream is ordered
                    # after the cast by the launch and before the cast back by the stream-level wait)
                    rec.cast_f32_bf16(eng.grad.data_ptr() + 4 * lo, self.stage.data_ptr() + 2 * lo, hi - lo)
                    average_range(self.stage, lo, hi, self.group, async_op=True).wait()
                    rec.cast_bf16_f32(self.stage.data_ptr() + 2 * lo, eng.grad.data_ptr() + 4 * lo, hi - lo)
                else:
                    w = average_range(eng.grad, lo, hi, self.group, async_op=True)
                    if tr is not None:
                        w.wait()      # so that the end event below marks the end of this all-reduce
                    else:
                        works.append(w)
                if rows is not None:
                    self._exchange_embedding_rows(eng, st, rows, emb)
                if tr is not None:
                    tr["comm"][-1][1].record(self.comm_stream)
        for w in works:
            w.wait()           # stream-level: the compute stream waits for NCCL, the host does not block
        cur.wait_stream(self.comm_stream)
        if tr is not None:
            tr["done"] = torch.cuda.Event(enable_timing=True)
            tr["done"].record(cur)


def _all_gather_slices(self, flat, shards):
    """In place: every rank contributes its slice of each shard range of `flat`, all ranks end with the full ranges.  The
    slice is staged in a scratch buffer (no aliasing between NCCL's send and receive buffers)."""
    for lo, bhi, olo, ohi in shards:
        n = ohi - olo
        if self._tmp is None or self._tmp.numel() < 4 * n or self._tmp.device != flat.device:
            self._tmp = torch.empty(4 * n, dtype=torch.uint8, device=flat.device)
        tmp = self._tmp[:n * flat.element_size()].view(flat.dtype)
        tmp.copy_(flat[olo:ohi])
        dist.all_gather_into_tensor(flat[lo:bhi], tmp, group=self.group)


def _after_update(self, eng, stream):
    """Called by VQAFusedAdamW.step behind the sharded update, on the optimizer stream: low-order halves of this rank's
    slices, then all-gather of the bf16 weights (and low-order halves) segment by segment in the order the forward needs
    them (T5 block 0's segment first, the SGA stack's last), with one event per segment: the next forward starts on the
    first segments while the later ones are still in flight (engine.forward)."""
    shards = eng.ddp_shards
    ctx = torch.cuda.stream(stream) if stream is not None else _Null()
    events = [None] * len(shards)
    with ctx:
        cur = torch.cuda.current_stream(eng.device)
        rec = eng.rec(None)
        for i in reversed(range(len(shards))):       # backward order reversed = forward order
            sh = shards[i]
            lo, bhi, olo, ohi = sh
            self._all_gather_slices(eng.shadow, [sh])
            if any(lo < r1 and r0 < bhi for r0, r1 in eng.lo_ranges):
                rec.split_lo_bf16(eng.master.data_ptr() + 4 * olo, eng.shadow_lo.data_ptr() + 2 * olo, ohi - olo)
                self._all_gather_slices(eng.shadow_lo, [sh])
            events[i] = torch.cuda.Event()
            events[i].record(cur)
    eng.shard_events = events
    eng.lo_fresh = True
    eng.master_stale = True
    eng.master_shards = list(shards)


def _sync_master(self, eng):
    """Collective: gather the owners' fp32 master slices so that every rank holds current values of every parameter."""
    shards = getattr(eng, "master_shards", None)
    if not shards:
        return
    eng.wait_optimizer()
    self._all_gather_slices(eng.master, shards)
    eng.master_stale = False


class _Null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def _gather_embedding(self, eng, st, rows, emb):
    """(current stream = communication stream) all-gather every rank's token ids and gradient rows, then scatter them into
    this rank's zeroed table gradient with an order-independent kernel: bit-identical on all ranks."""
    M, D = rows.shape
    T = M * self.world
    if self._emb is None or self._emb[0].shape[0] != T or self._emb[0].device != rows.device:
        self._emb = (torch.empty(T, D, dtype=torch.float32, device=rows.device),
                     torch.empty(T, dtype=torch.int64, device=rows.device),
                     torch.empty(2 * emb.shape[0], dtype=torch.int32, device=rows.device),
                     torch.empty(T * D, dtype=torch.int64, device=rows.device))
    rows_all, ids_all, first, acc = self._emb
    dist.all_gather_into_tensor(ids_all, st.ids.reshape(-1), group=self.group)
    dist.all_gather_into_tensor(rows_all, rows, group=self.group)
    eng.rec(None).embedding_scatter_ordered(ids_all, rows_all, eng.gp(emb), first, acc, T, D, emb.shape[0])


GradSync._exchange_embedding_rows = _gather_embedding
GradSync._all_gather_slices = _all_gather_slices
GradSync.after_update = _after_update
GradSync.sync_master = _sync_master


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
            eng._seed_rank = dist.get_rank()     # every replica draws its own dropout masks
            eng._reseed()
