"""Execution engine of the B200-native ResnetVQAModel step.

Host side only: owns the flat parameter / gradient / bf16-shadow buffers, the activation workspace and the
recorded launch plans (one per input shape and mode), and replays them through libvqa_b200.so.  All arithmetic
happens in the library's sm_100a kernels; torch is used for device memory, streams and autograd plumbing.

Reference path being replaced: ResnetVQAModel.forward (model/resnet_vqa_model.py:101-165) and its autograd
backward as driven by train_one_step (trainer/faster_rcnn_vqa_trainer.py:391-406).
"""
import ctypes
import math
import os

import torch

from . import lib as L

ALIGN = 64  # elements; every parameter starts on a 256-byte boundary of the flat fp32 buffers


def _align(n, a=ALIGN):
    return (n + a - 1) // a * a


def _env_flag(name, default):
    v = os.environ.get(name)
    if v is None:
        return default
    return v not in ("0", "false", "False", "")


def t5_relative_buckets(L_q, L_k, num_buckets=32, max_distance=128):
    """Bidirectional relative-position buckets, evaluated exactly as transformers does (hf:189-234), int32 [Lq, Lk]."""
    ctx = torch.arange(L_q, dtype=torch.long)[:, None]
    mem = torch.arange(L_k, dtype=torch.long)[None, :]
    rel = mem - ctx
    nb = num_buckets // 2
    buckets = (rel > 0).to(torch.long) * nb
    rel = torch.abs(rel)
    max_exact = nb // 2
    is_small = rel < max_exact
    large = max_exact + (torch.log(rel.float() / max_exact) / math.log(max_distance / max_exact)
                         * (nb - max_exact)).to(torch.long)
    large = torch.min(large, torch.full_like(large, nb - 1))
    buckets = buckets + torch.where(is_small, rel, large)
    return buckets.to(torch.int32)


class _Rec:
    """Thin typed front of the C ABI: appends launches to `plan` (or runs them now when plan is None)."""

    KS_WS_BYTES = 148 * 8 * 128 * 32 * 4   # largest cluster split-K workspace any launch can need (19.4 MB)

    def __init__(self, lib, plan, stream_fn, ws_store=None):
        self.lib, self.plan, self.stream_fn = lib, plan, stream_fn
        self._lane = 0
        self.wgrad_ctas = int(os.environ.get("VQA_B200_WGRAD_CTAS", "0"))
        # cluster split-K workspaces, one per lane: launches of different lanes / plans may run concurrently and
        # must not share one.  ws_store (a dict owned by the plan's State) keeps them alive with the plan.
        self._ws = ws_store if ws_store is not None else {}

    def _ks_workspace(self):
        key = (self.plan, self._lane)
        t = self._ws.get(key)
        if t is None:
            t = torch.empty(self.KS_WS_BYTES, dtype=torch.uint8, device="cuda")
            self._ws[key] = t
        return t

    def _s(self):
        return None if self.plan is not None else self.stream_fn()

    def _call(self, name, *args):
        L.check(getattr(self.lib, name)(self.plan, *args, self._s()), name)

    # ---- lanes (parallel branches of a plan) ----
    def lane(self, n):
        self._lane = n
        if self.plan is not None:
            L.check(self.lib.vqa_plan_set_lane(self.plan, n), "plan_set_lane")

    def fork(self):
        if self.plan is not None:
            L.check(self.lib.vqa_plan_fork(self.plan), "plan_fork")

    def join(self):
        if self.plan is not None:
            L.check(self.lib.vqa_plan_join(self.plan), "plan_join")

    def mark(self):
        """Id of lane 1's current position (None when launches run immediately)."""
        if self.plan is None:
            return None
        mid = self.lib.vqa_plan_mark(self.plan)
        if mid < 0:
            raise RuntimeError("vqa_plan_mark failed")
        return mid

    def wait(self, mid):
        if self.plan is not None and mid is not None:
            L.check(self.lib.vqa_plan_wait(self.plan, mid), "plan_wait")

    # ---- contractions ----
    def gemm(self, M, N, K, A, lda, a_mn, B, ldb, b_mn, out, ldo, out_fp32, bias=None, relu=0, relu_mask=None,
             ldm=0, drop_p=0.0, sid=0, rng=None, residual=None, ldr=0, res_fp32=1, alpha=1.0, accumulate=0,
             bn=None, split_k=1, pair=None, ksplit=None, b_lo=None, a_lo_col=0, max_ctas=0):
        nseg = 1 if b_lo is None else (3 if a_lo_col else 2)   # two-term operand split: the k-loop is nseg times as long
        if bn is None:
            bn, ks = pick_tile(M, N, K * nseg, allow_ksplit=not accumulate and split_k == 1 and not pair and nseg == 1)
            if ksplit is None:
                ksplit = ks
        ksplit = int(ksplit or 1)
        a = L.GemmArgs()
        a.M, a.N, a.K = M, N, K
        a.A, a.lda, a.a_mn = L.ptr(A), lda, a_mn
        a.B, a.ldb, a.b_mn = L.ptr(B), ldb, b_mn
        a.out, a.ldo, a.out_fp32 = L.ptr(out), ldo, int(out_fp32)
        a.bias, a.relu = L.ptr(bias), int(relu)
        a.relu_mask, a.ldm = L.ptr(relu_mask), ldm
        a.drop_p, a.drop_sid, a.rng = float(drop_p), sid, L.ptr(rng) if drop_p > 0 else None
        a.residual, a.ldr, a.res_fp32, a.res_first = L.ptr(residual), ldr, int(res_fp32), 0
        a.alpha, a.accumulate, a.bn, a.split_k = alpha, int(accumulate), bn, split_k
        a.cta_pair = int(bool(pair)) if pair is not None else 0
        a.ksplit = ksplit
        if ksplit > 1:
            ws = self._ks_workspace()
            a.ks_ws, a.ks_ws_bytes = ws.data_ptr(), ws.numel()
        a.B_lo, a.a_lo_col = L.ptr(b_lo), int(a_lo_col)
        a.max_ctas = int(max_ctas)
        L.check(self.lib.vqa_gemm_bf16(self.plan, ctypes.byref(a), self._s()), "gemm")

    def linear(self, X, M, K, ldx, W, N, out, ldo, out_fp32=0, **kw):
        """out[M,N] = epi(X[M,K] @ W[N,K]^T)"""
        self.gemm(M, N, K, X, ldx, 0, W, K, 0, out, ldo, out_fp32, **kw)

    def dgrad(self, dY, M, N, ldy, W, K, out, ldo, out_fp32=0, **kw):
        """out[M,K] = epi(dY[M,N] @ W[N,K])"""
        self.gemm(M, K, N, dY, ldy, 0, W, K, 1, out, ldo, out_fp32, **kw)

    def wgrad(self, dY, M, N, ldy, X, K, ldx, dW, **kw):
        """dW[N,K] (fp32) = dY[M,N]^T @ X[M,K]"""
        if self.wgrad_ctas and self._lane == 1 and "max_ctas" not in kw:
            kw = dict(kw, max_ctas=self.wgrad_ctas)      # side-lane weight gradients leave SMs to the main lane's chain
        if not kw:
            # few output tiles, deep contraction over the tokens: split K over CTAs (fp32 red.add into a zeroed dW)
            bn, split = pick_wgrad_split(N, K, M)
            if split > 1:
                self.memset_zero(dW, 4 * N * K)
                kw = dict(bn=bn, split_k=split)
        self.gemm(N, K, M, dY, ldy, 1, X, ldx, 1, dW, K, 1, **kw)

    def conv(self, N, H, W, Cin, Cout, R, stride, pad, x, w, out, bias=None, residual=None, relu=1, stem7=0,
             out_fp32=0, bn=None, pair=None, ksplit=None):
        Ho = (H + 2 * pad - R) // stride + 1
        Wo = (W + 2 * pad - R) // stride + 1
        if bn is None:
            bn = pick_conv_tile(N * Ho * Wo, Cout)
            if pair and bn < 128:
                bn = 128
        ksplit = int(ksplit or 1)
        a = L.ConvArgs()
        a.N, a.H, a.W, a.Cin, a.Cout, a.R, a.S = N, H, W, Cin, Cout, R, R
        a.stride, a.pad, a.Ho, a.Wo, a.stem7 = stride, pad, Ho, Wo, stem7
        a.x, a.w, a.out, a.out_fp32 = L.ptr(x), L.ptr(w), L.ptr(out), out_fp32
        a.bias, a.residual, a.relu, a.bn = L.ptr(bias), L.ptr(residual), int(relu), bn
        a.cta_pair = int(bool(pair)) if pair is not None else 0
        a.ksplit = ksplit
        if ksplit > 1:
            ws = self._ks_workspace()
            a.ks_ws, a.ks_ws_bytes = ws.data_ptr(), ws.numel()
        L.check(self.lib.vqa_conv2d_bf16(self.plan, ctypes.byref(a), self._s()), "conv2d")
        return Ho, Wo

    def conv_wgrad(self, N, H, W, Cin, Cout, dy, x, dw, bn, split_k, pair=None):
        a = L.ConvWgradArgs()
        a.N, a.H, a.W, a.Cin, a.Cout, a.R, a.S, a.pad = N, H, W, Cin, Cout, 3, 3, 1
        a.dy, a.x, a.dw, a.bn, a.split_k = L.ptr(dy), L.ptr(x), L.ptr(dw), bn, split_k
        a.cta_pair = int(bool(pair)) if pair is not None else 0
        L.check(self.lib.vqa_conv2d_wgrad_bf16(self.plan, ctypes.byref(a), self._s()), "conv2d_wgrad")

    def attn_fwd(self, B, H, Lq, Lk, hd, q, ldq, k, ldk, v, ldv, out, ldo, probs, bias, key_mask, scale, drop_p,
                 sid, rng, stats=None):
        a = L.AttnFwdArgs()
        a.B, a.H, a.Lq, a.Lk, a.hd = B, H, Lq, Lk, hd
        a.q, a.ldq, a.k, a.ldk, a.v, a.ldv = L.ptr(q), ldq, L.ptr(k), ldk, L.ptr(v), ldv
        a.out, a.ldo, a.probs, a.bias, a.key_mask = L.ptr(out), ldo, L.ptr(probs), L.ptr(bias), L.ptr(key_mask)
        a.scale, a.drop_p, a.sid, a.rng = scale, float(drop_p), sid, L.ptr(rng) if drop_p > 0 else None
        a.stats = L.ptr(stats)
        L.check(self.lib.vqa_attention_fwd(self.plan, ctypes.byref(a), self._s()), "attention_fwd")

    def attn_bwd(self, B, H, Lq, Lk, hd, q, ldq, k, ldk, v, ldv, probs, dout, ldo, dq, lddq, dk, lddk, dv, lddv,
                 dbias, scale, drop_p, sid, rng, stats=None, bias=None, key_mask=None):
        a = L.AttnBwdArgs()
        a.B, a.H, a.Lq, a.Lk, a.hd = B, H, Lq, Lk, hd
        a.q, a.ldq, a.k, a.ldk, a.v, a.ldv = L.ptr(q), ldq, L.ptr(k), ldk, L.ptr(v), ldv
        a.probs, a.dout, a.ldo = L.ptr(probs), L.ptr(dout), ldo
        a.dq, a.lddq, a.dk, a.lddk, a.dv, a.lddv = L.ptr(dq), lddq, L.ptr(dk), lddk, L.ptr(dv), lddv
        a.dbias, a.scale, a.drop_p, a.sid = L.ptr(dbias), scale, float(drop_p), sid
        a.rng = L.ptr(rng) if drop_p > 0 else None
        a.stats, a.bias, a.key_mask = L.ptr(stats), L.ptr(bias), L.ptr(key_mask)
        L.check(self.lib.vqa_attention_bwd(self.plan, ctypes.byref(a), self._s()), "attention_bwd")

    # ---- everything else: positional passthrough ----
    def __getattr__(self, name):
        fn_name = "vqa_" + name
        if fn_name not in L.SIGNATURES or fn_name.startswith("vqa_plan_"):
            raise AttributeError(name)

        def call(*args):
            self._call(fn_name, *[L.ptr(a) if isinstance(a, torch.Tensor) else a for a in args])
        return call


N_SM = 148


def _tile_cost(tiles, bn):
    """Relative time of a persistent launch: waves over the SMs x per-tile cost.  A tile is bound by the bytes its
    SM pulls from L2 per k-block (128 A rows + bn B rows) plus a fixed part (pipeline fill, epilogue tail); fitted
    to tools/gemm_bench.py on B200 (it reproduces the measured best width on all 13 GEMM and 14 conv shapes)."""
    waves = (tiles + N_SM - 1) // N_SM
    return waves * (bn + int(os.environ.get("VQA_B200_TILE_FIXED", "160")))


# Launch-time model of the tcgen05 GEMM on B200 (microseconds; tools/gemm_bench.py, graph-replayed, L2-warm):
# t = fixed(bn) + waves * k_blocks_per_CTA * per_kblock(bn) [+ exchange when a cluster splits K].  It reproduces the
# measured best configuration of the step's 13 GEMM shapes to within 10 %.
_GEMM_FIXED = {64: 3.2, 128: 3.3, 256: 4.0}
_GEMM_KB = {64: 0.117, 128: 0.175, 256: 0.278}
_GEMM_KSPLIT_EXCHANGE = 1.3


def pick_tile(M, N, K, allow_ksplit=True):
    """(tile width, cluster split-K factor) of the persistent 128 x bn tcgen05 GEMM (one CTA per SM)."""
    tm = (M + 127) // 128
    kb = (K + 63) // 64
    # Cluster split-K is opt-in: on the isolated K = 3072 / 2304 GEMMs it is 5..10 % faster (tools/gemm_bench.py), but inside
    # the step, where two plan lanes keep the SMs busy, co-scheduling whole clusters costs more than the shorter k-loops
    # save (bench.py on B200: 6.68 ms/step with it, 6.17 without).
    use_ks = allow_ksplit and _env_flag("VQA_B200_KSPLIT", False)
    best = None
    for bn in (256, 128, 64):
        if bn > 64 and N <= bn // 2:
            continue
        tiles = tm * ((N + bn - 1) // bn)
        waves = (tiles + N_SM - 1) // N_SM
        cands = [(1, _GEMM_FIXED[bn] + waves * kb * _GEMM_KB[bn])]
        if use_ks:
            for ks in (2, 3, 4):
                if tiles * ks <= N_SM and kb >= 4 * ks:
                    cands.append((ks, _GEMM_FIXED[bn] + -(-kb // ks) * _GEMM_KB[bn] + _GEMM_KSPLIT_EXCHANGE))
        for ks, cost in cands:
            if best is None or cost < best[0] - 1e-9:
                best = (cost, bn, ks)
    return best[1], best[2]


def pick_wgrad_split(N, K, M):
    """(tile width, split-K factor) for a weight gradient dW[N,K] over M tokens, or (None, 1) to leave the choice to
    pick_tile.  Same launch model; the split variant also pays for zeroing dW (~0.5 us per MB + a graph node).
    Opt-in (VQA_B200_WGRAD_SPLIT=1): the isolated launch gets faster, the step does not (5.63 vs 5.57 ms) - the weight
    gradients run on the side lane, where a 36-CTA launch that leaves 112 SMs to the data-gradient chain is worth more
    than a shorter 144-CTA one."""
    if not _env_flag("VQA_B200_WGRAD_SPLIT", False):
        return None, 1
    kb = (M + 63) // 64
    tn = (N + 127) // 128
    unsplit, best = None, None
    for bn in (128, 64):
        if K % bn:
            continue
        tiles = tn * (K // bn)
        base = _GEMM_FIXED[bn] + -(-tiles // N_SM) * kb * _GEMM_KB[bn]
        unsplit = base if unsplit is None else min(unsplit, base)
        for split in (2, 3, 4):
            if tiles * split > N_SM or kb < 8 * split:
                continue
            cost = _GEMM_FIXED[bn] + -(-kb // split) * _GEMM_KB[bn] + 1.0 + 4.0 * N * K / 2.0e6
            if best is None or cost < best[0]:
                best = (cost, bn, split)
    if best is None or unsplit is None or best[0] > unsplit - 1.0:
        return None, 1
    return best[1], best[2]


def pick_conv_tile(M, Cout):
    tm = (M + 127) // 128
    best = None
    for bn in (256, 128, 64):
        if bn > 64 and Cout <= bn // 2:
            continue
        cost = _tile_cost(tm * ((Cout + bn - 1) // bn), bn)
        if best is None or cost < best[0]:
            best = (cost, bn)
    return best[1]


class Engine:
    def __init__(self, model):
        self.model = model
        self.lib = L.load()
        self.device = None
        self.plans = {}
        self.run_id = 0
        self.use_graphs = _env_flag("VQA_B200_GRAPHS", True)
        self.use_lanes = _env_flag("VQA_B200_LANES", True)
        self.use_tc_attention = _env_flag("VQA_B200_TC_ATTENTION", True)
        # The fused optimizer runs on its own stream and only this engine's consumers of the parameters wait for it,
        # so the next step's frozen backbone (which needs nothing but the images) overlaps the HBM-bound AdamW pass.
        self.async_optimizer = _env_flag("VQA_B200_ASYNC_OPTIMIZER", True) and self.use_lanes
        self.early_backbone = _env_flag("VQA_B200_EARLY_BACKBONE", False)   # measured on B200: 5.19 ms/step with it, 5.15 without
        self.s_vis = None           # vision stream (backbone + projection of the forward)
        self.s_opt = None           # optimizer stream
        self.opt_event = None       # recorded after the last fused optimizer pass
        self.ev_vis = None
        self.vision_sig = None
        self.param_sig = None
        self.shadow_fresh = False
        self.shadow_stale = True    # the bf16 shadow must be re-cast from the fp32 master before the next forward
        self.proj_dirty = True
        self.lo_fresh = False       # low-order halves of the split-precision T5 blocks' weights are current
        # T5 blocks 0..n-1 run their forward GEMMs with two-term (hi + lo bf16) operands: their weights' and normalised
        # inputs' bf16 rounding is what pulls the deepest tensors' gradient cosine below the 0.999 bar
        # (tools/precision_probe.py).  0 = plain bf16 everywhere.
        self.t5_split_blocks = int(os.environ.get("VQA_B200_T5_SPLIT_BLOCKS", "2"))
        # the SGA stack's and the classifier's forward GEMMs add the weights' low-order term (x W_hi + x W_lo): the log-probs
        # of the flat random-init classifier then agree with fp32 closely enough for >= 99 % top-1 agreement with margin
        self.split_head = _env_flag("VQA_B200_SPLIT_HEAD", True)
        self._ddp = None
        self.shard_events = None    # per-segment "weights updated (and gathered)" events of the last fused optimizer pass
        # Opt-in experiment (single GPU / replica mode): the fused AdamW pass runs segment by segment in forward order and the
        # next forward's text parts wait per segment, so that the HBM-bound update of the later blocks could run under the
        # first blocks' GEMMs.  Measured on B200: 5.19 ms/step with it, 5.15 without - a GEMM CTA owns its SM (shared
        # memory, registers), so the two only time-share the SMs.  Under the sharded optimizer the same per-segment events
        # DO pay off (they hide the all-gather, which runs on NVLink, not on the SMs' HBM path).
        self.pipeline_update = _env_flag("VQA_B200_PIPELINE_UPDATE", False)
        self.ddp_shards = None      # sharded data parallelism: [(lo, big_hi, own_lo, own_hi)] of the CURRENT step, else None
        self.master_stale = False   # fp32 master of the GEMM weights is current only on the owning rank (ddp.py)
        self.master_shards = None
        self.fused_opt = None       # weakref to a VQAFusedAdamW that updates every parameter of this engine
        self.pending_clip = None    # max_norm of a clip_grad_norm_ whose scaling the fused optimizer will apply
        self.clip_sumsq = None      # device scalar: sum of squared gradients of the last clip_grad_norm_

    # ------------------------------------------------------------------------------------------
    # flat parameter layout
    # ------------------------------------------------------------------------------------------
    def _layout(self):
        m = self.model
        big, small = [], []
        cls = m.classification_layer
        big.append(cls.weight); small.append(cls.bias)
        pl = m.attention_pooler.attention[0]
        small += [pl.weight, pl.bias]
        for sga in reversed(list(m.sga_modules)):
            m1, m2, mlp = sga.mhatt1, sga.mhatt2, sga.ffn.mlp
            big += [m1.linear_v.weight, m1.linear_k.weight, m1.linear_q.weight, m1.linear_merge.weight,
                    m2.linear_v.weight, m2.linear_k.weight, m2.linear_q.weight, m2.linear_merge.weight,
                    mlp.fc1.weight, mlp.fc2.weight]
            small += [m1.linear_v.bias, m1.linear_k.bias, m1.linear_q.bias, m1.linear_merge.bias,
                      m2.linear_v.bias, m2.linear_k.bias, m2.linear_q.bias, m2.linear_merge.bias,
                      mlp.fc1.bias, mlp.fc2.bias,
                      sga.norm1.norm.weight, sga.norm1.norm.bias, sga.norm2.norm.weight, sga.norm2.norm.bias,
                      sga.norm3.norm.weight, sga.norm3.norm.bias]
        proj = m._projection()
        big.append(proj.weight); small.append(proj.bias)
        t5 = m.lang_model
        for blk in reversed(list(t5.block)):
            att, ff = blk.layer[0], blk.layer[1]
            sa = att.SelfAttention
            big += [sa.q.weight, sa.k.weight, sa.v.weight, sa.o.weight,
                    ff.DenseReluDense.wi.weight, ff.DenseReluDense.wo.weight]
            small += [att.layer_norm.weight, ff.layer_norm.weight]
            if hasattr(sa, "relative_attention_bias"):
                small.append(sa.relative_attention_bias.weight)
        small.append(t5.final_layer_norm.weight)
        small.append(t5.embed_tokens.weight)
        return big, small

    def _flatten(self, device):
        big, small = self._layout()
        params = big + small
        off, offs = 0, {}
        for p in params:
            offs[id(p)] = off
            off += _align(p.numel())
        self.n_big = offs[id(small[0])]
        self.total = off
        self.master = torch.zeros(off, dtype=torch.float32, device=device)
        self.grad = torch.zeros(off, dtype=torch.float32, device=device)
        self.shadow = torch.zeros(off, dtype=torch.bfloat16, device=device)
        with torch.no_grad():
            for p in params:
                o, n = offs[id(p)], p.numel()
                view = self.master[o:o + n].view(p.shape)
                view.copy_(p.data)
                p.data = view
        self.params = params
        self.param_index = {id(p): i for i, p in enumerate(params)}
        self.offs = offs
        self.device = device
        self.param_sig = None
        self.shadow_fresh = False
        self.shadow_stale = True
        self.proj_dirty = True
        self.lo_fresh = False
        # low-order bf16 halves of the split-precision GEMM weights, at the same offsets as the bf16 shadow.  Two contiguous
        # ranges of `big`: classifier + SGA stack (its head), and T5 blocks n-1 .. 0 (its tail: reverse execution order)
        self.lo_ranges = self._split_ranges(offs)
        self.shadow_lo = torch.zeros(self.n_big, dtype=torch.bfloat16, device=device) if self.lo_ranges else None
        self.plans = {}
        self.rng = torch.zeros(2, dtype=torch.int64, device=device)
        self._seed_src = None
        self._reseed()
        # vision caches are rebuilt whenever the frozen weights change
        self.vision_sig = None
        _REGISTRY[self.master.data_ptr()] = self

    def _split_ranges(self, offs):
        """[(lo, hi)] element ranges of `big` whose weights also get a low-order bf16 half (split precision)."""
        blocks = list(self.model.lang_model.block)
        nsplit = max(0, min(self.t5_split_blocks, len(blocks)))
        self.t5_split_blocks = nsplit
        ranges = []
        if self.split_head:
            ranges.append((0, offs[id(self.model._projection().weight)]))
        if nsplit > 0:
            ranges.append((offs[id(blocks[nsplit - 1].layer[0].SelfAttention.q.weight)], self.n_big))
        return ranges

    def _after_flatten(self, device):
        """Model-specific device buffers that follow the flat layout (here: the projection's conv-layout weight)."""
        proj = self.model._projection()
        self.proj_w = torch.empty(proj.weight.numel(), dtype=torch.bfloat16, device=device)
        self.proj_dirty = True
        from .ddp import maybe_enable
        maybe_enable(self)

    ddp_shardable = True          # the sharded optimizer's pipelined all-gather needs this engine's segmented text forward

    def embedding_param(self):
        """The token table (last tensor of the flat buffers; exchanged as rows under data parallelism)."""
        return self.model.lang_model.embed_tokens.weight

    def _reseed(self):
        """Dropout seed = torch's global seed (so a later torch.manual_seed is honoured), mixed with the data-parallel rank (each
        replica draws its own masks); VQA_B200_SEED pins it.  The Philox offset restarts at 0."""
        src = (torch.initial_seed(), getattr(self, "_seed_rank", 0))
        if src == self._seed_src:
            return
        self._seed_src = src
        seed = int(os.environ.get("VQA_B200_SEED", src[0]))
        seed = (seed ^ (src[1] * 0x9E3779B97F4A7C15)) & 0x7FFFFFFFFFFFFFFF
        self.rng.copy_(torch.tensor([seed, 0], dtype=torch.int64))

    def _params_on(self, device):
        p0 = self.model.classification_layer.weight
        return (self.device == device and p0.device == device
                and p0.data_ptr() == self.master.data_ptr() + 4 * self.offs[id(p0)])

    # pointers into the flat buffers
    def mp(self, p):
        return self.master.data_ptr() + 4 * self.offs[id(p)]

    def gp(self, p):
        return self.grad.data_ptr() + 4 * self.offs[id(p)]

    def sp(self, p):
        return self.shadow.data_ptr() + 2 * self.offs[id(p)]

    def lp(self, p):
        """Low-order bf16 half (bf16(w - bf16(w))) of a split-precision weight."""
        o = self.offs[id(p)]
        if not any(lo0 <= o < lo1 for lo0, lo1 in self.lo_ranges):
            raise RuntimeError("parameter is not in a split-precision range")
        return self.shadow_lo.data_ptr() + 2 * o

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def rec(self, plan=None, ws_store=None):
        """ws_store: dict that owns the cluster split-K workspaces of the plan being recorded (they must outlive it)."""
        return _Rec(self.lib, plan, self._stream, ws_store)

    # ------------------------------------------------------------------------------------------
    # weight preparation
    # ------------------------------------------------------------------------------------------
    def _vision_convs(self):
        """(conv, bn) pairs of the frozen backbone in execution order, with their role."""
        vm = self.model._resnet_body()
        out = [("stem", vm.conv1, vm.bn1)]
        for li, layer in enumerate([vm.layer1, vm.layer2, vm.layer3, vm.layer4]):
            for bi, blk in enumerate(layer):
                names = ["conv1", "conv2"] + (["conv3"] if hasattr(blk, "conv3") else [])
                for j, nm in enumerate(names):
                    out.append(("l%d.%d.%s" % (li, bi, nm), getattr(blk, nm), getattr(blk, "bn%d" % (j + 1))))
                if blk.downsample is not None:
                    out.append(("l%d.%d.ds" % (li, bi), blk.downsample[0], blk.downsample[1]))
        fpn = self.model._fpn()
        if fpn is not None:      # lateral (1x1) and output (3x3) convs with bias, no normalisation (bn = None)
            for i in range(len(fpn.inner_blocks)):
                out.append(("fpn.inner%d" % i, fpn.inner_blocks[i][0], None))
                out.append(("fpn.layer%d" % i, fpn.layer_blocks[i][0], None))
        return out

    def _prepare_vision(self):
        convs = self._vision_convs()
        sig = []
        for _, c, b in convs:
            for t in ((c.weight, b.weight, b.bias, b.running_mean, b.running_var) if b is not None else (c.weight, c.bias)):
                if t.device != self.device:
                    raise RuntimeError("vision_model parameters must live on %s" % (self.device,))
                sig.append((t.data_ptr(), t._version))
        sig = tuple(sig)
        if sig == self.vision_sig:
            return
        r = self.rec(None)
        # Re-folded IN PLACE: the recorded plans hold raw pointers to these tensors.  Only a backbone of another shape (or
        # device) gets new buffers, and then every recorded plan is dropped.
        if not hasattr(self, "vw"):
            self.vw, self.vb = {}, {}
        realloc = False
        for name, c, b in convs:
            O, I, R, S = c.weight.shape
            Sp, Ip = (8, 8) if name == "stem" else (S, I)
            w, bias = self.vw.get(name), self.vb.get(name)
            if w is None or w.numel() != O * R * Sp * Ip or w.device != self.device:
                w = torch.empty(O * R * Sp * Ip, dtype=torch.bfloat16, device=self.device)
                bias = torch.empty(O, dtype=torch.float32, device=self.device)
                self.vw[name], self.vb[name] = w, bias
                realloc = True
            if b is not None:
                r.fold_conv_bn(c.weight.detach().float().contiguous(), b.weight.detach(), b.bias.detach(),
                               b.running_mean, b.running_var, float(b.eps), w, bias, O, I, R, S, Sp, Ip)
            else:
                # no normalisation: identity scale (gamma 1, var 1, eps 0), shift = the conv's own bias (beta, mean 0)
                one, zero = self._fold_const(O)
                r.fold_conv_bn(c.weight.detach().float().contiguous(), one, c.bias.detach().float().contiguous(), zero, one,
                               0.0, w, bias, O, I, R, S, Sp, Ip)
        if realloc and self.plans:
            for st in self.plans.values():
                st.destroy(self.lib)
            self.plans = {}
        self.vision_sig = sig

    def _fold_const(self, n):
        key = ("fold_const", n)
        if not hasattr(self, "_consts"):
            self._consts = {}
        if key not in self._consts:
            self._consts[key] = (torch.ones(n, device=self.device), torch.zeros(n, device=self.device))
        return self._consts[key]

    def _param_signature(self):
        return tuple(p._version for p in self.params)

    def _note_param_changes(self):
        """Host-side check, made BEFORE anything of a forward is enqueued: did something other than the fused optimizer
        (torch.optim.*, load_state_dict, an in-place edit) change the parameters since the last forward?  Then the bf16
        shadow, the projection's conv-layout weights and the low-order halves are all stale.  (Writes through `p.data`
        do not move the version counter; call `invalidate()` after such edits.)"""
        sig = self._param_signature()
        if not (self.shadow_fresh and sig == self.param_sig):
            self.shadow_stale = True
            self.proj_dirty = True
            self.lo_fresh = False
        self.param_sig = sig

    def invalidate(self):
        """Declare every derived weight cache stale (bf16 shadow, folded backbone, projection layout, low-order halves):
        for callers that edit parameters or BatchNorm statistics through `.data`, which no version counter sees."""
        self.shadow_fresh = False
        self.shadow_stale = True
        self.proj_dirty = True
        self.lo_fresh = False
        self.vision_sig = None

    def sync_master(self):
        """Sharded data parallelism: make the fp32 master copy of every parameter current on this rank (COLLECTIVE: all
        ranks must call it).  No-op otherwise.  Runs before state_dict(); call it before reading parameter values."""
        if self.master_stale and self._ddp is not None:
            self._ddp.sync_master(self)

    def _refresh_shadow(self):
        """bf16 copies of the fp32 master weights (what the GEMMs read), on the current stream.  Only needed when
        something other than the fused optimizer changed the parameters (it writes the shadow itself)."""
        if self.shadow_stale and self.master_stale:
            self.sync_master()     # (all ranks take this branch together: they changed the same parameters)
        if self.shadow_stale:
            self.rec(None).cast_f32_bf16(self.master, self.shadow, self.total)
            self.shadow_stale = False
        self.shadow_fresh = True
        if self.lo_ranges and not self.lo_fresh:
            for lo0, lo1 in self.lo_ranges:
                self.rec(None).split_lo_bf16(self.master.data_ptr() + 4 * lo0, self.shadow_lo.data_ptr() + 2 * lo0, lo1 - lo0)
            self.lo_fresh = True

    def _refresh_projection(self):
        """The ConvTranspose2d weight re-expressed as the equivalent 3x3 conv weight (current stream)."""
        if self.proj_dirty:
            proj = self.model._projection()
            Cin, Cout = proj.weight.shape[0], proj.weight.shape[1]
            if self.master_stale:     # sharded update: the gathered bf16 shadow is the current copy on every rank
                self.rec(None).convT_weight_prep_bf16(self.sp(proj.weight), self.proj_w, Cin, Cout)
            else:
                self.rec(None).convT_weight_prep(self.mp(proj.weight), self.proj_w, Cin, Cout)
            self.proj_dirty = False

    # ------------------------------------------------------------------------------------------
    # streams
    # ------------------------------------------------------------------------------------------
    def optimizer_stream(self):
        """Stream the fused optimizer launches on (None: the caller's current stream)."""
        if not self.async_optimizer:
            return None
        if self.s_opt is None:
            self.s_opt = torch.cuda.Stream(device=self.device)
        return self.s_opt

    def note_optimizer_launched(self, stream):
        self.opt_event = torch.cuda.Event()
        self.opt_event.record(stream)

    def wait_optimizer(self, stream=None):
        """Order `stream` (default: current) after the last fused optimizer pass.  Called by every consumer of the
        parameters the engine controls: forward, weight preparation, state_dict(), optimizer.state_dict()."""
        if self.opt_event is not None:
            (stream or torch.cuda.current_stream(self.device)).wait_event(self.opt_event)

    def refresh_lo_after_update(self, stream):
        """Low-order halves of the split-precision weights, recomputed right behind the fused optimizer pass on ITS stream
        (None: the current one), i.e. off the next forward's critical path (which waits for the optimizer event anyway)."""
        if not self.lo_ranges or self.lo_fresh or not self.shadow_fresh or self.master_stale:
            return
        rec = _Rec(self.lib, None, (lambda: stream.cuda_stream) if stream is not None else self._stream)
        for lo0, lo1 in self.lo_ranges:
            rec.split_lo_bf16(self.master.data_ptr() + 4 * lo0, self.shadow_lo.data_ptr() + 2 * lo0, lo1 - lo0)
        self.lo_fresh = True

    def segment_ranges(self):
        """[(segment index, (lo, big_hi))]: the GEMM-weight range of each backward segment of the last recorded step (what the
        text forward's parts wait for), or None before the first step."""
        st = getattr(self, "last_state", None)
        if st is None:
            return None
        out = []
        for i, seg in enumerate(st.bwd_segments):
            lo, hi = seg.grad_lo, min(seg.grad_hi, self.n_big)
            if hi > lo:
                out.append((i, (lo, hi)))
        return out if len(out) == len(st.bwd_segments) else None

    def refresh_lo_segment(self, segs, si, stream):
        """Low-order halves of segment si's split-precision weights, on the optimizer stream behind that segment's update."""
        lo, bhi = dict(segs)[si]
        rec = _Rec(self.lib, None, lambda: stream.cuda_stream)
        for r0, r1 in self.lo_ranges:
            a, b = max(lo, r0), min(bhi, r1)
            if a < b:
                rec.split_lo_bf16(self.master.data_ptr() + 4 * a, self.shadow_lo.data_ptr() + 2 * a, b - a)

    def note_fused_update(self, covered):
        """Called by VQAFusedAdamW after it updated `covered` of this engine's parameters in place (raw
        pointers, so no version counters moved) and wrote their bf16 shadow itself."""
        self.lo_fresh = False
        if covered == len(self.params):
            self.proj_dirty = True
        else:
            self.shadow_fresh = False

    # ------------------------------------------------------------------------------------------
    # plans
    # ------------------------------------------------------------------------------------------
    def _ensure(self, device):
        if device.type != "cuda":
            raise RuntimeError("%s (B200-native) runs on CUDA only: there is no CPU fallback path; "
                               "move the model and its inputs to a cuda device" % type(self.model).__name__)
        if self.device is None or not self._params_on(device):
            for p in self.model.parameters():
                if p.device != device:
                    raise RuntimeError("model parameters are on %s but inputs are on %s" % (p.device, device))
            self._flatten(device)
            self._after_flatten(device)

    def prepare(self):
        """Refresh the folded frozen-backbone weights if their sources changed (the trainable caches are refreshed
        inside forward(), after the optimizer)."""
        self._prepare_vision()

    def get_plan(self, B, Lt, H, W, training, has_labels, want_features, u8_images=False):
        key = (B, Lt, H, W, bool(training), bool(has_labels), bool(want_features), bool(u8_images))
        st = self.plans.get(key)
        if st is None:
            from .plan_builder import build_state
            st = build_state(self, *key)
            self.plans[key] = st
        return st

    def run_plan(self, plan):
        L.check(self.lib.vqa_plan_run(plan, ctypes.c_void_p(self._stream())), "plan_run")

    def forward(self, st, ids, mask, labels, images):
        """Copy the inputs into the plans' static buffers and replay the four forward plans: the vision branch on
        its own stream (the backbone does not wait for the optimizer), the text branch on the caller's stream."""
        main = torch.cuda.current_stream(self.device)
        self._note_param_changes()      # sets proj_dirty before the vision-stream block below reads it
        if self.use_lanes:
            if self.s_vis is None:
                self.s_vis = torch.cuda.Stream(device=self.device)
                self.ev_vis = torch.cuda.Event()
            vis = self.s_vis
            if not images.is_cuda:
                # host (pinned) images: the upload only writes the static image buffer, whose last reader (the previous
                # step's backbone plan) is earlier on this same stream - so it need not wait for the previous step's
                # backward on `main` and overlaps it (38.5 MB per step at batch 64: ~0.4 ms of the end-to-end step)
                with torch.cuda.stream(vis):
                    st.images.copy_(images, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(main)              # whatever produced the inputs (and last step's backward) is on `main`
            # Early backbone (opt-in, VQA_B200_EARLY_BACKBONE=1): the frozen backbone reads only the images and writes only its
            # own buffers (backward reads a COPY of its last map, made by fwd_proj), so it need not wait for the previous
            # step's backward on `main`.  Measured on B200 with a caller that runs ahead: no gain (5.19 vs 5.15 ms/step) -
            # the backbone's persistent 148-CTA grids and the text GEMMs each want whole SMs (~200 KB of shared memory per
            # CTA), so two streams of them time-share the SMs instead of filling each other's idle ones.
            # Device-resident images must still be ordered behind their producer on `main` - unless this very tensor
            # (same storage, same version) was already waited for and read by an earlier forward.
            sig = (images.data_ptr(), images._version, tuple(images.shape)) if images.is_cuda else None
            early = (self.early_backbone and getattr(st, "early_backbone_ok", False)
                     and (not images.is_cuda or sig == getattr(st, "images_sig", None)))
            if not early:
                vis.wait_event(ev)
            st.images_sig = sig
            if images.is_cuda:
                images.record_stream(vis)
            with torch.cuda.stream(vis):
                if images.is_cuda:
                    st.images.copy_(images, non_blocking=True)
                self.run_plan(st.fwd_vis)
                if early:
                    vis.wait_event(ev)   # fwd_proj overwrites the vision tokens / map copy the previous backward reads
                self.wait_optimizer(vis)
                self._refresh_projection()
                self.run_plan(st.fwd_proj)
                self.ev_vis.record(vis)
        else:
            st.images.copy_(images, non_blocking=True)
        st.ids.copy_(ids, non_blocking=True)
        if mask is not None:
            st.mask.copy_(mask, non_blocking=True)
        else:
            st.mask.fill_(1)
        if labels is not None:
            st.labels.copy_(labels, non_blocking=True)
        # sharded optimizer: the all-gathers of the updated weights are still in flight on the optimizer stream; each part of
        # the text forward waits only for the segments it reads (ddp.after_update records one event per segment)
        pipelined = (self.shard_events is not None and self.use_lanes and not self.shadow_stale and self.lo_fresh
                     and len(st.fwd_text_parts) > 1 and all(s < len(self.shard_events) for _, sg in st.fwd_text_parts for s in sg))
        if pipelined:
            for sidx in st.fwd_text_parts[0][1]:
                main.wait_event(self.shard_events[sidx])
        else:
            self.wait_optimizer(main)
        self._refresh_shadow()
        if st.training:
            # one dropout stream position per TRAINING forward: its backward (and a second backward under retain_graph)
            # regenerates the masks from the same {seed, offset}, a forward without backward does not repeat them
            self._reseed()
            self.rec(None).rng_advance(self.rng)
        if not self.use_lanes:
            self.run_plan(st.fwd_vis)
            self._refresh_projection()
            self.run_plan(st.fwd_proj)
        for k, (plan, segs) in enumerate(st.fwd_text_parts):
            if pipelined and k > 0:
                for sidx in segs:
                    main.wait_event(self.shard_events[sidx])
            self.run_plan(plan)
        self.shard_events = None
        if self.use_lanes:
            main.wait_event(self.ev_vis)
        self.run_plan(st.fwd_fuse)
        self.run_id += 1
        st.run_id = self.run_id
        self.last_state = st

    def grad_views(self):
        """Per-parameter views of the flat fp32 gradient buffer (None for parameters that do not require grad), created
        once: a step assigns these same tensor objects as `.grad`, so the fused optimizer and clip_grad_norm_ can
        recognise them by identity instead of re-deriving 180 pointers per step."""
        key = (self.grad.data_ptr(), tuple(p.requires_grad for p in self.params))
        if getattr(self, "_grad_views_key", None) != key:
            views = []
            for p in self.params:
                if p.requires_grad:
                    o = self.offs[id(p)]
                    v = self.grad[o:o + p.numel()].view(p.shape)
                    v._vqa_flat_view = True
                    views.append(v)
                else:
                    views.append(None)
            self._grad_views, self._grad_views_key = views, key
        return self._grad_views

    def backward(self, st, gloss, glogp):
        if st.run_id != self.run_id:
            raise RuntimeError("backward() after another forward(): the saved activations were overwritten")
        if gloss is not None:
            st.gloss.copy_(gloss.reshape(1))
        else:
            self.rec(None).memset_zero(st.gloss, 4)
        if glogp is not None:
            st.glogp.copy_(glogp)
        elif st.glogp_used:
            self.rec(None).memset_zero(st.glogp, 4 * st.glogp.numel())      # (a memset node, not a fill kernel)
        self.pending_clip = None
        if self._ddp is not None:
            self._ddp.backward(self, st)
        else:
            for seg in st.bwd_segments:
                self.run_plan(seg.plan)

    def __del__(self):
        try:
            for st in self.plans.values():
                st.destroy(self.lib)
        except Exception:
            pass


_REGISTRY = {}


def engine_for_ptr(ptr):
    """Engine whose flat fp32 master buffer contains device address `ptr` (used by the fused optimizer)."""
    for base, eng in _REGISTRY.items():
        if eng.master is not None and base <= ptr < base + 4 * eng.total:
            return eng
    return None
