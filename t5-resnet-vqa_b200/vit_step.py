"""Engine and launch plans of the B200-native VitVQAModel step (reference: model/vit_vqa_model.py:127-227, driven by
trainer/vit_vqa_trainer.py:450-464; SURVEY.md 8f-4, BASELINE.json configs[4]).

Forward  = frozen ViT-B/16 -> pooler_output (torch.no_grad in the reference, :184-186)  |  T5 encoder over the question (:189-192)
           -> [pooled | encoder token 0] -> Linear(1536, 768) + ReLU + Dropout(0.5) (:198-203) -> T5 decoder over
           decoder_question_input_ids, cross-attending to that one fused token (:207-212) -> last un-padded decoder position
           (:215-219) -> Linear(768, answers) -> log_softmax -> NLLLoss (:221-227).
Backward = that graph in reverse, written out by hand; no gradient reaches the ViT.

The T5 stacks reuse the kernels of the ResnetVQAModel step (tcgen05 GEMMs with fused epilogues, tcgen05 flash attention with
relative-position bias, RMSNorm).  What is specific to this model: the decoder's causal position bias, its one-key
cross-attention (softmax over a single key is 1: the context is the value row of the sample, the q / k projections and
their RMSNorm receive exactly zero gradient), the tied token table (encoder and decoder both scatter into one gradient), and
the ViT (patch GEMM, LayerNorm, 197-token tcgen05 attention, exact GELU).  `hf:` = transformers/models/t5/modeling_t5.py,
`vit:` = transformers/models/vit/modeling_vit.py.
"""
import ctypes
import math
import os

import torch

from . import lib as L
from .engine import Engine, t5_relative_buckets
from .plan_builder import Segment, State, _Alloc


def t5_causal_buckets(L_q, L_k, num_buckets=32, max_distance=128):
    """Unidirectional relative-position buckets of the decoder's self-attention (hf:189-234 with bidirectional=False)."""
    ctx = torch.arange(L_q, dtype=torch.long)[:, None]
    mem = torch.arange(L_k, dtype=torch.long)[None, :]
    rel = -torch.min(mem - ctx, torch.zeros(1, dtype=torch.long))
    max_exact = num_buckets // 2
    large = max_exact + (torch.log(rel.float() / max_exact) / math.log(max_distance / max_exact)
                         * (num_buckets - max_exact)).to(torch.long)
    large = torch.min(large, torch.full_like(large, num_buckets - 1))
    return torch.where(rel < max_exact, rel, large).to(torch.int32)


class VitEngine(Engine):
    """Flat fp32 master / gradient / bf16 shadow buffers over the trainable parameters of VitVQAModel (T5 encoder + decoder,
    fusing layer, classifier) and bf16 / fp32 caches of the frozen ViT.  Data parallel as full replicas: the flat gradient is
    averaged with one all-reduce (bf16 wire) behind the single backward segment (ddp.GradSync, all-reduce mode); the sharded
    optimizer and the row exchange of the token table are built for the north-star ResnetVQAModel step only."""

    def __init__(self, model):
        super().__init__(model)
        self.t5_split_blocks = 0
        # VQA_B200_VIT_SPLIT=1: every forward GEMM of the two T5 stacks, the fusing layer and the classifier contracts two-term
        # (hi + lo bf16) operands; 0: plain bf16
        self.split_head = os.environ.get("VQA_B200_VIT_SPLIT", "0") == "1"

    # ---- flat layout: GEMM weights (classifier, fusing layer, decoder 11..0, encoder 11..0), then the small tensors ----
    def _layout(self):
        m = self.model
        big, small = [m.classification_layer.weight, m.fusing_layer[0].weight], [m.classification_layer.bias,
                                                                                 m.fusing_layer[0].bias]
        t5 = m.lang_model
        for stack in (t5.decoder, t5.encoder):
            for blk in reversed(list(stack.block)):
                att = blk.layer[0]
                sa = att.SelfAttention
                big += [sa.q.weight, sa.k.weight, sa.v.weight, sa.o.weight]
                small.append(att.layer_norm.weight)
                if hasattr(sa, "relative_attention_bias"):
                    small.append(sa.relative_attention_bias.weight)
                if stack.is_decoder:
                    ca = blk.layer[1].EncDecAttention
                    big += [ca.q.weight, ca.k.weight, ca.v.weight, ca.o.weight]
                    small.append(blk.layer[1].layer_norm.weight)
                ff = blk.layer[-1]
                big += [ff.DenseReluDense.wi.weight, ff.DenseReluDense.wo.weight]
                small.append(ff.layer_norm.weight)
            small.append(stack.final_layer_norm.weight)
        small.append(t5.shared.weight)
        return big, small

    def _split_ranges(self, offs):
        return [(0, self.n_big)] if self.split_head else []

    ddp_shardable = False         # replicas + one all-reduce of the flat gradient after backward (one backward segment)

    def embedding_param(self):
        return self.model.lang_model.shared.weight

    def _after_flatten(self, device):
        # a new device (or a re-flattened model): the frozen ViT's caches are rebuilt there by the next prepare()
        self.__dict__.pop("vit_w", None)
        self.__dict__.pop("vit_f", None)
        self.vision_sig = None
        from .ddp import maybe_enable
        maybe_enable(self)

    def _refresh_projection(self):
        pass

    # ---- frozen ViT: bf16 GEMM weights (q|k|v fused) and fp32 vectors in engine-owned buffers (stable pointers) ----
    def _vit_tensors(self):
        vm = self.model.vision_model
        w, f = [], []      # (name, [tensors to concatenate])
        emb = vm.embeddings
        w.append(("patch", [emb.patch_embeddings.projection.weight]))
        f += [("patch_b", [emb.patch_embeddings.projection.bias]), ("cls", [emb.cls_token]), ("pos", [emb.position_embeddings])]
        for i, lyr in enumerate(vm.encoder.layer):
            a = lyr.attention.attention
            w.append(("l%d.qkv" % i, [a.query.weight, a.key.weight, a.value.weight]))
            f.append(("l%d.qkv_b" % i, [a.query.bias, a.key.bias, a.value.bias]))
            w.append(("l%d.o" % i, [lyr.attention.output.dense.weight]))
            f.append(("l%d.o_b" % i, [lyr.attention.output.dense.bias]))
            w.append(("l%d.fc1" % i, [lyr.intermediate.dense.weight]))
            f.append(("l%d.fc1_b" % i, [lyr.intermediate.dense.bias]))
            w.append(("l%d.fc2" % i, [lyr.output.dense.weight]))
            f.append(("l%d.fc2_b" % i, [lyr.output.dense.bias]))
            f += [("l%d.ln1_w" % i, [lyr.layernorm_before.weight]), ("l%d.ln1_b" % i, [lyr.layernorm_before.bias]),
                  ("l%d.ln2_w" % i, [lyr.layernorm_after.weight]), ("l%d.ln2_b" % i, [lyr.layernorm_after.bias])]
        f += [("ln_w", [vm.layernorm.weight]), ("ln_b", [vm.layernorm.bias])]
        w.append(("pool", [vm.pooler.dense.weight]))
        f.append(("pool_b", [vm.pooler.dense.bias]))
        return w, f

    def _prepare_vision(self):
        w, f = self._vit_tensors()
        sig = []
        for _, ts in w + f:
            for t in ts:
                if t.device != self.device:
                    raise RuntimeError("vision_model parameters must live on %s" % (self.device,))
                sig.append((t.data_ptr(), t._version))
        sig = tuple(sig)
        if sig == self.vision_sig:
            return
        if not hasattr(self, "vit_w"):
            def layout(items):
                off, offs = 0, {}
                for name, ts in items:
                    offs[name] = off
                    off += (sum(t.numel() for t in ts) + 63) // 64 * 64
                return off, offs
            nw, self.vit_woff = layout(w)
            nf, self.vit_foff = layout(f)
            self.vit_w = torch.zeros(nw, dtype=torch.bfloat16, device=self.device)
            self.vit_f = torch.zeros(nf, dtype=torch.float32, device=self.device)
        r = self.rec(None)
        with torch.no_grad():
            for name, ts in f:
                o = self.vit_foff[name]
                for t in ts:
                    self.vit_f[o:o + t.numel()].copy_(t.detach().reshape(-1))
                    o += t.numel()
            for name, ts in w:
                o = self.vit_woff[name]
                for t in ts:
                    src = t.detach().float().contiguous()
                    r.cast_f32_bf16(src, self.vit_w.data_ptr() + 2 * o, src.numel())
                    o += src.numel()
        self.vision_sig = sig

    def vw(self, name):
        return self.vit_w.data_ptr() + 2 * self.vit_woff[name]

    def vf(self, name):
        return self.vit_f.data_ptr() + 4 * self.vit_foff[name]

    # ---- plans ----
    def get_plan(self, B, Lt, Ld, H, W, training, has_labels, want_attn=False):
        key = (B, Lt, Ld, H, W, bool(training), bool(has_labels), bool(want_attn))
        st = self.plans.get(key)
        if st is None:
            st = build_state(self, *key)
            self.plans[key] = st
        return st

    def forward(self, st, ids, mask, dec_ids, dec_mask, labels, pixels):
        """Inputs -> static buffers; ViT on the vision stream (it does not wait for the optimizer), the T5 encoder on the
        caller's stream behind the optimizer, then everything that needs both."""
        main = torch.cuda.current_stream(self.device)
        self._note_param_changes()
        if self.s_vis is None:
            self.s_vis = torch.cuda.Stream(device=self.device)
            self.ev_vis = torch.cuda.Event()
        vis = self.s_vis
        if not pixels.is_cuda:
            with torch.cuda.stream(vis):
                st.pixels.copy_(pixels, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(main)
        vis.wait_event(ev)
        if pixels.is_cuda:
            pixels.record_stream(vis)
        with torch.cuda.stream(vis):
            if pixels.is_cuda:
                st.pixels.copy_(pixels, non_blocking=True)
            self.run_plan(st.fwd_vis)
            self.ev_vis.record(vis)
        st.ids.copy_(ids, non_blocking=True)
        st.dec_ids.copy_(dec_ids, non_blocking=True)
        for dst, src in ((st.mask, mask), (st.dec_mask, dec_mask)):
            if src is not None:
                dst.copy_(src, non_blocking=True)
            else:
                dst.fill_(1)
        if labels is not None:
            st.labels.copy_(labels, non_blocking=True)
        self.wait_optimizer(main)
        self._refresh_shadow()
        if st.training:
            self._reseed()
            self.rec(None).rng_advance(self.rng)
        self.run_plan(st.fwd_text)
        main.wait_event(self.ev_vis)
        self.run_plan(st.fwd_fuse)
        self.shard_events = None
        self.run_id += 1
        st.run_id = self.run_id
        self.last_state = st


class _Side:
    """Weight / bias gradients go to plan lane 1 and overlap the data-gradient chain on lane 0 (as in plan_builder.py)."""

    def __init__(self, enabled):
        self.enabled, self.r, self.pending = enabled, None, {}

    def bind(self, rec):
        self.r, self.pending = rec, {}

    def leaf(self, reads, fn):
        if not self.enabled:
            fn()
            return
        self.r.fork()
        self.r.lane(1)
        fn()
        mid = self.r.mark()
        self.r.lane(0)
        for t in reads:
            self.pending[t.data_ptr()] = mid

    def before_write(self, *bufs):
        for t in bufs:
            mid = self.pending.pop(t.data_ptr(), None)
            if mid is not None:
                self.r.wait(mid)


class _T5Stack:
    """Forward / backward recorder of one T5 stack (hf:637-792): encoder, or decoder with the one-token cross-attention."""

    def __init__(self, eng, al, stack, B, Lq, ids, key_mask, training, new_sid, fused=None):
        self.eng, self.al, self.stack, self.B, self.L, self.ids, self.key_mask = eng, al, stack, B, Lq, ids, key_mask
        self.M = B * Lq
        self.cfg = stack.cfg
        self.p = 0.1 if training else 0.0
        self.new_sid = new_sid
        self.fused = fused            # bf16 [B, 768]: the decoder's encoder_hidden_states (one token per sample)
        self.dec = stack.is_decoder
        self.blocks = list(stack.block)

    def forward(self, r):
        eng, al, cfg, M, B, Lq, p = self.eng, self.al, self.cfg, self.M, self.B, self.L, self.p
        f32 = torch.float32
        D, nH, dkv, dff, vocab = cfg["d_model"], cfg["num_heads"], cfg["d_kv"], cfg["d_ff"], cfg["vocab"]
        inner = nH * dkv
        eps = float(cfg["eps"])
        rng = eng.rng
        n = len(self.blocks)
        self.hid = [al(M, D, dtype=f32) for _ in range(n + 1)]
        self.sid_embed = self.new_sid()
        r.embedding_fwd(self.ids, eng.mp(self.stack.embed_tokens.weight), self.hid[0], M, D, vocab, p, self.sid_embed, rng)
        if self.dec:
            bucket = t5_causal_buckets(Lq, Lq, cfg["num_buckets"], cfg["max_distance"])
        else:
            bucket = t5_relative_buckets(Lq, Lq, cfg["num_buckets"], cfg["max_distance"])
        self.bucket = bucket.to(eng.device).contiguous()
        al.keep.append(self.bucket)
        self.relw = self.blocks[0].layer[0].SelfAttention.relative_attention_bias.weight
        self.pos_bias = al(nH, Lq, Lq, dtype=f32)
        r.t5_bias_build(eng.mp(self.relw), self.bucket, self.pos_bias, nH, Lq, cfg["num_buckets"])
        if self.dec:
            r.t5_bias_causal(self.pos_bias, nH, Lq)
        self.saved = []
        for bi, blk in enumerate(self.blocks):
            att, ff = blk.layer[0], blk.layer[-1]
            sa, dd = att.SelfAttention, ff.DenseReluDense
            split = eng.split_head
            ldy = 2 * D if split else D
            sv = dict(y1=al(M, ldy), rstd1=al(M, dtype=f32), qkv=al(M, 3 * inner), stats=al(B * nH * Lq, 2, dtype=f32),
                      ctx=al(M, inner), hmid=al(M, D, dtype=f32), y2=al(M, ldy), rstd2=al(M, dtype=f32), h=al(M, dff),
                      ldy=ldy, sid_p=self.new_sid(), sid_o=self.new_sid(), sid_h=self.new_sid(), sid_f=self.new_sid())

            def norm(x, w, y, rstd):
                if split:
                    r.rmsnorm_fwd_split(x, eng.mp(w), y, rstd, M, D, eps)
                else:
                    r.rmsnorm_fwd(x, eng.mp(w), y, None, rstd, M, D, eps, 0.0, 0, None)

            def lo(w, with_a):
                return dict(b_lo=eng.lp(w), a_lo_col=D if with_a else 0) if split else {}
            norm(self.hid[bi], att.layer_norm.weight, sv["y1"], sv["rstd1"])
            r.linear(sv["y1"], M, D, ldy, eng.sp(sa.q.weight), 3 * inner, sv["qkv"], 3 * inner, **lo(sa.q.weight, True))
            qkv = sv["qkv"]
            r.attn_fwd(B, nH, Lq, Lq, dkv, qkv, 3 * inner, qkv.data_ptr() + 2 * inner, 3 * inner,
                       qkv.data_ptr() + 4 * inner, 3 * inner, sv["ctx"], inner, None, self.pos_bias, self.key_mask, 1.0,
                       p, sv["sid_p"], rng, stats=sv["stats"])
            r.linear(sv["ctx"], M, inner, inner, eng.sp(sa.o.weight), D, sv["hmid"], D, out_fp32=1,
                     drop_p=p, sid=sv["sid_o"], rng=rng, residual=self.hid[bi], ldr=D, res_fp32=1, **lo(sa.o.weight, False))
            x = sv["hmid"]
            if self.dec:
                # cross-attention onto the single fused token: ctx = dropout(1) * (fused W_v^T); q / k never run
                ca = blk.layer[1].EncDecAttention
                sv.update(vx=al(B, inner), ctxc=al(M, inner), hmid2=al(M, D, dtype=f32),
                          sid_pc=self.new_sid(), sid_co=self.new_sid())
                r.linear(self.fused, B, D, D, eng.sp(ca.v.weight), inner, sv["vx"], inner, bn=64, **lo(ca.v.weight, False))
                r.xattn1_fwd(sv["vx"], sv["ctxc"], B, nH, Lq, dkv, p, sv["sid_pc"], rng)
                r.linear(sv["ctxc"], M, inner, inner, eng.sp(ca.o.weight), D, sv["hmid2"], D, out_fp32=1,
                         drop_p=p, sid=sv["sid_co"], rng=rng, residual=sv["hmid"], ldr=D, res_fp32=1, **lo(ca.o.weight, False))
                x = sv["hmid2"]
            sv["xff"] = x
            norm(x, ff.layer_norm.weight, sv["y2"], sv["rstd2"])
            r.linear(sv["y2"], M, D, ldy, eng.sp(dd.wi.weight), dff, sv["h"], dff, relu=1, drop_p=p, sid=sv["sid_h"], rng=rng,
                     **lo(dd.wi.weight, True))
            r.linear(sv["h"], M, dff, dff, eng.sp(dd.wo.weight), D, self.hid[bi + 1], D, out_fp32=1,
                     drop_p=p, sid=sv["sid_f"], rng=rng, residual=x, ldr=D, res_fp32=1, **lo(dd.wo.weight, False))
            self.saved.append(sv)
        self.out_f32, self.rstd_f = al(M, D, dtype=f32), al(M, dtype=f32)
        self.sid_final = self.new_sid()
        r.rmsnorm_fwd(self.hid[n], eng.mp(self.stack.final_layer_norm.weight), None, self.out_f32, self.rstd_f, M, D, eps, p,
                      self.sid_final, rng)
        return self.out_f32

    def backward(self, r, side, d_out, scratch, dfused=None):
        """d_out: fp32 [M, D] gradient of the stack's output (after the final dropout).  scratch: shared bf16 / fp32 work
        buffers.  dfused: fp32 [B, D] accumulator of the gradient of the fused token (decoder; zeroed by the caller)."""
        eng, cfg, M, B, Lq, p = self.eng, self.cfg, self.M, self.B, self.L, self.p
        D, nH, dkv, dff, vocab = cfg["d_model"], cfg["num_heads"], cfg["d_kv"], cfg["d_ff"], cfg["vocab"]
        inner = nH * dkv
        rng = eng.rng
        n = len(self.blocks)
        g_pair, dpre, dsm, dqkv, dH, dbias_pos, dvx = (scratch[k] for k in ("g_pair", "dpre", "dsm", "dqkv", "dH",
                                                                               "dbias_pos", "dvx"))
        state = {"g": 0}

        def next_g():
            state["g"] ^= 1
            return g_pair[state["g"]]
        r.memset_zero(dbias_pos, 4 * nH * Lq * Lq)
        g_bf = next_g()
        side.before_write(g_bf)
        r.rmsnorm_bwd(d_out, 1, self.hid[n], eng.mp(self.stack.final_layer_norm.weight), self.rstd_f, None, dH,
                      eng.gp(self.stack.final_layer_norm.weight), M, D, p, self.sid_final, rng,
                      g_bf, p, self.saved[n - 1]["sid_f"])
        for bi in reversed(range(n)):
            blk, sv = self.blocks[bi], self.saved[bi]
            att, ff = blk.layer[0], blk.layer[-1]
            sa, dd = att.SelfAttention, ff.DenseReluDense
            # FFN sub-layer
            g0 = g_bf
            side.leaf([g0], lambda: r.wgrad(g0, M, D, D, sv["h"], dff, dff, eng.gp(dd.wo.weight)))
            side.before_write(dpre)
            r.dgrad(g_bf, M, D, D, eng.sp(dd.wo.weight), dff, dpre, dff, relu_mask=sv["h"], ldm=dff, drop_p=p,
                    sid=sv["sid_h"], rng=rng)
            side.leaf([dpre], lambda: r.wgrad(dpre, M, dff, dff, sv["y2"], D, sv["ldy"], eng.gp(dd.wi.weight)))
            side.before_write(dsm)
            r.dgrad(dpre, M, dff, dff, eng.sp(dd.wi.weight), D, dsm, D)
            g_bf = next_g()
            side.before_write(g_bf)
            r.rmsnorm_bwd(dsm, 0, sv["xff"], eng.mp(ff.layer_norm.weight), sv["rstd2"], dH, dH,
                          eng.gp(ff.layer_norm.weight), M, D, 0.0, 0, rng, g_bf, p, sv["sid_co"] if self.dec else sv["sid_o"])
            if self.dec:
                ca = blk.layer[1].EncDecAttention
                g1 = g_bf
                side.leaf([g1], lambda: r.wgrad(g1, M, D, D, sv["ctxc"], inner, inner, eng.gp(ca.o.weight)))
                side.before_write(dsm)
                r.dgrad(g_bf, M, D, D, eng.sp(ca.o.weight), inner, dsm, inner)
                side.before_write(dvx)
                r.xattn1_bwd(dsm, dvx, B, nH, Lq, dkv, p, sv["sid_pc"], rng)
                side.leaf([dvx], lambda: r.wgrad(dvx, B, inner, inner, self.fused, D, D, eng.gp(ca.v.weight), bn=64))
                r.dgrad(dvx, B, inner, inner, eng.sp(ca.v.weight), D, dfused, D, out_fp32=1, accumulate=1, bn=64)
                # the residual stream's gradient passes the cross-attention sub-layer unchanged (its RMSNorm only feeds q)
                g_bf = next_g()
                side.before_write(g_bf)
                r.dropout_cast(dH, g_bf, M, D, p, sv["sid_o"], rng)
            # self-attention sub-layer
            g2 = g_bf
            side.leaf([g2], lambda: r.wgrad(g2, M, D, D, sv["ctx"], inner, inner, eng.gp(sa.o.weight)))
            side.before_write(dsm)
            r.dgrad(g_bf, M, D, D, eng.sp(sa.o.weight), inner, dsm, inner)
            qkv = sv["qkv"]
            side.before_write(dqkv)
            r.attn_bwd(B, nH, Lq, Lq, dkv, qkv, 3 * inner, qkv.data_ptr() + 2 * inner, 3 * inner,
                       qkv.data_ptr() + 4 * inner, 3 * inner, None, dsm, inner,
                       dqkv, 3 * inner, dqkv.data_ptr() + 2 * inner, 3 * inner, dqkv.data_ptr() + 4 * inner, 3 * inner,
                       dbias_pos, 1.0, p, sv["sid_p"], rng, stats=sv["stats"], bias=self.pos_bias, key_mask=self.key_mask)
            side.leaf([dqkv], lambda: r.wgrad(dqkv, M, 3 * inner, 3 * inner, sv["y1"], D, sv["ldy"], eng.gp(sa.q.weight)))
            side.before_write(dsm)
            r.dgrad(dqkv, M, 3 * inner, 3 * inner, eng.sp(sa.q.weight), D, dsm, D)
            if bi > 0:
                g_bf = next_g()
                side.before_write(g_bf)
            r.rmsnorm_bwd(dsm, 0, self.hid[bi], eng.mp(att.layer_norm.weight), sv["rstd1"], dH, dH,
                          eng.gp(att.layer_norm.weight), M, D, 0.0, 0, rng,
                          g_bf if bi > 0 else None, p, self.saved[bi - 1]["sid_f"] if bi > 0 else 0)
        r.t5_bias_grad(dbias_pos, self.bucket, eng.gp(self.relw), nH, Lq, cfg["num_buckets"])
        # tied token table: the encoder's and the decoder's rows are both added into the one (zeroed) gradient
        r.embedding_bwd(self.ids, dH, eng.gp(self.stack.embed_tokens.weight), M, D, vocab, p, self.sid_embed, rng)


def build_state(eng, B, Lt, Ld, H, W, training, has_labels, want_attn=False):
    m, dev, lib = eng.model, eng.device, eng.lib
    if not eng.use_tc_attention or Lt > 64 or Ld > 64:
        raise RuntimeError("VitVQAModel (B200-native): question / decoder lengths up to 64 tokens (tcgen05 attention)")
    st = State()
    st.training, st.run_id = training, -1
    al = _Alloc(dev)
    st.alloc = al
    st.ks_store = {}
    f32, i64 = torch.float32, torch.int64
    rng = eng.rng
    D = 768
    A = m.classification_layer.weight.shape[0]
    Apad = (A + 7) // 8 * 8
    sid_counter = [0]

    def new_sid():
        sid_counter[0] += 1
        return sid_counter[0]

    st.ids, st.mask = al(B, Lt, dtype=i64, zero=True), al(B, Lt, dtype=i64, zero=True)
    st.dec_ids, st.dec_mask = al(B, Ld, dtype=i64, zero=True), al(B, Ld, dtype=i64, zero=True)
    st.labels = al(B, dtype=i64, zero=True)
    st.pixels = al(B, 3, H, W, dtype=f32, zero=True)
    st.logp, st.loss = al(B, A, dtype=f32, zero=True), al(1, dtype=f32, zero=True)
    st.gloss, st.glogp = al(1, dtype=f32, zero=True), al(B, A, dtype=f32, zero=True)
    st.glogp_used = False
    st.features = {}

    st.fwd_vis, st.fwd_text, st.fwd_fuse = lib.vqa_plan_create(), lib.vqa_plan_create(), lib.vqa_plan_create()
    st.fwd_plans = [st.fwd_vis, st.fwd_text, st.fwd_fuse]
    st.fwd_text_parts = [(st.fwd_text, [0])]

    # =============================================================================================
    # frozen ViT-B/16 (vit: ViTEmbeddings, ViTLayer x 12, layernorm, ViTPooler) -> pooled_pre fp32 [B, 768]
    # =============================================================================================
    vcfg = m.vision_model.cfg
    P, Dv, Hv, Iv = vcfg["patch"], vcfg["hidden"], vcfg["heads"], vcfg["inter"]
    if H != vcfg["image"] or W != vcfg["image"]:
        raise RuntimeError("VitVQAModel: pixel_values must be %dx%d (the position table has no interpolation path)"
                           % (vcfg["image"], vcfg["image"]))
    NP = (H // P) * (W // P)
    T = NP + 1
    Mv = B * T
    veps = float(vcfg["eps"])
    r = eng.rec(st.fwd_vis, st.ks_store)
    patches = al(B * NP, 3 * P * P)
    r.vit_patchify(st.pixels, patches, B, H, W, P)
    proj = al(B * NP, Dv, dtype=f32)
    r.linear(patches, B * NP, 3 * P * P, 3 * P * P, eng.vw("patch"), Dv, proj, Dv, out_fp32=1, bias=eng.vf("patch_b"))
    hid_a, hid_b = al(Mv, Dv, dtype=f32), al(Mv, Dv, dtype=f32)
    r.vit_assemble(proj, eng.vf("cls"), eng.vf("pos"), hid_a, B, NP, Dv)
    nrm, qkv, ctx, hbuf = al(Mv, Dv), al(Mv, 3 * Dv), al(Mv, Dv), al(Mv, Iv)
    mean_s, rstd_s = al(Mv, dtype=f32), al(Mv, dtype=f32)
    st.attentions = []      # generate_answers: the ViT's per-layer attention maps, fp32 [B, heads, T, T] (:240-243)
    for i in range(vcfg["layers"]):
        pre = "l%d." % i
        if want_attn:
            st.attentions.append(al(B, Hv, T, T, dtype=f32))
        r.layernorm_fwd(hid_a, eng.vf(pre + "ln1_w"), eng.vf(pre + "ln1_b"), nrm, None, mean_s, rstd_s, Mv, Dv, veps)
        r.linear(nrm, Mv, Dv, Dv, eng.vw(pre + "qkv"), 3 * Dv, qkv, 3 * Dv, bias=eng.vf(pre + "qkv_b"))
        r.attention_long_fwd(qkv, 3 * Dv, qkv.data_ptr() + 2 * Dv, 3 * Dv, qkv.data_ptr() + 4 * Dv, 3 * Dv, ctx, Dv,
                             B, Hv, T, Dv // Hv, 1.0 / math.sqrt(Dv // Hv), st.attentions[i] if want_attn else None)
        r.linear(ctx, Mv, Dv, Dv, eng.vw(pre + "o"), Dv, hid_b, Dv, out_fp32=1, bias=eng.vf(pre + "o_b"),
                 residual=hid_a, ldr=Dv, res_fp32=1)
        r.layernorm_fwd(hid_b, eng.vf(pre + "ln2_w"), eng.vf(pre + "ln2_b"), nrm, None, mean_s, rstd_s, Mv, Dv, veps)
        r.linear(nrm, Mv, Dv, Dv, eng.vw(pre + "fc1"), Iv, hbuf, Iv, bias=eng.vf(pre + "fc1_b"), relu=2)   # bias + exact GELU
        r.linear(hbuf, Mv, Iv, Iv, eng.vw(pre + "fc2"), Dv, hid_a, Dv, out_fp32=1, bias=eng.vf(pre + "fc2_b"),
                 residual=hid_b, ldr=Dv, res_fp32=1)
    r.layernorm_fwd(hid_a, eng.vf("ln_w"), eng.vf("ln_b"), nrm, None, mean_s, rstd_s, Mv, Dv, veps)
    pooled_pre = al(B, Dv, dtype=f32)
    # ViTPooler: dense over token 0 of every sample = rows b*T of the normalised sequence (row stride T*D)
    r.linear(nrm, B, Dv, T * Dv, eng.vw("pool"), Dv, pooled_pre, Dv, out_fp32=1, bias=eng.vf("pool_b"), bn=64)
    st.pooled = al(B, Dv, dtype=f32)

    # =============================================================================================
    # T5 encoder over the question (main stream, behind the optimizer)
    # =============================================================================================
    t5 = m.lang_model
    r = eng.rec(st.fwd_text, st.ks_store)
    enc = _T5Stack(eng, al, t5.encoder, B, Lt, st.ids, st.mask, training, new_sid)
    enc_out = enc.forward(r)

    # =============================================================================================
    # fusing layer, T5 decoder, answer token, classifier, loss
    # =============================================================================================
    r = eng.rec(st.fwd_fuse, st.ks_store)
    fl, cls = m.fusing_layer[0], m.classification_layer
    p_fuse = 0.5 if training else 0.0            # nn.Dropout(0.5), model/vit_vqa_model.py:153
    cat_b, fused_b = al(B, 2 * D), al(B, D)
    r.vit_fuse_concat(pooled_pre, enc_out, Lt, cat_b, st.pooled, B, D)
    sid_fuse = new_sid()
    hlo = (lambda w: dict(b_lo=eng.lp(w))) if eng.split_head else (lambda w: {})
    r.linear(cat_b, B, 2 * D, 2 * D, eng.sp(fl.weight), D, fused_b, D, bias=eng.mp(fl.bias), relu=1, drop_p=p_fuse,
             sid=sid_fuse, rng=rng, bn=64, **hlo(fl.weight))
    dec = _T5Stack(eng, al, t5.decoder, B, Ld, st.dec_ids, st.dec_mask, training, new_sid, fused=fused_b)
    dec_out = dec.forward(r)
    ans_b = al(B, D)
    r.gather_rows(dec_out, st.dec_mask, ans_b, None, B, Ld, D)
    logits = al(B, Apad, dtype=f32, zero=True)
    r.linear(ans_b, B, D, D, eng.sp(cls.weight), A, logits, Apad, out_fp32=1, bias=eng.mp(cls.bias), bn=64, **hlo(cls.weight))
    r.logsoftmax_nll_fwd(logits, Apad, st.labels if has_labels else None, st.logp, st.loss if has_labels else None, B, A)
    st.n_fwd_launches = sum(lib.vqa_plan_size(p) for p in st.fwd_plans)

    # =============================================================================================
    # backward (one segment)
    # =============================================================================================
    Mmax = B * max(Lt, Ld)
    dff = t5.cfg["d_ff"]
    nH = t5.cfg["num_heads"]
    Lmax = max(Lt, Ld)
    scratch = dict(g_pair=[al(Mmax, D), al(Mmax, D)], dpre=al(Mmax, dff), dsm=al(Mmax, D), dqkv=al(Mmax, 3 * D),
                   dH=al(Mmax, D, dtype=f32), dbias_pos=al(nH, Lmax, Lmax, dtype=f32), dvx=al(B, D))
    dlogits = al(B, Apad, zero=True)
    dans, dDec = al(B, D, dtype=f32), al(B * Ld, D, dtype=f32)
    dfused, dpre_f = al(B, D, dtype=f32), al(B, D)
    denc0, dEnc = al(B, D, dtype=f32), al(B * Lt, D, dtype=f32)
    bp = lib.vqa_plan_create()
    r = eng.rec(bp, st.ks_store)
    side = _Side(eng.use_lanes)
    side.bind(r)
    n_small = eng.total - eng.n_big
    r.memset_zero(eng.grad.data_ptr() + 4 * eng.n_big, 4 * n_small)
    r.memset_zero(dfused, 4 * B * D)
    r.logsoftmax_nll_bwd(st.logp, st.labels if has_labels else None, st.gloss, st.glogp, dlogits, Apad, B, A)
    st.glogp_used = True

    def head_leaf():
        r.colsum_bf16(dlogits, Apad, eng.gp(cls.bias), B, A)
        r.wgrad(dlogits, B, A, Apad, ans_b, D, D, eng.gp(cls.weight), bn=64)
    side.leaf([dlogits], head_leaf)
    r.dgrad(dlogits, B, A, Apad, eng.sp(cls.weight), D, dans, D, out_fp32=1, bn=64)
    r.scatter_rows(dans, st.dec_mask, dDec, B, Ld, D)
    dec.backward(r, side, dDec, scratch, dfused=dfused)
    # fusing layer: Dropout(0.5) o ReLU o Linear over [pooled | encoder token 0]; only the encoder half carries on
    r.relu_dropout_bwd(dfused, fused_b, dpre_f, 1.0 / (1.0 - p_fuse), B * D)

    def fuse_leaf():
        r.colsum_bf16(dpre_f, D, eng.gp(fl.bias), B, D)
        r.wgrad(dpre_f, B, D, D, cat_b, 2 * D, 2 * D, eng.gp(fl.weight), bn=64)
    side.leaf([dpre_f], fuse_leaf)
    # d(encoder token 0) = dpre W[:, 768:1536]: the weight's column slice as a [768, 768] operand with row stride 1536
    r.gemm(B, D, D, dpre_f, D, 0, eng.sp(fl.weight) + 2 * D, 2 * D, 1, denc0, D, 1, bn=64)
    r.scatter_rows(denc0, None, dEnc, B, Lt, D)
    enc.backward(r, side, dEnc, scratch)
    st.bwd_segments = [Segment(bp, 0, eng.total)]
    st.n_bwd_launches = lib.vqa_plan_size(bp)

    if eng.use_graphs:
        s2 = torch.cuda.Stream(device=dev)
        s2.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(s2):
            sp = ctypes.c_void_p(s2.cuda_stream)
            for fp in st.fwd_plans:
                L.check(lib.vqa_plan_run(fp, sp), "plan warm-up (fwd)")
            L.check(lib.vqa_plan_run(bp, sp), "plan warm-up (bwd)")
            s2.synchronize()
            for fp in st.fwd_plans:
                L.check(lib.vqa_plan_capture_graph(fp, sp), "graph capture (fwd)")
            L.check(lib.vqa_plan_capture_graph(bp, sp), "graph capture (bwd)")
            s2.synchronize()
        torch.cuda.current_stream(dev).wait_stream(s2)
    return st
