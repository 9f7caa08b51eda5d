"""Records the forward and backward launch plans of one ResnetVQAModel step for a fixed input shape.

Forward  = model/resnet_vqa_model.py:114-160 (frozen ResNet body -> ConvTranspose2d projection -> tokens;
           T5 encoder; 3 x SGA; AttentionPooler; classifier; log_softmax + NLL).
Backward = the autograd graph of that forward, written out by hand in reverse (no gradient reaches the frozen
           backbone: the reference runs it under torch.no_grad()).
Every buffer is allocated once here; the plans hold raw device pointers into them and into the engine's flat
parameter / gradient / bf16-shadow buffers, so a step replays with two host calls (or two CUDA graph launches).
"""
import ctypes
import math
import os

import torch

from . import lib as L
from .engine import t5_relative_buckets


class Segment:
    def __init__(self, plan, grad_lo, grad_hi):
        self.plan, self.grad_lo, self.grad_hi = plan, grad_lo, grad_hi


class State:
    def destroy(self, lib):
        for p in list(self.fwd_plans) + [s.plan for s in self.bwd_segments]:
            if p:
                lib.vqa_plan_destroy(p)
        self.fwd_plans, self.bwd_segments = [], []


class _Alloc:
    def __init__(self, device):
        self.device, self.keep = device, []

    def __call__(self, *shape, dtype=torch.bfloat16, zero=False):
        t = (torch.zeros if zero else torch.empty)(*shape, dtype=dtype, device=self.device)
        self.keep.append(t)
        return t


def build_state(eng, B, Lt, H, W, training, has_labels, want_features, u8_images=False):
    m = eng.model
    dev = eng.device
    lib = eng.lib
    st = State()
    st.training, st.run_id = training, -1
    al = _Alloc(dev)
    st.alloc = al
    st.ks_store = {}   # cluster split-K workspaces, one per (plan, lane); live as long as the plans
    f32, bf16, i64 = torch.float32, torch.bfloat16, torch.int64
    rng = eng.rng
    D = 768
    A = m.classification_layer.weight.shape[0]
    Apad = (A + 7) // 8 * 8
    p_t5 = 0.1 if training else 0.0     # T5Config.dropout_rate
    p_sga = 0.1 if training else 0.0    # TextConfiguration.DROPOUT_R
    sid_counter = [0]

    def new_sid():
        sid_counter[0] += 1
        return sid_counter[0]

    def attn_save(nheads, Lq, Lk):
        """What attention saves for backward: (probs, stats).  Up to 64 x 64 tokens (the 49 / 64 vision tokens of 224x224 /
        256x256 images included) the tcgen05 flash kernels run and keep only the softmax row statistics; longer key
        sequences (196 tokens at 448x448) use the SIMT kernel and its fp32 probs."""
        if Lq <= 64 and Lk <= 64 and eng.use_tc_attention:
            return None, al(B * nheads * Lq, 2, dtype=f32)
        return al(B * nheads * Lq * Lk, dtype=f32), None

    # ---- static inputs / outputs ----
    st.ids = al(B, Lt, dtype=i64, zero=True)
    st.mask = al(B, Lt, dtype=i64, zero=True)
    st.labels = al(B, dtype=i64, zero=True)
    # fp32 [B,3,H,W] in 0..1 (the reference collate's ToTensor output) or uint8 [B,H,W,3] (input edge: /255 on the device)
    st.images = al(B, H, W, 3, dtype=torch.uint8, zero=True) if u8_images else al(B, 3, H, W, dtype=f32, zero=True)
    st.logp = al(B, A, dtype=f32, zero=True)
    st.loss = al(1, dtype=f32, zero=True)
    st.gloss = al(1, dtype=f32, zero=True)
    st.glogp = al(B, A, dtype=f32, zero=True)
    st.glogp_used = False

    # The forward is four plans so that independent parts run on different streams (engine.forward):
    #   fwd_vis   frozen backbone: needs only the images -> starts while the previous step's optimizer still runs
    #   fwd_proj  channel projection (trainable weights)  -> vision stream, after the optimizer
    #   fwd_text  T5 encoder + the start of SGA layer 1    -> main stream, after the optimizer
    #   fwd_fuse  everything from the first use of the vision tokens on -> main stream, after fwd_proj
    # fwd_text is a LIST of plans, cut where the T5 blocks change backward segment: under the sharded optimizer (ddp.py) each
    # part only waits for the all-gather of the weights it reads, so the gathers of later blocks run under earlier blocks
    st.fwd_vis, st.fwd_proj = lib.vqa_plan_create(), lib.vqa_plan_create()
    st.fwd_fuse = lib.vqa_plan_create()
    st.fwd_text_parts = []        # [(plan, [backward segment indices whose weights it reads])]
    st.fwd_plans = [st.fwd_vis, st.fwd_proj, st.fwd_fuse]
    r = eng.rec(st.fwd_vis, st.ks_store)
    M = B * Lt

    # =============================================================================================
    # frozen ResNet body (tv:266-277 without avgpool/fc), NHWC bf16, BatchNorm folded
    # =============================================================================================
    two_lanes = eng.use_lanes
    vm = m._resnet_body()
    fpn = m._fpn()
    levels = []     # (map, C, H, W) after layer1..4: the FPN's inputs
    stem_in = al(B, H, W + 8, 8)
    if u8_images:
        r.image_u8_to_stem(st.images, stem_in, B, H, W)
    else:
        r.image_to_stem(st.images, stem_in, B, H, W)
    H1, W1 = (H + 6 - 7) // 2 + 1, (W + 6 - 7) // 2 + 1
    c1 = al(B, H1, W1, 64)
    r.conv(B, H, W, 8, 64, 7, 2, 3, stem_in, eng.vw["stem"], c1, bias=eng.vb["stem"], relu=1, stem7=1)
    Hc, Wc = (H1 + 2 - 3) // 2 + 1, (W1 + 2 - 3) // 2 + 1
    x = al(B, Hc, Wc, 64)
    r.maxpool3x3s2(c1, x, B, H1, W1, 64)
    C = 64
    for li, layer in enumerate([vm.layer1, vm.layer2, vm.layer3, vm.layer4]):
        for bi, blk in enumerate(layer):
            pre = "l%d.%d." % (li, bi)
            s = blk.stride
            if hasattr(blk, "conv3"):      # Bottleneck (tv:143-163)
                planes = blk.conv1.weight.shape[0]
                t1 = al(B, Hc, Wc, planes)
                r.conv(B, Hc, Wc, C, planes, 1, 1, 0, x, eng.vw[pre + "conv1"], t1, bias=eng.vb[pre + "conv1"])
                Ho, Wo = (Hc + 2 - 3) // s + 1, (Wc + 2 - 3) // s + 1
                t2 = al(B, Ho, Wo, planes)
                r.conv(B, Hc, Wc, planes, planes, 3, s, 1, t1, eng.vw[pre + "conv2"], t2, bias=eng.vb[pre + "conv2"])
                Cout = planes * 4
                if blk.downsample is not None:
                    idn = al(B, Ho, Wo, Cout)
                    r.conv(B, Hc, Wc, C, Cout, 1, s, 0, x, eng.vw[pre + "ds"], idn, bias=eng.vb[pre + "ds"], relu=0)
                else:
                    idn = x
                y = al(B, Ho, Wo, Cout)
                r.conv(B, Ho, Wo, planes, Cout, 1, 1, 0, t2, eng.vw[pre + "conv3"], y, bias=eng.vb[pre + "conv3"],
                       residual=idn, relu=1)
            else:                          # BasicBlock (tv:89-105)
                planes = blk.conv1.weight.shape[0]
                Ho, Wo = (Hc + 2 - 3) // s + 1, (Wc + 2 - 3) // s + 1
                t1 = al(B, Ho, Wo, planes)
                r.conv(B, Hc, Wc, C, planes, 3, s, 1, x, eng.vw[pre + "conv1"], t1, bias=eng.vb[pre + "conv1"])
                Cout = planes
                if blk.downsample is not None:
                    idn = al(B, Ho, Wo, Cout)
                    r.conv(B, Hc, Wc, C, Cout, 1, s, 0, x, eng.vw[pre + "ds"], idn, bias=eng.vb[pre + "ds"], relu=0)
                else:
                    idn = x
                y = al(B, Ho, Wo, Cout)
                r.conv(B, Ho, Wo, planes, Cout, 3, 1, 1, t1, eng.vw[pre + "conv2"], y, bias=eng.vb[pre + "conv2"],
                       residual=idn, relu=1)
            x, C, Hc, Wc = y, Cout, Ho, Wo
        levels.append((x, C, Hc, Wc))
    st.features = {}

    def export(name, t, c, h, w):
        st.features[name] = al(B, c, h, w, dtype=f32)
        r.nhwc_to_nchw_f32(t, st.features[name], B, h, w, c)

    if fpn is None:
        feat, Cf, Hf, Wf = x, C, Hc, Wc
        if want_features:
            export("features", feat, Cf, Hf, Wf)
    else:
        # FeaturePyramidNetwork + LastLevelMaxPool (torchvision feature_pyramid_network.py; the reference uses its 'pool'
        # output, model/faster_rcnn_vqa_model.py:102-108).  The top level receives no top-down term, and
        # max_pool2d(kernel 1, stride 2) of its 3x3 output conv is that conv evaluated at stride 2: two launches.
        Fo = fpn.out_channels
        c5, C5, H5, W5 = levels[3]
        inner = al(B, H5, W5, Fo)
        r.conv(B, H5, W5, C5, Fo, 1, 1, 0, c5, eng.vw["fpn.inner3"], inner, bias=eng.vb["fpn.inner3"], relu=0)
        Hp, Wp = (H5 + 2 - 3) // 2 + 1, (W5 + 2 - 3) // 2 + 1
        pool = al(B, Hp, Wp, Fo)
        r.conv(B, H5, W5, Fo, Fo, 3, 2, 1, inner, eng.vw["fpn.layer3"], pool, bias=eng.vb["fpn.layer3"], relu=0)
        feat, Cf, Hf, Wf = pool, Fo, Hp, Wp
        if want_features:     # generate_answers returns all five maps (model/faster_rcnn_vqa_model.py:150-154)
            export("pool", pool, Fo, Hp, Wp)
            out3 = al(B, H5, W5, Fo)
            r.conv(B, H5, W5, Fo, Fo, 3, 1, 1, inner, eng.vw["fpn.layer3"], out3, bias=eng.vb["fpn.layer3"], relu=0)
            export("3", out3, Fo, H5, W5)
            last, hl, wl = inner, H5, W5
            for idx in (2, 1, 0):
                ci, Ci, Hi, Wi = levels[idx]
                up = al(B, Hi, Wi, Fo)
                r.upsample_nearest_nhwc(last, up, B, hl, wl, Hi, Wi, Fo)
                lat = al(B, Hi, Wi, Fo)
                r.conv(B, Hi, Wi, Ci, Fo, 1, 1, 0, ci, eng.vw["fpn.inner%d" % idx], lat, bias=eng.vb["fpn.inner%d" % idx],
                       residual=up, relu=0)
                oi = al(B, Hi, Wi, Fo)
                r.conv(B, Hi, Wi, Fo, Fo, 3, 1, 1, lat, eng.vw["fpn.layer%d" % idx], oi, bias=eng.vb["fpn.layer%d" % idx],
                       relu=0)
                export(str(idx), oi, Fo, Hi, Wi)
                last, hl, wl = lat, Hi, Wi
    st.feat_shape = (B, Cf, Hf, Wf)

    # channel projection = ConvTranspose2d(k3,s1,p1) as a 3x3 same conv with flipped/transposed weights
    # (model/resnet_vqa_model.py:124,135) writing [B*hw, 768] tokens directly (:142-143)
    proj = m._projection()
    if proj.weight.shape[0] != Cf:
        raise RuntimeError("projection expects %d channels, backbone gives %d" % (proj.weight.shape[0], Cf))
    Ty = Hf * Wf
    My = B * Ty
    y0 = al(My, D)
    r = eng.rec(st.fwd_proj, st.ks_store)
    # The backbone's last map is copied out of the backbone's own buffers here: backward reads the copy (projection weight
    # gradient), so the NEXT step's backbone - which needs nothing but its images - may start while this step's text forward
    # and backward still run (engine.forward: fwd_vis is not ordered behind the previous backward).  12.8 MB at batch 64.
    feat_keep = al(B, Hf, Wf, Cf)
    r.memcpy_d2d(feat_keep, feat, 2 * B * Hf * Wf * Cf)
    feat = feat_keep
    st.early_backbone_ok = not want_features     # the exported maps are read by the caller's stream after the forward
    r.conv(B, Hf, Wf, Cf, D, 3, 1, 1, feat, eng.proj_w, y0, bias=eng.mp(proj.bias), relu=0)
    # T5 blocks per backward segment, from the LAST block down (= per gradient-exchange bucket under data parallelism).  The
    # final segments are short: the exchange of the last one cannot overlap anything, and under the sharded optimizer the
    # next forward starts as soon as block 0's weights are gathered.  VQA_B200_DDP_SEG_PLAN="3,3,3,2,1";
    # VQA_B200_DDP_BLOCKS_PER_SEG=<n> gives uniform segments.
    if os.environ.get("VQA_B200_DDP_BLOCKS_PER_SEG"):
        seg_plan = [max(1, int(os.environ["VQA_B200_DDP_BLOCKS_PER_SEG"]))] * 64
    else:
        seg_plan = [max(1, int(v)) for v in os.environ.get("VQA_B200_DDP_SEG_PLAN", "3,3,3,2,1").split(",")] + [1] * 64
    _seg_of = {}

    def new_text_part(segs):
        plan = lib.vqa_plan_create()
        st.fwd_text_parts.append((plan, list(segs)))
        st.fwd_plans.append(plan)
        return eng.rec(plan, st.ks_store)

    def seg_of_block(bi, nblk):    # backward segment holding T5 block bi's weights (segment 0 = head + SGA stack)
        if not _seg_of:
            b, sidx = nblk - 1, 1
            for n in seg_plan:
                for _ in range(n):
                    if b >= 0:
                        _seg_of[b] = sidx
                        b -= 1
                sidx += 1
                if b < 0:
                    break
        return _seg_of[bi]
    r = None

    # =============================================================================================
    # T5 encoder (hf:637-792)
    # =============================================================================================
    t5 = m.lang_model
    cfg = t5.cfg
    nH, dkv, dff, vocab = cfg["num_heads"], cfg["d_kv"], cfg["d_ff"], cfg["vocab"]
    inner = nH * dkv
    eps_t5 = float(cfg["eps"])
    blocks = list(t5.block)
    nblk = len(blocks)
    hid = [al(M, D, dtype=f32) for _ in range(nblk + 1)]
    r = new_text_part([seg_of_block(0, nblk)])
    sid_embed = new_sid()
    r.embedding_fwd(st.ids, eng.mp(t5.embed_tokens.weight), hid[0], M, D, vocab, p_t5, sid_embed, rng)
    bucket = t5_relative_buckets(Lt, Lt, cfg["num_buckets"], cfg["max_distance"]).to(dev).contiguous()
    al.keep.append(bucket)
    relw = blocks[0].layer[0].SelfAttention.relative_attention_bias.weight
    pos_bias = al(nH, Lt, Lt, dtype=f32)
    r.t5_bias_build(eng.mp(relw), bucket, pos_bias, nH, Lt, cfg["num_buckets"])
    saved_t5 = []
    nsplit = eng.t5_split_blocks
    for bi, blk in enumerate(blocks):
        att, ff = blk.layer[0], blk.layer[1]
        sa, dd = att.SelfAttention, ff.DenseReluDense
        t5_probs, t5_stats = attn_save(nH, Lt, Lt)
        # Blocks 0..nsplit-1 contract two-term (hi + lo bf16) operands in their forward GEMMs: the normalised inputs are
        # written as [M, 2D] = hi | lo, the weights' low-order halves come from the engine (engine.lp), and one k-loop adds
        # x_hi W_hi + x_lo W_hi + x_hi W_lo (o / wo: their bf16 inputs have no low half, two terms).  The backward still
        # uses the plain bf16 operands (backward rounding does not move the gradient cosine, tools/precision_probe.py).
        if bi > 0 and seg_of_block(bi, nblk) != seg_of_block(bi - 1, nblk):
            last = seg_of_block(bi, nblk) == 1      # the last part also runs the start of the SGA stack (segment 0)
            r = new_text_part([seg_of_block(bi, nblk)] + ([0] if last else []))
        split = bi < nsplit
        ldy = 2 * D if split else D
        sv = dict(y1=al(M, ldy), rstd1=al(M, dtype=f32), qkv=al(M, 3 * inner), probs=t5_probs, stats=t5_stats,
                  ctx=al(M, inner), hmid=al(M, D, dtype=f32), y2=al(M, ldy), rstd2=al(M, dtype=f32), h=al(M, dff),
                  ldy=ldy, sid_p=new_sid(), sid_o=new_sid(), sid_h=new_sid(), sid_f=new_sid())

        def norm(x, w, y, rstd):
            if split:
                r.rmsnorm_fwd_split(x, eng.mp(w), y, rstd, M, D, eps_t5)
            else:
                r.rmsnorm_fwd(x, eng.mp(w), y, None, rstd, M, D, eps_t5, 0.0, 0, None)

        def lo(w, with_a):
            return dict(b_lo=eng.lp(w), a_lo_col=D if with_a else 0) if split else {}
        norm(hid[bi], att.layer_norm.weight, sv["y1"], sv["rstd1"])
        # fused q|k|v projection: the three [768,768] weights are adjacent in the flat bf16 shadow
        r.linear(sv["y1"], M, D, ldy, eng.sp(sa.q.weight), 3 * inner, sv["qkv"], 3 * inner, **lo(sa.q.weight, True))
        qkv = sv["qkv"]
        r.attn_fwd(B, nH, Lt, Lt, dkv, qkv, 3 * inner, qkv.data_ptr() + 2 * inner, 3 * inner,
                   qkv.data_ptr() + 4 * inner, 3 * inner, sv["ctx"], inner, sv["probs"], pos_bias, st.mask, 1.0,
                   p_t5, sv["sid_p"], rng, stats=sv["stats"])
        r.linear(sv["ctx"], M, inner, inner, eng.sp(sa.o.weight), D, sv["hmid"], D, out_fp32=1,
                 drop_p=p_t5, sid=sv["sid_o"], rng=rng, residual=hid[bi], ldr=D, res_fp32=1, **lo(sa.o.weight, False))
        norm(sv["hmid"], ff.layer_norm.weight, sv["y2"], sv["rstd2"])
        r.linear(sv["y2"], M, D, ldy, eng.sp(dd.wi.weight), dff, sv["h"], dff, relu=1, drop_p=p_t5, sid=sv["sid_h"],
                 rng=rng, **lo(dd.wi.weight, True))
        r.linear(sv["h"], M, dff, dff, eng.sp(dd.wo.weight), D, hid[bi + 1], D, out_fp32=1,
                 drop_p=p_t5, sid=sv["sid_f"], rng=rng, residual=sv["hmid"], ldr=D, res_fp32=1,
                 **lo(dd.wo.weight, False))
        saved_t5.append(sv)
    text_f32, text_bf16, rstd_f = al(M, D, dtype=f32), al(M, D), al(M, dtype=f32)
    sid_final = new_sid()
    r.rmsnorm_fwd(hid[nblk], eng.mp(t5.final_layer_norm.weight), text_bf16, text_f32, rstd_f, M, D, eps_t5, p_t5,
                  sid_final, rng)

    # =============================================================================================
    # SGA stack (model/multi_head_vision_text_attn.py:145-158; x = text for every layer, y = previous output)
    # =============================================================================================
    Hs, hd = 8, D // 8
    scale = 1.0 / math.sqrt(hd)
    sgas = list(m.sga_modules)
    saved_sga = []

    def hlo(w):   # low-order weight term of the SGA / classifier forward GEMMs (engine.split_head)
        return dict(b_lo=eng.lp(w)) if eng.split_head else {}
    y_bf16, Ly = y0, Ty
    out_f32 = None
    for li, sga in enumerate(sgas):
        Myl = B * Ly
        m1, m2, mlp = sga.mhatt1, sga.mhatt2, sga.ffn.mlp
        probs1, stats1 = attn_save(Hs, Lt, Lt)
        probs2, stats2 = attn_save(Hs, Lt, Ly)
        sv = dict(y=y_bf16, Ly=Ly, stats1=stats1, stats2=stats2,
                  qkv1=al(M, 3 * D), probs1=probs1, ctx1=al(M, D), z1=al(M, D, dtype=f32),
                  mean1=al(M, dtype=f32), rstd1=al(M, dtype=f32), x1f=al(M, D, dtype=f32), x1b=al(M, D),
                  q2=al(M, D), vk2=al(Myl, 2 * D), probs2=probs2, ctx2=al(M, D),
                  z2=al(M, D, dtype=f32), mean2=al(M, dtype=f32), rstd2=al(M, dtype=f32),
                  x2f=al(M, D, dtype=f32), x2b=al(M, D), hm=al(M, D), z3=al(M, D, dtype=f32),
                  mean3=al(M, dtype=f32), rstd3=al(M, dtype=f32), of=al(M, D, dtype=f32), ob=al(M, D),
                  sid_p1=new_sid(), sid_r1=new_sid(), sid_p2=new_sid(), sid_r2=new_sid(), sid_h=new_sid(),
                  sid_r3=new_sid())
        # mhatt1(v=x, k=x, q=x): fused v|k|q projection
        r.linear(text_bf16, M, D, D, eng.sp(m1.linear_v.weight), 3 * D, sv["qkv1"], 3 * D,
                 bias=eng.mp(m1.linear_v.bias), **hlo(m1.linear_v.weight))
        q1 = sv["qkv1"]
        r.attn_fwd(B, Hs, Lt, Lt, hd, q1.data_ptr() + 4 * D, 3 * D, q1.data_ptr() + 2 * D, 3 * D, q1, 3 * D,
                   sv["ctx1"], D, sv["probs1"], None, None, scale, p_sga, sv["sid_p1"], rng, stats=sv["stats1"])
        r.linear(sv["ctx1"], M, D, D, eng.sp(m1.linear_merge.weight), D, sv["z1"], D, out_fp32=1,
                 bias=eng.mp(m1.linear_merge.bias), drop_p=p_sga, sid=sv["sid_r1"], rng=rng, residual=text_f32,
                 ldr=D, res_fp32=1, **hlo(m1.linear_merge.weight))
        r.layernorm_fwd(sv["z1"], eng.mp(sga.norm1.norm.weight), eng.mp(sga.norm1.norm.bias), sv["x1b"], sv["x1f"],
                        sv["mean1"], sv["rstd1"], M, D, float(sga.norm1.norm.eps))
        # mhatt2(v=y, k=y, q=x1)
        r.linear(sv["x1b"], M, D, D, eng.sp(m2.linear_q.weight), D, sv["q2"], D, bias=eng.mp(m2.linear_q.bias),
                 **hlo(m2.linear_q.weight))
        if li == 0:
            r = eng.rec(st.fwd_fuse, st.ks_store)    # the vision tokens (vision stream) are needed from here on
        r.linear(y_bf16, Myl, D, D, eng.sp(m2.linear_v.weight), 2 * D, sv["vk2"], 2 * D,
                 bias=eng.mp(m2.linear_v.bias), **hlo(m2.linear_v.weight))
        vk = sv["vk2"]
        r.attn_fwd(B, Hs, Lt, Ly, hd, sv["q2"], D, vk.data_ptr() + 2 * D, 2 * D, vk, 2 * D, sv["ctx2"], D,
                   sv["probs2"], None, None, scale, p_sga, sv["sid_p2"], rng, stats=sv["stats2"])
        r.linear(sv["ctx2"], M, D, D, eng.sp(m2.linear_merge.weight), D, sv["z2"], D, out_fp32=1,
                 bias=eng.mp(m2.linear_merge.bias), drop_p=p_sga, sid=sv["sid_r2"], rng=rng, residual=sv["x1f"],
                 ldr=D, res_fp32=1, **hlo(m2.linear_merge.weight))
        r.layernorm_fwd(sv["z2"], eng.mp(sga.norm2.norm.weight), eng.mp(sga.norm2.norm.bias), sv["x2b"], sv["x2f"],
                        sv["mean2"], sv["rstd2"], M, D, float(sga.norm2.norm.eps))
        # FFN
        r.linear(sv["x2b"], M, D, D, eng.sp(mlp.fc1.weight), D, sv["hm"], D, bias=eng.mp(mlp.fc1.bias), relu=1,
                 drop_p=p_sga, sid=sv["sid_h"], rng=rng, **hlo(mlp.fc1.weight))
        r.linear(sv["hm"], M, D, D, eng.sp(mlp.fc2.weight), D, sv["z3"], D, out_fp32=1, bias=eng.mp(mlp.fc2.bias),
                 drop_p=p_sga, sid=sv["sid_r3"], rng=rng, residual=sv["x2f"], ldr=D, res_fp32=1, **hlo(mlp.fc2.weight))
        r.layernorm_fwd(sv["z3"], eng.mp(sga.norm3.norm.weight), eng.mp(sga.norm3.norm.bias), sv["ob"], sv["of"],
                        sv["mean3"], sv["rstd3"], M, D, float(sga.norm3.norm.eps))
        saved_sga.append(sv)
        y_bf16, Ly, out_f32 = sv["ob"], Lt, sv["of"]

    # =============================================================================================
    # head: AttentionPooler + classifier + log_softmax/NLL (model/resnet_vqa_model.py:152-160)
    # =============================================================================================
    pl = m.attention_pooler.attention[0]
    cls = m.classification_layer
    pool_w, pooled_b = al(B, Lt, dtype=f32), al(B, D)
    r.pooler_fwd(out_f32, eng.mp(pl.weight), eng.mp(pl.bias), pool_w, None, pooled_b, B, Lt, D)
    logits = al(B, Apad, dtype=f32, zero=True)
    r.linear(pooled_b, B, D, D, eng.sp(cls.weight), A, logits, Apad, out_fp32=1, bias=eng.mp(cls.bias), bn=64,
             **hlo(cls.weight))
    r.logsoftmax_nll_fwd(logits, Apad, st.labels if has_labels else None, st.logp,
                         st.loss if has_labels else None, B, A)
    st.n_fwd_launches = sum(lib.vqa_plan_size(p) for p in st.fwd_plans)

    # =============================================================================================
    # backward
    # =============================================================================================
    st.bwd_segments = []
    segs = st.bwd_segments

    def new_segment():
        return lib.vqa_plan_create()

    def close_segment(plan, lo_param, hi_param_end):
        segs.append(Segment(plan, lo_param, hi_param_end))

    o = eng.offs
    # scratch shared by all layers
    g_pair = [al(M, D), al(M, D)]  # dropout-masked residual-branch gradient (bf16 GEMM operand), ping-pong
    g_idx = [0]
    dpre = al(M, max(dff, D))  # gradient before ReLU
    dsm = al(M, D)             # small bf16 [M,768] gradients (dy of a norm, dctx)
    dqkv = al(M, 3 * D)
    dvk = al(B * max(Ty, Lt), 2 * D)
    dH = al(M, D, dtype=f32)   # running gradient of the T5 residual stream / SGA x-stream
    dText = al(M, D, dtype=f32)
    dY = [al(M, D, dtype=f32), al(M, D, dtype=f32)]
    dX = al(M, D, dtype=f32)
    dZ = al(M, D, dtype=f32)
    dy0 = al(My, D)            # gradient of the vision tokens (bf16, wgrad operand)
    dlogits = al(B, Apad, zero=True)
    dpooled = al(B, D, dtype=f32)
    dbias_pos = al(nH, Lt, Lt, dtype=f32)
    proj_dw = al(D, 9 * Cf, dtype=f32)

    def next_g():
        g_idx[0] ^= 1
        return g_pair[g_idx[0]]

    # Weight / bias gradients are leaves of the backward graph: they go to lane 1 and overlap the data-gradient
    # chain on lane 0.  `side.leaf(reads, fn)` orders lane 1 after the producer of its operands and remembers a
    # mark; `side.before_write(buf)` makes lane 0 wait for that mark before it overwrites a scratch buffer a leaf
    # still reads (rarely a real stall: the scratch buffers are reused several launches later).
    class _Side:
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

    side = _Side(two_lanes)

    # ---- segment 0: head + SGA ----
    bp = new_segment()
    r = eng.rec(bp, st.ks_store)
    side.bind(r)
    n_small = eng.total - eng.n_big
    r.memset_zero(eng.grad.data_ptr() + 4 * eng.n_big, 4 * n_small)
    r.memset_zero(dbias_pos, 4 * nH * Lt * Lt)
    r.logsoftmax_nll_bwd(st.logp, st.labels if has_labels else None, st.gloss, st.glogp, dlogits, Apad, B, A)
    st.glogp_used = True

    def head_leaf():
        r.colsum_bf16(dlogits, Apad, eng.gp(cls.bias), B, A)
        r.wgrad(dlogits, B, A, Apad, pooled_b, D, D, eng.gp(cls.weight), bn=64)
    side.leaf([dlogits], head_leaf)
    r.dgrad(dlogits, B, A, Apad, eng.sp(cls.weight), D, dpooled, D, out_fp32=1, bn=64)
    dOut = dY[0]
    r.pooler_bwd(out_f32, eng.mp(pl.weight), pool_w, dpooled, dOut, eng.gp(pl.weight), eng.gp(pl.bias), B, Lt, D)

    first_text = True
    for li in reversed(range(len(sgas))):
        sga, sv = sgas[li], saved_sga[li]
        m1, m2, mlp = sga.mhatt1, sga.mhatt2, sga.ffn.mlp
        Ly = sv["Ly"]
        Myl = B * Ly
        # norm3 / FFN
        # LayerNorm backward also emits the dropout-masked bf16 branch gradient and its column sums (bias grad)
        g_bf = next_g()
        side.before_write(g_bf)
        r.layernorm_bwd(dOut, sv["z3"], eng.mp(sga.norm3.norm.weight), sv["mean3"], sv["rstd3"], dZ,
                        eng.gp(sga.norm3.norm.weight), eng.gp(sga.norm3.norm.bias), M, D,
                        g_bf, p_sga, sv["sid_r3"], rng, eng.gp(mlp.fc2.bias))
        side.leaf([g_bf], lambda: r.wgrad(g_bf, M, D, D, sv["hm"], D, D, eng.gp(mlp.fc2.weight)))
        side.before_write(dpre)
        r.dgrad(g_bf, M, D, D, eng.sp(mlp.fc2.weight), D, dpre, D, relu_mask=sv["hm"], ldm=D, drop_p=p_sga,
                sid=sv["sid_h"], rng=rng)

        def fc1_leaf():
            r.colsum_bf16(dpre, D, eng.gp(mlp.fc1.bias), M, D)
            r.wgrad(dpre, M, D, D, sv["x2b"], D, D, eng.gp(mlp.fc1.weight))
        side.leaf([dpre], fc1_leaf)
        r.dgrad(dpre, M, D, D, eng.sp(mlp.fc1.weight), D, dX, D, out_fp32=1, residual=dZ, ldr=D, res_fp32=1)
        # norm2 / mhatt2
        g_bf = next_g()
        side.before_write(g_bf)
        r.layernorm_bwd(dX, sv["z2"], eng.mp(sga.norm2.norm.weight), sv["mean2"], sv["rstd2"], dZ,
                        eng.gp(sga.norm2.norm.weight), eng.gp(sga.norm2.norm.bias), M, D,
                        g_bf, p_sga, sv["sid_r2"], rng, eng.gp(m2.linear_merge.bias))
        side.leaf([g_bf], lambda: r.wgrad(g_bf, M, D, D, sv["ctx2"], D, D, eng.gp(m2.linear_merge.weight)))
        r.dgrad(g_bf, M, D, D, eng.sp(m2.linear_merge.weight), D, dsm, D)
        vk = sv["vk2"]
        side.before_write(dqkv, dvk)
        r.attn_bwd(B, Hs, Lt, Ly, hd, sv["q2"], D, vk.data_ptr() + 2 * D, 2 * D, vk, 2 * D, sv["probs2"], dsm, D,
                   dqkv, D, dvk.data_ptr() + 2 * D, 2 * D, dvk, 2 * D, None, scale, p_sga, sv["sid_p2"], rng,
                   stats=sv["stats2"])

        def att2_leaf():
            r.colsum_bf16(dqkv, D, eng.gp(m2.linear_q.bias), M, D)
            r.colsum_bf16(dvk, 2 * D, eng.gp(m2.linear_v.bias), Myl, 2 * D)
            r.wgrad(dqkv, M, D, D, sv["x1b"], D, D, eng.gp(m2.linear_q.weight))
            r.wgrad(dvk, Myl, 2 * D, 2 * D, sv["y"], D, D, eng.gp(m2.linear_v.weight))
        side.leaf([dqkv, dvk], att2_leaf)
        r.dgrad(dqkv, M, D, D, eng.sp(m2.linear_q.weight), D, dX, D, out_fp32=1, residual=dZ, ldr=D, res_fp32=1)
        if li > 0:
            dPrev = dY[1] if dOut is dY[0] else dY[0]
            r.dgrad(dvk, Myl, 2 * D, 2 * D, eng.sp(m2.linear_v.weight), D, dPrev, D, out_fp32=1)
        else:
            dPrev = None
            r.dgrad(dvk, Myl, 2 * D, 2 * D, eng.sp(m2.linear_v.weight), D, dy0, D)
        # norm1 / mhatt1
        g_bf = next_g()
        side.before_write(g_bf)
        r.layernorm_bwd(dX, sv["z1"], eng.mp(sga.norm1.norm.weight), sv["mean1"], sv["rstd1"], dZ,
                        eng.gp(sga.norm1.norm.weight), eng.gp(sga.norm1.norm.bias), M, D,
                        g_bf, p_sga, sv["sid_r1"], rng, eng.gp(m1.linear_merge.bias))
        side.leaf([g_bf], lambda: r.wgrad(g_bf, M, D, D, sv["ctx1"], D, D, eng.gp(m1.linear_merge.weight)))
        r.dgrad(g_bf, M, D, D, eng.sp(m1.linear_merge.weight), D, dsm, D)
        q1 = sv["qkv1"]
        side.before_write(dqkv)
        r.attn_bwd(B, Hs, Lt, Lt, hd, q1.data_ptr() + 4 * D, 3 * D, q1.data_ptr() + 2 * D, 3 * D, q1, 3 * D,
                   sv["probs1"], dsm, D, dqkv.data_ptr() + 4 * D, 3 * D, dqkv.data_ptr() + 2 * D, 3 * D, dqkv, 3 * D,
                   None, scale, p_sga, sv["sid_p1"], rng, stats=sv["stats1"])

        def att1_leaf():
            r.colsum_bf16(dqkv, 3 * D, eng.gp(m1.linear_v.bias), M, 3 * D)
            r.wgrad(dqkv, M, 3 * D, 3 * D, text_bf16, D, D, eng.gp(m1.linear_v.weight))
        side.leaf([dqkv], att1_leaf)
        # x is the T5 output for every layer: its gradient accumulates over the three layers
        r.dgrad(dqkv, M, 3 * D, 3 * D, eng.sp(m1.linear_v.weight), D, dText, D, out_fp32=1, residual=dZ, ldr=D,
                res_fp32=1, accumulate=0 if first_text else 1)
        first_text = False
        dOut = dPrev
    close_segment(bp, 0, o[id(proj.weight)])

    # ---- T5 encoder backward, a few blocks per segment; the projection's gradients ride on lane 1 ----
    bp = new_segment()
    r = eng.rec(bp, st.ks_store)
    side.bind(r)

    def proj_leaf():
        # bias grad + wgrad of the equivalent 3x3 conv, mapped back to the ConvTranspose2d layout
        r.colsum_bf16(dy0, D, eng.gp(proj.bias), My, D)
        wg_bn = 256 if Cf % 256 == 0 else (128 if Cf % 128 == 0 else 64)
        n_kb = (My + 63) // 64
        tiles = ((D + 127) // 128) * (9 * Cf // wg_bn)
        split = 1
        if tiles < 148 and n_kb >= 8:
            split = min(4, max(1, 296 // max(tiles, 1)), n_kb // 4)
        if split > 1:
            r.memset_zero(proj_dw, 4 * D * 9 * Cf)
        r.conv_wgrad(B, Hf, Wf, Cf, D, dy0, feat, proj_dw, wg_bn, split)
        r.convT_wgrad_unprep(proj_dw, eng.gp(proj.weight), Cf, D)
    side.leaf([dy0], proj_leaf)
    # every RMSNorm backward also writes the dropout-masked bf16 gradient the NEXT residual branch consumes
    g_bf = next_g()
    r.rmsnorm_bwd(dText, 1, hid[nblk], eng.mp(t5.final_layer_norm.weight), rstd_f, None, dH,
                  eng.gp(t5.final_layer_norm.weight), M, D, p_t5, sid_final, rng,
                  g_bf, p_t5, saved_t5[nblk - 1]["sid_f"])
    seg_lo = o[id(proj.weight)]
    for bi in reversed(range(nblk)):
        blk, sv = blocks[bi], saved_t5[bi]
        att, ff = blk.layer[0], blk.layer[1]
        sa, dd = att.SelfAttention, ff.DenseReluDense
        # FFN sub-layer
        side.leaf([g_bf], lambda: r.wgrad(g_bf, M, D, D, sv["h"], dff, dff, eng.gp(dd.wo.weight)))
        side.before_write(dpre)
        r.dgrad(g_bf, M, D, D, eng.sp(dd.wo.weight), dff, dpre, dff, relu_mask=sv["h"], ldm=dff, drop_p=p_t5,
                sid=sv["sid_h"], rng=rng)
        side.leaf([dpre], lambda: r.wgrad(dpre, M, dff, dff, sv["y2"], D, sv["ldy"], eng.gp(dd.wi.weight)))
        r.dgrad(dpre, M, dff, dff, eng.sp(dd.wi.weight), D, dsm, D)
        g_bf = next_g()
        side.before_write(g_bf)
        r.rmsnorm_bwd(dsm, 0, sv["hmid"], eng.mp(ff.layer_norm.weight), sv["rstd2"], dH, dH,
                      eng.gp(ff.layer_norm.weight), M, D, 0.0, 0, rng, g_bf, p_t5, sv["sid_o"])
        # self-attention sub-layer
        side.leaf([g_bf], lambda: r.wgrad(g_bf, M, D, D, sv["ctx"], inner, inner, eng.gp(sa.o.weight)))
        r.dgrad(g_bf, M, D, D, eng.sp(sa.o.weight), inner, dsm, inner)
        qkv = sv["qkv"]
        side.before_write(dqkv)
        r.attn_bwd(B, nH, Lt, Lt, dkv, qkv, 3 * inner, qkv.data_ptr() + 2 * inner, 3 * inner,
                   qkv.data_ptr() + 4 * inner, 3 * inner, sv["probs"], dsm, inner,
                   dqkv, 3 * inner, dqkv.data_ptr() + 2 * inner, 3 * inner, dqkv.data_ptr() + 4 * inner, 3 * inner,
                   dbias_pos, 1.0, p_t5, sv["sid_p"], rng, stats=sv["stats"], bias=pos_bias, key_mask=st.mask)
        side.leaf([dqkv], lambda: r.wgrad(dqkv, M, 3 * inner, 3 * inner, sv["y1"], D, sv["ldy"], eng.gp(sa.q.weight)))
        r.dgrad(dqkv, M, 3 * inner, 3 * inner, eng.sp(sa.q.weight), D, dsm, D)
        if bi > 0:
            g_bf = next_g()
            side.before_write(g_bf)
        r.rmsnorm_bwd(dsm, 0, hid[bi], eng.mp(att.layer_norm.weight), sv["rstd1"], dH, dH,
                      eng.gp(att.layer_norm.weight), M, D, 0.0, 0, rng,
                      g_bf if bi > 0 else None, p_t5, saved_t5[bi - 1]["sid_f"] if bi > 0 else 0)
        if bi > 0 and seg_of_block(bi, nblk) != seg_of_block(bi - 1, nblk):
            hi = o[id(blocks[bi - 1].layer[0].SelfAttention.q.weight)]
            close_segment(bp, seg_lo, hi)
            seg_lo = hi
            bp = new_segment()
            r = eng.rec(bp, st.ks_store)
            side.bind(r)
    r.t5_bias_grad(dbias_pos, bucket, eng.gp(relw), nH, Lt, cfg["num_buckets"])
    st.emb_rows = None
    if eng._ddp is not None and eng._ddp.sparse_embedding:
        # data parallel: the token rows leave the rank (scaled by 1 / world) instead of a dense 99 MB table; ddp.GradSync
        # gathers every rank's (ids, rows) and scatters them into the zeroed table in one fixed order
        st.emb_rows = al(M, D, dtype=f32)
        r.embedding_bwd_rows(dH, st.emb_rows, M, D, p_t5, sid_embed, rng, 1.0 / eng._ddp.world)
    else:
        r.embedding_bwd(st.ids, dH, eng.gp(t5.embed_tokens.weight), M, D, vocab, p_t5, sid_embed, rng)
    close_segment(bp, seg_lo, eng.total)
    st.n_bwd_launches = sum(lib.vqa_plan_size(s.plan) for s in segs)

    # warm every kernel once outside capture (function attributes, module loading), then capture graphs
    if eng.use_graphs:
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            sp = ctypes.c_void_p(side.cuda_stream)
            for fp in st.fwd_plans:
                L.check(lib.vqa_plan_run(fp, sp), "plan warm-up (fwd)")
            for s in segs:
                L.check(lib.vqa_plan_run(s.plan, sp), "plan warm-up (bwd)")
            side.synchronize()
            for fp in st.fwd_plans:
                L.check(lib.vqa_plan_capture_graph(fp, sp), "graph capture (fwd)")
            for s in segs:
                L.check(lib.vqa_plan_capture_graph(s.plan, sp), "graph capture (bwd)")
            side.synchronize()
        torch.cuda.current_stream(dev).wait_stream(side)
    return st
