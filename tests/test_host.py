"""CPU: host-side logic of the drop-in module - reference state_dict layout, flat parameter layout, tile
heuristics, optimizer registration, loud failure without CUDA, and the data-parallel gradient averaging over
gloo with world_size 2."""
import os
import socket

import pytest
import torch


@pytest.fixture(scope="module")
def model(pkg):
    os.environ["VQA_B200_PRETRAINED"] = "0"
    torch.manual_seed(0)
    return pkg.ResnetVQAModel("resnet34", "t5-base", answer_spaces=170)


def test_state_dict_layout_matches_reference(model):
    from oracle import vqa_oracle as O
    spec = O.state_dict_spec("resnet34", 170)
    sd = model.state_dict()
    assert list(sd.keys()) == [k for k, _, _ in spec]
    assert len(sd) == 403                                      # SURVEY 8b [probe]
    for k, shape, _ in spec:
        assert tuple(sd[k].shape) == tuple(shape), k
    assert sum(p.numel() for p in model.parameters()) == 166_985_555
    for v in sd.values():
        assert v.dtype in (torch.float32, torch.int64)
    # strict load of a reference-layout state_dict, both directions
    ref = O.random_state_dict("resnet34", 170, seed=0)
    model.load_state_dict(ref, strict=True)
    assert torch.equal(model.state_dict()["sga_modules.1.mhatt2.linear_k.bias"], ref["sga_modules.1.mhatt2.linear_k.bias"])


def test_faster_rcnn_state_dict_layout_matches_reference(pkg):
    """FasterRcnnVQAModel: the state_dict the unmodified reference class produced when tests/golden/frcnn_*.pt were made
    (torchvision BackboneWithFPN with FrozenBatchNorm2d + the shared head), key for key, and strict loading both ways."""
    import torch
    from oracle import vqa_oracle as O
    os.environ["VQA_B200_PRETRAINED"] = "0"
    gold = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "frcnn_b2_256_l16.pt"),
                      weights_only=False)
    m = pkg.FasterRcnnVQAModel("faster-rcnn", "t5-base", answer_spaces=170)
    assert list(m.state_dict().keys()) == gold["state_dict_keys"] and len(gold["state_dict_keys"]) == 464
    sd = O.random_state_dict("faster-rcnn", 170, seed=0)
    assert list(sd.keys()) == gold["state_dict_keys"]
    m.load_state_dict(sd, strict=True)
    assert all(torch.equal(v, sd[k]) for k, v in m.state_dict().items())
    frozen = [k for k, p in m.named_parameters() if not p.requires_grad]
    assert frozen and all(k.startswith(("vision_model.body.conv1", "vision_model.body.layer1")) for k in frozen)
    assert m._projection() is m.upscale_layer and not hasattr(m, "downscale_layer")
    with pytest.raises(ValueError):
        pkg.FasterRcnnVQAModel("resnet50", "t5-base", answer_spaces=170)


def test_trainer_facing_attributes(model):
    for name in ("vision_model", "lang_model", "upscale_layer", "downscale_layer", "sga_modules", "attention_pooler",
                 "classification_layer"):
        assert len(list(getattr(model, name).parameters())) > 0
    assert model.vision_model_name == "resnet34" and model.device == "cpu"
    import inspect
    sig = list(inspect.signature(model.forward).parameters)
    assert sig[:6] == ["question_input_ids", "decoder_question_input_ids", "question_attention_masks",
                       "decoder_question_attention_masks", "annotation_ids", "image_tensors"]
    assert set(sig[6:]) == {"answer_input_ids", "pixel_values", "answer_attention_masks", "question_type_ids"}


def test_flat_layout_covers_exactly_the_trainable_tensors(model):
    from oracle import vqa_oracle as O
    eng = model._engine
    big, small = eng._layout()
    names = {id(p): k for k, p in model.named_parameters()}
    got = sorted(names[id(p)] for p in big + small)
    want = sorted(O.trainable_keys(model.state_dict(), "resnet34"))
    assert got == want
    assert len(set(id(p) for p in big + small)) == len(big) + len(small)
    # fused projections need their weights adjacent and in this order
    order = [names[id(p)] for p in big]
    i = order.index("lang_model.block.3.layer.0.SelfAttention.q.weight")
    assert order[i + 1].endswith("block.3.layer.0.SelfAttention.k.weight")
    assert order[i + 2].endswith("block.3.layer.0.SelfAttention.v.weight")
    j = order.index("sga_modules.1.mhatt1.linear_v.weight")
    assert order[j + 1].endswith("mhatt1.linear_k.weight") and order[j + 2].endswith("mhatt1.linear_q.weight")
    # the embedding table is last: its dense gradient is all-reduced in the final bucket
    assert names[id(small[-1])] == "lang_model.embed_tokens.weight"


def test_no_cpu_fallback(model):
    from oracle import vqa_oracle as O
    b = O.synthetic_batch(1, 16, 64, 64, 170, seed=1)
    with pytest.raises(RuntimeError, match="CUDA only"):
        model(b["question_input_ids"], None, b["question_attention_masks"], None, b["annotation_ids"],
              b["image_tensors"])
    with pytest.raises(RuntimeError, match="parameter container"):
        model.sga_modules[0](torch.zeros(1, 4, 768), torch.zeros(1, 4, 768))


def test_optimizer_is_registered_where_the_trainer_looks(pkg):
    assert getattr(torch.optim, "VQAFusedAdamW") is pkg.VQAFusedAdamW
    p = torch.nn.Parameter(torch.zeros(8))
    opt = torch.optim.VQAFusedAdamW([{"params": [p], "lr": 1e-3, "model_name": "x"}], weight_decay=0.1, amsgrad=True)
    assert opt.param_groups[0]["model_name"] == "x" and opt.param_groups[0]["amsgrad"] is True
    p.grad = torch.ones(8)
    with pytest.raises(RuntimeError, match="CUDA"):
        opt.step()


def test_tile_heuristics_and_buckets(pkg):
    from t5_resnet_vqa_b200 import engine as E
    from oracle import vqa_oracle as O
    for M, N in [(2048, 768), (2048, 2304), (64, 170), (200704, 64), (3136, 2048), (128, 768)]:
        bn, ks = E.pick_tile(M, N, 768, allow_ksplit=False)
        assert bn in (64, 128, 256) and ks == 1
    assert E.pick_tile(64, 170, 768, allow_ksplit=False)[0] <= 128
    # cluster split-K is chosen only where a cluster per tile fits the 148 SMs, and for deep contractions
    for M, N, K in [(2048, 768, 3072), (2048, 768, 2304), (768, 768, 2048), (2048, 768, 768), (2048, 3072, 768)]:
        bn, ks = E.pick_tile(M, N, K)
        tiles = ((M + 127) // 128) * ((N + bn - 1) // bn)
        assert ks == 1 or (tiles * ks <= 148 and (K + 63) // 64 >= 4 * ks)
    assert E.pick_tile(2048, 768, 3072)[1] == 1      # opt-in (VQA_B200_KSPLIT=1): slower inside the step
    os.environ["VQA_B200_KSPLIT"] = "1"
    try:
        assert E.pick_tile(2048, 768, 3072)[1] > 1 and E.pick_tile(2048, 3072, 768)[1] == 1
    finally:
        del os.environ["VQA_B200_KSPLIT"]
    for L in (16, 32, 20):
        assert torch.equal(E.t5_relative_buckets(L, L).long(), O.t5_buckets(L, L))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _ddp_worker(rank, world, port, out):
    import torch.distributed as dist
    from t5_resnet_vqa_b200 import ddp
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)

    class Seg:
        def __init__(self, lo, hi):
            self.grad_lo, self.grad_hi = lo, hi
    total = 1000
    segs = [Seg(0, 300), Seg(300, 640), Seg(640, 1000)]
    ranges = ddp.segment_ranges(segs, total)
    g = torch.Generator().manual_seed(100 + rank)
    flat = torch.randn(total, generator=g)
    mine = flat.clone()
    for lo, hi in ranges:
        ddp.average_range(flat, lo, hi)
    others = [torch.randn(total, generator=torch.Generator().manual_seed(100 + r)) for r in range(world)]
    want = sum(others) / world
    ok = torch.allclose(flat, want, atol=1e-6) and not torch.equal(mine, flat)
    try:
        ddp.segment_ranges([Seg(0, 300), Seg(310, 1000)], total)
        ok = False
    except RuntimeError:
        pass
    # bf16 wire format: cast down, average the staging range, cast back; every rank ends with the SAME fp32 values, equal to
    # the mean of the ranks' bf16-rounded gradients
    flat2, stage = mine.clone(), torch.zeros(total, dtype=torch.bfloat16)
    for lo, hi in ranges:
        ddp.average_range_via(flat2, stage, lo, hi, lambda src, dst: dst.copy_(src), lambda src, dst: dst.copy_(src))
    want2 = (sum(o.bfloat16() for o in others) / world).float()
    ok = ok and torch.allclose(flat2, want2, atol=2e-2, rtol=1e-2) and torch.allclose(flat2, want, atol=3e-2)
    gathered = [torch.zeros_like(flat2) for _ in range(world)]
    dist.all_gather(gathered, flat2)
    ok = ok and all(torch.equal(gathered[0], t) for t in gathered)
    # sharded optimizer: every segment's GEMM-weight range is split evenly over the ranks; the slices of all ranks tile it
    import weakref

    class Opt:
        max_grad_norm = None
    opt = Opt()

    class Eng:
        n_big = 64 * 30
        fused_opt = weakref.ref(opt)

    class St:
        bwd_segments = [Seg(0, 64 * 8), Seg(64 * 8, 64 * 20), Seg(64 * 20, 64 * 30 + 500)]
    sync = ddp.GradSync()
    sh = sync.shards_for(Eng, St)
    ok = ok and sh is not None and len(sh) == 3 and sh[2][1] == Eng.n_big
    mine_sl = torch.tensor([[s[2], s[3]] for s in sh])
    all_sl = [torch.zeros_like(mine_sl) for _ in range(world)]
    dist.all_gather(all_sl, mine_sl)
    for k, (lo, bhi, olo, ohi) in enumerate(sh):
        cuts = sorted((int(a[k][0]), int(a[k][1])) for a in all_sl)
        ok = ok and cuts[0][0] == lo and cuts[-1][1] == bhi and all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
        ok = ok and (ohi - olo) % 8 == 0
    Eng.fused_opt = None
    ok = ok and sync.shards_for(Eng, St) is None         # no fused optimizer over the engine: plain all-reduce
    out.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_gradient_averaging_two_ranks_gloo(pkg):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_ddp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


def test_tile_model_agrees_with_the_measured_table(pkg):
    """The launch-time model behind pick_tile is fitted to profiles/r1_gemm_bench.txt (B200, graph-replayed): for every
    measured GEMM shape its choice among the single-CTA tile widths must be within 12 % of the fastest measured one."""
    import re
    from t5_resnet_vqa_b200 import engine as E
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r1_gemm_bench.txt")
    rows = []
    for line in open(path):
        m = re.match(r"(fwd|dgrad|wgrad)\s+M(\d+)\s+N(\d+)\s+K(\d+)\s+([\d.]+)\s+([\d.]+)\s+([\d.]+)", line)
        if m:
            rows.append((m.group(1), int(m.group(2)), int(m.group(3)), int(m.group(4)),
                         {64: float(m.group(5)), 128: float(m.group(6)), 256: float(m.group(7))}))
    assert len(rows) >= 12
    for kind, M, N, K, t in rows:
        bn, ks = E.pick_tile(M, N, K)
        assert ks == 1
        assert t[bn] <= 1.12 * min(t.values()), (kind, M, N, K, bn, t)


def test_vit_vqa_model_surface_and_tied_table(pkg):
    """VitVQAModel (model/vit_vqa_model.py:127-227): constructor arguments, state_dict layout of the reference (464 entries,
    the tied token table under four names), named_parameters lists it once, strict load both ways, no CPU path."""
    import os
    os.environ["VQA_B200_PRETRAINED"] = "0"
    gold = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "vit_b2_l16.pt"), weights_only=False)
    m = pkg.VitVQAModel("google/vit-base-patch16-224-in21k", "t5-base", answer_spaces=170)
    assert list(m.state_dict().keys()) == gold["state_dict_keys"] and len(gold["state_dict_keys"]) == 464
    assert [k for k, _ in m.named_parameters()] == gold["param_keys"]
    sd = m.state_dict()
    assert sd["lang_model.shared.weight"].data_ptr() == sd["lang_model.decoder.embed_tokens.weight"].data_ptr()
    assert sd["lang_model.shared.weight"].data_ptr() == sd["lang_model.lm_head.weight"].data_ptr()
    from oracle import vit_oracle as V
    m.load_state_dict(V.random_state_dict(170, seed=0), strict=True)
    for attr in ("vision_model", "lang_model", "fusing_layer", "classification_layer"):   # trainer/vit_vqa_trainer.py:300-318
        assert len(list(getattr(m, attr).parameters())) > 0
    # flat layout: exactly the tensors the reference trains (oracle.trainable_keys), each once; fused q|k|v adjacent and in
    # order in both stacks; GEMM weights first, the tied token table last
    eng = m._engine
    big, small = eng._layout()
    names = {id(p): k for k, p in m.named_parameters()}
    assert sorted(names[id(p)] for p in big + small) == sorted(V.trainable_keys(V.random_state_dict(170, seed=0)))
    assert len(set(id(p) for p in big + small)) == len(big) + len(small) == 261
    order = [names[id(p)] for p in big]
    for stack in ("encoder", "decoder"):
        i = order.index("lang_model.%s.block.5.layer.0.SelfAttention.q.weight" % stack)
        assert order[i + 1].endswith("%s.block.5.layer.0.SelfAttention.k.weight" % stack)
        assert order[i + 2].endswith("%s.block.5.layer.0.SelfAttention.v.weight" % stack)
    assert all(p.dim() == 2 and p.numel() % 64 == 0 for p in big[2:]) and names[id(small[-1])] == "lang_model.shared.weight"
    with pytest.raises(ValueError):
        pkg.VitVQAModel("resnet50", "t5-base", 170)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(2, 16, dtype=torch.long), torch.zeros(2, 20, dtype=torch.long), pixel_values=torch.zeros(2, 3, 224, 224))
