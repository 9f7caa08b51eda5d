"""The reference's own trainer drives the B200 model unchanged (SURVEY.md 8b "trainer runs unchanged", 8c recipe).

`FasterRcnnVQATrainer` is imported from the UNMODIFIED reference (tests/trainer_harness.py: stubs for the uninstalled nltk /
rouge_score / albumentations, object.__new__ instead of the DAQUAR data pipeline) and its `_init_optimizer` (JSON "type" =
"AdamW" and "VQAFusedAdamW"), `_init_lr_scheduler` and `train_one_step` run against `ResnetVQAModel` with the collate
function's dict (dataset_utils/resnet_vqa_daquar_dataset.py:216-227, incl. the keys the model ignores).

The reference does not travel to the GPU box (no /root/reference there).  So:
  * CPU (here): the real trainer class runs everything up to the device boundary - optimizer groups, schedule, and
    `train_one_step` up to `self.model(**data_items)`, which binds and then refuses to run without a CUDA device (there is no
    CPU fallback) - and the contract it exposes is frozen in tests/golden/trainer_contract.json;
  * GPU: with a staged reference (VQA_REFERENCE_DIR) the real `train_one_step` runs three steps; without one, the same three
    steps run through bench.py's restatement of the loop and are checked against the frozen contract.
"""
import inspect
import json
import os

import pytest
import torch

import trainer_harness as H

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "trainer_contract.json")
TOTAL = 20


def _model(pkg, vision="resnet50"):
    os.environ["VQA_B200_PRETRAINED"] = "0"
    return pkg.ResnetVQAModel(vision, "t5-base", answer_spaces=170)


def _contract(trainer, steps=6):
    groups = [dict(model_name=g["model_name"], initial_lr=g["initial_lr"], n_params=len(g["params"]),
                   weight_decay=g["weight_decay"], amsgrad=g["amsgrad"]) for g in trainer.optimizer.param_groups]
    lrs = []
    for _ in range(steps):
        lrs.append([g["lr"] for g in trainer.optimizer.param_groups])
        trainer.optimizer._opt_called = True      # silence the "scheduler before optimizer" warning: no step is taken here
        trainer.lr_scheduler.step()
    return dict(groups=groups, lr_by_step=lrs)


@pytest.mark.skipif(not H.reference_available(), reason="the reference is only present in the build container")
@pytest.mark.parametrize("opt_type", ["AdamW", "VQAFusedAdamW"])
def test_reference_trainer_builds_optimizer_and_schedule_on_our_model(pkg, opt_type):
    cls = H.load_reference_trainer_class()
    model = _model(pkg)
    tr = H.make_trainer(cls, model, total_train_batch=TOTAL, optimizer_type=opt_type)
    assert type(tr.optimizer).__name__ == opt_type
    names = [g["model_name"] for g in tr.optimizer.param_groups]
    assert names == ["Vision Model", "Language Model", "DownScaler Layer", "Self-Guided Attention Module", "Attention Pooler",
                     "Classifier Layer"]
    c = _contract(tr)
    if os.environ.get("VQA_WRITE_GOLDEN") == "1" and opt_type == "AdamW":
        with open(GOLD, "w") as f:
            json.dump(c, f, indent=1)
    with open(GOLD) as f:
        gold = json.load(f)
    assert c == gold          # the fused optimizer exposes exactly what torch.optim.AdamW exposes to the trainer


@pytest.mark.skipif(not H.reference_available(), reason="the reference is only present in the build container")
def test_reference_train_one_step_reaches_the_model_call(pkg):
    """train_one_step: zero_grad -> self.model(**data_items) with the whole collate dict.  On a machine without a CUDA device the
    call must bind (every key accepted, the four ignored ones included) and then fail loudly - no CPU fallback."""
    from oracle import vqa_oracle as O
    cls = H.load_reference_trainer_class()
    model = _model(pkg, "resnet18")
    tr = H.make_trainer(cls, model, total_train_batch=TOTAL, optimizer_type="VQAFusedAdamW")
    batch = H.collate_batch(O.synthetic_batch(2, 16, 64, 64, 170, seed=1))
    inspect.signature(model.forward).bind(**batch)
    if torch.cuda.is_available():
        pytest.skip("covered by the GPU test")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        tr.train_one_step(batch)


def _restated_trainer(model, opt_type):
    """bench.py's restatement of _init_optimizer / _init_lr_scheduler / train_one_step (used where the reference is absent)."""
    import bench

    class T:
        pass
    t = T()
    os.environ["VQA_BENCH_OPTIMIZER"] = opt_type
    t.model = model
    t.optimizer, t.lr_scheduler = bench.build_trainer_objects(model, TOTAL)
    t.train_one_step = lambda batch: (lambda loss: (loss, None))(bench.train_one_step(model, t.optimizer, t.lr_scheduler,
                                                                                     batch, read_loss=True))
    return t


@pytest.mark.gpu
def test_trainer_drives_three_steps(pkg, cuda):
    """Three train_one_step calls (real trainer when staged, else the restated loop), "AdamW" vs "VQAFusedAdamW" from the same
    start with dropout off: same losses and the same parameters afterwards."""
    from oracle import vqa_oracle as O
    cls = H.load_reference_trainer_class()
    sd = O.random_state_dict("resnet50", 170, seed=0)
    batch = H.collate_batch(O.synthetic_batch(4, 32, 224, 224, 170, seed=1, masked_tail=5))
    with open(GOLD) as f:
        gold = json.load(f)
    results = {}
    for opt_type in ("AdamW", "VQAFusedAdamW"):
        model = _model(pkg)
        model.load_state_dict(sd, strict=True)
        model.to(cuda).eval()          # train_one_step does not touch the mode; dropout off makes the two runs comparable
        tr = H.make_trainer(cls, model, total_train_batch=TOTAL, optimizer_type=opt_type) if cls is not None \
            else _restated_trainer(model, opt_type)
        assert [g["model_name"] for g in tr.optimizer.param_groups] == [g["model_name"] for g in gold["groups"]]
        assert [len(g["params"]) for g in tr.optimizer.param_groups] == [g["n_params"] for g in gold["groups"]]
        # the trainer's loop moves tensor values to the device and passes the rest through (:327-329)
        data = {k: (v.to(cuda) if torch.is_tensor(v) else v) for k, v in batch.items()}
        losses = []
        for step in range(3):
            assert [g["lr"] for g in tr.optimizer.param_groups] == pytest.approx(gold["lr_by_step"][step])
            loss, logits = tr.train_one_step(data)
            assert isinstance(loss, float) and loss == loss
            if logits is not None:
                assert logits.shape == (4, 170)
                assert tr.convert_logits_to_predictions(logits).shape == (4,)
            losses.append(loss)
        torch.cuda.synchronize()
        results[opt_type] = (losses, {k: p.detach().float().cpu().clone() for k, p in model.named_parameters()})
        assert all(g is None for g in (p.grad for p in model.vision_model.parameters()))
    la, lb = results["AdamW"][0], results["VQAFusedAdamW"][0]
    assert la[0] == lb[0] and la[1] == la[0] and lb[1] == lb[0]    # the warm-up's step 0 runs at lr = 0: nothing moves yet
    assert la[2] != la[0] and lb[2] != lb[0]                       # step 1 (lr > 0) has moved the weights
    assert max(abs(a - b) for a, b in zip(la, lb)) < 2e-3 * abs(la[0])
    pa, pb = results["AdamW"][1], results["VQAFusedAdamW"][1]
    # same update, tensor by tensor (element-exact equality of the kernel with torch.optim.AdamW on the SAME gradients is
    # test_kernels_gpu.py's job; here the two runs' gradients differ by bf16 re-rounding of weights that moved by 1e-7)
    moved = 0
    for k in pa:
        da, db = pa[k] - sd[k].float(), pb[k] - sd[k].float()
        if k.startswith(("vision_model.", "upscale_layer.")):
            assert float(da.abs().max()) == 0.0 and float(db.abs().max()) == 0.0, k      # frozen / unused: untouched
            continue
        assert float(da.norm()) > 0, k
        moved += 1
        if k.endswith("linear_k.bias") or k == "attention_pooler.attention.0.bias":
            continue       # gradient is mathematically zero (softmax is shift invariant): Adam normalises pure rounding noise
        assert float((da - db).norm()) <= 0.05 * float(da.norm()), (k, float((da - db).norm() / da.norm()))
    assert moved == 183
    out = os.path.join(os.path.dirname(GOLD), "..", "..", "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "trainer_dropin.json"), "w") as f:
        json.dump(dict(trainer="reference FasterRcnnVQATrainer (staged at %s)" % H.REFERENCE_DIR if cls is not None
                       else "restated loop (bench.py); reference not on this machine",
                       losses=dict(AdamW=la, VQAFusedAdamW=lb)), f)


# --------------------------------------------------------------------------------------------------
# VitVQAModel under the reference's ViTVQATrainer (trainer/vit_vqa_trainer.py:300-341,450-464)
# --------------------------------------------------------------------------------------------------
def _vit_model(pkg):
    os.environ["VQA_B200_PRETRAINED"] = "0"
    return pkg.VitVQAModel("google/vit-base-patch16-224-in21k", "t5-base", answer_spaces=170)


@pytest.mark.skipif(not H.reference_available(), reason="the reference is only present in the build container")
@pytest.mark.parametrize("opt_type", ["AdamW", "VQAFusedAdamW"])
def test_reference_vit_trainer_builds_optimizer_and_schedule_on_our_model(pkg, opt_type):
    cls = H.load_reference_trainer_class("vit")
    tr = H.make_trainer(cls, _vit_model(pkg), total_train_batch=TOTAL, optimizer_type=opt_type)
    assert type(tr.optimizer).__name__ == opt_type
    groups = tr.optimizer.param_groups
    assert [g["model_name"] for g in groups] == ["Vision Model", "Language Model", "Fusion Layer", "Classifier Layer"]
    assert [len(g["params"]) for g in groups] == [200, 257, 2, 2]          # the tied token table once, under lang_model
    assert [g["initial_lr"] for g in groups] == [0.008, 0.005, 0.00001, 0.00001]
    assert all(g["weight_decay"] == 0.1 and g["amsgrad"] for g in groups)


@pytest.mark.gpu
def test_vit_trainer_drives_three_steps(pkg, cuda):
    """Three `train_one_step` calls of the UNMODIFIED ViTVQATrainer on the collate dict, "AdamW" vs "VQAFusedAdamW" from the same
    start with dropout off: same losses, same parameter updates.  Needs a staged reference (VQA_REFERENCE_DIR) on the GPU box."""
    if not H.reference_available():
        pytest.skip("no staged reference on this machine (the optimizer / schedule contract is checked on CPU)")
    from oracle import vit_oracle as V
    cls = H.load_reference_trainer_class("vit")
    sd = V.random_state_dict(170, seed=0)
    batch = H.vit_collate_batch(V.synthetic_batch(4, 16, 20, 170, seed=1, masked_tail=3))
    results = {}
    for opt_type in ("AdamW", "VQAFusedAdamW"):
        model = _vit_model(pkg)
        model.load_state_dict(sd, strict=True)
        model.to(cuda).eval()
        tr = H.make_trainer(cls, model, total_train_batch=TOTAL, optimizer_type=opt_type)
        data = {k: (v.to(cuda) if torch.is_tensor(v) else v) for k, v in batch.items()}
        losses = []
        for _ in range(3):
            loss, logits = tr.train_one_step(data)
            assert isinstance(loss, float) and loss == loss and logits.shape == (4, 170)
            losses.append(loss)
        torch.cuda.synchronize()
        results[opt_type] = (losses, {k: p.detach().float().cpu().clone() for k, p in model.named_parameters()})
        assert all(p.grad is None for p in model.vision_model.parameters())
    la, lb = results["AdamW"][0], results["VQAFusedAdamW"][0]
    assert la[0] == lb[0] and la[1] == la[0]            # warm-up step 0 runs at lr = 0
    assert la[2] != la[0] and lb[2] != lb[0]
    assert max(abs(a - b) for a, b in zip(la, lb)) < 5e-3 * abs(la[0])
    pa, pb = results["AdamW"][1], results["VQAFusedAdamW"][1]
    for k in pa:
        da, db = pa[k] - sd[k].float(), pb[k] - sd[k].float()
        if k.startswith("vision_model."):
            assert float(da.abs().max()) == 0.0 and float(db.abs().max()) == 0.0, k
        else:
            assert float(da.norm()) > 0 and float((da - db).norm()) <= 0.1 * float(da.norm()), k
    out = os.path.join(os.path.dirname(GOLD), "..", "..", "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "vit_trainer_dropin.json"), "w") as f:
        json.dump(dict(trainer="reference ViTVQATrainer (staged at %s)" % H.REFERENCE_DIR,
                       losses=dict(AdamW=la, VQAFusedAdamW=lb)), f)
