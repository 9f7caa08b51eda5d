"""TEST INFRASTRUCTURE: loads the reference's own trainer class (trainer/faster_rcnn_vqa_trainer.py) unmodified and builds an
instance without its DAQUAR data pipeline, so that its `_init_optimizer`, `_init_lr_scheduler` and `train_one_step` can drive
the B200 model (SURVEY.md 8b/8c "trainer runs unchanged").

The reference lives at /root/reference in the build container and does not travel to the GPU box; VQA_REFERENCE_DIR points at
another copy when one is staged (e.g. unpacked into /tmp by the gpurun command that carries it).  Modules the trainer imports
but that are neither installed nor used by the three methods (nltk, rouge_score, albumentations) are stubbed in sys.modules;
wandb runs with WANDB_MODE=disabled.  Nothing here is imported by the package.
"""
import importlib
import os
import sys
import types

REFERENCE_DIR = os.environ.get("VQA_REFERENCE_DIR", "/root/reference")

# vit_daquar_config.json of the reference: optimizer / schedule / trainer settings the three methods read
OPTIMIZER_KWARGS = {"type": "AdamW", "kwargs": {"weight_decay": 0.1, "amsgrad": True}, "lm_encoder_lr": 0.005,
                    "lm_decoder_lr": 0.0001, "vision_lr": 0.008, "classifier_lr": 0.00001, "default_lr": 0.00005}
LR_SCHEDULER_KWARGS = {"num_warmup_steps": -1, "num_training_steps": -1, "max_warmup_steps": 10000}
GRADIENT_CLIPPING = 1.0


def reference_available():
    return os.path.exists(os.path.join(REFERENCE_DIR, "trainer", "faster_rcnn_vqa_trainer.py"))


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__vqa_stub__ = True
    sys.modules[name] = m
    return m


def load_reference_trainer_class(kind="cnn"):
    """FasterRcnnVQATrainer (kind "cnn") or ViTVQATrainer (kind "vit") from the unmodified reference file (None when the
    reference is not on this machine)."""
    if not reference_available():
        return None
    os.environ.setdefault("WANDB_MODE", "disabled")
    for name in ("nltk", "rouge_score", "albumentations"):
        try:
            importlib.import_module(name)
        except ImportError:
            if name == "nltk":
                _stub("nltk")
                _stub("nltk.translate")
                _stub("nltk.translate.bleu_score", sentence_bleu=None, corpus_bleu=None)
                _stub("nltk.corpus", wordnet=None)
            elif name == "rouge_score":
                _stub("rouge_score", rouge_scorer=types.SimpleNamespace())
            else:
                alb = _stub("albumentations")

                def _any_transform(attr):   # any transform constructor (dataset_utils/enums.py builds a table of them)
                    if attr.startswith("__"):
                        raise AttributeError(attr)
                    return lambda *a, **k: None
                alb.__getattr__ = _any_transform
                _stub("albumentations.pytorch", ToTensorV2=lambda *a, **k: None)
    if REFERENCE_DIR not in sys.path:
        sys.path.insert(0, REFERENCE_DIR)
    if kind == "vit":      # trainer/vit_vqa_trainer.py:23 (drives VitVQAModel; same three methods)
        # trainer/vit_vqa_trainer.py:10 imports dataset_utils.vit_vqa_dataset, a module the reference repository does not
        # contain (its OK-VQA pipeline was never committed): stubbed, like the uninstalled packages above
        if not os.path.exists(os.path.join(REFERENCE_DIR, "dataset_utils", "vit_vqa_dataset.py")):
            importlib.import_module("dataset_utils")
            _stub("dataset_utils.vit_vqa_dataset", VitT5CollateFn=None, OKVQADataset=None)
        return importlib.import_module("trainer.vit_vqa_trainer").ViTVQATrainer
    mod = importlib.import_module("trainer.faster_rcnn_vqa_trainer")
    return mod.FasterRcnnVQATrainer


class _NullLogger:
    def __getattr__(self, name):
        return lambda *a, **k: None


def make_trainer(cls, model, total_train_batch=20, epochs=1, optimizer_type="AdamW", output_dir="/tmp/vqa_b200_trainer"):
    """An instance of the reference trainer with exactly the attributes its __init__ would have set before calling
    _init_optimizer / _init_lr_scheduler (trainer/faster_rcnn_vqa_trainer.py:44-124), minus dataloaders and wandb."""
    t = object.__new__(cls)
    t.model = model
    t.epochs = epochs
    t.gradient_clipping = GRADIENT_CLIPPING
    t.output_dir = output_dir
    t.logger = _NullLogger()
    t.total_train_batch = total_train_batch
    t.num_training_steps = total_train_batch * epochs
    t.num_warmup_steps = min(t.num_training_steps // 10, LR_SCHEDULER_KWARGS["max_warmup_steps"])
    okw = dict(OPTIMIZER_KWARGS, type=optimizer_type)
    t._init_optimizer(okw, False)
    t._init_lr_scheduler(dict(LR_SCHEDULER_KWARGS))
    return t


def collate_batch(batch, L_dec=20):
    """The dict DaquarFasterRcnnT5CollateFn returns in training mode (dataset_utils/resnet_vqa_daquar_dataset.py:216-227),
    including the keys the model must accept and ignore, from an oracle.synthetic_batch."""
    import torch
    B = batch["question_input_ids"].shape[0]
    g = torch.Generator().manual_seed(7)
    dec = torch.randint(2, 32100, (B, L_dec), generator=g)
    return {"question_input_ids": batch["question_input_ids"], "decoder_question_input_ids": dec,
            "question_attention_masks": batch["question_attention_masks"],
            "decoder_question_attention_masks": torch.ones_like(dec), "annotation_ids": batch["annotation_ids"],
            "pixel_values": None, "image_tensors": batch["image_tensors"], "question_type_ids": None,
            "answer_input_ids": torch.randint(2, 32100, (B, 5), generator=g),
            "answer_attention_masks": torch.ones(B, 5, dtype=torch.long)}


def vit_collate_batch(batch):
    """The dict DaquarVitT5CollateFn returns in training mode (dataset_utils/vit_vqa_daquar_dataset.py:168-178) from an
    oracle.vit_oracle.synthetic_batch: the keys VitVQAModel.forward uses plus the ones it must accept and ignore."""
    import torch
    B = batch["question_input_ids"].shape[0]
    g = torch.Generator().manual_seed(7)
    return {"question_input_ids": batch["question_input_ids"],
            "decoder_question_input_ids": batch["decoder_question_input_ids"],
            "question_attention_masks": batch["question_attention_masks"],
            "decoder_question_attention_masks": batch["decoder_question_attention_masks"],
            "annotation_ids": batch["annotation_ids"], "pixel_values": batch["pixel_values"], "image_tensors": None,
            "answer_input_ids": torch.randint(2, 32100, (B, 20), generator=g),
            "answer_attention_masks": torch.ones(B, 20, dtype=torch.long)}
