"""Drop-in `ResnetVQAModel` (reference: model/resnet_vqa_model.py:28-231) whose training step runs in
hand-written sm_100a kernels.  Same constructor, forward / generate_answers signatures, attribute names and
state_dict keys as the reference, so trainer/faster_rcnn_vqa_trainer.py drives it unchanged.
"""
import os
import warnings
from collections import defaultdict

import torch
import torch.nn as nn

from . import modules as M
from .engine import Engine


_DIRECT_GRADS = os.environ.get("VQA_B200_DIRECT_GRADS", "1") != "0"


class _StepFunction(torch.autograd.Function):
    """One autograd node for the whole model: forward replays the forward plan, backward the backward plan.
    Gradients come back as views of the engine's flat fp32 gradient buffer (no per-parameter copies)."""

    @staticmethod
    def forward(ctx, model, st, has_labels, *params):
        ctx.model, ctx.st = model, st
        ctx.set_materialize_grads(False)
        logp = st.logp.clone()
        if has_labels:
            return logp, st.loss.reshape(()).clone()
        return logp

    @staticmethod
    def backward(ctx, glogp, gloss=None):
        model, st = ctx.model, ctx.st
        eng = model._engine
        params = eng.params
        # gradient accumulation across steps: an existing .grad that aliases our flat buffer must be detached
        # from it before the buffer is overwritten (checked per parameter: any of them may be frozen or cleared)
        views = eng.grad_views()
        for p, v in zip(params, views):
            g = p.grad
            if g is not None and v is not None and (g is v or g.data_ptr() == v.data_ptr()):
                p.grad = g.clone()
        eng.backward(st, gloss, glogp)
        # The fast path assigns .grad itself, so it is only taken when autograd would do exactly that for every parameter
        # (loss.backward(), no hooks).  torch.autograd.grad(loss, some_params) / backward(inputs=some_params) set
        # needs_input_grad only for the requested tensors and expect the gradients to be RETURNED: they take the autograd
        # path below.  (Asking torch.autograd.grad for ALL parameters at once cannot be told apart from backward() here:
        # set VQA_B200_DIRECT_GRADS=0 for that; INTEGRATION.md.)
        wanted = ctx.needs_input_grad[3:]
        direct = (_DIRECT_GRADS and all(w == p.requires_grad for w, p in zip(wanted, params))
                  and not any(p._backward_hooks or getattr(p, "_post_accumulate_grad_hooks", None) for p in params))
        if direct:
            # Host fast path: hand out the engine's cached gradient views directly instead of returning 180 fresh views
            # for autograd's AccumulateGrad nodes (~1 ms of host time per step, which the end-to-end step - host inputs,
            # loss read back every step - cannot hide).  Same result: `.grad` is a view of the flat gradient buffer, or,
            # when a gradient already exists (accumulation over micro-batches), that gradient plus this step's.
            # Parameters with hooks keep the autograd path below.
            for p, v in zip(params, eng.grad_views()):
                if v is None:
                    continue
                if p.grad is None:
                    p.grad = v
                else:
                    p.grad = p.grad + v
            return (None, None, None) + (None,) * len(params)
        grads = []
        g = eng.grad
        for p in params:
            if p.requires_grad:
                o = eng.offs[id(p)]
                grads.append(g[o:o + p.numel()].view(p.shape))
            else:
                grads.append(None)
        return (None, None, None) + tuple(grads)


class _VQAModelBase(nn.Module):
    """What the two CNN + T5 + SGA models of the reference share (model/resnet_vqa_model.py, model/faster_rcnn_vqa_model.py:
    identical heads and forward / generate_answers signatures; they differ in the frozen backbone and the scaling layer)."""

    def _init_common(self, vision_model_name, language_model_name, answer_spaces, fine_tune_lm_encoder,
                     fine_tune_lm_decoder, fine_tune_vision, num_attention_blocks, device):
        if language_model_name != "t5-base":
            raise ValueError("language_model_name must be 't5-base'")
        self.vision_model_name = vision_model_name
        self.language_model_name = language_model_name
        self._build_vision()            # registration order = the reference's state_dict key order
        self.lang_model = M.T5Encoder()
        self._build_scalers()
        self.sga_modules = nn.ModuleList([M.SGA() for _ in range(num_attention_blocks)])
        self.classification_layer = M.Linear(768, answer_spaces)
        self.attention_pooler = M.AttentionPooler(768)
        self.fine_tune_lm_encoder = fine_tune_lm_encoder
        self.fine_tune_lm_decoder = fine_tune_lm_decoder
        self.fine_tune_vision = fine_tune_vision
        self.device = device
        self.num_beams = 2
        self.max_answer_length = 5
        self.temperature_scaler = 1.5
        object.__setattr__(self, "_engine", Engine(self))
        # the fused optimizer may still be running on its own stream: anything that reads the parameters through
        # state_dict() (checkpoints, trainer/callbacks.py:34-46) is ordered after it
        self.register_state_dict_pre_hook(lambda module, prefix, keep_vars: (module._engine.wait_optimizer(),
                                                                              module._engine.sync_master()))
        self._load_pretrained()

    # the reference always starts from pretrained torchvision / HF weights (model/resnet_vqa_model.py:51-62);
    # offline they cannot be downloaded, so they are only used when already cached locally.
    def _load_pretrained(self):
        mode = os.environ.get("VQA_B200_PRETRAINED", "auto")
        if mode == "0":
            return
        loaded = []
        try:
            import torchvision
            name = self.vision_model_name if self._fpn() is None else "fasterrcnn_resnet50_fpn"
            weights = torchvision.models.get_model_weights(name).DEFAULT
            hub = os.path.join(torch.hub.get_dir(), "checkpoints", os.path.basename(weights.url))
            if os.path.exists(hub):
                sd = torch.load(hub, map_location="cpu")
                if self._fpn() is not None:     # the detector's checkpoint: keep its `backbone.` entries
                    sd = {k[len("backbone."):]: v for k, v in sd.items() if k.startswith("backbone.")}
                self.vision_model.load_state_dict(sd)
                loaded.append("vision")
        except Exception as e:  # pragma: no cover - depends on local caches
            if mode == "1":
                raise
            warnings.warn("pretrained %s not loaded: %s" % (self.vision_model_name, e))
        try:
            from transformers import T5ForQuestionAnswering
            hf = T5ForQuestionAnswering.from_pretrained(self.language_model_name, local_files_only=True)
            self.lang_model.load_state_dict(hf.encoder.state_dict())
            loaded.append("t5")
        except Exception as e:  # pragma: no cover - depends on local caches
            if mode == "1":
                raise
        if mode == "1" and len(loaded) != 2:
            raise RuntimeError("VQA_B200_PRETRAINED=1 but pretrained weights are not available offline")
        self.pretrained_loaded = loaded

    def _build_vision(self):
        raise NotImplementedError

    def _build_scalers(self):
        raise NotImplementedError

    def _resnet_body(self):
        """The container holding conv1 / bn1 / layer1..4 of the frozen backbone."""
        return self.vision_model

    def _fpn(self):
        """FeaturePyramidNetwork container behind the body, or None."""
        return None

    def named_parameters(self, *args, **kwargs):
        """Same iterator as nn.Module's; reading parameters through it (or parameters(), state_dict()) is ordered
        after a fused optimizer pass that may still be running on the engine's optimizer stream."""
        self._engine.wait_optimizer()
        return super().named_parameters(*args, **kwargs)

    def _projection(self):
        raise NotImplementedError

    # ------------------------------------------------------------------------------------------
    def _run(self, question_input_ids, question_attention_masks, annotation_ids, image_tensors, want_features):
        if question_input_ids.dim() != 2 or image_tensors.dim() != 4:
            raise ValueError("expected question_input_ids [B, L] and image_tensors [B, 3, H, W]")
        # the model's device decides (inputs may still be in pinned host memory: they are copied straight into
        # the plan's static device buffers, which is the trainer's `v.to(device)` folded into the step)
        dev = self.classification_layer.weight.device
        eng = self._engine
        eng._ensure(dev)
        self.vision_model.eval()  # side effect of the reference forward (model/resnet_vqa_model.py:116,127)
        B, Lt = question_input_ids.shape
        # Two image formats: float [B,3,H,W] in 0..1 (the reference collate: cv2 -> ToTensor, dataset_utils/
        # resnet_vqa_daquar_dataset.py:153-171) or, for the input edge, the uint8 RGB [B,H,W,3] array cv2 produced BEFORE
        # ToTensor - a quarter of the upload, the /255 and the layout change happen in the stem-packing kernel.
        u8 = image_tensors.dtype == torch.uint8
        if u8:
            if image_tensors.shape[3] != 3:
                raise ValueError("uint8 image_tensors must be [B, H, W, 3] (RGB, as cv2 produces them)")
            H, W = image_tensors.shape[1], image_tensors.shape[2]
        else:
            if image_tensors.shape[1] != 3:
                raise ValueError("float image_tensors must be [B, 3, H, W]")
            H, W = image_tensors.shape[2], image_tensors.shape[3]
        has_labels = annotation_ids is not None
        eng.prepare()
        st = eng.get_plan(B, Lt, H, W, self.training, has_labels, want_features, u8)
        eng.forward(st, question_input_ids, question_attention_masks, annotation_ids, image_tensors)
        needs_grad = torch.is_grad_enabled() and any(p.requires_grad for p in eng.params)
        if needs_grad:
            out = _StepFunction.apply(self, st, has_labels, *eng.params)
            logp, loss = (out if has_labels else (out, None))
        else:
            logp = st.logp.clone()
            loss = st.loss.reshape(()).clone() if has_labels else None
        feats = {k: v.clone() for k, v in st.features.items()} if want_features else None
        return logp, loss, feats

    def forward(self, question_input_ids: torch.Tensor, decoder_question_input_ids: torch.Tensor = None,
                question_attention_masks: torch.Tensor = None, decoder_question_attention_masks: torch.Tensor = None,
                annotation_ids: torch.Tensor = None, image_tensors: torch.Tensor = None,
                answer_input_ids: torch.Tensor = None, pixel_values: torch.Tensor = None,
                answer_attention_masks: torch.Tensor = None, question_type_ids: torch.Tensor = None):
        logp, loss, _ = self._run(question_input_ids, question_attention_masks, annotation_ids, image_tensors, False)
        return logp, loss

    def generate_answers(self, question_input_ids: torch.Tensor, decoder_question_input_ids: torch.Tensor = None,
                         question_attention_masks: torch.Tensor = None,
                         decoder_question_attention_masks: torch.Tensor = None, image_tensors: torch.Tensor = None,
                         annotation_ids: torch.Tensor = None, answer_input_ids: torch.Tensor = None,
                         pixel_values: torch.Tensor = None, answer_attention_masks: torch.Tensor = None,
                         question_type_ids: torch.Tensor = None):
        logp, loss, feats = self._run(question_input_ids, question_attention_masks, annotation_ids, image_tensors,
                                      True)
        return logp, loss, self._feature_dict(feats)


class ResnetVQAModel(_VQAModelBase):
    """ResNet-18/34/50 (frozen, eval) -> ConvTranspose2d 512/2048->768 -> tokens; T5-base encoder over the
    question; 3 x SGA; AttentionPooler; Linear(768 -> answer_spaces); log_softmax; NLLLoss
    (model/resnet_vqa_model.py:28-231)."""

    def __init__(self, vision_model_name: str, language_model_name: str, answer_spaces: int,
                 fine_tune_lm_encoder: bool = True, fine_tune_lm_decoder: bool = True,
                 fine_tune_vision: bool = True, num_attention_blocks=3, device="cpu"):
        super().__init__()
        if vision_model_name not in M.ResNet.CFG:
            raise ValueError("vision_model_name must be one of %s" % sorted(M.ResNet.CFG))
        self._init_common(vision_model_name, language_model_name, answer_spaces, fine_tune_lm_encoder,
                          fine_tune_lm_decoder, fine_tune_vision, num_attention_blocks, device)

    def _build_vision(self):
        self.vision_model = M.ResNet(self.vision_model_name)

    def _build_scalers(self):
        self.upscale_layer = M.ConvTranspose2d(512, 768, 3)
        self.downscale_layer = M.ConvTranspose2d(2048, 768, 3)

    def _projection(self):
        return self.downscale_layer if self.vision_model_name == "resnet50" else self.upscale_layer

    def _feature_dict(self, feats):
        d = defaultdict()
        d["features"] = feats["features"]            # backbone map [B, C, h, w] fp32 (model/resnet_vqa_model.py:189,201)
        return d


class FasterRcnnVQAModel(_VQAModelBase):
    """`fasterrcnn_resnet50_fpn().backbone` (frozen, eval; FrozenBatchNorm ResNet-50 + FPN) -> its 'pool' level (256 channels)
    -> ConvTranspose2d 256->768 -> tokens; the same T5 / SGA / pooler / classifier head (model/faster_rcnn_vqa_model.py:28-197).
    generate_answers returns the FPN's five feature maps ('0'..'3', 'pool'), fp32 NCHW."""

    def __init__(self, vision_model_name: str, language_model_name: str, answer_spaces: int,
                 fine_tune_lm_encoder: bool = True, fine_tune_lm_decoder: bool = True,
                 fine_tune_vision: bool = True, num_attention_blocks=3, device="cpu"):
        super().__init__()
        if vision_model_name != "faster-rcnn":
            raise ValueError("vision_model_name must be 'faster-rcnn'")
        self._init_common(vision_model_name, language_model_name, answer_spaces, fine_tune_lm_encoder,
                          fine_tune_lm_decoder, fine_tune_vision, num_attention_blocks, device)

    def _build_vision(self):
        self.vision_model = M.BackboneWithFPN()

    def _build_scalers(self):
        self.upscale_layer = M.ConvTranspose2d(256, 768, 3)

    def _resnet_body(self):
        return self.vision_model.body

    def _fpn(self):
        return self.vision_model.fpn

    def _projection(self):
        return self.upscale_layer

    def _feature_dict(self, feats):
        d = defaultdict()
        for k in ("0", "1", "2", "3", "pool"):       # model/faster_rcnn_vqa_model.py:150-154
            d[k] = feats[k]
        return d


class VitVQAModel(nn.Module):
    """Drop-in `VitVQAModel` (reference: model/vit_vqa_model.py:127-351; SURVEY.md 8f-4, BASELINE.json configs[4]): frozen
    ViT-B/16 pooled output + T5 encoder token 0 -> Linear(1536, 768) + ReLU + Dropout(0.5) -> T5 decoder over the decoder
    question with a one-token cross-attention -> last un-padded position -> Linear(768, answers) -> log_softmax -> NLLLoss.
    Same constructor / forward signature, attribute names (`vision_model`, `lang_model`, `fusing_layer`,
    `classification_layer`: what trainer/vit_vqa_trainer.py:300-318 builds its parameter groups from) and state_dict keys
    (464 entries, the tied token table under its four names) as the reference."""

    def __init__(self, vision_model_name: str, language_model_name: str, answer_spaces: int,
                 fine_tune_lm_encoder: bool = True, fine_tune_lm_decoder: bool = True, fine_tune_vision: bool = True,
                 device="cpu"):
        super().__init__()
        if vision_model_name != "google/vit-base-patch16-224-in21k":
            raise ValueError("vision_model_name must be 'google/vit-base-patch16-224-in21k'")
        if language_model_name != "t5-base":
            raise ValueError("language_model_name must be 't5-base'")
        from .vit_step import VitEngine
        self.vision_model_name = vision_model_name
        self.language_model_name = language_model_name
        self.vision_model = M.ViTModel()
        self.lang_model = M.T5ForConditionalGeneration()
        self.fusing_layer = nn.ModuleList([M.Linear(768 + 768, 768)])      # keys fusing_layer.0.{weight,bias} (nn.Sequential)
        self.classification_layer = M.Linear(768, answer_spaces)
        self.fine_tune_lm_encoder = fine_tune_lm_encoder
        self.fine_tune_lm_decoder = fine_tune_lm_decoder
        self.fine_tune_vision = fine_tune_vision
        self.device = device
        self.num_beams = 2
        self.max_answer_length = 5
        object.__setattr__(self, "_engine", VitEngine(self))
        self.register_state_dict_pre_hook(lambda module, prefix, keep_vars: (module._engine.wait_optimizer(),
                                                                              module._engine.sync_master()))
        self._load_pretrained()

    def _load_pretrained(self):
        """The reference starts from pretrained HF weights (:146-150); offline they are used only when cached locally."""
        mode = os.environ.get("VQA_B200_PRETRAINED", "auto")
        self.pretrained_loaded = []
        if mode == "0":
            return
        try:
            from transformers import T5ForConditionalGeneration, ViTModel
            vit = ViTModel.from_pretrained(self.vision_model_name, local_files_only=True)
            self.vision_model.load_state_dict(vit.state_dict())
            self.pretrained_loaded.append("vision")
            t5 = T5ForConditionalGeneration.from_pretrained(self.language_model_name, local_files_only=True)
            self.lang_model.load_state_dict(t5.state_dict())
            self.pretrained_loaded.append("t5")
        except Exception:  # pragma: no cover - depends on local caches
            if mode == "1":
                raise

    def named_parameters(self, *args, **kwargs):
        self._engine.wait_optimizer()
        return super().named_parameters(*args, **kwargs)

    def _run(self, question_input_ids, decoder_question_input_ids, question_attention_masks,
             decoder_question_attention_masks, annotation_ids, pixel_values, want_attn=False):
        if pixel_values is None or decoder_question_input_ids is None:
            raise ValueError("VitVQAModel needs pixel_values and decoder_question_input_ids")
        if question_input_ids.dim() != 2 or decoder_question_input_ids.dim() != 2 or pixel_values.dim() != 4:
            raise ValueError("expected question_input_ids [B, L], decoder_question_input_ids [B, Ld], pixel_values [B,3,H,W]")
        dev = self.classification_layer.weight.device
        eng = self._engine
        eng._ensure(dev)
        B, Lt = question_input_ids.shape
        Ld = decoder_question_input_ids.shape[1]
        has_labels = annotation_ids is not None
        eng.prepare()
        st = eng.get_plan(B, Lt, Ld, pixel_values.shape[2], pixel_values.shape[3], self.training, has_labels, want_attn)
        eng.forward(st, question_input_ids, question_attention_masks, decoder_question_input_ids,
                    decoder_question_attention_masks, annotation_ids, pixel_values.float())
        if torch.is_grad_enabled() and any(p.requires_grad for p in eng.params):
            out = _StepFunction.apply(self, st, has_labels, *eng.params)
            logp, loss = (out if has_labels else (out, None))
        else:
            logp = st.logp.clone()
            loss = st.loss.reshape(()).clone() if has_labels else None
        return logp, loss, st

    def forward(self, question_input_ids: torch.Tensor, decoder_question_input_ids: torch.Tensor = None,
                question_attention_masks: torch.Tensor = None, decoder_question_attention_masks: torch.Tensor = None,
                annotation_ids: torch.Tensor = None, pixel_values: torch.Tensor = None, image_tensors: torch.Tensor = None,
                answer_input_ids: torch.Tensor = None, answer_attention_masks: torch.Tensor = None,
                question_type_ids: torch.Tensor = None):
        logp, loss, _ = self._run(question_input_ids, decoder_question_input_ids, question_attention_masks,
                                  decoder_question_attention_masks, annotation_ids, pixel_values)
        return logp, loss

    def vision_pooler_output(self):
        """pooler_output of the frozen ViT for the last forward's batch, fp32 [B, 768] (what :186 feeds the fusing layer)."""
        return self._engine.last_state.pooled.clone()

    def generate_answers(self, question_input_ids: torch.Tensor, decoder_question_input_ids: torch.Tensor = None,
                         question_attention_masks: torch.Tensor = None,
                         decoder_question_attention_masks: torch.Tensor = None, pixel_values: torch.Tensor = None,
                         image_tensors: torch.Tensor = None, answer_input_ids: torch.Tensor = None,
                         answer_attention_masks: torch.Tensor = None, annotation_ids: torch.Tensor = None,
                         question_type_ids: torch.Tensor = None):
        """(log_probs, loss or None, attentions) as model/vit_vqa_model.py:229-293: attentions = the ViT's twelve per-layer
        attention maps, fp32 [B, 12, 197, 197] (ViTModel(..., output_attentions=True), :240-243; what ViT_vqa_heatmap.py rolls
        out), written by the attention kernel itself on this path."""
        logp, loss, st = self._run(question_input_ids, decoder_question_input_ids, question_attention_masks,
                                   decoder_question_attention_masks, annotation_ids, pixel_values, want_attn=True)
        return logp, loss, tuple(t.clone() for t in st.attentions)
