"""Parameter containers that reproduce the reference's module tree and state_dict keys.

These classes hold parameters/buffers only: every forward of the hot path is executed by
`engine.Engine` through libvqa_b200.so.  Names, shapes and initialisation follow
  - torchvision ResNet (torchvision/models/resnet.py:59-285) for `vision_model.*`
  - transformers T5Stack encoder (modeling_t5.py:46-792) for `lang_model.*`
  - model/multi_head_vision_text_attn.py:26-158 for `sga_modules.*`
  - model/resnet_vqa_model.py:14-26,64-95 for the pooler, scaling layers and classifier.
Calling one of these sub-modules directly raises: there is no torch-op fallback path.
"""
import math

import torch
import torch.nn as nn


class _Holder(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover - guard
        raise RuntimeError("%s is a parameter container; the computation runs inside ResnetVQAModel.forward "
                           "through libvqa_b200.so (no torch-op fallback)" % type(self).__name__)


# ------------------------------------------------------------------------------------------------
# leaves
# ------------------------------------------------------------------------------------------------
class Linear(_Holder):
    """nn.Linear-shaped parameters; default init = torch's (kaiming_uniform(a=sqrt(5)) => U(+-1/sqrt(fan_in)))."""

    def __init__(self, in_features, out_features, bias=True, std=None):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.weight = nn.Parameter(torch.empty(out_features, in_features))
        if bias:
            self.bias = nn.Parameter(torch.empty(out_features))
        else:
            self.register_parameter("bias", None)
        with torch.no_grad():
            if std is not None:
                self.weight.normal_(0.0, std)
            else:
                bound = 1.0 / math.sqrt(in_features)
                self.weight.uniform_(-bound, bound)
            if bias:
                bound = 1.0 / math.sqrt(in_features)
                self.bias.uniform_(-bound, bound)


class Embedding(_Holder):
    def __init__(self, num, dim, std=1.0):
        super().__init__()
        self.num_embeddings, self.embedding_dim = num, dim
        self.weight = nn.Parameter(torch.empty(num, dim).normal_(0.0, std))


class Conv2d(_Holder):
    """Bias-free Conv2d weight [O, I, k, k]; torchvision init: kaiming_normal_(fan_out, relu) (tv:208-210)."""

    def __init__(self, cin, cout, k, stride=1, padding=0):
        super().__init__()
        self.in_channels, self.out_channels = cin, cout
        self.kernel_size, self.stride, self.padding = k, stride, padding
        std = math.sqrt(2.0 / (cout * k * k))
        self.weight = nn.Parameter(torch.empty(cout, cin, k, k).normal_(0.0, std))


class BatchNorm2d(_Holder):
    def __init__(self, c, eps=1e-5):
        super().__init__()
        self.num_features, self.eps = c, eps
        self.weight = nn.Parameter(torch.ones(c))
        self.bias = nn.Parameter(torch.zeros(c))
        self.register_buffer("running_mean", torch.zeros(c))
        self.register_buffer("running_var", torch.ones(c))
        self.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))


class ConvTranspose2d(_Holder):
    """nn.ConvTranspose2d(k=3, s=1, p=1) parameters: weight [Cin, Cout, 3, 3], bias [Cout]
    (model/resnet_vqa_model.py:64-78); torch default init (fan_in = Cout * 9)."""

    def __init__(self, cin, cout, k=3):
        super().__init__()
        self.in_channels, self.out_channels, self.kernel_size = cin, cout, k
        bound = 1.0 / math.sqrt(cout * k * k)
        self.weight = nn.Parameter(torch.empty(cin, cout, k, k).uniform_(-bound, bound))
        self.bias = nn.Parameter(torch.empty(cout).uniform_(-bound, bound))


class LayerNormParams(_Holder):
    def __init__(self, d, eps=1e-5):
        super().__init__()
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(d))
        self.bias = nn.Parameter(torch.zeros(d))


class RMSNormParams(_Holder):
    def __init__(self, d, eps=1e-6):
        super().__init__()
        self.variance_epsilon = eps
        self.weight = nn.Parameter(torch.ones(d))


# ------------------------------------------------------------------------------------------------
# ResNet body (torchvision key layout)
# ------------------------------------------------------------------------------------------------
class BasicBlock(_Holder):
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1 = Conv2d(inplanes, planes, 3, stride, 1)
        self.bn1 = BatchNorm2d(planes)
        self.conv2 = Conv2d(planes, planes, 3, 1, 1)
        self.bn2 = BatchNorm2d(planes)
        self.downsample = downsample
        self.stride = stride


class Bottleneck(_Holder):
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1 = Conv2d(inplanes, planes, 1)
        self.bn1 = BatchNorm2d(planes)
        self.conv2 = Conv2d(planes, planes, 3, stride, 1)  # ResNet v1.5: stride on the 3x3 (tv:116-120)
        self.bn2 = BatchNorm2d(planes)
        self.conv3 = Conv2d(planes, planes * 4, 1)
        self.bn3 = BatchNorm2d(planes * 4)
        self.downsample = downsample
        self.stride = stride


class _Seq(nn.Sequential):
    def forward(self, *a, **k):  # pragma: no cover - guard
        raise RuntimeError("parameter container; computation runs through libvqa_b200.so")


class ResNet(_Holder):
    """conv1/bn1/layer1..4/fc with torchvision's names; avgpool/fc are never evaluated
    (model/resnet_vqa_model.py:119-121) but fc's parameters stay in the state_dict."""

    CFG = {"resnet18": (BasicBlock, [2, 2, 2, 2]), "resnet34": (BasicBlock, [3, 4, 6, 3]),
           "resnet50": (Bottleneck, [3, 4, 6, 3])}

    def __init__(self, name):
        super().__init__()
        block, layers = self.CFG[name]
        self.block_type = block
        self.inplanes = 64
        self.conv1 = Conv2d(3, 64, 7, 2, 3)
        self.bn1 = BatchNorm2d(64)
        self.layer1 = self._make_layer(block, 64, layers[0], 1)
        self.layer2 = self._make_layer(block, 128, layers[1], 2)
        self.layer3 = self._make_layer(block, 256, layers[2], 2)
        self.layer4 = self._make_layer(block, 512, layers[3], 2)
        self.fc = Linear(512 * block.expansion, 1000)
        self.out_channels = 512 * block.expansion

    def _make_layer(self, block, planes, n, stride):
        downsample = None
        if stride != 1 or self.inplanes != planes * block.expansion:
            downsample = _Seq(Conv2d(self.inplanes, planes * block.expansion, 1, stride, 0),
                              BatchNorm2d(planes * block.expansion))
        blocks = [block(self.inplanes, planes, stride, downsample)]
        self.inplanes = planes * block.expansion
        for _ in range(1, n):
            blocks.append(block(self.inplanes, planes))
        return _Seq(*blocks)


# ------------------------------------------------------------------------------------------------
# T5 encoder (HF key layout); init follows T5PreTrainedModel._init_weights with factor 1.0 (hf:541-593)
# ------------------------------------------------------------------------------------------------
class T5Attention(_Holder):
    def __init__(self, d_model, d_kv, n_heads, has_bias, num_buckets):
        super().__init__()
        inner = d_kv * n_heads
        self.q = Linear(d_model, inner, bias=False, std=(d_model * d_kv) ** -0.5)
        self.k = Linear(d_model, inner, bias=False, std=d_model ** -0.5)
        self.v = Linear(d_model, inner, bias=False, std=d_model ** -0.5)
        self.o = Linear(inner, d_model, bias=False, std=inner ** -0.5)
        if has_bias:
            self.relative_attention_bias = Embedding(num_buckets, n_heads, std=d_model ** -0.5)


class T5LayerSelfAttention(_Holder):
    def __init__(self, cfg, has_bias):
        super().__init__()
        self.SelfAttention = T5Attention(cfg["d_model"], cfg["d_kv"], cfg["num_heads"], has_bias,
                                         cfg["num_buckets"])
        self.layer_norm = RMSNormParams(cfg["d_model"], cfg["eps"])


class T5DenseActDense(_Holder):
    def __init__(self, d_model, d_ff):
        super().__init__()
        self.wi = Linear(d_model, d_ff, bias=False, std=d_model ** -0.5)
        self.wo = Linear(d_ff, d_model, bias=False, std=d_ff ** -0.5)


class T5LayerFF(_Holder):
    def __init__(self, cfg):
        super().__init__()
        self.DenseReluDense = T5DenseActDense(cfg["d_model"], cfg["d_ff"])
        self.layer_norm = RMSNormParams(cfg["d_model"], cfg["eps"])


class T5Block(_Holder):
    def __init__(self, cfg, has_bias):
        super().__init__()
        self.layer = nn.ModuleList([T5LayerSelfAttention(cfg, has_bias), T5LayerFF(cfg)])


T5_BASE = dict(vocab=32128, d_model=768, d_kv=64, d_ff=3072, num_layers=12, num_heads=12, num_buckets=32,
               max_distance=128, dropout=0.1, eps=1e-6)


class T5Encoder(_Holder):
    """`T5ForQuestionAnswering.from_pretrained("t5-base").encoder` (model/resnet_vqa_model.py:60-62)."""

    def __init__(self, cfg=None):
        super().__init__()
        cfg = dict(T5_BASE if cfg is None else cfg)
        self.cfg = cfg
        self.embed_tokens = Embedding(cfg["vocab"], cfg["d_model"], std=1.0)
        self.block = nn.ModuleList([T5Block(cfg, i == 0) for i in range(cfg["num_layers"])])
        self.final_layer_norm = RMSNormParams(cfg["d_model"], cfg["eps"])


# ------------------------------------------------------------------------------------------------
# SGA stack (model/multi_head_vision_text_attn.py)
# ------------------------------------------------------------------------------------------------
HIDDEN, HEADS, FF, DROPOUT = 768, 8, 768, 0.1


class MHAtt(_Holder):
    def __init__(self):
        super().__init__()
        self.linear_v = Linear(HIDDEN, HIDDEN)
        self.linear_k = Linear(HIDDEN, HIDDEN)
        self.linear_q = Linear(HIDDEN, HIDDEN)
        self.linear_merge = Linear(HIDDEN, HIDDEN)


class MLP(_Holder):
    def __init__(self):
        super().__init__()
        self.fc1 = Linear(HIDDEN, FF)
        self.fc2 = Linear(FF, HIDDEN)


class FFN(_Holder):
    def __init__(self):
        super().__init__()
        self.mlp = MLP()


class LayerNorm(_Holder):
    def __init__(self):
        super().__init__()
        self.norm = LayerNormParams(HIDDEN)


class SGA(_Holder):
    def __init__(self):
        super().__init__()
        self.mhatt1 = MHAtt()
        self.mhatt2 = MHAtt()
        self.ffn = FFN()
        self.norm1 = LayerNorm()
        self.norm2 = LayerNorm()
        self.norm3 = LayerNorm()


class AttentionPooler(_Holder):
    def __init__(self, hidden_size):
        super().__init__()
        self.hidden_size = hidden_size
        self.attention = nn.ModuleList([Linear(hidden_size, 1)])  # key: attention.0.{weight,bias}


# ------------------------------------------------------------------------------------------------
# Faster R-CNN backbone: ResNet-50 body with FrozenBatchNorm2d + FeaturePyramidNetwork (torchvision key layout of
# `fasterrcnn_resnet50_fpn(pretrained=True).backbone`, model/faster_rcnn_vqa_model.py:51-53)
# ------------------------------------------------------------------------------------------------
class FrozenBatchNorm2d(_Holder):
    """torchvision.ops.misc.FrozenBatchNorm2d: weight / bias / running statistics are BUFFERS (no num_batches_tracked)."""

    def __init__(self, c, eps=1e-5):
        super().__init__()
        self.num_features, self.eps = c, eps
        self.register_buffer("weight", torch.ones(c))
        self.register_buffer("bias", torch.zeros(c))
        self.register_buffer("running_mean", torch.zeros(c))
        self.register_buffer("running_var", torch.ones(c))


class Conv2dBias(_Holder):
    """Conv2d with bias and no norm (the FPN's lateral / output convs); torchvision FPN init: kaiming_uniform_(a=1), bias 0."""

    def __init__(self, cin, cout, k, stride=1, padding=0):
        super().__init__()
        self.in_channels, self.out_channels = cin, cout
        self.kernel_size, self.stride, self.padding = k, stride, padding
        bound = math.sqrt(3.0 / (cin * k * k))      # gain sqrt(2 / (1 + a^2)) = 1 with a = 1
        self.weight = nn.Parameter(torch.empty(cout, cin, k, k).uniform_(-bound, bound))
        self.bias = nn.Parameter(torch.zeros(cout))


class _FrozenBottleneck(_Holder):
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1 = Conv2d(inplanes, planes, 1)
        self.bn1 = FrozenBatchNorm2d(planes)
        self.conv2 = Conv2d(planes, planes, 3, stride, 1)
        self.bn2 = FrozenBatchNorm2d(planes)
        self.conv3 = Conv2d(planes, planes * 4, 1)
        self.bn3 = FrozenBatchNorm2d(planes * 4)
        self.downsample = downsample
        self.stride = stride


class ResNet50Body(_Holder):
    """IntermediateLayerGetter(resnet50(norm_layer=FrozenBatchNorm2d), layer1..4): no avgpool / fc."""

    def __init__(self):
        super().__init__()
        self.block_type = _FrozenBottleneck
        self.inplanes = 64
        self.conv1 = Conv2d(3, 64, 7, 2, 3)
        self.bn1 = FrozenBatchNorm2d(64)
        self.layer1 = self._make_layer(64, 3, 1)
        self.layer2 = self._make_layer(128, 4, 2)
        self.layer3 = self._make_layer(256, 6, 2)
        self.layer4 = self._make_layer(512, 3, 2)
        self.out_channels = 2048
        # fasterrcnn_resnet50_fpn(pretrained=True): trainable_backbone_layers = 3 -> conv1 and layer1 are frozen
        for mod in (self.conv1, self.layer1):
            for p in mod.parameters():
                p.requires_grad_(False)

    def _make_layer(self, planes, n, stride):
        downsample = None
        if stride != 1 or self.inplanes != planes * 4:
            downsample = _Seq(Conv2d(self.inplanes, planes * 4, 1, stride, 0), FrozenBatchNorm2d(planes * 4))
        blocks = [_FrozenBottleneck(self.inplanes, planes, stride, downsample)]
        self.inplanes = planes * 4
        for _ in range(1, n):
            blocks.append(_FrozenBottleneck(self.inplanes, planes))
        return _Seq(*blocks)


class FeaturePyramidNetwork(_Holder):
    def __init__(self, in_channels=(256, 512, 1024, 2048), out_channels=256):
        super().__init__()
        self.inner_blocks = nn.ModuleList([_Seq(Conv2dBias(c, out_channels, 1)) for c in in_channels])
        self.layer_blocks = nn.ModuleList([_Seq(Conv2dBias(out_channels, out_channels, 3, 1, 1)) for _ in in_channels])
        self.out_channels = out_channels


class BackboneWithFPN(_Holder):
    """`fasterrcnn_resnet50_fpn(pretrained=True).backbone`: keys body.* and fpn.{inner,layer}_blocks.N.0.{weight,bias}."""

    def __init__(self):
        super().__init__()
        self.body = ResNet50Body()
        self.fpn = FeaturePyramidNetwork()
        self.out_channels = 256


# ------------------------------------------------------------------------------------------------
# VitVQAModel (model/vit_vqa_model.py:127-227): transformers ViTModel and T5ForConditionalGeneration key layouts
# ------------------------------------------------------------------------------------------------
class T5LayerCrossAttention(_Holder):
    def __init__(self, cfg):
        super().__init__()
        self.EncDecAttention = T5Attention(cfg["d_model"], cfg["d_kv"], cfg["num_heads"], False, cfg["num_buckets"])
        self.layer_norm = RMSNormParams(cfg["d_model"], cfg["eps"])


class T5DecoderBlock(_Holder):
    def __init__(self, cfg, has_bias):
        super().__init__()
        self.layer = nn.ModuleList([T5LayerSelfAttention(cfg, has_bias), T5LayerCrossAttention(cfg), T5LayerFF(cfg)])


class T5Stack(_Holder):
    """encoder / decoder of T5ForConditionalGeneration; `embed_tokens` IS the model's shared table (same module object)."""

    def __init__(self, cfg, shared, is_decoder):
        super().__init__()
        self.cfg = cfg
        self.is_decoder = is_decoder
        self.embed_tokens = shared
        blk = T5DecoderBlock if is_decoder else T5Block
        self.block = nn.ModuleList([blk(cfg, i == 0) for i in range(cfg["num_layers"])])
        self.final_layer_norm = RMSNormParams(cfg["d_model"], cfg["eps"])


class T5ForConditionalGeneration(_Holder):
    """`T5ForConditionalGeneration.from_pretrained("t5-base")` (model/vit_vqa_model.py:149-150): state_dict keys shared.weight,
    encoder.embed_tokens.weight, encoder.block.*, decoder.embed_tokens.weight, decoder.block.*, lm_head.weight; the four
    token-table names are ONE tied parameter (named_parameters() lists it once, as shared.weight).  lm_head is never
    evaluated on this path (the model classifies from the decoder's last hidden state)."""

    def __init__(self, cfg=None):
        super().__init__()
        cfg = dict(T5_BASE if cfg is None else cfg)
        self.cfg = cfg
        self.shared = Embedding(cfg["vocab"], cfg["d_model"], std=1.0)
        self.encoder = T5Stack(cfg, self.shared, False)
        self.decoder = T5Stack(cfg, self.shared, True)
        self.lm_head = Linear(cfg["d_model"], cfg["vocab"], bias=False, std=1.0)
        self.lm_head.weight = self.shared.weight


VIT_BASE = dict(hidden=768, heads=12, layers=12, inter=3072, patch=16, image=224, eps=1e-12, init_range=0.02)


class _Dense(_Holder):
    def __init__(self, cin, cout, std):
        super().__init__()
        self.dense = Linear(cin, cout, std=std)
        with torch.no_grad():
            self.dense.bias.zero_()


class ViTSelfAttention(_Holder):
    def __init__(self, d, std):
        super().__init__()
        for nm in ("query", "key", "value"):
            lin = Linear(d, d, std=std)
            with torch.no_grad():
                lin.bias.zero_()
            setattr(self, nm, lin)


class ViTAttention(_Holder):
    def __init__(self, d, std):
        super().__init__()
        self.attention = ViTSelfAttention(d, std)
        self.output = _Dense(d, d, std)


class ViTLayer(_Holder):
    def __init__(self, cfg):
        super().__init__()
        d, std = cfg["hidden"], cfg["init_range"]
        self.attention = ViTAttention(d, std)
        self.intermediate = _Dense(d, cfg["inter"], std)
        self.output = _Dense(cfg["inter"], d, std)
        self.layernorm_before = LayerNormParams(d, cfg["eps"])
        self.layernorm_after = LayerNormParams(d, cfg["eps"])


class ViTEncoder(_Holder):
    def __init__(self, cfg):
        super().__init__()
        self.layer = nn.ModuleList([ViTLayer(cfg) for _ in range(cfg["layers"])])


class Conv2dPatch(_Holder):
    def __init__(self, cin, cout, k, std):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(cout, cin, k, k).normal_(0.0, std))
        self.bias = nn.Parameter(torch.zeros(cout))


class ViTPatchEmbeddings(_Holder):
    def __init__(self, cfg):
        super().__init__()
        self.projection = Conv2dPatch(3, cfg["hidden"], cfg["patch"], cfg["init_range"])


class ViTEmbeddings(_Holder):
    def __init__(self, cfg):
        super().__init__()
        n = (cfg["image"] // cfg["patch"]) ** 2
        self.cls_token = nn.Parameter(torch.empty(1, 1, cfg["hidden"]).normal_(0.0, cfg["init_range"]))
        self.position_embeddings = nn.Parameter(torch.empty(1, n + 1, cfg["hidden"]).normal_(0.0, cfg["init_range"]))
        self.patch_embeddings = ViTPatchEmbeddings(cfg)


class ViTModel(_Holder):
    """`ViTModel.from_pretrained("google/vit-base-patch16-224-in21k")` (model/vit_vqa_model.py:146-147): embeddings, 12 pre-LN
    layers, final layernorm, tanh pooler.  Frozen on this path (the reference evaluates it under torch.no_grad())."""

    def __init__(self, cfg=None):
        super().__init__()
        cfg = dict(VIT_BASE if cfg is None else cfg)
        self.cfg = cfg
        self.embeddings = ViTEmbeddings(cfg)
        self.encoder = ViTEncoder(cfg)
        self.layernorm = LayerNormParams(cfg["hidden"], cfg["eps"])
        self.pooler = _Dense(cfg["hidden"], cfg["hidden"], cfg["init_range"])
