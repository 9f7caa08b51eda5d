"""B200-native ResnetVQAModel training step (drop-in for shiv-vignesh/T5-Resnet-VQA's hot path).

Host code is PyTorch (device memory, streams, torch.distributed); every arithmetic op on the path runs
in hand-written sm_100a CUDA kernels inside libvqa_b200.so, reached through the C ABI of
include/vqa_b200.h via ctypes.  There is no CPU or torch-op fallback: without the library or without
a GPU the compute entry points raise.
"""
from . import lib  # noqa: F401
from .model import FasterRcnnVQAModel, ResnetVQAModel, VitVQAModel  # noqa: F401
from .optim import VQAFusedAdamW, register as _register_optim  # noqa: F401

_register_optim()

__all__ = ["lib", "ResnetVQAModel", "FasterRcnnVQAModel", "VitVQAModel", "VQAFusedAdamW"]
