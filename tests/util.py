import ctypes

import torch


def rel_fro(a, b):
    a, b = a.float(), b.float()
    return float((a - b).norm() / (b.norm() + 1e-30))


def cosine(a, b):
    a, b = a.flatten().double(), b.flatten().double()
    return float((a @ b) / (a.norm() * b.norm() + 1e-300))


class Caller:
    """Immediate-mode (plan = NULL) calls into libvqa_b200.so on torch's current stream."""

    def __init__(self, pkg):
        from t5_resnet_vqa_b200.engine import _Rec
        self.lib = pkg.lib.load()
        self.rec = _Rec(self.lib, None, lambda: torch.cuda.current_stream().cuda_stream)

    def __getattr__(self, name):
        return getattr(self.rec, name)
