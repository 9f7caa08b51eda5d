"""ctypes binding of libvqa_b200.so (include/vqa_b200.h).  Loud failure when the library is missing."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvqa_b200.so")

c_int, c_ll, c_vp, c_f, c_u32 = ctypes.c_int, ctypes.c_longlong, ctypes.c_void_p, ctypes.c_float, ctypes.c_uint32


class GemmArgs(ctypes.Structure):
    _fields_ = [("M", c_int), ("N", c_int), ("K", c_int),
                ("A", c_vp), ("lda", c_ll), ("a_mn", c_int),
                ("B", c_vp), ("ldb", c_ll), ("b_mn", c_int),
                ("out", c_vp), ("ldo", c_ll), ("out_fp32", c_int),
                ("bias", c_vp), ("relu", c_int),
                ("relu_mask", c_vp), ("ldm", c_ll),
                ("drop_p", c_f), ("drop_sid", c_u32), ("rng", c_vp),
                ("residual", c_vp), ("ldr", c_ll), ("res_fp32", c_int),
                ("alpha", c_f), ("bn", c_int), ("split_k", c_int)]


class ConvArgs(ctypes.Structure):
    _fields_ = [("N", c_int), ("H", c_int), ("W", c_int), ("Cin", c_int), ("Cout", c_int),
                ("R", c_int), ("S", c_int), ("stride", c_int), ("pad", c_int), ("Ho", c_int),
                ("Wo", c_int), ("stem7", c_int),
                ("x", c_vp), ("w", c_vp), ("out", c_vp), ("out_fp32", c_int),
                ("bias", c_vp), ("residual", c_vp), ("relu", c_int), ("bn", c_int)]


class ConvWgradArgs(ctypes.Structure):
    _fields_ = [("N", c_int), ("H", c_int), ("W", c_int), ("Cin", c_int), ("Cout", c_int),
                ("R", c_int), ("S", c_int), ("pad", c_int),
                ("dy", c_vp), ("x", c_vp), ("dw", c_vp), ("bn", c_int), ("split_k", c_int)]


_lib = None


def load():
    """Return the loaded library; raises RuntimeError if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "libvqa_b200.so is missing (%s): run `python -c 'import __graft_entry__ as g; g.build()'`; "
                "there is no fallback path" % LIB_PATH)
        import torch  # noqa: F401  (loads libcudart.so.12 the library links against)
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.vqa_last_error.restype = ctypes.c_char_p
        for name in EXPORTS:
            if name != "vqa_last_error":
                getattr(_lib, name).restype = c_int
    return _lib


# every symbol include/vqa_b200.h declares (tests check the library exports all of them)
EXPORTS = ["vqa_last_error", "vqa_version", "vqa_debug_set_umma", "vqa_gemm_bf16", "vqa_conv2d_bf16", "vqa_conv2d_wgrad_bf16"]


def check(rc, what=""):
    if rc != 0:
        raise RuntimeError("libvqa_b200 %s failed (rc=%d): %s" % (what, rc, load().vqa_last_error().decode()))


def stream_ptr():
    import torch
    return c_vp(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    return c_vp(t.data_ptr()) if t is not None else c_vp(0)
