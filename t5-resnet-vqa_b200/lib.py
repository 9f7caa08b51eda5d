"""ctypes binding of libvqa_b200.so (include/vqa_b200.h).  Loud failure when the library is missing."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# VQA_B200_LIB selects another build of the same sources (the diagnostic libvqa_b200_dbg.so); never a fallback
LIB_PATH = os.environ.get("VQA_B200_LIB") or os.path.join(_HERE, "libvqa_b200.so")

c_int, c_ll, c_vp, c_f, c_u32, c_d = (ctypes.c_int, ctypes.c_longlong, ctypes.c_void_p, ctypes.c_float,
                                      ctypes.c_uint32, ctypes.c_double)


class GemmArgs(ctypes.Structure):
    _fields_ = [("M", c_int), ("N", c_int), ("K", c_int),
                ("A", c_vp), ("lda", c_ll), ("a_mn", c_int),
                ("B", c_vp), ("ldb", c_ll), ("b_mn", c_int),
                ("out", c_vp), ("ldo", c_ll), ("out_fp32", c_int),
                ("bias", c_vp), ("relu", c_int),
                ("relu_mask", c_vp), ("ldm", c_ll),
                ("drop_p", c_f), ("drop_sid", c_u32), ("rng", c_vp),
                ("residual", c_vp), ("ldr", c_ll), ("res_fp32", c_int), ("res_first", c_int),
                ("alpha", c_f), ("accumulate", c_int), ("bn", c_int), ("split_k", c_int), ("cta_pair", c_int),
                ("ksplit", c_int), ("ks_ws", c_vp), ("ks_ws_bytes", c_ll),
                ("B_lo", c_vp), ("a_lo_col", c_ll), ("max_ctas", c_int)]


class ConvArgs(ctypes.Structure):
    _fields_ = [("N", c_int), ("H", c_int), ("W", c_int), ("Cin", c_int), ("Cout", c_int),
                ("R", c_int), ("S", c_int), ("stride", c_int), ("pad", c_int), ("Ho", c_int),
                ("Wo", c_int), ("stem7", c_int),
                ("x", c_vp), ("w", c_vp), ("out", c_vp), ("out_fp32", c_int),
                ("bias", c_vp), ("residual", c_vp), ("relu", c_int), ("bn", c_int), ("cta_pair", c_int),
                ("ksplit", c_int), ("ks_ws", c_vp), ("ks_ws_bytes", c_ll)]


class ConvWgradArgs(ctypes.Structure):
    _fields_ = [("N", c_int), ("H", c_int), ("W", c_int), ("Cin", c_int), ("Cout", c_int),
                ("R", c_int), ("S", c_int), ("pad", c_int),
                ("dy", c_vp), ("x", c_vp), ("dw", c_vp), ("bn", c_int), ("split_k", c_int), ("cta_pair", c_int)]


class AttnFwdArgs(ctypes.Structure):
    _fields_ = [("B", c_int), ("H", c_int), ("Lq", c_int), ("Lk", c_int), ("hd", c_int),
                ("q", c_vp), ("ldq", c_ll), ("k", c_vp), ("ldk", c_ll), ("v", c_vp), ("ldv", c_ll),
                ("out", c_vp), ("ldo", c_ll), ("probs", c_vp), ("bias", c_vp), ("key_mask", c_vp),
                ("scale", c_f), ("drop_p", c_f), ("sid", c_u32), ("rng", c_vp), ("stats", c_vp)]


class AttnBwdArgs(ctypes.Structure):
    _fields_ = [("B", c_int), ("H", c_int), ("Lq", c_int), ("Lk", c_int), ("hd", c_int),
                ("q", c_vp), ("ldq", c_ll), ("k", c_vp), ("ldk", c_ll), ("v", c_vp), ("ldv", c_ll),
                ("probs", c_vp), ("dout", c_vp), ("ldo", c_ll),
                ("dq", c_vp), ("lddq", c_ll), ("dk", c_vp), ("lddk", c_ll), ("dv", c_vp), ("lddv", c_ll),
                ("dbias", c_vp), ("scale", c_f), ("drop_p", c_f), ("sid", c_u32), ("rng", c_vp),
                ("stats", c_vp), ("bias", c_vp), ("key_mask", c_vp)]


# every symbol include/vqa_b200.h declares, with its argument types (tests check the library exports all)
_P = c_vp
SIGNATURES = {
    "vqa_last_error": (ctypes.c_char_p, []),
    "vqa_version": (c_int, []),
    "vqa_debug_set_umma": (c_int, [c_int, c_int, c_int, c_int]),
    "vqa_debug_gemm_timing": (c_int, [_P]),
    "vqa_gemm_ksplit_workspace": (c_ll, [c_int, c_int, c_int, c_int]),
    "vqa_plan_create": (c_vp, []),
    "vqa_plan_destroy": (c_int, [_P]),
    "vqa_plan_size": (c_int, [_P]),
    "vqa_plan_run": (c_int, [_P, _P]),
    "vqa_plan_capture_graph": (c_int, [_P, _P]),
    "vqa_plan_set_lane": (c_int, [_P, c_int]),
    "vqa_plan_fork": (c_int, [_P]),
    "vqa_plan_join": (c_int, [_P]),
    "vqa_plan_mark": (c_int, [_P]),
    "vqa_plan_wait": (c_int, [_P, c_int]),
    "vqa_plan_profile": (c_int, [_P, _P, _P, c_int]),
    "vqa_plan_time_ops": (c_int, [_P, _P, ctypes.c_char_p, c_int, ctypes.POINTER(c_f), ctypes.POINTER(c_d), ctypes.POINTER(c_int)]),
    "vqa_plan_op_info": (c_int, [_P, c_int, ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(c_d),
                                 ctypes.POINTER(c_d)]),
    "vqa_gemm_bf16": (c_int, [_P, ctypes.POINTER(GemmArgs), _P]),
    "vqa_conv2d_bf16": (c_int, [_P, ctypes.POINTER(ConvArgs), _P]),
    "vqa_conv2d_wgrad_bf16": (c_int, [_P, ctypes.POINTER(ConvWgradArgs), _P]),
    "vqa_cast_f32_bf16": (c_int, [_P, _P, _P, c_ll, _P]),
    "vqa_cast_bf16_f32": (c_int, [_P, _P, _P, c_ll, _P]),
    "vqa_split_lo_bf16": (c_int, [_P, _P, _P, c_ll, _P]),
    "vqa_memset_zero": (c_int, [_P, _P, c_ll, _P]),
    "vqa_memcpy_d2d": (c_int, [_P, _P, _P, c_ll, _P]),
    "vqa_axpy_f32": (c_int, [_P, _P, _P, c_f, c_ll, _P]),
    "vqa_fold_conv_bn": (c_int, [_P, _P, _P, _P, _P, _P, c_f, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P]),
    "vqa_convT_weight_prep": (c_int, [_P, _P, _P, c_int, c_int, _P]),
    "vqa_convT_weight_prep_bf16": (c_int, [_P, _P, _P, c_int, c_int, _P]),
    "vqa_convT_wgrad_unprep": (c_int, [_P, _P, _P, c_int, c_int, _P]),
    "vqa_image_to_stem": (c_int, [_P, _P, _P, c_int, c_int, c_int, _P]),
    "vqa_image_u8_to_stem": (c_int, [_P, _P, _P, c_int, c_int, c_int, _P]),
    "vqa_upsample_nearest_nhwc": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P]),
    "vqa_nhwc_to_nchw_f32": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, _P]),
    "vqa_maxpool3x3s2": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, _P]),
    "vqa_embedding_fwd": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, c_f, c_u32, _P, _P]),
    "vqa_embedding_bwd": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, c_f, c_u32, _P, _P]),
    "vqa_embedding_bwd_rows": (c_int, [_P, _P, _P, c_int, c_int, c_f, c_u32, _P, c_f, _P]),
    "vqa_embedding_scatter_ordered": (c_int, [_P, _P, _P, _P, _P, _P, c_int, c_int, c_int, _P]),
    "vqa_rmsnorm_fwd": (c_int, [_P, _P, _P, _P, _P, _P, c_int, c_int, c_f, c_f, c_u32, _P, _P]),
    "vqa_rmsnorm_fwd_split": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_f, _P]),
    "vqa_rmsnorm_bwd": (c_int, [_P, _P, c_int, _P, _P, _P, _P, _P, _P, c_int, c_int, c_f, c_u32, _P, _P, c_f, c_u32,
                                _P]),
    "vqa_t5_bias_build": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, _P]),
    "vqa_t5_bias_grad": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, _P]),
    "vqa_attention_fwd": (c_int, [_P, ctypes.POINTER(AttnFwdArgs), _P]),
    "vqa_attention_bwd": (c_int, [_P, ctypes.POINTER(AttnBwdArgs), _P]),
    "vqa_debug_attn_timing": (c_int, [_P]),
    "vqa_layernorm_fwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_f, _P]),
    "vqa_layernorm_bwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, c_int, c_int, _P, c_f, c_u32, _P, _P, _P]),
    "vqa_dropout_cast": (c_int, [_P, _P, _P, c_ll, c_int, c_f, c_u32, _P, _P]),
    "vqa_colsum_bf16": (c_int, [_P, _P, c_ll, _P, c_int, c_int, _P]),
    "vqa_pooler_fwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, _P]),
    "vqa_pooler_bwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, _P]),
    "vqa_logsoftmax_nll_fwd": (c_int, [_P, _P, c_ll, _P, _P, _P, c_int, c_int, _P]),
    "vqa_logsoftmax_nll_bwd": (c_int, [_P, _P, _P, _P, _P, _P, c_ll, c_int, c_int, _P]),
    "vqa_attention_long_fwd": (c_int, [_P, _P, c_ll, _P, c_ll, _P, c_ll, _P, c_ll, c_int, c_int, c_int, c_int, c_f, _P, _P]),
    "vqa_vit_patchify": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, _P]),
    "vqa_vit_assemble": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, _P]),
    "vqa_gelu_bf16": (c_int, [_P, _P, c_ll, _P]),
    "vqa_vit_fuse_concat": (c_int, [_P, _P, _P, c_int, _P, _P, c_int, c_int, _P]),
    "vqa_xattn1_fwd": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, c_f, c_u32, _P, _P]),
    "vqa_xattn1_bwd": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, c_f, c_u32, _P, _P]),
    "vqa_gather_rows": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, _P]),
    "vqa_scatter_rows": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, _P]),
    "vqa_t5_bias_causal": (c_int, [_P, _P, c_int, c_int, _P]),
    "vqa_relu_dropout_bwd": (c_int, [_P, _P, _P, _P, c_f, c_ll, _P]),
    "vqa_sumsq_f32": (c_int, [_P, _P, c_ll, _P, _P]),
    "vqa_clip_scale_f32": (c_int, [_P, _P, c_ll, _P, c_f, _P]),
    "vqa_adamw_amsgrad": (c_int, [_P, _P, _P, _P, _P, _P, _P, c_ll, c_d, c_d, c_d, c_d, c_d, c_d, c_d, _P, c_f,
                                  c_int, _P]),
    "vqa_rng_advance": (c_int, [_P, _P, _P]),
}
EXPORTS = list(SIGNATURES)

_lib = None


def load():
    """Return the loaded library; raises RuntimeError if it has not been built (there is no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "libvqa_b200.so is missing (%s): run `python -c 'import __graft_entry__ as g; g.build()'`; "
                "there is no CPU or torch fallback path" % LIB_PATH)
        import torch  # noqa: F401  (loads libcudart.so.12 the library links against)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc, what=""):
    if rc != 0:
        raise RuntimeError("libvqa_b200 %s failed (rc=%d): %s" % (what, rc, load().vqa_last_error().decode()))


def stream_ptr():
    import torch
    return c_vp(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    """Device pointer of a tensor (or None / int passthrough)."""
    if t is None:
        return None
    if isinstance(t, int):
        return t
    return t.data_ptr()
