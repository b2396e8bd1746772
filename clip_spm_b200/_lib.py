"""ctypes binding of include/clipspm_b200.h.  The library is required: nothing in this package falls back to
PyTorch or CPU code when it is missing -- loading fails loudly instead."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# SPM_LIB overrides the library file (A/B comparisons of two builds on the same GPU box)
LIB_PATH = os.environ.get("SPM_LIB") or os.path.join(_HERE, "lib", "libclipspm_b200.so")

c_void_p = ctypes.c_void_p
c_int = ctypes.c_int
c_float = ctypes.c_float
c_ll = ctypes.c_longlong


class SpmConfig(ctypes.Structure):
    """Mirror of `spm_config` (include/clipspm_b200.h)."""
    _fields_ = [
        ("backbone", c_int),
        ("seq_len", c_int),
        ("n_text_classes", c_int),
        ("mid_dim_text", c_float),
        ("mid_dim_vision", c_float),
        ("negative_slope", c_float),
        ("alpha", c_float),
        ("single_direct", c_int),
        ("precision", c_int),
        ("max_episodes", c_int),
        ("max_support", c_int),
        ("max_query", c_int),
        ("max_way", c_int),
        ("head", c_int),
        ("cls_value", c_float),
        ("motion_residual_ratio", c_float),
        ("lambdas", c_float * 4),
        ("motion_coeff", c_float),
        ("normal_coeff", c_float),
        ("use_classification", c_int),
        ("fsar_depth", c_int),
        ("fsar_merge_before", c_int),
    ]


# name -> (restype, argtypes); every symbol declared in include/clipspm_b200.h
SIGNATURES = {
    "spm_last_error": (ctypes.c_char_p, []),
    "spm_abi_version": (c_int, []),
    "spm_launch_count": (c_ll, []),
    "spm_profile_begin": (c_int, [c_int]),
    "spm_profile_disarm": (c_int, []),
    "spm_profile_end": (c_int, [ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double),
                                ctypes.POINTER(c_int)]),
    "spm_create": (c_int, [ctypes.POINTER(SpmConfig), ctypes.POINTER(c_void_p)]),
    "spm_destroy": (c_int, [c_void_p]),
    "spm_load_weights": (c_int, [c_void_p, c_void_p, c_int, ctypes.POINTER(ctypes.c_char_p),
                                 ctypes.POINTER(c_void_p), ctypes.POINTER(ctypes.c_int64)]),
    "spm_set_text_features": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int]),
    "spm_set_text_features_train": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int]),
    "spm_class_logits": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "spm_cpm2c_outputs": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "spm_encode_frames": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "spm_head": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int] + [c_void_p] * 7),
    "spm_head_stage": (c_int, [c_void_p, c_void_p, ctypes.c_char_p, c_void_p, c_ll, ctypes.POINTER(c_ll)]),
    "spm_forward": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int] + [c_void_p] * 7),
    "spm_eval": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int] + [c_void_p] * 6 + [c_float] + [c_void_p] * 5),
    "spm_eval_host": (c_int, [c_void_p, c_int, c_int, c_int, c_int] + [c_void_p] * 6 + [c_float] + [c_void_p] * 5),
    "spm_eval_host_set_next": (c_int, [c_void_p, c_void_p, c_void_p, c_int]),
    "spm_otam_distance": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int,
                                  c_float, c_float, c_void_p]),
    "spm_otam_distance_backward": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int,
                                           c_float, c_void_p, c_void_p, c_void_p]),
    "spm_tv1_create": (c_int, [c_int, c_int, c_int, c_int, c_int, ctypes.POINTER(c_void_p)]),
    "spm_tv1_destroy": (c_int, [c_void_p]),
    "spm_tv1_load_weights": (c_int, [c_void_p, c_void_p] + [c_void_p] * 11),
    "spm_tv1_forward": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "spm_tv1_set_dropout": (c_int, [c_void_p, c_float, c_float, ctypes.c_ulonglong]),
    "spm_dropout_seed_source": (c_int, [c_void_p]),
    "spm_dropout": (c_int, [c_void_p, c_void_p, c_ll, c_float, ctypes.c_ulonglong, ctypes.c_uint, c_void_p]),
    "spm_tv1_backward": (c_int, [c_void_p, c_void_p] + [c_void_p] * 13),
    "spm_vitblock_create": (c_int, [c_int, ctypes.POINTER(c_void_p)]),
    "spm_vitblock_load_weights": (c_int, [c_void_p, c_void_p] + [c_void_p] * 12),
    "spm_vitblock_forward": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "spm_vitblock_backward": (c_int, [c_void_p, c_void_p] + [c_void_p] * 14),
    "spm_layernorm_forward": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "spm_layernorm_backward": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int] + [c_void_p] * 4),
    "spm_linear_backward_workspace": (c_ll, [c_int, c_int, c_int]),
    "spm_linear_backward": (c_int, [c_void_p, c_int] + [c_void_p] * 5 + [c_int, c_int, c_int, c_int, c_float]
                            + [c_void_p] * 4 + [c_ll]),
    "spm_adam_create": (c_int, [c_int, ctypes.POINTER(c_void_p), ctypes.POINTER(ctypes.c_longlong), ctypes.POINTER(c_void_p)]),
    "spm_adam_destroy": (c_int, [c_void_p]),
    "spm_adam_step": (c_int, [c_void_p, c_void_p, ctypes.POINTER(c_void_p), ctypes.c_double, ctypes.c_double,
                              ctypes.c_double, ctypes.c_double, ctypes.c_double, c_void_p]),
    "spm_sgd_step": (c_int, [c_void_p, c_void_p, ctypes.POINTER(c_void_p), ctypes.c_double, ctypes.c_double, ctypes.c_double,
                             c_void_p]),
    "spm_adam_state": (c_int, [c_void_p, c_int, ctypes.POINTER(c_void_p), ctypes.POINTER(c_void_p), ctypes.POINTER(c_void_p)]),
    "spm_scaler_update": (c_int, [c_void_p, c_void_p, c_float, c_float, c_int]),
    "spm_softdtw_forward": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_float, c_float, c_void_p, c_void_p]),
    "spm_softdtw_backward": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_float, c_float, c_void_p]),
    "spm_vit_attention": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int]),
    "spm_gemm": (c_int, [c_void_p, c_int, c_void_p, c_ll, c_void_p, c_ll, c_int, c_int, c_int, c_void_p, c_int,
                         c_float, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int]),
    "spm_text_create": (c_int, [c_int, c_int, ctypes.POINTER(c_void_p)]),
    "spm_text_destroy": (c_int, [c_void_p]),
    "spm_text_load_weights": (c_int, [c_void_p, c_void_p, c_int, ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(c_void_p),
                                      ctypes.POINTER(ctypes.c_int64)]),
    "spm_text_encode": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "spm_text_class_features": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "spm_transform_frames": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "spm_transform_frames_train": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "spm_frame_geometry": (c_int, [c_int, c_int] + [ctypes.POINTER(c_int)] * 4),
    "spm_encode_frames_u8": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "spm_eval_u8": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int] + [c_void_p] * 6 + [c_float]
                    + [c_void_p] * 5),
    "spm_jpeg_info": (c_int, [c_void_p, c_ll] + [ctypes.POINTER(c_int)] * 4),
    "spm_jpeg_decode": (c_int, [c_void_p, c_int, ctypes.POINTER(c_void_p), ctypes.POINTER(ctypes.c_int64), c_int, c_int,
                                c_void_p]),
    "spm_eval_host_u8": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int] + [c_void_p] * 6 + [c_float]
                         + [c_void_p] * 5),
}

_lib = None


def load():
    """dlopen the in-tree library (built by __graft_entry__.build()) and bind every declared symbol."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "clip_spm_b200: %s is missing -- run `python __graft_entry__.py` (nvcc, sm_100a) first; "
            "there is no CPU or PyTorch fallback for this path" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status):
    if status != 0:
        raise RuntimeError("clipspm_b200: " + load().spm_last_error().decode("utf-8", "replace"))
