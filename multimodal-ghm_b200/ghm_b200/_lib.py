"""ctypes binding of libghm_b200.so (the C ABI declared in include/ghm_b200.h).

There is no fallback: if the shared library is missing or a call fails, a
RuntimeError is raised.  Build it with ``python multimodal-ghm_b200/build.py``.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GHM_LIB_PATH") or os.path.join(_HERE, "libghm_b200.so")   # (override: A/B builds, development only)

_lib = None

c_i64, c_u64, c_int, c_vp, c_f = C.c_int64, C.c_uint64, C.c_int, C.c_void_p, C.c_float

# name -> (restype, argtypes); mirrors include/ghm_b200.h one-to-one
SIGNATURES = {
    "ghm_last_error": (C.c_char_p, []),
    "ghm_version": (C.c_char_p, []),
    "ghm_device_count": (c_int, []),
    "ghm_model_create": (c_int, [C.POINTER(c_vp), c_int, c_int, c_int, c_int, c_vp, c_vp, c_int]),
    "ghm_model_destroy": (c_int, [c_vp]),
    "ghm_model_update": (c_int, [c_vp, c_vp, c_vp, c_vp]),
    "ghm_model_table_bytes": (c_i64, [c_vp]),
    "ghm_model_info": (c_int, [c_vp] + [C.POINTER(c_int)] * 4 + [C.POINTER(c_i64)] * 2),
    "ghm_model_status": (c_int, [c_vp, c_vp, C.POINTER(c_int)]),
    "ghm_sample": (c_int, [c_vp, c_i64, c_int, c_vp, c_vp, c_u64, c_u64, c_vp, c_vp, c_int, c_vp, c_vp, c_vp]),
    "ghm_sample_mixed": (c_int, [c_vp, c_i64, c_i64, c_vp, c_u64, c_u64, c_vp, c_vp, c_int, c_vp, c_vp, c_vp]),
    "ghm_model_set_gemm_mode": (c_int, [c_vp, c_int]),
    "ghm_bp_cls_workspace_bytes": (c_i64, [c_vp, c_i64]),
    "ghm_sample_paired": (c_int, [c_vp, c_i64, c_i64, c_u64, c_u64, c_u64, c_vp, c_vp, c_int, c_vp, c_vp, c_vp]),
    "ghm_clip_bayes": (c_int, [c_vp, c_vp, c_i64, c_int, c_i64, c_i64, c_u64, c_u64, c_vp, c_vp, c_vp, c_int, c_vp, c_vp, c_vp,
                               c_vp, c_vp]),
    "ghm_sample_blocked": (c_int, [c_vp, c_i64, c_i64, c_i64, c_int, c_i64, c_vp, c_u64, c_u64, c_u64, c_vp, c_vp, c_int, c_vp,
                                   c_vp, c_vp]),
    "ghm_bp_cls": (c_int, [c_vp, c_i64, c_vp, c_int, c_vp, c_vp, c_vp, c_vp]),
    "ghm_bp_dns_workspace_bytes": (c_i64, [c_vp, c_i64]),
    "ghm_bp_dns": (c_int, [c_vp, c_i64, c_vp, c_f, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "ghm_bp_nwp_workspace_bytes": (c_i64, [c_vp, c_i64]),
    "ghm_bp_nwp": (c_int, [c_vp, c_i64, c_vp, c_int, c_vp, c_vp, c_vp, c_vp]),
    "ghm_guides_cls": (c_int, [c_vp, c_i64, c_vp, c_int, C.POINTER(c_vp), c_vp, c_vp, c_vp]),
    "ghm_guides_dns_workspace_bytes": (c_i64, [c_vp, c_i64]),
    "ghm_guides_dns": (c_int, [c_vp, c_i64, c_vp, c_f, c_vp, C.POINTER(c_vp), c_vp, c_vp, c_vp]),
    "ghm_guides_nwp_workspace_bytes": (c_i64, [c_vp, c_i64]),
    "ghm_guides_nwp": (c_int, [c_vp, c_i64, c_vp, c_int, c_vp, C.POINTER(c_vp), c_vp, c_vp, c_vp]),
    "ghm_risk_clip": (c_int, [c_vp, c_vp, c_i64, c_int, c_int, c_i64, c_i64, c_vp, c_vp]),
    "ghm_risk_cdm": (c_int, [c_vp, c_vp, c_int, c_i64, c_i64, c_vp, c_vp]),
    "ghm_risk_ce": (c_int, [c_vp, c_vp, c_int, c_i64, c_int, c_i64, c_i64, c_i64, c_vp, c_vp]),
    "ghm_risk_zsc": (c_int, [c_vp, c_i64, c_vp, c_vp, c_int, c_vp, c_vp]),
    "ghm_gauss_noise": (c_int, [c_vp, c_i64, c_vp, c_int, c_f, c_u64, c_u64, c_vp, c_vp]),
    "ghm_host_clip_bayes": (c_int, [c_vp, c_vp, c_i64, c_int, c_u64, c_u64, c_vp, c_vp, c_vp, c_int, c_vp, c_vp]),
}


def get_lib():
    """Load libghm_b200.so (once) and attach the prototypes.  Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "libghm_b200.so not found at %s -- build it with `python multimodal-ghm_b200/build.py` "
            "(there is no CPU/PyTorch fallback for this path)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is missing: fail loudly
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(code):
    if code != 0:
        msg = get_lib().ghm_last_error()
        raise RuntimeError("libghm_b200 error %d: %s" % (code, msg.decode() if msg else "?"))
