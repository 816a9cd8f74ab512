"""
ctypes binding of libevcplm.so (C ABI declared in include/evcplm.h).

There is NO fallback: if the CUDA library is missing or cannot be loaded the
import of the engine fails loudly (EngineUnavailableError).  Build it in-tree
with ``python -c "import __graft_entry__ as g; g.build()"`` or
``evcouplings_b200/csrc/build.sh``.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libevcplm.so")
ABI_VERSION = 1


class EngineUnavailableError(RuntimeError):
    """libevcplm.so (the sm_100a CUDA engine) is missing / unloadable / has no device."""


class EngineError(RuntimeError):
    """A libevcplm call returned non-zero."""


c_void_p = ctypes.c_void_p
c_i32 = ctypes.c_int32
c_i64 = ctypes.c_int64
c_f32 = ctypes.c_float

# name -> (restype, argtypes); mirrors include/evcplm.h one to one
PROTOTYPES = {
    "evc_abi_version": (ctypes.c_int, []),
    "evc_last_error": (ctypes.c_char_p, []),
    "evc_device_count": (ctypes.c_int, []),
    "evc_device_info": (ctypes.c_int, [c_i32, c_void_p, c_void_p, c_void_p, c_void_p]),
    "evc_hamming_counts": (ctypes.c_int, [c_void_p, c_i64, c_i32, c_i32, c_i32, c_void_p]),
    "evc_hamming_plane_words": (c_i64, [c_i64, c_i32]),
    "evc_hamming_num_tiles": (c_i64, [c_i64]),
    "evc_hamming_pack": (ctypes.c_int, [c_void_p, c_i64, c_i32, c_void_p, c_void_p]),
    "evc_hamming_count_tiles": (ctypes.c_int, [c_void_p, c_i64, c_i32, c_i32, c_i64, c_i64, c_void_p, c_void_p]),
    "evc_plm_create": (ctypes.c_int, [c_void_p, c_void_p, c_i64, c_i32, c_i32, c_i32, c_void_p, c_i32]),
    "evc_plm_destroy": (None, [c_void_p]),
    "evc_plm_num_params": (c_i64, [c_void_p]),
    "evc_plm_eval_data": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "evc_plm_set_backward": (ctypes.c_int, [c_void_p, c_i32]),
    "evc_plm_set_forward": (ctypes.c_int, [c_void_p, c_i32]),
    "evc_plm_set_profiling": (ctypes.c_int, [c_void_p, c_i32]),
    "evc_plm_last_stage_ms": (ctypes.c_int, [c_void_p, c_void_p]),
    "evc_plm_add_regulariser": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_f32, c_f32, c_void_p]),
    "evc_plm_eval_host": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_f32, c_f32]),
    "evc_plm_weighted_counts": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "evc_vec_dot": (ctypes.c_int, [c_void_p, c_void_p, c_i64, c_void_p, c_void_p]),
    "evc_vec_axpby": (ctypes.c_int, [c_void_p, c_void_p, c_f32, c_f32, c_i64, c_void_p]),
    "evc_vec_copy": (ctypes.c_int, [c_void_p, c_void_p, c_i64, c_void_p]),
    "evc_vec_sub": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_i64, c_void_p]),
    "evc_lbfgs_direction": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                           c_i64, c_i32, c_i32, c_i32, c_void_p]),
    "evc_lbfgs_update_pair": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                             c_void_p, c_void_p, c_i64, c_void_p]),
    "evc_ec_scores": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_i32, c_i32, c_void_p, c_void_p, c_void_p, c_void_p]),
    "evc_plm_energies": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "evc_fn_scores": (ctypes.c_int, [c_void_p, c_i32, c_i32, c_void_p, c_void_p]),
}

_lib = None


def load():
    """Load libevcplm.so and bind every symbol of the C ABI (no device needed)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EngineUnavailableError(
            "CUDA engine library not built: %s is missing. There is no CPU fallback; "
            "build it with evcouplings_b200/csrc/build.sh (nvcc, sm_100a)." % LIB_PATH)
    try:
        lib = ctypes.CDLL(LIB_PATH)
    except OSError as e:
        raise EngineUnavailableError("cannot load %s: %s" % (LIB_PATH, e))
    for name, (restype, argtypes) in PROTOTYPES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError:
            raise EngineUnavailableError("%s does not export %s" % (LIB_PATH, name))
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.evc_abi_version() != ABI_VERSION:
        raise EngineUnavailableError("libevcplm ABI version mismatch")
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().evc_last_error()
        raise EngineError("%s failed: %s" % (what or "libevcplm call", msg.decode() if msg else "unknown error"))


def require_device():
    """Raise unless at least one CUDA device is usable through the library."""
    lib = load()
    n = lib.evc_device_count()
    if n <= 0:
        raise EngineUnavailableError(
            "libevcplm found no CUDA device; the PLM engine has no CPU fallback")
    return n
