"""ctypes binding of libsake_b200.so (the C ABI declared in include/sake_b200.h)."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# SAKE_B200_LIB selects another build of the same library (e.g. the development build with the mbarrier
# watchdog: make -C sake_b200/csrc WATCHDOG=1 -> libsake_b200_wd.so); the default is the product build.
LIB_PATH = os.environ.get("SAKE_B200_LIB") or os.path.join(_HERE, "libsake_b200.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "(or `make -C sake_b200/csrc`). sake_b200 has no CPU fallback.")

lib = C.CDLL(LIB_PATH)

# enums (include/sake_b200.h)
SAKE_UPDATE, SAKE_HAS_V, SAKE_HAS_MASK, SAKE_NO_SPATIAL, SAKE_DEFER_DW, SAKE_COSINE_CUTOFF = 1, 2, 4, 8, 16, 32
SAKE_DEFER_REDUCE = 128
SAKE_WEIGHTS_PREPARED = 64
ENGINE_AUTO, ENGINE_FP32, ENGINE_TF32X3, ENGINE_BF16, ENGINE_F16X2 = 0, 1, 2, 3, 4
ENGINES = {"auto": ENGINE_AUTO, "fp32": ENGINE_FP32, "tf32x3": ENGINE_TF32X3, "bf16": ENGINE_BF16,
           "f16x2": ENGINE_F16X2}
ENGINE_NAMES = {v: k for k, v in ENGINES.items()}


class SakeDims(C.Structure):
    _fields_ = ([(n, C.c_int32) for n in ("B", "N", "H", "A", "K", "flags", "engine", "reserved")] +
                [("cutoff_lower", C.c_float), ("cutoff_upper", C.c_float)])


PARAM_FIELDS = (
    "rbf_means", "rbf_betas", "mlp_in_kernel", "mlp_in_bias", "mlp_out0_kernel", "mlp_out0_bias",
    "mlp_out2_kernel", "mlp_out2_bias", "sem_kernel", "sem_bias", "x_mixing_kernel", "post0_kernel",
    "post0_bias", "post2_kernel", "post2_bias", "node0_kernel", "node0_bias", "node2_kernel",
    "node2_bias", "v_mixing_kernel", "vel0_kernel", "vel0_bias", "vel2_kernel",
)


class SakeLayerParams(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in PARAM_FIELDS]


class SakeLayerGrads(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in PARAM_FIELDS]


class SakePairTerms(C.Structure):
    """`he` edge features as per-pair additive terms (include/sake_b200.h)."""
    _fields_ = [(n, C.c_void_p) for n in ("u", "p", "g_u", "g_p")]


_vp, _sz, _i32, _i64 = C.c_void_p, C.c_size_t, C.c_int32, C.c_int64
_DP = C.POINTER(SakeDims)

lib.sake_version.restype = C.c_char_p
lib.sake_last_error.restype = C.c_char_p
lib.sake_resolve_engine.argtypes = [_DP]
lib.sake_resolve_engine.restype = C.c_int
lib.sake_layer_saved_bytes.argtypes = [_DP]
lib.sake_layer_saved_bytes.restype = _sz
lib.sake_layer_scratch_bytes.argtypes = [_DP, C.c_int, C.c_int]
lib.sake_layer_scratch_bytes.restype = _sz
lib.sake_ragged_bytes.argtypes = [_i32, _i32]
lib.sake_ragged_bytes.restype = _sz
lib.sake_ragged_prepare.argtypes = [_i32, _i32, _vp, _vp, _sz, _vp]
lib.sake_ragged_prepare.restype = C.c_int
lib.sake_ragged_gather.argtypes = [_vp, _i32, _i32, _i32, _vp, _vp, _vp]
lib.sake_ragged_gather.restype = C.c_int
lib.sake_ragged_scatter.argtypes = [_vp, _i32, _i32, _i32, C.c_float, _vp, _vp, _vp]
lib.sake_ragged_scatter.restype = C.c_int
lib.sake_layer_fwd.argtypes = [_DP, C.POINTER(SakeLayerParams), _vp, _vp, _vp, _vp, _vp, C.POINTER(SakePairTerms),
                               _vp, _vp, _vp, _vp, _sz, _vp, _sz, _vp]
lib.sake_layer_fwd.restype = C.c_int
lib.sake_layer_prepare.argtypes = [_DP, C.POINTER(SakeLayerParams), _vp, _sz, _vp]
lib.sake_layer_prepare.restype = C.c_int
lib.sake_layer_bwd.argtypes = [_DP, C.POINTER(SakeLayerParams), _vp, _vp, _vp, _vp, _vp, C.POINTER(SakePairTerms),
                               _vp, _sz,
                               _vp, _vp, _vp, _vp, _vp, _vp, C.POINTER(SakeLayerGrads), _vp, _sz, _vp]
lib.sake_layer_bwd.restype = C.c_int
lib.sake_dw_sync.argtypes = [_vp]
lib.sake_dw_sync.restype = C.c_int
lib.sake_dense_fwd.argtypes = [_i64, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp]
lib.sake_dense_fwd.restype = C.c_int
lib.sake_dense_bwd.argtypes = [_i64, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]
lib.sake_dense_bwd.restype = C.c_int
lib.sake_energy_head.argtypes = [_i32, _i32, _i32, _i32, _vp, _vp, _vp, C.c_float, C.c_float, _vp, _vp, _vp, _vp, _vp]
lib.sake_energy_head.restype = C.c_int
lib.sake_flow_pre.argtypes = [_i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp]
lib.sake_flow_pre.restype = C.c_int
lib.sake_flow_post.argtypes = [_i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]
lib.sake_flow_post.restype = C.c_int
lib.sake_flow_logprob.argtypes = [_i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp]
lib.sake_flow_logprob.restype = C.c_int
lib.sake_adam_step.argtypes = [_i64, _vp, _vp, _vp, _vp, _i32] + [C.c_float] * 7 + [_vp]
lib.sake_adam_step.restype = C.c_int
lib.sake_profile_begin.argtypes = [_i32]
lib.sake_profile_begin.restype = C.c_int
lib.sake_profile_collect.argtypes = [C.POINTER(C.c_float), C.POINTER(_i32), C.POINTER(_i64), _i32]
lib.sake_profile_collect.restype = C.c_int
lib.sake_selftest_xtg.argtypes = [_i32, _i64, _i32, _i32, _vp, _vp, _vp, _vp]
lib.sake_selftest_xtg.restype = C.c_int
lib.sake_debug_counters.argtypes = [C.POINTER(C.c_ulonglong)]
lib.sake_debug_counters.restype = C.c_int
lib.sake_debug_counters_bwd.argtypes = [C.POINTER(C.c_ulonglong)]
lib.sake_debug_counters_bwd.restype = C.c_int
lib.sake_launch_count.restype = C.c_ulonglong
lib.sake_selftest_tcgen05.argtypes = [C.POINTER(C.c_float), _vp]
lib.sake_selftest_tcgen05.restype = C.c_int

EXPORTS = ("sake_layer_prepare", "sake_dw_sync", "sake_flow_pre", "sake_flow_post", "sake_flow_logprob", "sake_ragged_bytes", "sake_ragged_prepare", "sake_ragged_gather", "sake_ragged_scatter", "sake_version", "sake_last_error", "sake_resolve_engine", "sake_layer_saved_bytes",
           "sake_layer_scratch_bytes", "sake_layer_fwd", "sake_layer_bwd", "sake_dense_fwd",
           "sake_dense_bwd", "sake_selftest_tcgen05", "sake_energy_head", "sake_adam_step",
           "sake_profile_begin", "sake_profile_collect", "sake_launch_count", "sake_selftest_xtg", "sake_debug_counters", "sake_debug_counters_bwd")


class SakeError(RuntimeError):
    pass


def check(rc, what):
    if rc != 0:
        raise SakeError(f"{what} failed (code {rc}): {lib.sake_last_error().decode()}")


def version():
    return lib.sake_version().decode()
