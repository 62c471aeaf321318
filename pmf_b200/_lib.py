"""ctypes binding of libpmf_b200.so (C-ABI declared in include/pmfb.h).

The library is the product; there is no fallback.  If the shared object is missing, or a compute entry point is
called without an sm_100 device, a RuntimeError carrying ``pmfb_last_error()`` is raised.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpmf_b200.so")
ABI_VERSION = 6
MAX_TAPS = 9

ACT_NONE, ACT_RELU, ACT_LEAKY, ACT_SIGMOID = 0, 1, 2, 3
DT_F32, DT_F16, DT_BF16 = 0, 1, 2

f32p = C.c_void_p  # raw device pointers are passed as integers
i32, i64 = C.c_int32, C.c_int64


class View(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("sn", i64), ("sy", i64), ("sx", i64)]


class Epilogue(C.Structure):
    _fields_ = [("alpha1", C.c_void_p), ("beta1", C.c_void_p), ("alpha2", C.c_void_p), ("beta2", C.c_void_p),
                ("r1", View), ("mul", View), ("r2", View), ("act", i32), ("round_out", i32)]


class TmaSrc(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("dims", C.c_uint64 * 5), ("strides", C.c_uint64 * 4)]


class ConvDesc(C.Structure):
    _fields_ = [("x", TmaSrc), ("w", C.c_void_p), ("c_in", i32), ("c_out", i32), ("n_taps", i32),
                ("tap_dc", i32 * MAX_TAPS), ("tap_dw", i32 * MAX_TAPS), ("tap_dp", i32 * MAX_TAPS),
                ("tap_dh", i32 * MAX_TAPS), ("tap_wi", i32 * MAX_TAPS), ("use_tap_wi", i32), ("n_batch", i32),
                ("out_h", i32), ("out_w", i32), ("tile_w", i32),
                ("tile_h", i32), ("n_tile", i32), ("out", C.c_void_p), ("o_sn", i64), ("o_sy", i64), ("o_sx", i64),
                ("epi", Epilogue), ("bn_stats", C.c_void_p), ("dtype", i32), ("out_half", i32)]


class BnFuse(C.Structure):
    """pmfb_bn_fuse (include/pmfb.h): BatchNorm finalisation fused into the BN-apply pass."""
    _fields_ = [("sums", C.c_void_p), ("count", i64), ("gamma", C.c_void_p), ("beta", C.c_void_p), ("running_mean", C.c_void_p),
                ("running_var", C.c_void_p), ("momentum", C.c_float), ("eps", C.c_float), ("alpha_out", C.c_void_p),
                ("beta_out", C.c_void_p), ("mean_out", C.c_void_p), ("invstd_out", C.c_void_p)]


class WgradDesc(C.Structure):
    _fields_ = [("x", TmaSrc), ("dy", TmaSrc), ("c_in", i32), ("c_out", i32), ("n_taps", i32),
                ("tap_dc", i32 * MAX_TAPS), ("tap_dw", i32 * MAX_TAPS), ("tap_dp", i32 * MAX_TAPS),
                ("tap_dh", i32 * MAX_TAPS), ("n_batch", i32), ("out_h", i32), ("out_w", i32), ("ptile_w", i32),
                ("ptile_h", i32), ("n_tile", i32), ("ksplit", i32), ("dw", C.c_void_p), ("dtype", i32), ("reserved", i32)]


class WeightJob(C.Structure):
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("dst2", C.c_void_p), ("c_out", i32), ("c_in", i32), ("kh", i32),
                ("kw", i32), ("stem", i32), ("c_out_p", i32), ("c_in_p", i32), ("accumulate", i32), ("no_round", i32),
                ("reserved", i32), ("start", i64)]


VP = C.POINTER(View)
vp = C.c_void_p

_SIGNATURES = {
    "pmfb_abi_version": ([], C.c_int),
    "pmfb_sm_count": ([], C.c_int),
    "pmfb_last_error": ([], C.c_char_p),
    "pmfb_init": ([], C.c_int),
    "pmfb_conv_fwd": ([C.POINTER(ConvDesc), vp], C.c_int),
    "pmfb_conv_fused_stats_ok": ([C.POINTER(ConvDesc)], C.c_int),
    "pmfb_conv_wgrad": ([C.POINTER(WgradDesc), vp], C.c_int),
    "pmfb_conv16_ok": ([C.POINTER(ConvDesc)], C.c_int),
    "pmfb_wgrad16_ok": ([C.POINTER(WgradDesc)], C.c_int),
    "pmfb_pointwise16": ([VP, vp, i64, i64, i64, i32, i32, i32, i32, C.POINTER(Epilogue), vp, i32, vp, i32, vp], C.c_int),
    "pmfb_pointwise16_bn": ([VP, vp, i64, i64, i64, i32, i32, i32, i32, C.POINTER(Epilogue), vp, i32, vp, i32, C.POINTER(BnFuse), vp],
                            C.c_int),
    "pmfb_bn_bwd_reduce16": ([VP, VP, VP, i32, VP, vp, vp, vp, vp, i32, i32, i32, i32, vp, i32, vp], C.c_int),
    "pmfb_convert16": ([VP, i32, i32, i32, i32, vp, i64, i64, i64, i32, vp], C.c_int),
    "pmfb_bn_bwd_apply16": ([VP, VP, VP, i32, VP, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, vp, i64, i64, i64,
                             i32, vp, vp, vp, vp, i64, i64, i64, i32, vp, i32, vp], C.c_int),
    "pmfb_memset_zero": ([vp, C.c_size_t, vp], C.c_int),
    "pmfb_pack_input": ([vp, i64, i64, i64, i64, i32, i32, i32, i32, i32, vp, i32, i64, i32, vp], C.c_int),
    "pmfb_nhwc_to_nchw": ([VP, i32, i32, i32, i32, vp, vp], C.c_int),
    "pmfb_pack_weight": ([vp, i32, i32, i32, i32, i32, i32, i32, vp, vp, i32, vp], C.c_int),
    "pmfb_split_tf32": ([VP, i32, i32, i32, i32, vp, i64, i64, i64, i32, vp], C.c_int),
    "pmfb_unpack_wgrad": ([vp, i32, i32, i32, i32, i32, i32, i32, vp, i32, vp], C.c_int),
    "pmfb_weight_jobs": ([i32, vp, i32, i64, vp], C.c_int),
    "pmfb_pixel_mask": ([VP, i32, i32, i32, i32, vp, vp], C.c_int),
    "pmfb_mask_maxpool": ([vp, i32, i32, i32, i32, i32, i32, i32, vp, vp], C.c_int),
    "pmfb_pixel_scale": ([VP, i32, i32, i32, i32, vp, i32, vp, vp, VP, vp, vp, i64, i64, i64, i32, vp], C.c_int),
    "pmfb_pointwise": ([VP, vp, i64, i64, i64, i32, i32, i32, i32, C.POINTER(Epilogue), vp], C.c_int),
    "pmfb_bn_stats": ([VP, i32, i32, i32, i32, vp, vp], C.c_int),
    "pmfb_bn_finalize": ([vp, i64, i32, vp, vp, vp, vp, C.c_float, C.c_float, vp, vp, vp, vp, vp], C.c_int),
    "pmfb_bn_bwd_reduce": ([VP, VP, VP, i32, VP, vp, vp, vp, vp, i32, i32, i32, i32, vp, vp], C.c_int),
    "pmfb_bn_bwd_apply": ([VP, VP, VP, i32, VP, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, vp, i64, i64, i64,
                           i32, vp, vp, vp, vp, i64, i64, i64, i32, vp], C.c_int),
    "pmfb_colsum": ([VP, i32, i32, i32, i32, i32, vp, vp], C.c_int),
    "pmfb_d2f": ([vp, vp, i64, C.c_float, i32, i32, vp], C.c_int),
    "pmfb_pool3s2": ([i32, VP, i32, i32, i32, i32, vp, vp, i64, i64, i64, vp, i32, vp, vp, vp], C.c_int),
    "pmfb_pool3s2_bwd": ([i32, VP, i32, i32, i32, i32, vp, vp, i64, i64, i64, vp, i32, vp], C.c_int),
    "pmfb_pixel_shuffle": ([VP, i32, i32, i32, i32, vp, vp, i64, i64, i64, i32, vp, vp, vp], C.c_int),
    "pmfb_pixel_shuffle_bwd": ([VP, i32, i32, i32, i32, vp, vp, i64, i64, i64, i32, i32, vp], C.c_int),
    "pmfb_upsample2x": ([VP, i32, i32, i32, i32, vp, i64, i64, i64, i32, vp, vp, vp], C.c_int),
    "pmfb_upsample2x_bwd": ([VP, i32, i32, i32, i32, vp, i64, i64, i64, i32, vp], C.c_int),
    "pmfb_softmax_nchw": ([VP, i32, i32, i32, i32, vp, vp], C.c_int),
    "pmfb_softmax_nchw_bwd": ([vp, vp, i32, i32, i32, i32, vp, i64, i64, i64, i32, vp], C.c_int),
    "pmfb_loss_head": ([vp, vp, vp, i32, i32, i32, i32, vp, C.c_float, C.c_float, C.c_float, C.c_float, vp, vp, vp, vp], C.c_int),
    "pmfb_lovasz_workspace_bytes": ([i64, i32, i32], C.c_size_t),
    "pmfb_lovasz": ([vp, vp, vp, i32, i32, i32, i32, i32, C.c_float, vp, vp, vp, vp, C.c_size_t, vp], C.c_int),
    "pmfb_knn_vote": ([vp, vp, i32, i32, vp, vp, vp, i64, vp, i32, i32, C.c_float, i32, vp, vp], C.c_int),
    "pmfb_knn_vote_batched": ([vp, vp, i32, i32, i32, vp, vp, vp, vp, i64, vp, i32, i32, C.c_float, i32, vp, vp], C.c_int),
    "pmfb_argmax_nchw": ([vp, i32, i32, i32, i32, i32, i32, i32, i32, vp, vp, vp], C.c_int),
    "pmfb_lut_remap": ([vp, i64, vp, i32, vp, vp], C.c_int),
    "pmfb_merge_cameras": ([vp, vp, vp, vp, i64, i64, vp, vp, vp], C.c_int),
    "pmfb_confusion_add": ([vp, vp, i64, i32, vp, vp], C.c_int),
    "pmfb_project_scatter": ([vp, vp, i64, C.POINTER(C.c_double), i32, i32, vp, vp, vp, vp, vp, vp, vp, vp], C.c_int),
}

EXPORTS = tuple(_SIGNATURES.keys())

_lib = None
launches = 0  # number of C-ABI compute calls issued (each enqueues >= 1 kernel); read by bench.py


class PmfbError(RuntimeError):
    pass


def lib():
    """Load libpmf_b200.so (once).  Fails loudly when the extension has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PmfbError("libpmf_b200.so is not built (%s); run `python -c 'import __graft_entry__ as g; g.build()'` "
                            "or `make -C pmf_b200/csrc` — there is no CPU fallback" % LIB_PATH)
        l = C.CDLL(LIB_PATH)
        for name, (argtypes, restype) in _SIGNATURES.items():
            fn = getattr(l, name)
            fn.argtypes = argtypes
            fn.restype = restype
        if l.pmfb_abi_version() != ABI_VERSION:
            raise PmfbError("libpmf_b200.so ABI %d != binding ABI %d" % (l.pmfb_abi_version(), ABI_VERSION))
        _lib = l
        trace = os.environ.get("PMFB_TRACE_LOADS")  # acceptance tests: prove which native library a process loaded
        if trace:
            with open(trace, "a") as f:
                f.write("%d %s\n" % (os.getpid(), LIB_PATH))
    return _lib


def last_error():
    return lib().pmfb_last_error().decode("utf-8", "replace")


def check(rc, what):
    if rc != 0:
        msg = last_error()
        if what in ("pmfb_knn_vote", "pmfb_knn_vote_batched") and "odd number" in msg:
            raise ValueError(msg)  # knn.py:73-74 raises ValueError
        raise PmfbError("%s failed (%d): %s" % (what, rc, msg))


def call(name, *args):
    global launches
    launches += 1
    check(getattr(lib(), name)(*args), name)


def query(name, *args):
    """Entry points that answer a question (return value is the answer, not a status)."""
    return int(getattr(lib(), name)(*args))


# ---- precision mode of the tensor-core convolutions (process-wide; pmf_b200.precision(...) switches it)
#   "tf32"   : kind::tf32 operands (rounded where they are produced), one UMMA per K step — the arithmetic class of the
#              reference's own GPU path (cuDNN with allow_tf32); what every eval-mode forward runs
#   "3xtf32" : hi/lo operand split, three UMMAs per K step into the same TMEM accumulator (pmfb_split_tf32): fp32-class
#              results, ~3x the tensor time; the parity mode for train-mode (batch-statistics) BatchNorm
#   "f16"    : THE DEFAULT.  In training passes the stride-1 convolutions with >= 64 channels read 16-bit shadows of
#              their operands — fp16 activations and weights in the forward pass (the same 10-bit mantissa as tf32), bf16
#              output gradients / weights / activations in dgrad and wgrad — through kind::f16 UMMAs with fp32 accumulation:
#              K = 16 channels per instruction instead of 8 at the same operand bytes.  Everything else (BatchNorm,
#              elementwise, thin layers, storage, eval-mode forwards) stays fp32 / tf32
PRECISIONS = ("tf32", "3xtf32", "f16")
_precision = os.environ.get("PMFB_PRECISION", "f16").lower()
if _precision not in PRECISIONS:
    raise PmfbError("PMFB_PRECISION must be one of %s, got %r" % (PRECISIONS, _precision))


def get_precision():
    return _precision


def set_precision(mode):
    global _precision
    mode = str(mode).lower()
    if mode not in PRECISIONS:
        raise ValueError("precision must be one of %s, got %r" % (PRECISIONS, mode))
    prev, _precision = _precision, mode
    return prev


_inited = False


def require_device():
    """pmfb_init(): an sm_100 device must be present."""
    global _inited
    if not _inited:
        check(lib().pmfb_init(), "pmfb_init")
        _inited = True
