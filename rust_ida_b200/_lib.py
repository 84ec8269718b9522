"""ctypes loader for libb200ida.so (the C ABI declared in include/b200ida.h).

There is no CPU fallback: if the library is missing, or a compute call is made without a CUDA device,
this raises.  `build()` compiles the library in-tree with nvcc for sm_100a.
"""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
SO_PATH = os.environ.get("B200IDA_LIB") or os.path.join(CSRC, "libb200ida.so")  # env override: kernel experiments only

MEM_DEVICE, MEM_HOST = 0, 1
LAYOUT_SOA, LAYOUT_AOS = 0, 1
LU_DMMA = 0x100  # OR into the layout of b200ida_getrf_batched: tensor-core trailing update (tolerance mode)
LU_TEAM_PANEL = 0x200  # OR into the layout: the CTA-team kernel's shared panel + work-queue update (bit-identical)
OK, ERR_ARG, ERR_CUDA, ERR_NO_DEVICE, ERR_UNSUPPORTED = 0, -1, -2, -3, -4

# every symbol include/b200ida.h declares (tests check that the .so exports exactly these)
SYMBOLS = (
    "b200ida_last_error", "b200ida_version", "b200ida_device_count",
    "b200ida_getrf_batched", "b200ida_getrs_batched", "b200ida_wrms_batched", "b200ida_wrms_masked_batched", "b200ida_newton_batched",
    "b200ida_create", "b200ida_destroy", "b200ida_reinit", "b200ida_set_kernel", "b200ida_set_default_reorder", "b200ida_solve", "b200ida_solve_ladder", "b200ida_set_trace", "b200ida_set_roots", "b200ida_set_id", "b200ida_set_regroup",
    "b200ida_get_state", "b200ida_get_counters", "b200ida_get_step_info", "b200ida_get_dky", "b200ida_sync",
    "b200ida_fp64_issue_rate", "b200ida_detpow_batched", "b200ida_fastdiv_batched", "b200ida_kernel_launches", "b200ida_get_fallback_count",
)


class B200IdaError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("b200ida error %d: %s" % (code, msg))
        self.code = code


def build(force=False):
    """Compile csrc/*.cu into libb200ida.so (nvcc cross-compiles for sm_100a without a GPU)."""
    args = ["make", "-C", CSRC, "-s"]
    if force:
        args.append("-B")
    subprocess.check_call(args)
    return SO_PATH


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise ImportError(
            "%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` or `make -C %s`. "
            "There is no CPU fallback." % (SO_PATH, CSRC))
    L = C.CDLL(SO_PATH)
    vp, i32, i64, dbl = C.c_void_p, C.c_int, C.c_int64, C.c_double
    L.b200ida_last_error.restype = C.c_char_p
    L.b200ida_version.restype = i32
    L.b200ida_device_count.restype = i32
    L.b200ida_getrf_batched.argtypes = [vp, vp, vp, i32, i64, i32, i32, vp]
    L.b200ida_getrs_batched.argtypes = [vp, vp, vp, i32, i64, i32, i32, vp]
    L.b200ida_wrms_batched.argtypes = [vp, vp, vp, i32, i64, i32, i32, vp]
    L.b200ida_wrms_masked_batched.argtypes = [vp, vp, vp, vp, i32, i64, i32, i32, vp]
    L.b200ida_create.argtypes = [C.POINTER(vp), i32, i32, i64, i32, vp, vp, vp, dbl, vp, i32, i32, vp]
    L.b200ida_destroy.argtypes = [vp]
    L.b200ida_set_trace.argtypes = [vp, i64, vp, vp, vp, i32]
    L.b200ida_set_roots.argtypes = [vp, i32, vp, vp, vp, vp]
    L.b200ida_reinit.argtypes = [vp, vp, vp, vp, i32]
    L.b200ida_set_kernel.argtypes = [vp, i32]
    L.b200ida_set_default_reorder.argtypes = [i32]
    L.b200ida_solve.argtypes = [vp, dbl, i32, vp, i32]
    L.b200ida_solve_ladder.argtypes = [vp, vp, i32, i32, vp, vp, i32]
    L.b200ida_get_state.argtypes = [vp, vp, vp, vp, i32]
    L.b200ida_get_counters.argtypes = [vp] + [vp] * 8 + [i32]
    L.b200ida_get_step_info.argtypes = [vp] + [vp] * 6 + [i32]
    L.b200ida_get_dky.argtypes = [vp, dbl, i32, vp, vp, i32]
    L.b200ida_sync.argtypes = [vp]
    L.b200ida_fp64_issue_rate.argtypes = [i32, i32, C.POINTER(dbl)]
    L.b200ida_detpow_batched.argtypes = [vp, vp, vp, i64, i32, vp]
    L.b200ida_fastdiv_batched.argtypes = [vp, vp, vp, vp, i64]
    L.b200ida_set_id.argtypes = [vp, vp, i32]
    L.b200ida_set_regroup.argtypes = [vp, i32, i32]
    L.b200ida_newton_batched.argtypes = [i32, i32, i64, vp, vp, dbl, i32, i32, vp, vp, vp, vp, i32, vp]
    L.b200ida_kernel_launches.restype = C.c_longlong
    L.b200ida_get_fallback_count.argtypes = [vp, vp, vp]
    for name in SYMBOLS:
        getattr(L, name).restype = C.c_char_p if name == "b200ida_last_error" else i32
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise B200IdaError(rc, lib().b200ida_last_error().decode(errors="replace"))


def device_count():
    return lib().b200ida_device_count()


def require_device():
    if device_count() < 1:
        raise B200IdaError(ERR_NO_DEVICE, "no CUDA device: the engine has no CPU fallback")
