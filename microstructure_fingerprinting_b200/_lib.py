"""ctypes binding of libmfb200.so (C ABI in include/mfb200.h).

There is no CPU fallback: if the CUDA library is missing or no B200-class device is
visible, every compute entry point raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MFB_LIB") or os.path.join(_HERE, "libmfb200.so")   # MFB_LIB: A/B builds

MFB_OK, MFB_EINVAL, MFB_ECUDA, MFB_ENOMEM, MFB_EUNSUPPORTED = 0, -1, -2, -3, -4
MFB_ABI_VERSION = 4          # include/mfb200.h; a library built from other sources is refused
MFB_F64, MFB_F32 = 0, 1

c_dp = ctypes.POINTER(ctypes.c_double)
c_ip = ctypes.POINTER(ctypes.c_int32)
c_lp = ctypes.POINTER(ctypes.c_int64)
c_bp = ctypes.POINTER(ctypes.c_uint8)
c_vp = ctypes.c_void_p

# every symbol include/mfb200.h declares: (restype, argtypes)
SYMBOLS = {
    "mfb_version": (ctypes.c_int, []),
    "mfb_last_error": (ctypes.c_char_p, []),
    "mfb_launch_count": (ctypes.c_int64, []),
    "mfb_plan_create": (c_vp, [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                               c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp,
                               ctypes.c_int]),
    "mfb_plan_destroy": (None, [c_vp]),
    "mfb_rotate_multishell": (ctypes.c_int, [c_vp, ctypes.c_int64, c_vp, c_vp, ctypes.c_int64, c_vp]),
    "mfb_lerp_rows": (ctypes.c_int, [ctypes.c_int, ctypes.c_int64, ctypes.c_int, ctypes.c_int, c_vp, c_vp,
                                     c_vp, c_vp, c_vp, c_vp, c_vp, ctypes.c_int64, c_vp]),
    "mfb_solve_batch": (ctypes.c_int, [ctypes.c_int, ctypes.c_int64, ctypes.c_int, ctypes.c_int, c_vp,
                                       c_vp, ctypes.c_int64, ctypes.c_int64, c_vp, c_vp, c_vp, c_vp,
                                       c_vp, ctypes.c_int, c_vp]),
    "mfb_fit": (ctypes.c_int, [c_vp, ctypes.c_int64, c_vp, c_vp, c_vp, c_vp, c_vp, ctypes.c_int,
                               ctypes.c_int, ctypes.c_int, c_vp, ctypes.c_int, c_vp]),
    "mfb_fit_host": (ctypes.c_int, [c_vp, ctypes.c_int64, c_vp, c_vp, c_vp, c_vp, c_vp, ctypes.c_int,
                                    ctypes.c_int, ctypes.c_int, c_vp, ctypes.c_int]),
    "mfb_fit_volume": (ctypes.c_int, [c_vp, ctypes.c_int64, c_vp, ctypes.c_int, c_vp, ctypes.c_int64,
                                      ctypes.c_int64, c_vp, c_vp, c_vp, c_vp, ctypes.c_int, ctypes.c_int,
                                      ctypes.c_int, c_vp, ctypes.c_int]),
    "mfb_fit_stats": (ctypes.c_int, [c_vp, c_vp, ctypes.c_int]),
    "mfb_solve_stats": (ctypes.c_int, [c_vp, ctypes.c_int, ctypes.c_int]),
    "mfb_trim": (ctypes.c_int, [ctypes.c_int]),
    "mfb_plan2d": (ctypes.c_int, [ctypes.c_int, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                  c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, ctypes.c_double,
                                  c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp,
                                  c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "mfb_mc_average": (ctypes.c_int, [ctypes.c_int, ctypes.c_int64, ctypes.c_int, c_vp, ctypes.c_int64, c_vp,
                                      c_vp, ctypes.c_double, ctypes.c_int64, c_vp, c_vp]),
}

_lib = None


class MFBError(RuntimeError):
    pass


def load():
    """Load libmfb200.so (building it first if only the sources are present)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        try:
            from . import build as _build
            _build.build()
        except Exception as exc:  # pragma: no cover - depends on toolchain
            raise MFBError(
                "libmfb200.so is missing and could not be built (%s). Run "
                "`python -m microstructure_fingerprinting_b200.build`; this package has no "
                "CPU fallback." % (exc,))
    lib = ctypes.CDLL(LIB_PATH)
    lib.mfb_version.restype = ctypes.c_int
    if lib.mfb_version() != MFB_ABI_VERSION:
        raise MFBError("%s was built for ABI version %d, this package binds version %d: rebuild it "
                       "(`python -m microstructure_fingerprinting_b200.build --force`)."
                       % (LIB_PATH, lib.mfb_version(), MFB_ABI_VERSION))
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error():
    return load().mfb_last_error().decode("utf-8", "replace")


def check(rc, what=""):
    if rc == MFB_OK:
        return
    msg = "%s%s" % (what + ": " if what else "", last_error())
    if rc == MFB_EINVAL:
        raise ValueError(msg)
    if rc == MFB_ENOMEM:
        raise MemoryError(msg)
    if rc == MFB_EUNSUPPORTED:
        raise NotImplementedError(msg)
    raise MFBError(msg)


def solve_stats(reset=False):
    """Counters of mfb_solve_batch (include/mfb200.h): [screened, redone, no candidate,
    ill-conditioned, near tie, fewer active columns]."""
    out = (ctypes.c_int64 * 6)()
    check(load().mfb_solve_stats(out, 6, 1 if reset else 0), "mfb_solve_stats")
    return [int(x) for x in out]


def trim(device=0):
    """Free the device workspace mfb_solve_batch keeps between calls (mfb_trim)."""
    check(load().mfb_trim(int(device)), "mfb_trim")


def launch_count():
    return int(load().mfb_launch_count())


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise MFBError("No CUDA device visible: microstructure_fingerprinting_b200 runs its hot "
                       "path only on NVIDIA B200 (sm_100a) GPUs and has no CPU fallback.")
    return torch
