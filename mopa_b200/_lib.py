"""ctypes binding of libmopa_scn.so (the C ABI declared in include/mopa_scn.h).

There is NO CPU fallback: if the library is missing or an entry point fails, the call raises. The prototypes below
are the single source of truth for the Python side; tests/test_abi.py checks them against the header.
"""
import ctypes
import os
import warnings

from . import _build

_c = ctypes
_p = _c.c_void_p
_i64 = _c.c_int64
_int = _c.c_int
_f = _c.c_float
_sz = _c.c_size_t

PREC_FP32 = 0
PREC_TF32 = 1
ABI_VERSION = 2  # MOPA_SCN_ABI_VERSION in include/mopa_scn.h

# name -> (restype, argtypes); order and types mirror include/mopa_scn.h
PROTOTYPES = {
    "mopa_scn_abi_version": (_int, []),
    "mopa_scn_last_error": (_c.c_char_p, []),
    "mopa_scn_Metadata_new": (_p, [_int, _int]),
    "mopa_scn_Metadata_delete": (None, [_p]),
    "mopa_scn_InputLayer_setLocations": (_int, [_p, _i64, _p, _i64, _int, _int, _int, _p, _c.POINTER(_i64)]),
    "mopa_scn_InputLayer_updateOutput": (_int, [_p, _p, _i64, _int, _p, _i64, _p]),
    "mopa_scn_InputLayer_updateGradInput": (_int, [_p, _p, _i64, _p, _i64, _int, _p]),
    "mopa_scn_OutputLayer_updateOutput": (_int, [_p, _p, _i64, _int, _p, _i64, _p]),
    "mopa_scn_OutputLayer_updateGradInput": (_int, [_p, _p, _i64, _p, _i64, _int, _p]),
    "mopa_scn_Metadata_prepareSubmanifold": (_int, [_p, _i64, _int, _p, _c.POINTER(_i64)]),
    "mopa_scn_Metadata_prepareConvolution": (_int, [_p, _i64, _i64, _int, _int, _p, _c.POINTER(_i64)]),
    "mopa_scn_Metadata_getNActive": (_i64, [_p, _i64]),
    "mopa_scn_Metadata_getNPoints": (_i64, [_p]),
    "mopa_scn_Metadata_getSpatialLocations": (_int, [_p, _i64, _p]),
    "mopa_scn_Metadata_getPointToVoxel": (_int, [_p, _p]),
    "mopa_scn_Metadata_getInputRules": (_int, [_p, _p, _p]),
    "mopa_scn_Metadata_getSubmanifoldRuleBook": (_int, [_p, _i64, _p, _p]),
    "mopa_scn_Metadata_getConvolutionRuleBook": (_int, [_p, _i64, _p, _p]),
    "mopa_scn_Metadata_getTileRuleBook": (_int, [_p, _i64, _int, _c.POINTER(_i64), _p, _p]),
    "mopa_scn_packedWeightFloats": (_i64, [_int, _int, _int, _int]),
    "mopa_scn_packWeights": (_int, [_p, _int, _int, _int, _int, _int, _int, _p, _p]),
    "mopa_scn_SubmanifoldConvolution_updateOutput": (_int, [_p, _i64, _int, _p, _i64, _p, _i64, _p, _p, _int, _int, _int, _p]),
    "mopa_scn_Convolution_updateOutput": (_int, [_p, _i64, _i64, _int, _int, _p, _i64, _p, _i64, _p, _p, _int, _int, _int, _p]),
    "mopa_scn_Deconvolution_updateOutput": (_int, [_p, _i64, _i64, _int, _int, _p, _i64, _p, _i64, _p, _p, _int, _int, _int, _p]),
    "mopa_scn_backwardWorkspaceBytes": (_sz, [_int, _int, _int, _i64]),
    "mopa_scn_debug_dweightPlan": (_int, [_int, _int, _i64, ctypes.POINTER(ctypes.c_int)]),
    "mopa_scn_SubmanifoldConvolution_backward": (_int, [_p, _i64, _int, _p, _i64, _p, _i64, _p, _i64, _p, _p, _p, _int, _int, _int, _p, _sz, _p]),
    "mopa_scn_Convolution_backward": (_int, [_p, _i64, _i64, _int, _int, _p, _i64, _p, _i64, _p, _i64, _p, _p, _p, _int, _int, _int, _p, _sz, _p]),
    "mopa_scn_Deconvolution_backward": (_int, [_p, _i64, _i64, _int, _int, _p, _i64, _p, _i64, _p, _i64, _p, _p, _p, _int, _int, _int, _p, _sz, _p]),
    "mopa_scn_Metadata_getSubmanifoldRuleCount": (_i64, [_p, _i64]),
    "mopa_scn_bnWorkspaceBytes": (_sz, [_int]),
    "mopa_scn_BatchNormalization_updateOutput": (_int, [_p, _i64, _p, _i64, _p, _p, _p, _p, _p, _p, _f, _f, _int, _f, _i64, _int, _p, _sz, _p]),
    "mopa_scn_BatchNormalization_backward": (_int, [_p, _i64, _p, _i64, _p, _i64, _p, _p, _p, _p, _p, _p, _f, _int, _i64, _int, _p, _sz, _p]),
    "mopa_scn_Program_new": (_p, [_p, _int, _p, _int, _int, _int, _int, _i64, _int, _int]),
    "mopa_scn_Program_delete": (None, [_p]),
    "mopa_scn_Program_prepare": (_int, [_p, _p, _p, _i64, _int, _int, _int, _p, _p, _p]),
    "mopa_scn_Program_forward": (_int, [_p, _p, _p, _i64, _p, _int, _int, _p, _p, _p, _i64, _p]),
    "mopa_scn_Program_backward": (_int, [_p, _p, _p, _p, _int, _int, _p, _p, _p, _p, _i64, _p, _i64, _p]),
    "mopa_scn_kernelLaunchCount": (_i64, []),
    # include/mopa_xm.h
    "mopa_xm_checkAsyncError": (_int, [_p]),
    "mopa_xm_PixelGatherHeads_updateOutput": (_int, [_p, _int, _int, _int, _int, _p, _p, _i64, _p, _p, _p, _p, _int, _p, _p, _p, _p]),
    "mopa_xm_pixelGatherWorkspaceBytes": (_sz, [_int, _int]),
    "mopa_xm_PixelGatherHeads_backward": (_int, [_p, _p, _p, _int, _int, _int, _int, _i64, _p, _p, _int, _p, _p, _p, _p, _p, _p, _p,
                                                 _p, _p, _sz, _p]),
    "mopa_xm_KLDivLoss_updateOutput": (_int, [_p, _p, _i64, _int, _p, _p, _f, _p]),
    "mopa_xm_maskConsStatsBytes": (_sz, [_int, _int, _int]),
    "mopa_xm_MaskConsLoss_updateOutput": (_int, [_p, _p, _int, _i64, _int, _int, _int, _f, _p, _p, _p]),
    "mopa_xm_MaskConsLoss_backward": (_int, [_p, _p, _int, _i64, _int, _int, _int, _f, _p, _f, _p, _p]),
    "mopa_xm_vgiWorkspaceBytes": (_sz, [_i64, _int, _int]),
    "mopa_xm_VgiPostProcess": (_int, [_p, _p, _i64, _int, _c.c_double, _c.c_double, _int, _int, _p, _p, _c.c_double, _i64, _int,
                                      _p, _p, _p, _p, _p, _p, _sz, _p]),
    "mopa_scn_Profile_enable": (_int, [_int]),
    "mopa_scn_Profile_count": (_i64, []),
    "mopa_scn_Profile_read": (_int, [_p, _i64]),
}


class ProfileRecord(ctypes.Structure):
    """mopa_scn_profile_record (include/mopa_scn.h)"""
    _fields_ = [("tag", _c.c_int32), ("volume", _c.c_int32), ("c_in", _c.c_int32), ("c_out", _c.c_int32),
                ("rows_out", _i64), ("rows_in", _i64), ("rules", _i64), ("ms", _f), ("reserved", _c.c_int32)]

_lib = None


class ScnError(RuntimeError):
    pass


def _nvcc_present():
    try:
        _build._nvcc()
        return True
    except RuntimeError:
        return False


def library_path():
    return _build.LIB


def load():
    """Load (building in-tree first if the .so is absent or stale and nvcc exists). Raises if unavailable."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB
    if not os.path.exists(path) or _build.stale():
        try:
            _build.build()  # takes a file lock: one builder per checkout, atomic replace of the .so
        except Exception as e:  # no nvcc on this box, or the sources do not compile
            if not os.path.exists(path):
                raise ScnError("libmopa_scn.so is missing and could not be built (%s); mopa_b200 has no CPU fallback" % e)
            if os.environ.get("MOPA_SCN_ALLOW_STALE") != "1" and _nvcc_present():
                raise ScnError("libmopa_scn.so is older than its sources and the rebuild failed (%s); set "
                               "MOPA_SCN_ALLOW_STALE=1 to load the old binary anyway" % e)
            warnings.warn("libmopa_scn.so is older than its sources and could not be rebuilt (%s): loading the existing "
                          "binary" % e, RuntimeWarning)
    lib = _c.CDLL(path)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.mopa_scn_abi_version() != ABI_VERSION:
        raise ScnError("libmopa_scn.so reports ABI version %d, this binding was written for %d: rebuild the library"
                       % (lib.mopa_scn_abi_version(), ABI_VERSION))
    _lib = lib
    return lib


def check(status):
    if status != 0:
        raise ScnError(load().mopa_scn_last_error().decode("utf-8", "replace"))


def kernel_launches():
    return int(load().mopa_scn_kernelLaunchCount())


def profile_enable(on):
    check(load().mopa_scn_Profile_enable(1 if on else 0))


def profile_read():
    """List of dicts, one per op launched since profile_enable(True) (synchronises the device)."""
    lib = load()
    n = int(lib.mopa_scn_Profile_count())
    recs = (ProfileRecord * max(n, 1))()
    check(lib.mopa_scn_Profile_read(ctypes.cast(recs, _p), n))
    return [{f: getattr(recs[i], f) for f, _ in ProfileRecord._fields_ if f != "reserved"} for i in range(n)]
