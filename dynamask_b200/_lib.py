"""ctypes binding of ``libdynamask_sm100.so`` (the C ABI declared in ``include/dynamask_sm100.h``).

There is deliberately no fallback: if the shared library is missing or a symbol is absent the
import of the first op raises, so a GPU run can never silently degrade to a PyTorch/CPU path.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# DYNAMASK_LIB selects another build of the same library (kernel experiments); there is still no fallback
LIB_PATH = os.environ.get('DYNAMASK_LIB') or os.path.join(_HERE, 'lib', 'libdynamask_sm100.so')

_c_f32p = ctypes.c_void_p  # raw device / host addresses are passed as integers
_vp = ctypes.c_void_p
_i = ctypes.c_int
_i64 = ctypes.c_int64
_f = ctypes.c_float

# name -> (restype, argtypes); mirrors include/dynamask_sm100.h one to one
SIGNATURES = {
    'dm_version': (_i, []),
    'dm_error_string': (ctypes.c_char_p, [_i]),
    'dm_last_cuda_error': (ctypes.c_char_p, []),
    'dm_launch_count': (_i64, []),
    'dm_assign': (_i, [_vp, _i, _vp, _i, _i, _f, _vp, _vp, _vp, _vp, _vp]),
    'dm_roi_align_fwd': (_i, [_vp, _vp, _vp, _vp, _i, _vp, _i, _vp, _vp, _vp, _i, _vp, _vp, _vp,
                              _i, _i, _vp, _vp]),
    'dm_roi_align_bwd': (_i, [_vp, _vp, _vp, _vp, _i, _vp, _i, _vp, _vp, _vp, _i, _vp, _vp, _vp,
                              _i, _i, _i, _vp, _vp]),
    'dm_paste_masks': (_i, [_vp, _i64, _i64, _vp, _i, _i, _i, _i, _vp, _i, _i, _i, _i, _i, _i, _f,
                            _i, _vp, _vp]),
    'dm_paste_masks_select': (_i, [_vp, _i64, _i64, _vp, _i, _i, _i, _i, _vp, _i, _i, _i, _i, _i, _i, _f,
                                   _i, _vp, _i, _i, _vp, _vp]),
    'dm_mask_target': (_i, [_vp, _vp, _vp, _i, _vp, _vp, _vp, _i, _i, _vp, _i, _vp, _vp]),
    'dm_paste_rle': (_i, [_vp, _i64, _i64, _vp, _i, _i, _i, _i, _vp, _i, _i, _i, _i, _i, _i, _f,
                          _i, _vp, _vp, _vp, _vp, _vp]),
    'dm_rle_from_canvas': (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    'dm_rle_strings': (_i, [_vp, _vp, _i, _i64, _vp, _vp, _vp, _vp, _vp, _vp]),
    'dm_paste_rle_strings_workspace': (_i64, [_i, _i, _i, _i64]),
    'dm_paste_rle_strings': (_i, [_vp, _i64, _i64, _vp, _i, _i, _i, _i, _vp, _i, _i, _i, _i, _i, _i, _f,
                                  _i, _vp, _i64, _vp, _vp, _vp]),
    'dm_rle_compress_host': (_i64, [_vp, _i64, _i64, _vp, _i64]),
    'dm_rle_compress_batch_host': (_i64, [_vp, _vp, _i64, _i64, _vp, _i64, _vp]),
    'dm_polygon_target': (_i, [_vp, _vp, _vp, _i, _vp, _i, _vp, _vp, _vp, _i, _i, _vp, _i, _vp, _vp]),
    'dm_refine_stages': (_i, [_vp, _vp, _i, _i, _vp, _vp]),
    'dm_simple_roi_align_fwd': (_i, [_vp, _vp, _vp, _f, _vp, _i, _i, _i, _vp, _vp, _i, _vp, _vp]),
    'dm_simple_roi_align_bwd': (_i, [_vp, _vp, _vp, _f, _vp, _i, _i, _i, _vp, _vp, _i, _i, _vp, _vp]),
}

DM_ECUDA = -2   # include/dynamask_sm100.h

_lib = None


class DynaMaskLibraryError(RuntimeError):
    pass


def build_library(verbose=False):
    """Compile the sm_100a library in-tree (nvcc cross-compiles without a GPU)."""
    import subprocess
    cmd = ['make', '-C', os.path.join(_HERE, 'csrc'), '-j4']
    if not verbose:
        cmd.append('-s')
    subprocess.check_call(cmd)
    return LIB_PATH


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DynaMaskLibraryError(
            '%s is missing: build it with `python -c "import __graft_entry__ as g; g.build()"` or '
            '`make -C dynamask_b200/csrc`. dynamask_b200 has no CPU / PyTorch fallback.' % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise DynaMaskLibraryError('symbol %s missing from %s' % (name, LIB_PATH)) from e
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        lib = load()
        msg = lib.dm_error_string(rc).decode()
        # the CUDA text is thread-local and never cleared: only a DM_ECUDA return owns it
        cuda = lib.dm_last_cuda_error().decode() if rc == DM_ECUDA else ''
        raise RuntimeError('%s failed: %s%s' % (what, msg, (' [' + cuda + ']') if cuda else ''))


def launch_count():
    return int(load().dm_launch_count())
