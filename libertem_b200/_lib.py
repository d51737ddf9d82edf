"""ctypes binding of libltb200.so (C ABI: include/ltb200.h).  Fails loudly when missing."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, '_native', 'libltb200.so')

LTB_F32, LTB_U16, LTB_U8, LTB_I16, LTB_F64, LTB_I32, LTB_U32, LTB_I64, LTB_U64, LTB_I8 = range(10)

_c = ctypes
_vp, _i64, _int, _sz, _u32 = _c.c_void_p, _c.c_int64, _c.c_int, _c.c_size_t, _c.c_uint32

#: every symbol include/ltb200.h declares, with its signature
SIGNATURES = {
    'ltb200_abi_version': (_int, []),
    'ltb200_last_error': (_c.c_char_p, []),
    'ltb200_device_info': (_int, [_int, _c.POINTER(_int), _c.POINTER(_int), _c.POINTER(_int),
                                  _c.POINTER(_i64)]),
    'ltb200_masks_dense_workspace': (_sz, [_i64, _i64, _int, _int]),
    'ltb200_masks_dense': (_int, [_vp, _int, _i64, _i64, _i64, _vp, _int, _i64, _vp, _i64, _int,
                                  _vp, _vp, _sz, _vp]),
    'ltb200_masks_dense_f64': (_int, [_vp, _int, _i64, _i64, _i64, _vp, _int, _i64, _vp, _i64,
                                      _int, _vp]),
    'ltb200_masks_dense_tc_workspace': (_sz, [_i64, _i64, _int]),
    'ltb200_masks_dense_tc': (_int, [_vp, _i64, _i64, _i64, _vp, _int, _i64, _vp, _i64, _int,
                                     _int, _vp, _sz, _vp]),
    'ltb200_masks_dense_tc_u16_workspace': (_sz, [_i64, _i64, _int, _int]),
    'ltb200_masks_dense_tc_u16': (_int, [_vp, _i64, _i64, _i64, _vp, _int, _i64, _vp, _i64, _int,
                                         _int, _vp, _vp, _sz, _vp]),
    'ltb200_masks_dense_i8_workspace': (_sz, [_int, _i64, _i64, _int, _int]),
    'ltb200_masks_dense_i8': (_int, [_vp, _int, _i64, _i64, _i64, _vp, _int, _i64, _vp, _i64,
                                     _int, _vp, _vp, _sz, _vp]),
    'ltb200_set_k1_variant': (_int, [_int]),
    'ltb200_last_kernel': (_int, []),
    'ltb200_launch_count': (_i64, [_int]),
    'ltb200_masks_csc': (_int, [_vp, _int, _i64, _i64, _i64, _vp, _vp, _vp, _int, _vp, _i64,
                                _int, _vp]),
    'ltb200_masks_shifted': (_int, [_vp, _int, _i64, _int, _int, _i64, _vp, _int, _i64, _vp, _int,
                                    _vp, _i64, _int, _vp]),
    'ltb200_masks_shifted_banded_workspace': (_sz, [_i64, _int, _int, _int, _int]),
    'ltb200_masks_shifted_banded': (_int, [_vp, _int, _i64, _int, _int, _i64, _vp, _int, _i64, _vp,
                                           _int, _vp, _int, _vp, _i64, _int, _vp, _sz, _vp]),
    'ltb200_masks_shifted_f64': (_int, [_vp, _int, _i64, _int, _int, _i64, _vp, _int, _i64, _vp,
                                        _int, _vp, _i64, _int, _vp]),
    'ltb200_group_masks_workspace': (_sz, []),
    'ltb200_group_masks': (_int, [_vp, _int, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _int, _int,
                                  _vp, _i64, _int, _vp, _sz, _vp]),
    'ltb200_group_masks_tc_columns': (_int, [_int]),
    'ltb200_group_masks_tc': (_int, [_vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _int, _int, _vp,
                                     _i64, _int, _int, _vp, _sz, _vp]),
    'ltb200_group_masks_tc_workspace': (_sz, [_i64, _int, _int, _int]),
    'ltb200_group_masks_tc_banded': (_int, [_vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _int, _int,
                                            _int, _vp, _i64, _int, _int, _vp, _sz, _vp]),
    'ltb200_group_masks_tc_sym': (_int, [_vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _int, _int,
                                         _int, _vp, _i64, _int, _int, _vp, _sz, _vp]),
    'ltb200_group_masks_walk_workspace': (_sz, [_i64, _int, _int, _int]),
    'ltb200_group_masks_walk': (_int, [_vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                       _vp, _vp, _vp, _vp, _vp, _int, _int, _int, _vp, _i64,
                                       _int, _vp, _sz, _vp]),
    'ltb200_synth_fill': (_int, [_vp, _int, _i64, _i64, _u32, _vp]),
    'ltb200_probe_read': (_int, [_vp, _sz, _int, _vp, _vp]),
    'ltb200_com_workspace': (_sz, [_int, _int]),
    'ltb200_com_postprocess': (_int, [_vp, _i64, _vp, _vp, _int, _int, _c.c_double, _c.c_double,
                                      _vp, _int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                      _sz, _vp]),
    'ltb200_com_gradient_gram': (_int, [_vp, _vp, _int, _int, _int, _int, _int, _int, _vp, _vp,
                                        _sz, _vp]),
    'ltb200_com_divergence_stats': (_int, [_vp, _vp, _int, _int, _int, _int, _int, _int, _vp, _int,
                                           _c.c_double, _vp, _vp, _vp, _vp]),
}


class LTB200Error(RuntimeError):
    pass


_lib = None


def get_lib():
    """Load libltb200.so; raise (never fall back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise LTB200Error(
                f'{LIB_PATH} not found: build the CUDA extension first '
                '(python -c "import __graft_entry__ as g; g.build()" or make -C libertem_b200/csrc). '
                'There is no CPU fallback.'
            )
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc):
    if rc != 0:
        msg = get_lib().ltb200_last_error().decode('utf8', 'replace')
        raise LTB200Error(f'libltb200 error {rc}: {msg}')
