"""ctypes binding of libdiffvg_b200.so (C ABI in include/diffvg_b200.h).

There is deliberately no fallback: if the CUDA library has not been built, importing this
module raises, and every render call needs a CUDA device."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('DVG_B200_LIB') or os.path.join(_HERE, 'libdiffvg_b200.so')  # env override: A/B builds

if not os.path.exists(LIB_PATH):
    raise ImportError('diffvg_b200: %s is missing. Build it with `python -c "import __graft_entry__ as g; g.build()"` '
                      '(nvcc, sm_100a); there is no CPU fallback.' % LIB_PATH)

lib = ctypes.CDLL(LIB_PATH)

_vp = ctypes.c_void_p
_i = ctypes.c_int
_u64 = ctypes.c_uint64
_i64 = ctypes.c_int64

lib.dvg_abi_version.restype = _i
lib.dvg_last_error.restype = ctypes.c_char_p
lib.dvg_kernel_launch_count.restype = _i64
lib.dvg_scene_create.argtypes = [_vp, _i64, _i, ctypes.POINTER(_vp)]
lib.dvg_scene_create.restype = _i
lib.dvg_scene_set_params.argtypes = [_vp, _vp, _i64, _i, _vp]
lib.dvg_scene_set_params.restype = _i
lib.dvg_render_forward.argtypes = [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _u64, _i, _vp, _i, _vp]
lib.dvg_render_forward.restype = _i
lib.dvg_render_backward.argtypes = [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _u64, _i, _vp, _i, _vp, _vp, _vp,
                                    ctypes.c_uint32, _vp]
lib.dvg_render_backward.restype = _i
lib.dvg_render_forward_rows.argtypes = [_vp, _vp, _vp, _i, _i, _i, _i, _u64, _i, _i, _i, _vp]
lib.dvg_render_forward_rows.restype = _i
lib.dvg_render_backward_rows.argtypes = [_vp, _vp, _vp, _i, _i, _i, _i, _u64, _i, _i, _i, _vp, _vp, ctypes.c_uint32, _vp]
lib.dvg_render_backward_rows.restype = _i
lib.dvg_scene_create_batch.argtypes = [_vp, _i64, _i, _i, ctypes.POINTER(_vp)]
lib.dvg_scene_create_batch.restype = _i
lib.dvg_render_forward_batch.argtypes = [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp]
lib.dvg_render_forward_batch.restype = _i
lib.dvg_render_backward_batch.argtypes = [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, ctypes.c_uint32, _vp]
lib.dvg_render_backward_batch.restype = _i
lib.dvg_scene_row_costs.argtypes = [_vp, _i, _i, _i, _i, _i, _vp, _i, ctypes.POINTER(_i), _vp]
lib.dvg_scene_row_costs.restype = _i
lib.dvg_scene_destroy.argtypes = [_vp]
lib.dvg_scene_destroy.restype = _i
lib.dvg_scene_dump.argtypes = [_vp, _i, _i, _vp, _i64, _vp]
lib.dvg_scene_dump.restype = _i64

lib.dvg_debug_set_limits.argtypes = [_i64, _i64]
lib.dvg_debug_set_limits.restype = _i
lib.dvg_debug_set_prefilter_inline.argtypes = [_i]
lib.dvg_debug_set_prefilter_inline.restype = _i
lib.dvg_profile_enable.argtypes = [_i]
lib.dvg_profile_enable.restype = _i
lib.dvg_profile_report.argtypes = [ctypes.c_char_p, _i64]
lib.dvg_profile_report.restype = _i64
lib.dvg_measure_peak.argtypes = [_i, _i, ctypes.POINTER(ctypes.c_double)]
lib.dvg_measure_peak.restype = _i

DVG_ERR_UNSUPPORTED = 4
DVG_BWD_SKIP_XFORM_GRAD = 1
DVG_BWD_ACCUMULATE = 2
DVG_BWD_SKIP_FILTER_GRAD = 4


def check(rc):
    if rc != 0:
        msg = lib.dvg_last_error().decode()
        if rc == DVG_ERR_UNSUPPORTED:
            raise NotImplementedError(msg)
        raise RuntimeError(msg)


def launch_count():
    return int(lib.dvg_kernel_launch_count())


def profile_enable(on):
    check(lib.dvg_profile_enable(1 if on else 0))


def profile_report():
    """{kernel name: (launches, total_ms)} for the launches since the last report."""
    buf = ctypes.create_string_buffer(1 << 16)
    n = lib.dvg_profile_report(buf, len(buf))
    if n < 0:
        raise RuntimeError(lib.dvg_last_error().decode())
    out = {}
    for line in buf.value.decode().splitlines():
        name, cnt, ms = line.rsplit(',', 2)
        out[name] = (int(cnt), float(ms))
    return out


def measure_peak(which, device=0):
    v = ctypes.c_double()
    check(lib.dvg_measure_peak(which, device, ctypes.byref(v)))
    return v.value
