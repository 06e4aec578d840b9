"""TEST INFRASTRUCTURE (oracle) -- ctypes front end of oracle/liboracle.so, the plain-C restatement
of the reference algorithm (oracle/dvg_oracle.c).  Same call surface as ref_oracle.py.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; the product package never does.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, 'liboracle.so')
_lib = None


def available():
    return os.path.exists(_SO)


def _load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_SO):
        raise RuntimeError('oracle/liboracle.so not built: run `make -C oracle oracle`')
    lib = ctypes.CDLL(_SO)
    fp = ctypes.POINTER(ctypes.c_float)
    ip = ctypes.POINTER(ctypes.c_int32)
    lib.dvgo_render.argtypes = [ip, fp, fp, fp, fp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                ctypes.c_uint64, fp, fp, fp, fp, ctypes.c_int, fp, ctypes.c_int, fp, ctypes.c_int]
    lib.dvgo_render.restype = ctypes.c_int
    lib.dvgo_scene_dump.argtypes = [ip, fp, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_uint32), ctypes.c_int64]
    lib.dvgo_scene_dump.restype = ctypes.c_int64
    lib.dvgo_pcg.argtypes = [ctypes.c_int, ctypes.c_uint64, ctypes.POINTER(ctypes.c_uint64), fp, fp]
    lib.dvgo_last_error.restype = ctypes.c_char_p
    _lib = lib
    return lib


def _f(a):
    if a is None:
        return None
    assert a.dtype == np.float32 and a.flags['C_CONTIGUOUS']
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def render(topo, params, width, height, nsx, nsy, seed, background=None, d_render_image=None,
           d_render_sdf=None, want_image=True, want_sdf=False, use_prefiltering=False,
           eval_positions=None, want_d_translation=False, nthreads=0):
    lib = _load()
    topo = np.ascontiguousarray(topo, dtype=np.int32)
    params = np.ascontiguousarray(params, dtype=np.float32)
    n_eval = 0 if eval_positions is None else eval_positions.shape[0]
    backward = d_render_image is not None or d_render_sdf is not None
    image = sdf = d_params = d_bg = d_tr = None
    if not backward:
        if want_image:
            image = np.zeros((height, width, 4), dtype=np.float32)
        if want_sdf:
            sdf = np.zeros((n_eval, 1) if n_eval else (height, width, 1), dtype=np.float32)
    else:
        d_params = np.zeros(params.shape[0], dtype=np.float32)
        if background is not None:
            d_bg = np.zeros((height, width, 4), dtype=np.float32)
        if want_d_translation:
            d_tr = np.zeros((height, width, 2), dtype=np.float32)
    rc = lib.dvgo_render(topo.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), _f(params), _f(background),
                         _f(image), _f(sdf), width, height, nsx, nsy, int(seed), _f(d_bg),
                         _f(d_render_image), _f(d_render_sdf), _f(d_tr), int(use_prefiltering),
                         _f(eval_positions), n_eval, _f(d_params), nthreads or (os.cpu_count() or 1))
    if rc != 0:
        raise RuntimeError(lib.dvgo_last_error().decode())
    return dict(image=image, sdf=sdf, d_params=d_params, d_background=d_bg, d_translation=d_tr)


def scene_dump(topo, params, what, index=0, cap=1 << 22):
    lib = _load()
    topo = np.ascontiguousarray(topo, dtype=np.int32)
    params = np.ascontiguousarray(params, dtype=np.float32)
    buf = np.zeros(cap, dtype=np.uint32)
    n = lib.dvgo_scene_dump(topo.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), _f(params), what, index,
                            buf.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)), cap)
    if n < 0:
        raise RuntimeError(lib.dvgo_last_error().decode())
    return buf[:n].copy()


def pcg(idx, seed):
    lib = _load()
    st = ctypes.c_uint64()
    rx = ctypes.c_float()
    ry = ctypes.c_float()
    lib.dvgo_pcg(idx, seed, ctypes.byref(st), ctypes.byref(rx), ctypes.byref(ry))
    return st.value, rx.value, ry.value
