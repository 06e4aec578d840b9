"""TEST INFRASTRUCTURE (oracle) -- ctypes front end of oracle/_ref/diffvg*.so, the UNMODIFIED
reference CPU renderer driven through oracle/ref_capi.cpp.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module; the product package never does.
"""
import ctypes
import glob
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_libs = {}

# Two builds of the same unmodified sources (oracle/Makefile):
#   'det'   -ftrivial-auto-var-init=zero: pins the reference's uninitialised `intervals[0]` read
#            to "no extra split point"; the PARITY checker (default).
#   'plain' stock flags; what the CPU baseline times.
_PATTERN = {'det': 'diffvg_det.so', 'plain': 'diffvg.cpython*.so'}


def available():
    return all(len(glob.glob(os.path.join(_HERE, '_ref', p))) > 0 for p in _PATTERN.values())


def _load(variant='det'):
    if variant in _libs:
        return _libs[variant]
    cands = sorted(glob.glob(os.path.join(_HERE, '_ref', _PATTERN[variant])))
    if not cands:
        raise RuntimeError('oracle/_ref/%s not built: run `make -C oracle ref` where /root/reference exists' % _PATTERN[variant])
    lib = ctypes.CDLL(cands[0])
    fp = ctypes.POINTER(ctypes.c_float)
    ip = ctypes.POINTER(ctypes.c_int32)
    lib.dvgref_render.argtypes = [ip, fp, fp, fp, fp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                  ctypes.c_uint64, fp, fp, fp, fp, ctypes.c_int, fp, ctypes.c_int, fp]
    lib.dvgref_render.restype = ctypes.c_int
    lib.dvgref_scene_dump.argtypes = [ip, fp, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_uint32),
                                      ctypes.c_int64]
    lib.dvgref_scene_dump.restype = ctypes.c_int64
    lib.dvgref_last_error.restype = ctypes.c_char_p
    _libs[variant] = lib
    return lib


def _f(a):
    if a is None:
        return None
    assert a.dtype == np.float32 and a.flags['C_CONTIGUOUS']
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def render(topo, params, width, height, nsx, nsy, seed, background=None, d_render_image=None,
           d_render_sdf=None, want_image=True, want_sdf=False, use_prefiltering=False,
           eval_positions=None, want_d_translation=False, variant='det'):
    """One reference `Scene(...)` + `render(...)` call (diffvg.cpp:1477).

    Forward: returns dict(image=[H,W,4] and/or sdf).  Backward (d_render_image or
    d_render_sdf given): returns dict(d_params=..., d_background=..., d_translation=...)."""
    lib = _load(variant)
    topo = np.ascontiguousarray(topo, dtype=np.int32)
    params = np.ascontiguousarray(params, dtype=np.float32)
    n_eval = 0 if eval_positions is None else eval_positions.shape[0]
    backward = d_render_image is not None or d_render_sdf is not None
    out = {}
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
    rc = lib.dvgref_render(topo.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), _f(params), _f(background),
                           _f(image), _f(sdf), width, height, nsx, nsy, int(seed), _f(d_bg),
                           _f(d_render_image), _f(d_render_sdf), _f(d_tr), int(use_prefiltering),
                           _f(eval_positions), n_eval, _f(d_params))
    if rc != 0:
        raise RuntimeError(lib.dvgref_last_error().decode())
    out.update(image=image, sdf=sdf, d_params=d_params, d_background=d_bg, d_translation=d_tr)
    return out


def scene_dump(topo, params, what, index=0, cap=1 << 22):
    lib = _load()
    topo = np.ascontiguousarray(topo, dtype=np.int32)
    params = np.ascontiguousarray(params, dtype=np.float32)
    buf = np.zeros(cap, dtype=np.uint32)
    n = lib.dvgref_scene_dump(topo.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), _f(params), what, index,
                              buf.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)), cap)
    if n < 0:
        raise RuntimeError(lib.dvgref_last_error().decode())
    return buf[:n].copy()


def cuda_available():
    return os.path.exists(os.path.join(_HERE, '_ref', 'diffvg_cuda.so'))


def cuda_bench(topo, params, width, height, nsx, nsy, seed0, d_render_image, warmup, steps):
    """Times the reference's own CUDA path (oracle/_ref/diffvg_cuda.so, `make -C oracle ref_cuda`): per step
    Scene() + forward render + Scene() + backward render on the GPU.  -> (ms_per_step, image, d_params)."""
    lib = _libs.get('cuda')
    if lib is None:
        lib = ctypes.CDLL(os.path.join(_HERE, '_ref', 'diffvg_cuda.so'))
        fp = ctypes.POINTER(ctypes.c_float)
        lib.dvgref_cuda_bench.argtypes = [ctypes.POINTER(ctypes.c_int32), fp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                          ctypes.c_uint64, fp, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_double), fp, fp]
        lib.dvgref_cuda_bench.restype = ctypes.c_int
        lib.dvgref_last_error.restype = ctypes.c_char_p
        _libs['cuda'] = lib
    topo = np.ascontiguousarray(topo, dtype=np.int32)
    params = np.ascontiguousarray(params, dtype=np.float32)
    d_img = np.ascontiguousarray(d_render_image, dtype=np.float32)
    image = np.zeros((height, width, 4), np.float32)
    d_params = np.zeros(params.shape[0], np.float32)
    ms = ctypes.c_double()
    rc = lib.dvgref_cuda_bench(topo.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), _f(params), width, height, nsx, nsy,
                               int(seed0), _f(d_img), int(warmup), int(steps), ctypes.byref(ms), _f(image), _f(d_params))
    if rc != 0:
        raise RuntimeError(lib.dvgref_last_error().decode())
    return ms.value, image, d_params
