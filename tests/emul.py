"""ctypes front end of tests/host_emul/libemul.so (test-only host build of the product's
host/device arithmetic headers; see tests/host_emul/emul.cpp)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, 'host_emul', 'libemul.so')
_lib = None


def build():
    src = os.path.join(_HERE, 'host_emul', 'emul.cpp')
    deps = [src] + [os.path.join(_HERE, '..', 'diffvg_b200', 'csrc', f)
                    for f in os.listdir(os.path.join(_HERE, '..', 'diffvg_b200', 'csrc')) if f.endswith('.cuh')]
    if os.path.exists(_SO) and all(os.path.getmtime(_SO) >= os.path.getmtime(d) for d in deps):
        return
    subprocess.check_call(['g++', '-O2', '-std=c++17', '-ffp-contract=off', '-mfma', '-DDVG_FMA_QUINTIC', '-fPIC', '-shared', '-fvisibility=hidden',
                           '-o', _SO, src, '-lpthread'])


def _load():
    global _lib
    if _lib is None:
        build()
        lib = ctypes.CDLL(_SO)
        fp = ctypes.POINTER(ctypes.c_float)
        ip = ctypes.POINTER(ctypes.c_int32)
        lib.emul_render.argtypes = [ip, fp, fp, fp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_uint64,
                                    fp, fp, fp, ctypes.c_int, ctypes.c_int]
        lib.emul_render.restype = ctypes.c_int
        lib.emul_scene_dump.argtypes = [ip, fp, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_uint32), ctypes.c_int64]
        lib.emul_scene_dump.restype = ctypes.c_int64
        lib.emul_pcg.argtypes = [ctypes.c_int, ctypes.c_uint64, ctypes.POINTER(ctypes.c_uint64), fp, fp]
        _lib = lib
    return _lib


def _f(a):
    return None if a is None else a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def render(topo, params, width, height, nsx, nsy, seed, background=None, d_render_image=None, skip_xform_grad=False,
           nthreads=8):
    lib = _load()
    topo = np.ascontiguousarray(topo, dtype=np.int32)
    params = np.ascontiguousarray(params, dtype=np.float32)
    out = {}
    if d_render_image is None:
        img = np.zeros((height, width, 4), np.float32)
        rc = lib.emul_render(topo.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), _f(params), _f(background), _f(img),
                             width, height, nsx, nsy, int(seed), None, None, None, 0, nthreads)
        out['image'] = img
    else:
        d_params = np.zeros_like(params)
        d_bg = np.zeros((height, width, 4), np.float32) if background is not None else None
        rc = lib.emul_render(topo.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), _f(params), _f(background), None,
                             width, height, nsx, nsy, int(seed), _f(np.ascontiguousarray(d_render_image, dtype=np.float32)),
                             _f(d_params), _f(d_bg), int(skip_xform_grad), nthreads)
        out['d_params'] = d_params
        out['d_background'] = d_bg
    if rc != 0:
        raise RuntimeError('emul_render failed: %d' % rc)
    return out


def scene_dump(topo, params, what, index=0, cap=1 << 20):
    lib = _load()
    topo = np.ascontiguousarray(topo, dtype=np.int32)
    params = np.ascontiguousarray(params, dtype=np.float32)
    buf = np.zeros(cap, np.uint32)
    n = lib.emul_scene_dump(topo.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), _f(params), what, index,
                            buf.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)), cap)
    if n < 0:
        raise RuntimeError('emul_scene_dump failed')
    return buf[:n].copy()


def pcg(idx, seed):
    lib = _load()
    st = ctypes.c_uint64()
    rx = ctypes.c_float()
    ry = ctypes.c_float()
    lib.emul_pcg(idx, seed, ctypes.byref(st), ctypes.byref(rx), ctypes.byref(ry))
    return st.value, rx.value, ry.value


def render_pf(topo, params, width, height, nsx, nsy, seed, background=None, d_render_image=None,
              want_d_translation=False, nthreads=8):
    """use_prefiltering=True colour render (forward, or backward when d_render_image is given)."""
    lib = _load()
    fp = ctypes.POINTER(ctypes.c_float)
    ip = ctypes.POINTER(ctypes.c_int32)
    lib.emul_render_pf.argtypes = [ip, fp, fp, fp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_uint64,
                                   fp, fp, fp, fp, ctypes.c_int]
    lib.emul_render_pf.restype = ctypes.c_int
    topo = np.ascontiguousarray(topo, dtype=np.int32)
    params = np.ascontiguousarray(params, dtype=np.float32)
    out = {}
    tp = topo.ctypes.data_as(ip)
    if d_render_image is None:
        img = np.zeros((height, width, 4), np.float32)
        rc = lib.emul_render_pf(tp, _f(params), _f(background), _f(img), width, height, nsx, nsy, int(seed),
                                None, None, None, None, nthreads)
        out['image'] = img
    else:
        d_params = np.zeros_like(params)
        d_bg = np.zeros((height, width, 4), np.float32) if background is not None else None
        d_tr = np.zeros((height, width, 2), np.float32) if want_d_translation else None
        rc = lib.emul_render_pf(tp, _f(params), _f(background), None, width, height, nsx, nsy, int(seed),
                                _f(np.ascontiguousarray(d_render_image, dtype=np.float32)), _f(d_params), _f(d_bg), _f(d_tr),
                                nthreads)
        out.update(d_params=d_params, d_background=d_bg, d_translation=d_tr)
    if rc != 0:
        raise RuntimeError('emul_render_pf failed: %d' % rc)
    return out


def sdf(topo, params, width, height, nsx, nsy, seed, eval_positions=None, d_render_sdf=None, want_d_translation=False):
    lib = _load()
    fp = ctypes.POINTER(ctypes.c_float)
    ip = ctypes.POINTER(ctypes.c_int32)
    lib.emul_sdf.argtypes = [ip, fp, fp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_uint64,
                             fp, ctypes.c_int, fp, fp, fp]
    lib.emul_sdf.restype = ctypes.c_int
    topo = np.ascontiguousarray(topo, dtype=np.int32)
    params = np.ascontiguousarray(params, dtype=np.float32)
    n_eval = 0 if eval_positions is None else eval_positions.shape[0]
    ep = None if eval_positions is None else np.ascontiguousarray(eval_positions, dtype=np.float32)
    out = {}
    if d_render_sdf is None:
        s = np.zeros((n_eval, 1) if n_eval else (height, width, 1), np.float32)
        rc = lib.emul_sdf(topo.ctypes.data_as(ip), _f(params), _f(s), width, height, nsx, nsy, int(seed), _f(ep), n_eval,
                          None, None, None)
        out['sdf'] = s
    else:
        d_params = np.zeros_like(params)
        d_tr = np.zeros((height, width, 2), np.float32) if want_d_translation else None
        rc = lib.emul_sdf(topo.ctypes.data_as(ip), _f(params), None, width, height, nsx, nsy, int(seed), _f(ep), n_eval,
                          _f(np.ascontiguousarray(d_render_sdf, dtype=np.float32)), _f(d_params), _f(d_tr))
        out.update(d_params=d_params, d_translation=d_tr)
    if rc != 0:
        raise RuntimeError('emul_sdf failed: %d' % rc)
    return out
