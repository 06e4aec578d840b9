"""Ad-hoc: raw per-primitive predicates of a few pixels, device vs host build of the same headers (run under gpurun)."""
import ctypes, os, sys, warnings
import numpy as np, torch
warnings.simplefilter('ignore')
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')):
    sys.path.insert(0, p)
import emul
from diffvg_b200 import _native as n

def run(name, W, H, ns, seed, pixels):
    g = np.load(os.path.join(ROOT, 'tests', 'golden_svg', name + '.npz'))
    topo = np.ascontiguousarray(g['topo'], np.int32); params = np.ascontiguousarray(g['params'], np.float32)
    h = ctypes.c_void_p()
    n.check(n.lib.dvg_scene_create(topo.ctypes.data, topo.shape[0], 0, ctypes.byref(h)))
    n.check(n.lib.dvg_scene_set_params(h, params.ctypes.data, params.shape[0], 0, None))
    lib = emul._load()
    ip, fp = ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_float)
    lib.emul_debug_prim_tests.argtypes = [ip, fp] + [ctypes.c_int] * 4 + [ctypes.c_uint64, ctypes.c_int, ctypes.c_int, ip, fp]
    n.lib.dvg_debug_prim_tests.argtypes = [ctypes.c_void_p] + [ctypes.c_int] * 4 + [ctypes.c_uint64, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    # number of primitives = sum of segments over group shapes
    ns_, ng = int(topo[3]), int(topo[4])
    srec = topo[topo[10]:][:ns_ * 8].reshape(ns_, 8); gsh = topo[topo[13]:][:topo[9]]
    nprims = int(sum(srec[s, 5] if srec[s, 0] == 2 else 1 for s in gsh))
    for (x, y) in pixels:
        a = np.zeros((ns * ns, nprims), np.int32); pa = np.zeros((ns * ns, 2), np.float32)
        b = np.zeros_like(a); pb = np.zeros_like(pa)
        n.check(n.lib.dvg_debug_prim_tests(h, W, H, ns, ns, seed, x, y, a.ctypes.data, pa.ctypes.data, None))
        lib.emul_debug_prim_tests(topo.ctypes.data_as(ip), params.ctypes.data_as(fp), W, H, ns, ns, seed, x, y, b.ctypes.data_as(ip), pb.ctypes.data_as(fp))
        diff = np.argwhere((a & ~8) != b)
        print('%s px (%d,%d): pos equal %s, %d raw predicate differences' % (name, x, y, np.array_equal(pa, pb), len(diff)), flush=True)
        for s, e in diff[:10]:
            print('   sample %d pos %s prim %d: device %#x host %#x' % (s, pa[s], e, a[s, e], b[s, e]))
        # primitives whose predicate is non-trivial for some sample but which the tile bin does not hold
        live = ((b & 1) != 0) | (((b >> 8) & 0xff) != 0)
        missing = np.argwhere(live & ((a & 8) == 0))
        print('   live primitives missing from the bin: %d' % len(missing), [(int(s), int(e), hex(int(b[s, e]))) for s, e in missing[:10]])
    n.lib.dvg_scene_destroy(h)

run('tiger', 495, 510, 4, 0, [(87, 226), (232, 301), (207, 405), (100, 100)])
run('flower', 512, 554, 2, 1, [(90, 417), (281, 490), (100, 100)])
