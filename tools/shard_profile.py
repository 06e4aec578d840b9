#!/usr/bin/env python
"""Where a row shard's time goes (one GPU): per-kernel CUDA-event times of the whole 2048^2 render against ONE band of
1/8 of its rows through dvg_render_*_rows, for the two strong-scaling workloads of bench.py.  What does not shrink with
the band is what every rank of an 8-GPU run repeats.

    python tools/shard_profile.py"""
import ctypes
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import scenes  # noqa: E402
import util  # noqa: E402
from diffvg_b200 import _native as n  # noqa: E402

W = H = 2048


def run(label, topo, params, pf, r0, r1):
    h = ctypes.c_void_p()
    topo = np.ascontiguousarray(topo, np.int32)
    n.check(n.lib.dvg_scene_create(topo.ctypes.data, topo.shape[0], 0, ctypes.byref(h)))
    p = torch.from_numpy(np.ascontiguousarray(params, np.float32)).cuda()
    g = torch.empty_like(p)
    img = torch.zeros(H, W, 4, device='cuda')
    dimg = torch.rand(H, W, 4, device='cuda') - 0.5
    st = torch.cuda.current_stream().cuda_stream

    def step(seed):
        n.check(n.lib.dvg_scene_set_params(h, p.data_ptr(), p.numel(), 1, st))
        n.check(n.lib.dvg_render_forward_rows(h, None, img.data_ptr(), W, H, 2, 2, seed, pf, r0, r1, st))
        n.check(n.lib.dvg_render_backward_rows(h, None, dimg.data_ptr(), W, H, 2, 2, seed, pf, r0, r1, g.data_ptr(), None, 1, st))
    for i in range(3):
        step(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record()
    for i in range(5):
        step(3 + i)
    b.record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / 5 * 1e3
    ms = a.elapsed_time(b) / 5
    n.profile_enable(True)
    step(9)
    torch.cuda.synchronize()
    n.profile_report()
    for i in range(3):
        step(10 + i)
    torch.cuda.synchronize()
    rep = n.profile_report()
    n.profile_enable(False)
    n.lib.dvg_scene_destroy(h)
    print('%-34s rows [%4d, %4d)  %7.3f ms device  %7.3f ms wall' % (label, r0, r1, ms, wall))
    return {k: v[1] / 3 for k, v in rep.items()}


def main():
    import warnings
    warnings.simplefilter('ignore')
    flower = np.load(os.path.join(ROOT, 'tests', 'golden_svg', 'flower.npz'))
    ptopo, pparams = util.pack(scenes.painterly())
    for label, topo, params, pf in (('C4 flower prefilter 2x2', flower['topo'], flower['params'], 1),
                                    ('painterly sampled 2x2', ptopo, pparams, 0)):
        full = run(label, topo, params, pf, 0, H)
        band = run(label, topo, params, pf, 3 * H // 8, 4 * H // 8)
        print('   %-30s %9s %9s %7s' % ('kernel', 'full ms', 'band ms', 'band/full'))
        for k in sorted(full, key=lambda k: -band.get(k, 0.0)):
            print('   %-30s %9.3f %9.3f %7.2f' % (k, full[k], band.get(k, 0.0), band.get(k, 0.0) / max(full[k], 1e-9)))
        print('   %-30s %9.3f %9.3f' % ('sum of kernels', sum(full.values()), sum(band.values())))


if __name__ == '__main__':
    main()
