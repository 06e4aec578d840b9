#!/usr/bin/env python
"""How even the cost-balanced row bands are (one GPU): device time of every band of an 8-way split through
dvg_render_*_rows, equal-height bands against sharded.balanced_row_partition, for the strong-scaling workloads.

    python tools/band_balance.py [per_tile]"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import scenes  # noqa: E402
import util  # noqa: E402
from diffvg_b200 import _native as n, sharded  # noqa: E402

W = H = 2048
WORLD = 8


def main():
    import warnings
    warnings.simplefilter('ignore')
    per_tile = float(sys.argv[1]) if len(sys.argv) > 1 else 4.0
    flower = np.load(os.path.join(ROOT, 'tests', 'golden_svg', 'flower.npz'))
    ptopo, pparams = util.pack(scenes.painterly())
    for label, topo, params, pf in (('C4 flower prefilter 2x2', flower['topo'], flower['params'], 1),
                                    ('painterly sampled 2x2', ptopo, pparams, 0)):
        h = ctypes.c_void_p()
        topo = np.ascontiguousarray(topo, np.int32)
        n.check(n.lib.dvg_scene_create(topo.ctypes.data, topo.shape[0], 0, ctypes.byref(h)))
        p = torch.from_numpy(np.ascontiguousarray(params, np.float32)).cuda()
        g = torch.empty_like(p)
        img = torch.zeros(H, W, 4, device='cuda')
        dimg = torch.rand(H, W, 4, device='cuda') - 0.5
        st = torch.cuda.current_stream().cuda_stream
        n.check(n.lib.dvg_scene_set_params(h, p.data_ptr(), p.numel(), 1, st))
        costs = np.zeros(H + 1, np.float32)
        th = ctypes.c_int()
        n.check(n.lib.dvg_scene_row_costs(h, W, H, 2, 2, pf, costs.ctypes.data, H + 1, ctypes.byref(th), st))
        units = (H + th.value - 1) // th.value

        def band_ms(r0, r1):
            def step(seed):
                n.check(n.lib.dvg_scene_set_params(h, p.data_ptr(), p.numel(), 1, st))
                n.check(n.lib.dvg_render_forward_rows(h, None, img.data_ptr(), W, H, 2, 2, seed, pf, r0, r1, st))
                n.check(n.lib.dvg_render_backward_rows(h, None, dimg.data_ptr(), W, H, 2, 2, seed, pf, r0, r1, g.data_ptr(), None, 1, st))
            for i in range(2):
                step(i)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for i in range(4):
                step(2 + i)
            b.record()
            torch.cuda.synchronize()
            return a.elapsed_time(b) / 4
        whole = band_ms(0, H)
        for name, bands in (('equal rows', sharded.row_partition(H, WORLD, th.value)),
                            ('balanced', sharded.balanced_row_partition(costs[:units], H, WORLD, th.value, per_unit=per_tile * ((W + 7) // 8)))):
            ms = [band_ms(b, e) for b, e in bands]
            print('%-26s %-11s whole %.2f ms; bands max %.3f mean %.3f -> %.1f%% of ideal;  %s' % (
                label, name, whole, max(ms), sum(ms) / len(ms), 100 * whole / WORLD / max(ms), ' '.join('%.2f' % m for m in ms)))
        bands = sharded.balanced_row_partition(costs[:units], H, WORLD, th.value, per_unit=per_tile * ((W + 7) // 8))
        for rnd in range(3):     # what bench.py does during warm-up: re-cut from the measured per-band times
            ms = [band_ms(b, e) for b, e in bands]
            bands = sharded.rebalance_bands(bands, ms, H, th.value)
            ms = [band_ms(b, e) for b, e in bands]
            print('%-26s %-11s whole %.2f ms; bands max %.3f mean %.3f -> %.1f%% of ideal;  %s' % (
                label, 'rebalanced%d' % (rnd + 1), whole, max(ms), sum(ms) / len(ms), 100 * whole / WORLD / max(ms), ' '.join('%.2f' % m for m in ms)))
        n.lib.dvg_scene_destroy(h)


if __name__ == '__main__':
    main()
