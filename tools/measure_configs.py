#!/usr/bin/env python
"""Device-resident fwd+bwd times of the other BASELINE.json configurations (the bench line is the painterly one):

    python tools/measure_configs.py [--cpu]      # --cpu also times the reference CPU path on a bounded sample

  C1  single circle 256^2, 2x2 spp              (launch-latency floor)
  C2  tiger.svg at its own size, 4x4 spp         (scene pack of tests/golden_svg)
  C2' 1024 closed cubic blobs 512^2, 4x4 spp    (synthetic fill-heavy scene)
  C4  flower.svg 2048^2, 2x2 spp, prefiltered    (scene pack of tests/golden_svg; no boundary pass)
  C4' the blobs at 2048^2, 2x2 spp, prefiltered
  C5  512 scenes x 16 strokes 64^2, 2x2 spp      (512 native scenes in a loop, then ONE batch scene)
Prints one line per config: ms per fwd+bwd (CUDA events, median of 5 after 2 warm-ups)."""
import ctypes
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests'), os.path.join(ROOT, 'oracle')):
    if p not in sys.path:
        sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import scenes  # noqa: E402
import util  # noqa: E402
from diffvg_b200 import _native as n  # noqa: E402


class Scene:
    def __init__(self, scene):
        if isinstance(scene, tuple):
            self.topo, self.params = util.pack(scene)
        else:   # a scene pack of tests/golden_svg (topology + parameters of a parsed SVG asset)
            self.topo, self.params = np.ascontiguousarray(scene['topo'], np.int32), np.ascontiguousarray(scene['params'], np.float32)
        self.h = ctypes.c_void_p()
        n.check(n.lib.dvg_scene_create(self.topo.ctypes.data, self.topo.shape[0], 0, ctypes.byref(self.h)))
        self.p = torch.from_numpy(self.params).cuda()
        self.g = torch.empty_like(self.p)

    def step(self, W, H, nsx, nsy, seed, pf, img, dimg):
        st = torch.cuda.current_stream().cuda_stream
        n.check(n.lib.dvg_scene_set_params(self.h, self.p.data_ptr(), self.p.numel(), 1, st))
        n.check(n.lib.dvg_render_forward(self.h, None, img.data_ptr(), None, W, H, nsx, nsy, seed, pf, None, 0, st))
        torch.mul(img, 2.0 / img.numel(), out=dimg)
        n.check(n.lib.dvg_render_backward(self.h, None, dimg.data_ptr(), None, W, H, nsx, nsy, seed, pf, None, 0,
                                          self.g.data_ptr(), None, None, 0, st))


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    out = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        out.append(a.elapsed_time(b))
    return float(np.median(out))


def main():
    cpu = '--cpu' in sys.argv
    import warnings
    warnings.simplefilter('ignore')
    rows = []
    assets = {k: np.load(os.path.join(ROOT, 'tests', 'golden_svg', k + '.npz')) for k in ('tiger', 'flower')}
    tiger_wh = (int(assets['tiger']['topo'][1]), int(assets['tiger']['topo'][2]))
    c5_only = '--c5-only' in sys.argv    # only the C5 batch scene (for an ncu capture of its kernels)
    for name, scene, (W, H, nsx, nsy, pf) in () if c5_only else (
            ('C1 single_circle 256^2 2x2', scenes.single_circle(), (256, 256, 2, 2, 0)),
            ('C2 tiger.svg %dx%d 4x4' % tiger_wh, assets['tiger'], tiger_wh + (4, 4, 0)),
            ("C2' blobs1024 512^2 4x4", scenes.blobs(), (512, 512, 4, 4, 0)),
            ('C4 flower.svg 2048^2 2x2 prefilter', assets['flower'], (2048, 2048, 2, 2, 1)),
            ("C4' blobs1024 2048^2 2x2 prefilter", scenes.blobs(), (2048, 2048, 2, 2, 1))):
        s = Scene(scene)
        img = torch.empty(H, W, 4, device='cuda'); dimg = torch.empty_like(img)
        ms = timed(lambda: s.step(W, H, nsx, nsy, 0, pf, img, dimg))
        line = '%-38s %8.3f ms fwd+bwd  %8.1f Msamples/s' % (name, ms, W * H * nsx * nsy * (1 if pf else 2) / ms / 1e3)
        if cpu:
            import oracle_check
            rows_s = min(H, 256)
            t0 = time.perf_counter()
            ref = oracle_check.render(s.topo, s.params, W, rows_s, nsx, nsy, 0, use_prefiltering=bool(pf))['image']
            oracle_check.render(s.topo, s.params, W, rows_s, nsx, nsy, 0, use_prefiltering=bool(pf),
                                d_render_image=(2.0 * ref / ref.size).astype(np.float32))
            t = (time.perf_counter() - t0) * (H / rows_s)
            line += '   reference CPU %.2f s (from a %d-row sample, %d threads) -> %.0fx' % (t, rows_s, os.cpu_count(), t * 1e3 / ms)
        print(line, flush=True)
        n.lib.dvg_scene_destroy(s.h)
    batch = [Scene(scenes.batched_strokes(b)) for b in range(512)]
    img = torch.empty(64, 64, 4, device='cuda'); dimg = torch.empty_like(img)
    if not c5_only:
        ms = timed(lambda: [s.step(64, 64, 2, 2, b, 0, img, dimg) for b, s in enumerate(batch)], reps=3, warm=1)
        print('%-38s %8.3f ms fwd+bwd for 512 scenes (%.3f ms per scene, sequential native scenes)' % ('C5 512x16 strokes 64^2 2x2', ms, ms / 512), flush=True)
    # the same 512 scenes as ONE batch scene (dvg_scene_create_batch)
    B = len(batch)
    rows = torch.stack([s.p for s in batch]).contiguous()
    for s in batch:
        n.lib.dvg_scene_destroy(s.h)
    h = ctypes.c_void_p()
    n.check(n.lib.dvg_scene_create_batch(batch[0].topo.ctypes.data, batch[0].topo.shape[0], 0, B, ctypes.byref(h)))
    seeds = np.arange(B, dtype=np.uint64)
    imgs = torch.empty(B, 64, 64, 4, device='cuda'); dimgs = torch.empty_like(imgs)
    grads = torch.empty_like(rows)

    def bstep():
        st = torch.cuda.current_stream().cuda_stream
        n.check(n.lib.dvg_scene_set_params(h, rows.data_ptr(), rows.numel(), 1, st))
        n.check(n.lib.dvg_render_forward_batch(h, None, imgs.data_ptr(), 64, 64, 2, 2, seeds.ctypes.data, st))
        torch.mul(imgs, 2.0 / imgs[0].numel(), out=dimgs)
        n.check(n.lib.dvg_render_backward_batch(h, None, dimgs.data_ptr(), 64, 64, 2, 2, seeds.ctypes.data, grads.data_ptr(), None, 1, st))
    ms = timed(bstep, reps=5, warm=2)
    print('%-38s %8.3f ms fwd+bwd for 512 scenes (%.4f ms per scene, one batch scene)' % ('C5 512x16 strokes 64^2 2x2 BATCH', ms, ms / 512), flush=True)
    n.profile_enable(True)
    bstep(); torch.cuda.synchronize()
    n.profile_report()
    bstep(); torch.cuda.synchronize()
    rep = n.profile_report()
    n.profile_enable(False)
    print('   per kernel (ms): ' + ', '.join('%s %.3f' % (k.replace('k_wave_', 'w_').replace('k_', ''), v[1]) for k, v in sorted(rep.items(), key=lambda kv: -kv[1][1])[:12]))
    n.lib.dvg_scene_destroy(h)


if __name__ == '__main__':
    main()
