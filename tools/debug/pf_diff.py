"""words vs inline prefilter renders: where do they differ?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, 'tests'), os.path.join(ROOT, 'oracle')):
    sys.path.insert(0, p)
import warnings; warnings.simplefilter('ignore')
import numpy as np
import scenes, util
from diffvg_b200 import _native as n
for name, scene in (('zoo', scenes.zoo()), ('blobs', scenes.blobs())):
    topo, params = util.pack(scene)
    for (W, H, nsx, nsy) in ((96, 96, 2, 2), (70, 50, 3, 1), (70, 50, 1, 1), (96, 96, 3, 1), (96, 96, 1, 3), (64, 64, 4, 4), (70, 50, 2, 2), (96, 96, 3, 3)):
        a = util.gpu_render(topo, params, W, H, nsx, nsy, 0, use_prefiltering=True)['image']
        n.lib.dvg_debug_set_prefilter_inline(1)
        b = util.gpu_render(topo, params, W, H, nsx, nsy, 0, use_prefiltering=True)['image']
        n.lib.dvg_debug_set_prefilter_inline(0)
        d = np.abs(a - b).max(axis=2)
        ys, xs = np.nonzero(d)
        print(name, W, H, nsx, nsy, 'differing pixels', len(ys), 'max', d.max(), list(zip(xs[:6], ys[:6])), flush=True)
