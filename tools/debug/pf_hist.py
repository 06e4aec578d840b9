import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, 'tests'), os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tools')):
    sys.path.insert(0, p)
import warnings; warnings.simplefilter('ignore')
import numpy as np, torch
import scenes, util
import measure_configs as mc
from diffvg_b200 import _native as n
n.lib.dvg_debug_set_boundary_dump.argtypes = [ctypes.c_void_p]
assets = {k: np.load(os.path.join(ROOT, 'tests', 'golden_svg', k + '.npz')) for k in ('tiger', 'flower')}
for name, scene, (W, H, ns) in (('flower', assets['flower'], (2048, 2048, 2)), ('blobs', scenes.blobs(), (2048, 2048, 2)), ('tiger', assets['tiger'], (495, 510, 2)),
                                 ('painterly', scenes.painterly(), (512, 512, 2))):
    s = mc.Scene(scene)
    hist = torch.zeros(64, device='cuda')
    img = torch.empty(H, W, 4, device='cuda')
    st = torch.cuda.current_stream().cuda_stream
    n.check(n.lib.dvg_scene_set_params(s.h, s.p.data_ptr(), s.p.numel(), 1, st))
    n.lib.dvg_debug_set_boundary_dump(hist.data_ptr())
    n.check(n.lib.dvg_render_forward(s.h, None, img.data_ptr(), None, W, H, ns, ns, 0, 1, None, 0, st))
    torch.cuda.synchronize()
    n.lib.dvg_debug_set_boundary_dump(None)
    h = hist.cpu().numpy(); tot = h.sum()
    cum = np.cumsum(h) / tot
    print(name, 'samples', int(tot), 'mean frags %.2f' % ((h * np.arange(64)).sum() / tot), 'cum<=k:', ' '.join('%d:%.4f' % (k, cum[k]) for k in (0, 1, 2, 3, 4, 6, 8, 12, 16, 24, 32)), flush=True)
