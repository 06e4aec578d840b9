#!/usr/bin/env python
"""A few forward+backward steps of BASELINE C4 (flower.svg, 2048x2048, use_prefiltering, 2x2 spp) through the C ABI: the
command ncu is pointed at for the prefiltered path (tools/gpu_prof.sh style), and a plain timing when run alone."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from diffvg_b200 import _native as n  # noqa: E402

W = H = 2048
flower = np.load(os.path.join(ROOT, 'tests', 'golden_svg', 'flower.npz'))
topo = np.ascontiguousarray(flower['topo'], np.int32)
h = ctypes.c_void_p()
n.check(n.lib.dvg_scene_create(topo.ctypes.data, topo.shape[0], 0, ctypes.byref(h)))
p = torch.from_numpy(np.ascontiguousarray(flower['params'], np.float32)).cuda()
g = torch.empty_like(p)
img = torch.zeros(H, W, 4, device='cuda')
dimg = torch.rand(H, W, 4, device='cuda') - 0.5
st = torch.cuda.current_stream().cuda_stream
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
for i in range(steps):
    ev[i].record()
    n.check(n.lib.dvg_scene_set_params(h, p.data_ptr(), p.numel(), 1, st))
    n.check(n.lib.dvg_render_forward(h, None, img.data_ptr(), None, W, H, 2, 2, i, 1, None, 0, st))
    n.check(n.lib.dvg_render_backward(h, None, dimg.data_ptr(), None, W, H, 2, 2, i, 1, None, 0, g.data_ptr(), None, None, 1, st))
ev[steps].record()
torch.cuda.synchronize()
print('C4 flower 2048^2 prefilter 2x2: ' + ' '.join('%.2f' % ev[i].elapsed_time(ev[i + 1]) for i in range(steps)) + ' ms per fwd+bwd')
