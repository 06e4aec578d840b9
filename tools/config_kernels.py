#!/usr/bin/env python
"""Per-kernel CUDA-event times of one fwd+bwd step of a scene pack of tests/golden_svg:  python tools/config_kernels.py tiger|flower [spp] [pf]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tools'))
import measure_configs as mc  # noqa: E402
import numpy as np  # noqa: E402
import torch  # noqa: E402
from diffvg_b200 import _native as n  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else 'tiger'
ns = int(sys.argv[2]) if len(sys.argv) > 2 else 4
pf = int(sys.argv[3]) if len(sys.argv) > 3 else 0
a = np.load(os.path.join(ROOT, 'tests', 'golden_svg', name + '.npz'))
s = mc.Scene(a)
W, H = (int(a['topo'][1]), int(a['topo'][2])) if name == 'tiger' else (2048, 2048)
img = torch.empty(H, W, 4, device='cuda'); dimg = torch.empty_like(img)
for i in range(3):
    s.step(W, H, ns, ns, i, pf, img, dimg)
torch.cuda.synchronize()
n.profile_enable(True)
s.step(W, H, ns, ns, 5, pf, img, dimg); torch.cuda.synchronize(); n.profile_report()
for i in range(3):
    s.step(W, H, ns, ns, 6 + i, pf, img, dimg)
torch.cuda.synchronize()
rep = n.profile_report()
tot = sum(v[1] for v in rep.values()) / 3
print('%s %dx%d %dx%d spp pf=%d: %.3f ms in kernels' % (name, W, H, ns, ns, pf, tot))
for k, (c, ms) in sorted(rep.items(), key=lambda kv: -kv[1][1]):
    print('  %-34s %3d launches %8.3f ms' % (k, c // 3, ms / 3))
