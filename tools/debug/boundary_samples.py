"""Ad-hoc: per-boundary-sample comparison GPU vs host emulation."""
import sys, os, warnings, ctypes
import numpy as np, torch
warnings.simplefilter('ignore')
R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [R, os.path.join(R, 'oracle'), os.path.join(R, 'tests')]
import scenes, util, emul
from diffvg_b200 import _native as n
topo, params = util.pack(scenes.painterly())
W = H = 512
got = util.gpu_render(topo, params, W, H, 4, 4, 0)['image']
target = torch.rand(512, 512, 4, generator=torch.Generator().manual_seed(1234)).numpy()
d_img = (2.0 * (got - target) / got.size).astype(np.float32)
N = W * H * 16
dbg = torch.zeros(N, 4, device='cuda')
n.lib.dvg_debug_set_boundary_dump.argtypes = [ctypes.c_void_p]
n.lib.dvg_debug_set_boundary_dump(dbg.data_ptr())
b = util.gpu_render(topo, params, W, H, 4, 4, 0, d_render_image=d_img)['d_params']
n.lib.dvg_debug_set_boundary_dump(None)
g = dbg.cpu().numpy()
lib = emul._load()
cpu = np.zeros((N, 4), np.float32)
lib.emul_set_boundary_dump.argtypes = [ctypes.c_void_p]
lib.emul_set_boundary_dump(cpu.ctypes.data)
c = emul.render(topo, params, W, H, 4, 4, 0, d_render_image=d_img, nthreads=16)['d_params']
lib.emul_set_boundary_dump(None)
print('rel gpu vs emul', util.rel_l2(c, b))
d = np.abs(g - cpu)
bad = np.nonzero((d[:, 1] != 0) | (d[:, 0] > 1e-4 * np.abs(cpu[:, 0]) + 1e-9))[0]
print('samples differing:', len(bad))
for i in bad[:20]:
    print(i, 'gpu', g[i], 'cpu', cpu[i])
w = np.argsort(-np.abs(cpu[:, 0]))[:5]
print('largest contribs', w, cpu[w, 0], g[w, 0])
