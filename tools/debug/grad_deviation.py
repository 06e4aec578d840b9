"""Ad-hoc: where does the C3 gradient deviate from the oracle?"""
import sys, os, warnings
import numpy as np, torch
warnings.simplefilter('ignore')
R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [R, os.path.join(R, 'oracle'), os.path.join(R, 'tests')]
import scenes, util, ref_oracle
topo, params = util.pack(scenes.painterly())
got = util.gpu_render(topo, params, 512, 512, 4, 4, 0)['image']
target = torch.rand(512, 512, 4, generator=torch.Generator().manual_seed(1234)).numpy()
d_img = (2.0 * (got - target) / got.size).astype(np.float32)
a = ref_oracle.render(topo, params, 512, 512, 4, 4, 0, d_render_image=d_img)['d_params'].astype(np.float64)
for rep in range(2):
    b = util.gpu_render(topo, params, 512, 512, 4, 4, 0, d_render_image=d_img)['d_params'].astype(np.float64)
    print('all   rel', util.rel_l2(a, b))
    print('no-xf rel', util.rel_l2(a[:-9], b[:-9]))
    print('xf ref', a[-9:]); print('xf gpu', b[-9:])
    d = np.abs(a - b)[:-9]; w = np.argsort(-d)[:8]
    print(w, a[w], b[w])
# float64 reference of the matrix gradient: sum of the per-group contributions is not available, so use
# two different d_img scalings to see whether the reference or the GPU is the noisy one
for scale in (1.0, 1024.0):
    a2 = ref_oracle.render(topo, params, 512, 512, 4, 4, 0, d_render_image=(d_img * scale).astype(np.float32))['d_params'].astype(np.float64) / scale
    b2 = util.gpu_render(topo, params, 512, 512, 4, 4, 0, d_render_image=(d_img * scale).astype(np.float32))['d_params'].astype(np.float64) / scale
    print('scale', scale, 'ref xf', a2[-9:][[0,2,6,8]], 'gpu xf', b2[-9:][[0,2,6,8]])
