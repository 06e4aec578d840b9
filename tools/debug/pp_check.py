import os, sys, warnings
import numpy as np, torch
warnings.simplefilter('ignore')
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')):
    sys.path.insert(0, p)
import scenes, util, ref_oracle
from diffvg_b200 import pydiffvg
pydiffvg.set_use_gpu(True); pydiffvg.set_device(torch.device('cuda', 0))
target = torch.rand(128, 128, 4, generator=torch.Generator().manual_seed(3)).cuda()
cw, ch, shapes, groups = scenes.painterly(96, 128)
topo, params = util.pack((cw, ch, shapes, groups))
ref = ref_oracle.render(topo, params, 128, 128, 2, 2, 5)['image']
def cmp(tag, img):
    d = np.abs(img.detach().cpu().numpy() - ref)
    print(tag, 'vs oracle: max', d.max(), 'bad px', int((d.max(axis=2) > 1e-5).sum()), flush=True)
leaves = [s.points.requires_grad_(True) for s in shapes] + [s.stroke_width.requires_grad_(True) for s in shapes] + [g.stroke_color.requires_grad_(True) for g in groups]
args = pydiffvg.RenderFunction.serialize_scene(cw, ch, shapes, groups)
img0 = pydiffvg.RenderFunction.apply(128, 128, 2, 2, 5, None, *args); cmp('stock fwd #1', img0)
img0b = pydiffvg.RenderFunction.apply(128, 128, 2, 2, 5, None, *args); cmp('stock fwd #2', img0b)
(img0 - target).pow(2).mean().backward()
cmp('stock fwd #1 after backward', img0)
img0c = pydiffvg.RenderFunction.apply(128, 128, 2, 2, 5, None, *args); cmp('stock fwd #3 (after a backward)', img0c)
for t in leaves: t.requires_grad_(False)
pp = pydiffvg.PackedParams(cw, ch, shapes, groups, device=torch.device('cpu'))
img1 = pydiffvg.RenderFunction.apply(128, 128, 2, 2, 5, None, *pp.scene_args()); cmp('packed fwd', img1)
a = pp.scene_args()
print('params equal', np.array_equal(a[1].detach().numpy(), params), 'topo equal', np.array_equal(a[0].topo, topo))
