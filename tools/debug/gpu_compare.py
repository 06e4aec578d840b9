"""Ad-hoc GPU comparison script (run under gpurun while developing)."""
import sys, time, os, warnings
import numpy as np, torch
warnings.simplefilter('ignore')
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), 'oracle'))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), 'tests'))
import scenes, util, ref_oracle

def fwd(name, scene, W, H, nsx, nsy, seed, ftype=0, frad=0.5, bg=None):
    topo, params = util.pack(scene, ftype, frad)
    t0 = time.time(); ref = ref_oracle.render(topo, params, W, H, nsx, nsy, seed, background=bg)['image']; t1 = time.time()
    got = util.gpu_render(topo, params, W, H, nsx, nsy, seed, background=bg)['image']; t2 = time.time()
    d = np.abs(ref - got)
    print('FWD %-16s ref %.2fs gpu %.2fs sum ref %.6f gpu %.6f maxabs %.3g bad px %d' % (
        name, t1 - t0, t2 - t1, ref.astype(np.float64).sum(), got.astype(np.float64).sum(), d.max(), (d.max(axis=2) > 1e-5).sum()), flush=True)

def bwd(name, scene, W, H, nsx, nsy, seed, ftype=0, frad=0.5, bg=None):
    topo, params = util.pack(scene, ftype, frad)
    d_img = (np.random.RandomState(1).rand(H, W, 4).astype(np.float32) - 0.5)
    t0 = time.time(); ref = ref_oracle.render(topo, params, W, H, nsx, nsy, seed, background=bg, d_render_image=d_img); t1 = time.time()
    got = util.gpu_render(topo, params, W, H, nsx, nsy, seed, background=bg, d_render_image=d_img); t2 = time.time()
    a, b = ref['d_params'].astype(np.float64), got['d_params'].astype(np.float64)
    w = np.argmax(np.abs(a - b))
    print('BWD %-16s ref %.2fs gpu %.2fs |g| %.4g relL2 %.3g maxabs %.3g (ref %.6g gpu %.6g @%d) nnz %d/%d' % (
        name, t1 - t0, t2 - t1, np.linalg.norm(a), util.rel_l2(a, b), np.abs(a - b).max(), a[w], b[w], w, (a != 0).sum(), (b != 0).sum()), flush=True)
    if bg is not None:
        print('    d_bg maxabs', np.abs(ref['d_background'] - got['d_background']).max())

print(torch.cuda.get_device_name(0), 'cpus', os.cpu_count(), flush=True)
bg = np.random.RandomState(0).rand(128, 128, 4).astype(np.float32)
fwd('circle', scenes.single_circle(), 256, 256, 2, 2, 0)
fwd('stroke', scenes.single_stroke(), 256, 256, 2, 2, 0)
fwd('stroke thick', scenes.single_stroke([10., 5., 4., 20.], fill=False), 256, 256, 2, 2, 0)
fwd('circle hann8', scenes.single_circle(), 256, 256, 2, 2, 0, ftype=3, frad=8.0)
fwd('zoo', scenes.zoo(), 128, 128, 2, 2, 3)
fwd('zoo 3x3 tent', scenes.zoo(), 160, 96, 3, 3, 5, ftype=1, frad=1.5)
fwd('zoo 1spp', scenes.zoo(), 128, 128, 1, 1, 3)
fwd('zoo bg', scenes.zoo(), 128, 128, 2, 2, 3, bg=bg)
fwd('painterly256', scenes.painterly(256, 256), 256, 256, 4, 4, 0)
fwd('blobs128', scenes.blobs(128, 256), 256, 256, 2, 2, 0)
fwd('painterly C3', scenes.painterly(), 512, 512, 4, 4, 0)
bwd('circle', scenes.single_circle(), 256, 256, 2, 2, 0)
bwd('stroke', scenes.single_stroke(), 256, 256, 2, 2, 0)
bwd('stroke thick', scenes.single_stroke([10., 5., 4., 20.], fill=False), 256, 256, 2, 2, 0)
bwd('zoo', scenes.zoo(), 128, 128, 2, 2, 3)
bwd('zoo bg 1spp', scenes.zoo(), 128, 128, 1, 1, 3, bg=bg)
bwd('zoo tent', scenes.zoo(), 128, 128, 2, 2, 5, ftype=1, frad=1.5)
bwd('painterly256', scenes.painterly(256, 256), 256, 256, 4, 4, 0)
bwd('blobs128', scenes.blobs(128, 256), 256, 256, 2, 2, 0)
bwd('painterly C3', scenes.painterly(), 512, 512, 4, 4, 0)

# timing of the C3 config through the pydiffvg API
from diffvg_b200 import pydiffvg
cw, ch, shapes, groups = scenes.painterly()
for p in shapes:
    p.points.requires_grad_(True); p.stroke_width.requires_grad_(True)
for g in groups:
    g.stroke_color.requires_grad_(True)
target = torch.rand(512, 512, 4, generator=torch.Generator().manual_seed(1234)).cuda()
for it in range(6):
    torch.cuda.synchronize(); t0 = time.time()
    args = pydiffvg.RenderFunction.serialize_scene(cw, ch, shapes, groups)
    t1 = time.time()
    img = pydiffvg.RenderFunction.apply(512, 512, 4, 4, it, None, *args)
    torch.cuda.synchronize(); t2 = time.time()
    loss = (img - target).pow(2).mean()
    loss.backward()
    torch.cuda.synchronize(); t3 = time.time()
    print('iter %d: serialize %.1f ms fwd %.1f ms bwd %.1f ms loss %.6f' % (it, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, loss.item()), flush=True)
