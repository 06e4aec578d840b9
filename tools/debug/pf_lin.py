import os, sys, warnings, time
import numpy as np, torch
warnings.simplefilter('ignore')
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')):
    sys.path.insert(0, p)
import util, ref_oracle
g = np.load(os.path.join(ROOT, 'tests', 'golden_svg', 'flower.npz'))
topo, params = g['topo'], g['params']
W = H = 2048
def rl(a, b):
    a = a.astype(np.float64).copy(); b = b.astype(np.float64).copy(); a[-1] = b[-1] = 0
    return np.linalg.norm(a - b) / np.linalg.norm(a)
for scale in (1.0 / (W * H * 4), 1.0):
    d_img = (np.random.RandomState(6).rand(H, W, 4).astype(np.float32) - 0.5) * np.float32(scale)
    g1 = util.gpu_render(topo, params, W, H, 2, 2, 0, use_prefiltering=True, d_render_image=d_img)['d_params']
    g1b = util.gpu_render(topo, params, W, H, 2, 2, 0, use_prefiltering=True, d_render_image=d_img)['d_params']
    g2 = util.gpu_render(topo, params, W, H, 2, 2, 0, use_prefiltering=True, d_render_image=2 * d_img)['d_params']
    print('scale %g: run-to-run %.3g, linearity %.3g, |g| %.4g, radius entry %g %g %g' % (scale, rl(g1, g1b), rl(2 * g1, g2), np.linalg.norm(g1[:-1]), g1[-1], g1b[-1], g2[-1]), flush=True)
    w = np.argsort(-np.abs(g1[:-1].astype(np.float64) - g1b[:-1]))[:5]
    print('   worst entries', [(int(i), float(g1[i]), float(g1b[i])) for i in w])
t = time.time(); rb = ref_oracle.render(topo, params, W, H, 2, 2, 0, use_prefiltering=True, d_render_image=d_img)['d_params']; print('ref bwd %.1fs' % (time.time() - t))
print('ref vs gpu rel-L2 %.3g (radius ref %g gpu %g)' % (rl(rb, g1), rb[-1], g1[-1]))
rb2 = ref_oracle.render(topo, params, W, H, 2, 2, 0, use_prefiltering=True, d_render_image=d_img)['d_params']
print('ref run-to-run %.3g' % rl(rb, rb2))
