"""Ad-hoc: where do the GPU and the reference differ on the asset packs? (run under gpurun; writes gpurun_out/asset_diff_*.npz)"""
import os, sys, time, warnings
import numpy as np
warnings.simplefilter('ignore')
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')):
    sys.path.insert(0, p)
import ref_oracle, util

def run(name, W, H, ns, seed, pf, tag):
    g = np.load(os.path.join(ROOT, 'tests', 'golden_svg', name + '.npz'))
    topo, params = g['topo'], g['params']
    ref = ref_oracle.render(topo, params, W, H, ns, ns, seed, use_prefiltering=pf)['image']
    got = util.gpu_render(topo, params, W, H, ns, ns, seed, use_prefiltering=pf)['image']
    d = np.abs(ref - got).max(axis=2)
    ys, xs = np.nonzero(d > 1e-5)
    print('%s %s %dx%d spp %d pf %d: %d px differ, max %g' % (tag, name, W, H, ns * ns, pf, len(ys), d.max()), flush=True)
    for y, x in list(zip(ys.tolist(), xs.tolist()))[:12]:
        print('   px (x=%d, y=%d) ref %s got %s' % (x, y, ref[y, x], got[y, x]))
    np.savez_compressed(os.path.join(ROOT, 'gpurun_out', 'asset_diff_%s_%s.npz' % (tag, name)), ys=ys, xs=xs, ref=ref[ys, xs], got=got[ys, xs])

tag = sys.argv[1] if len(sys.argv) > 1 else 'default'
run('tiger', 495, 510, 4, 0, False, tag)
run('flower', 512, 554, 2, 1, False, tag)
run('flower', 1024, 1024, 1, 0, True, tag)
