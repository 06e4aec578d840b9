"""Generates tests/golden_svg/{tiger,flower}.npz: the scene PACKS of the two bundled assets BASELINE.json's
configs[1] / configs[3] name (apps/imgs/tiger.svg, apps/imgs/flower.svg) as produced by
diffvg_b200.pydiffvg.svg_to_scene, plus outputs of the UNMODIFIED reference renderer (oracle/_ref) on
them at a reduced size.  The assets themselves live only under /root/reference; the packs let the GPU
box (where /root/reference does not exist) render the real scenes at full size.  Run in the build
container only:

    make -C oracle ref && python tests/golden/make_svg_golden.py
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)

import ref_oracle  # noqa: E402
import util  # noqa: E402
from golden.make_golden import d_image_for  # noqa: E402

ASSETS = '/root/reference/apps/imgs'
OUT = os.path.join(os.path.dirname(HERE), 'golden_svg')
# name -> (W, H, nsx, nsy, seed, use_prefiltering) of the reduced-size reference outputs stored with the pack
CASES = {
    'tiger': (124, 128, 2, 2, 0, False),     # finite_difference_comp.py / render_svg.py asset, sampled
    'flower': (160, 173, 1, 1, 0, True),     # refine_svg.py asset, SDF prefiltering (finite_difference_comp.py: 1x1)
}


def load_pack(name):
    from diffvg_b200 import pydiffvg
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        scene = pydiffvg.svg_to_scene(os.path.join(ASSETS, name + '.svg'))
        return util.pack(scene)


def main():
    os.makedirs(OUT, exist_ok=True)
    for name, (W, H, nsx, nsy, seed, pf) in CASES.items():
        topo, params = load_pack(name)
        fwd = ref_oracle.render(topo, params, W, H, nsx, nsy, seed, use_prefiltering=pf)
        bwd = ref_oracle.render(topo, params, W, H, nsx, nsy, seed, use_prefiltering=pf, d_render_image=d_image_for(name, H, W))
        np.savez_compressed(os.path.join(OUT, name + '.npz'), topo=topo, params=params,
                            config=np.asarray([W, H, nsx, nsy, seed, 1 if pf else 0], np.int64),
                            image=fwd['image'], d_params=bwd['d_params'])
        print('%-7s shapes %d groups %d params %d | %dx%d image sum %.6f |d_params| %.6g' % (
            name, topo[3], topo[4], params.shape[0], W, H, fwd['image'].astype(np.float64).sum(),
            np.linalg.norm(bwd['d_params'].astype(np.float64))))


if __name__ == '__main__':
    main()
