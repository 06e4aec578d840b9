#!/usr/bin/env python
"""The reference's gradient-validation recipe (apps/finite_difference_comp.py:79-125, 173-187) as a
function with numbers out: translate every shape by +-epsilon along one axis, central difference of the
two renders summed over the channels, against `RenderFunction.render_grad(ones)` (the per-pixel
translation gradient image).  The reference writes both as colour-mapped PNGs and leaves the comparison
to the eye; here the two fields are also compared numerically (correlation / rel-L2 of block sums).

    python tools/finite_difference_comp.py some.svg [--size_scale 1.0] [--num_spp 4] [--use_prefiltering] [--out DIR]
"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def perturb_scene(pydiffvg, shapes, shape_groups, axis, epsilon):
    """finite_difference_comp.py:80-96, in place."""
    for s in shapes:
        if isinstance(s, (pydiffvg.Circle, pydiffvg.Ellipse)):
            s.center[axis] += epsilon
        elif isinstance(s, (pydiffvg.Path, pydiffvg.Polygon)):
            s.points[:, axis] += epsilon
        elif isinstance(s, pydiffvg.Rect):
            s.p_min[axis] += epsilon
            s.p_max[axis] += epsilon
    for g in shape_groups:
        if isinstance(g.fill_color, pydiffvg.LinearGradient):
            g.fill_color.begin[axis] += epsilon
            g.fill_color.end[axis] += epsilon


def finite_difference_comp(scene, w, h, num_spp=4, use_prefiltering=False, epsilon=0.1, seed=0):
    """-> dict(fd=[H, W, 2], grad=[H, W, 2]): central differences and render_grad(ones), as numpy arrays."""
    from diffvg_b200 import pydiffvg
    canvas_width, canvas_height, shapes, shape_groups = scene
    nsx = nsy = 1 if use_prefiltering else num_spp

    def render():
        args = pydiffvg.RenderFunction.serialize_scene(canvas_width, canvas_height, shapes, shape_groups,
                                                       use_prefiltering=use_prefiltering)
        return pydiffvg.RenderFunction.apply(w, h, nsx, nsy, seed, None, *args), args

    def moved_tensors():
        out = []
        for s in shapes:
            out += [getattr(s, a) for a in ('center', 'points', 'p_min', 'p_max') if hasattr(s, a)]
        for g in shape_groups:
            if isinstance(g.fill_color, pydiffvg.LinearGradient):
                out += [g.fill_color.begin, g.fill_color.end]
        return out

    fd = []
    with torch.no_grad():
        saved = [t.clone() for t in moved_tensors()]
        for axis in (0, 1):
            perturb_scene(pydiffvg, shapes, shape_groups, axis, epsilon)
            img0, _ = render()
            perturb_scene(pydiffvg, shapes, shape_groups, axis, -2 * epsilon)
            img1, _ = render()
            for t, t0 in zip(moved_tensors(), saved):   # the reference adds +epsilon back, which leaves the scene an ulp off
                t.copy_(t0)
            fd.append(((img0 - img1) / (2 * epsilon)).sum(dim=2))
        _, args = render()
        grad = pydiffvg.RenderFunction.render_grad(torch.ones(h, w, 4, device=pydiffvg.get_device()), w, h, nsx, nsy, seed, None, *args)
    return dict(fd=torch.stack(fd, dim=2).cpu().numpy(), grad=grad.cpu().numpy())


def block_agreement(fd, grad, block=8):
    """Per axis: (correlation, rel-L2) between the two fields after summing over block x block pixels."""
    out = []
    H, W = fd.shape[:2]
    h, w = (H // block) * block, (W // block) * block
    for axis in (0, 1):
        f = fd[:h, :w, axis].reshape(h // block, block, w // block, block).sum(axis=(1, 3)).astype(np.float64)
        a = grad[:h, :w, axis].reshape(h // block, block, w // block, block).sum(axis=(1, 3)).astype(np.float64)
        out.append((float(np.corrcoef(f.ravel(), a.ravel())[0, 1]), float(np.linalg.norm(f - a) / max(np.linalg.norm(f), 1e-30))))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('svg_file')
    ap.add_argument('--size_scale', type=float, default=1.0)
    ap.add_argument('--clamping_factor', type=float, default=0.1)
    ap.add_argument('--num_spp', type=int, default=4)
    ap.add_argument('--use_prefiltering', action='store_true')
    ap.add_argument('--out', default='results/finite_difference_comp')
    args = ap.parse_args()
    from diffvg_b200 import pydiffvg
    scene = pydiffvg.svg_to_scene(args.svg_file)
    w, h = int(scene[0] * args.size_scale), int(scene[1] * args.size_scale)
    r = finite_difference_comp(scene, w, h, args.num_spp, args.use_prefiltering)

    def normalize(x, lo, hi):
        rng = max(abs(lo), abs(hi), 1e-30)
        return (x + rng) / (2 * rng)

    for axis, name in ((0, 'x'), (1, 'y')):
        lo, hi = r['fd'][:, :, axis].min() * args.clamping_factor, r['fd'][:, :, axis].max() * args.clamping_factor
        pydiffvg.imwrite(normalize(r['fd'][:, :, axis], lo, hi), os.path.join(args.out, 'finite_%s_diff.png' % name), gamma=1.0)
        pydiffvg.imwrite(normalize(r['grad'][:, :, axis], lo, hi), os.path.join(args.out, 'ours_%s_diff.png' % name), gamma=1.0)
    for axis, (corr, rel) in zip('xy', block_agreement(r['fd'], r['grad'])):
        print('d/d%s: correlation of 8x8 block sums %.4f, rel-L2 %.3f' % (axis, corr, rel))


if __name__ == '__main__':
    main()
