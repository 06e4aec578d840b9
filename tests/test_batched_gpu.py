"""Batched scenes (reference front end: apps/generative_models/rendering.py:170-307): `batch` scenes of one
topology through dvg_scene_create_batch / dvg_render_*_batch must give, scene by scene, what the single-scene entry
points give for the same parameters and seed -- and those are checked against the compiled reference."""
import numpy as np
import pytest
import torch

import oracle_check
import ref_oracle
import scenes
import util
from diffvg_b200 import pydiffvg

pytestmark = pytest.mark.gpu


def _stroke_batch(indices):
    packs = [util.pack(scenes.batched_strokes(b)) for b in indices]
    topo = packs[0][0]
    for t, _ in packs:
        assert np.array_equal(t, topo)
    return topo, np.stack([p for _, p in packs])


def test_batch_equals_single_scenes_c5():
    """BASELINE configs[4] shape: 16 one-segment cubic strokes per 64x64 scene, 2x2 spp, seed = scene index."""
    idx = list(range(24)) + [100, 511]
    topo, rows = _stroke_batch(idx)
    got = util.gpu_render_batch(topo, rows, 64, 64, 2, 2, idx)['image']
    rng = np.random.RandomState(3)
    d_imgs = (rng.rand(len(idx), 64, 64, 4).astype(np.float32) - 0.5)
    gb = util.gpu_render_batch(topo, rows, 64, 64, 2, 2, idx, d_render_images=d_imgs)['d_params']
    for k, b in enumerate(idx):
        one = util.gpu_render(topo, rows[k], 64, 64, 2, 2, b)['image']
        assert np.array_equal(one, got[k]), 'scene %d' % b
        g1 = util.gpu_render(topo, rows[k], 64, 64, 2, 2, b, d_render_image=d_imgs[k])['d_params']
        assert util.rel_l2(g1, gb[k]) <= 2e-5, 'scene %d' % b      # float atomics arrive in a different order


@pytest.mark.skipif(not ref_oracle.available(), reason='oracle/_ref not built')
def test_batch_c5_vs_oracle():
    idx = [0, 1, 2, 3, 100, 511]
    topo, rows = _stroke_batch(idx)
    got = util.gpu_render_batch(topo, rows, 64, 64, 2, 2, idx)['image']
    d_imgs = (2.0 * got / got[0].size).astype(np.float32)
    gb = util.gpu_render_batch(topo, rows, 64, 64, 2, 2, idx, d_render_images=d_imgs)['d_params']
    for k, b in enumerate(idx):
        ref = oracle_check.render(topo, rows[k], 64, 64, 2, 2, b)['image']
        assert np.abs(ref - got[k]).max() <= 1e-5
        rb = oracle_check.render(topo, rows[k], 64, 64, 2, 2, b, d_render_image=d_imgs[k])
        assert util.rel_l2(rb['d_params'], gb[k]) <= 1e-4


def test_batch_of_mixed_scenes_with_backgrounds_and_transforms():
    """Every primitive type, fills, gradients colours, groups with transforms, a background per scene: the zoo scene with
    per-scene parameter perturbations (topology unchanged), transform gradients on."""
    topo, p0 = util.pack(scenes.zoo())
    rng = np.random.RandomState(11)
    B = 5
    rows = np.stack([p0] + [p0 + (rng.rand(p0.shape[0]).astype(np.float32) - 0.5) * 0.4 * (np.abs(p0) > 2.0) for _ in range(B - 1)])
    rows[:, -1] = p0[-1]                      # the pixel-filter radius is scene 0's for the whole batch
    seeds = [7, 7, 9, 2 ** 40 + 5, 0]
    bgs = rng.rand(B, 128, 128, 4).astype(np.float32)
    got = util.gpu_render_batch(topo, rows, 128, 128, 2, 2, seeds, backgrounds=bgs)['image']
    d_imgs = (rng.rand(B, 128, 128, 4).astype(np.float32) - 0.5)
    gb = util.gpu_render_batch(topo, rows, 128, 128, 2, 2, seeds, backgrounds=bgs, d_render_images=d_imgs)
    for k in range(B):
        one = util.gpu_render(topo, rows[k], 128, 128, 2, 2, seeds[k], background=bgs[k])
        assert np.array_equal(one['image'], got[k]), 'scene %d' % k
        g1 = util.gpu_render(topo, rows[k], 128, 128, 2, 2, seeds[k], background=bgs[k], d_render_image=d_imgs[k])
        assert util.rel_l2(g1['d_params'], gb['d_params'][k]) <= 2e-5, 'scene %d' % k
        assert np.abs(g1['d_background'] - gb['d_background'][k]).max() <= 1e-6


def test_batch_of_one_and_argument_checks():
    import ctypes
    from diffvg_b200 import _native as n
    topo, rows = _stroke_batch([5])
    got = util.gpu_render_batch(topo, rows, 64, 64, 2, 2, [5])['image']
    assert np.array_equal(got[0], util.gpu_render(topo, rows[0], 64, 64, 2, 2, 5)['image'])
    h = ctypes.c_void_p()
    t = np.ascontiguousarray(topo, np.int32)
    assert n.lib.dvg_scene_create_batch(t.ctypes.data, t.shape[0], 0, 0, ctypes.byref(h)) != 0
    n.check(n.lib.dvg_scene_create_batch(t.ctypes.data, t.shape[0], 0, 3, ctypes.byref(h)))
    try:
        p = np.ascontiguousarray(np.tile(rows[0], 3))
        assert n.lib.dvg_scene_set_params(h, p.ctypes.data, rows.shape[1], 0, None) != 0       # one scene's worth
        n.check(n.lib.dvg_scene_set_params(h, p.ctypes.data, p.size, 0, None))
        img = torch.empty(3, 64, 64, 4, device='cuda')
        # the single-scene entry points refuse a batch (no seeds), and so do prefiltering / SDF
        assert n.lib.dvg_render_forward(h, None, img.data_ptr(), None, 64, 64, 2, 2, 0, 0, None, 0, None) != 0
        sd = np.arange(3, dtype=np.uint64)
        assert n.lib.dvg_render_forward_batch(h, None, img.data_ptr(), 64, 64, 2, 2, None, None) != 0
        n.check(n.lib.dvg_render_forward_batch(h, None, img.data_ptr(), 64, 64, 2, 2, sd.ctypes.data, None))
        torch.cuda.synchronize()
    finally:
        n.lib.dvg_scene_destroy(h)


def test_bezier_and_line_render_match_per_sample_renders():
    """The vectorised front ends against the reference's per-sample loop restated with this package's single-scene
    RenderFunction (rendering.py:239-307): same images, same gradients on the batch tensors."""
    pydiffvg.set_use_gpu(True)
    g = torch.Generator().manual_seed(4)
    for fn, npts, ncp in ((pydiffvg.bezier_render, 7, [2, 2]), (pydiffvg.line_render, 2, [0])):
        bs, ns, canvas = 6, 5, 32
        pts = (torch.rand(bs, ns, npts, 2, generator=g) * 1.6 - 0.8).requires_grad_(True)
        widths = (0.5 + 2 * torch.rand(bs, ns, generator=g)).requires_grad_(True)
        alphas = torch.rand(bs, ns, generator=g).requires_grad_(True)
        colors = torch.rand(bs, ns, 3, generator=g).requires_grad_(True)
        seeds = list(range(40, 40 + bs))
        torch.manual_seed(0)
        out, scs = fn(pts, widths, alphas, canvas_size=canvas, colors=colors, seeds=seeds)
        assert tuple(out.shape) == (bs, 3, canvas, canvas) and len(scs) == bs
        tgt = torch.rand(bs, 3, canvas, canvas, generator=g)
        ((out - tgt) ** 2).mean().backward()
        grads = [t.grad.clone() for t in (pts, widths, alphas, colors)]
        for t in (pts, widths, alphas, colors):
            t.grad = None
        # per-sample loop
        torch.manual_seed(0)
        p2 = 0.5 * (pts + 1.0) * canvas
        p2 = p2 + 1e-4 * torch.randn_like(p2)
        outs = []
        for k in range(bs):
            shapes, groups = [], []
            for s in range(ns):
                shapes.append(pydiffvg.Path(num_control_points=torch.tensor(ncp, dtype=torch.int32), points=p2[k, s],
                                            stroke_width=widths[k, s], is_closed=False))
                groups.append(pydiffvg.ShapeGroup(shape_ids=torch.tensor([s]), fill_color=None,
                                                  stroke_color=torch.cat([colors[k, s], alphas[k, s].view(1)])))
            args = pydiffvg.RenderFunction.serialize_scene(canvas, canvas, shapes, groups)
            r = pydiffvg.RenderFunction.apply(canvas, canvas, 2, 2, seeds[k], None, *args).permute(2, 0, 1)
            outs.append(r[:3] * r[3:4])
            cw, ch, sh, gr = scs[k]
            assert len(sh) == ns and torch.allclose(sh[0].points, p2[k, 0].detach().cpu())
        ref = torch.stack(outs)
        assert torch.equal(ref.cpu(), out.cpu())
        ((ref.cpu() - tgt) ** 2).mean().backward()
        for a, t in zip(grads, (pts, widths, alphas, colors)):
            assert util.rel_l2(t.grad.numpy(), a.numpy()) <= 1e-5
