"""GPU parity tests (`-m gpu`) for SDF prefiltering, SDF output / eval_positions and the translation
gradient of render_grad, through the C ABI, against the committed reference fixtures
(tests/golden_modes/) and, where oracle/_ref is available, against the reference at larger sizes."""
import glob
import os

import numpy as np
import pytest
import torch

import oracle_check
from diffvg_b200 import scene_pack
import ref_oracle
import scenes
import util
from golden import make_golden as mg

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIX = sorted(glob.glob(os.path.join(ROOT, 'tests', 'golden_modes', '*.npz')))


def _check(g, out):
    for k in ('image', 'sdf'):
        if k in g.files:
            assert np.abs(out[k] - g[k]).max() <= 1e-5, k            # forward: 1e-5 absolute
    if np.linalg.norm(g['d_params']) > 0:
        assert util.rel_l2(g['d_params'], out['d_params']) <= 1e-4   # gradients: 1e-4 relative (rel-L2)
    else:
        assert np.all(out['d_params'] == 0)
    ref_t = g['d_translation']
    assert util.rel_l2(ref_t, out['d_translation']) <= 1e-4 or np.abs(ref_t).max() == 0
    if 'd_background' in g.files and int(g['config'][2]) * int(g['config'][3]) == 1:
        assert np.abs(out['d_background'] - g['d_background']).max() <= 1e-5


@pytest.mark.parametrize('path', FIX, ids=[os.path.basename(p)[:-4] for p in FIX])
def test_mode_fixtures(path):
    g = np.load(path)
    name = os.path.basename(path)[:-4]
    mode, _, W, H, nsx, nsy, seed, ft, fr, use_bg = mg.MODE_CASES[name]
    out = mg.run_mode(util.gpu_render, name, mode, g['topo'], g['params'], W, H, nsx, nsy, seed, use_bg)
    _check(g, out)


@pytest.mark.skipif(not ref_oracle.available(), reason='oracle/_ref not built')
def test_prefilter_blobs_512_vs_oracle():
    """Fill-heavy proxy of the flower config: 1024 closed cubic blobs, 512^2, 2x2 spp, prefiltered."""
    topo, params = util.pack(scenes.blobs())
    W = H = 512
    ref = oracle_check.render(topo, params, W, H, 2, 2, 0, use_prefiltering=True)['image']
    got = util.gpu_render(topo, params, W, H, 2, 2, 0, use_prefiltering=True)['image']
    assert np.abs(ref - got).max() <= 1e-5
    assert abs(ref.astype(np.float64).sum() - 356627.6572) < 0.05   # SURVEY 8d known answer
    d_img = np.random.RandomState(1).rand(H, W, 4).astype(np.float32) - 0.5
    rb = oracle_check.render(topo, params, W, H, 2, 2, 0, use_prefiltering=True, d_render_image=d_img)
    gb = util.gpu_render(topo, params, W, H, 2, 2, 0, use_prefiltering=True, d_render_image=d_img)
    assert util.rel_l2(rb['d_params'], gb['d_params']) <= 1e-4


@pytest.mark.skipif(not ref_oracle.available(), reason='oracle/_ref not built')
def test_prefilter_painterly_strokes_vs_oracle():
    topo, params = util.pack(scenes.painterly(256, 256))
    W = H = 256
    ref = oracle_check.render(topo, params, W, H, 2, 2, 3, use_prefiltering=True)['image']
    got = util.gpu_render(topo, params, W, H, 2, 2, 3, use_prefiltering=True)['image']
    assert np.abs(ref - got).max() <= 1e-5
    d_img = np.random.RandomState(2).rand(H, W, 4).astype(np.float32) - 0.5
    rb = oracle_check.render(topo, params, W, H, 2, 2, 3, use_prefiltering=True, d_render_image=d_img)
    gb = util.gpu_render(topo, params, W, H, 2, 2, 3, use_prefiltering=True, d_render_image=d_img)
    assert util.rel_l2(rb['d_params'], gb['d_params']) <= 1e-4


def test_pydiffvg_sdf_and_eval_positions_api():
    """test_eval_positions.py of the reference: SDF at explicit positions equals the SDF image sampled there."""
    from diffvg_b200 import pydiffvg
    pydiffvg.set_use_gpu(True)
    pydiffvg.set_device(torch.device('cuda', 0))
    cw, ch, shapes, groups = scenes.single_stroke()
    shapes[0].points.requires_grad_(True)
    args = pydiffvg.RenderFunction.serialize_scene(cw, ch, shapes, groups, output_type=pydiffvg.OutputType.sdf)
    sdf = pydiffvg.RenderFunction.apply(256, 256, 1, 1, 0, None, *args)
    assert sdf.shape == (256, 256, 1)
    ep = torch.tensor([[100.5, 40.5], [30.25, 200.75], [128.0, 128.0]])
    args2 = pydiffvg.RenderFunction.serialize_scene(cw, ch, shapes, groups, output_type=pydiffvg.OutputType.sdf,
                                                    eval_positions=ep)
    vals = pydiffvg.RenderFunction.apply(256, 256, 1, 1, 0, None, *args2)
    assert vals.shape == (3, 1)
    vals.sum().backward()
    assert shapes[0].points.grad is not None and torch.isfinite(shapes[0].points.grad).all()
    assert shapes[0].points.grad.abs().sum() > 0
    # render_grad: translation gradient image
    args3 = pydiffvg.RenderFunction.serialize_scene(cw, ch, shapes, groups)
    tg = pydiffvg.RenderFunction.render_grad(torch.ones(256, 256, 4), 256, 256, 2, 2, 0, None, *args3)
    assert tg.shape == (256, 256, 2) and torch.isfinite(tg).all() and tg.abs().sum() > 0


def test_c4_like_prefilter_2048_size_independent_properties():
    """BASELINE configs[3] shape (flower.svg, 2048x2048, use_prefiltering, 2x2 spp) on the fill-heavy proxy scene:
    at the full size the oracle takes minutes, so parity is carried by the 512^2 test above and the full size is
    checked through properties: the row-sharded render equals the whole one (what the multi-GPU split relies
    on), the backward pass is linear in d_image, and results repeat."""
    topo, params = util.pack(scenes.blobs())
    W = H = 2048
    whole = util.gpu_render(topo, params, W, H, 2, 2, 0, use_prefiltering=True)['image']
    assert np.isfinite(whole).all() and whole[..., 3].min() >= 0 and whole[..., 3].max() <= 1 + 1e-5
    again = util.gpu_render(topo, params, W, H, 2, 2, 0, use_prefiltering=True)['image']
    assert np.abs(whole - again).max() <= 1e-6
    rng = np.random.RandomState(5)
    d1 = (rng.rand(H, W, 4).astype(np.float32) - 0.5)
    d2 = (rng.rand(H, W, 4).astype(np.float32) - 0.5)
    parts = util.gpu_render_rows(topo, params, W, H, 2, 2, 0, [(0, 512), (512, 1536), (1536, 2048)], d_render_image=d1,
                                 use_prefiltering=True)
    assert np.abs(parts['image'] - whole).max() <= 1e-6
    g1 = util.gpu_render(topo, params, W, H, 2, 2, 0, use_prefiltering=True, d_render_image=d1)['d_params'].astype(np.float64)
    g2 = util.gpu_render(topo, params, W, H, 2, 2, 0, use_prefiltering=True, d_render_image=d2)['d_params'].astype(np.float64)
    g12 = util.gpu_render(topo, params, W, H, 2, 2, 0, use_prefiltering=True,
                          d_render_image=(0.5 * d1 - 2.0 * d2).astype(np.float32))['d_params'].astype(np.float64)
    assert util.rel_l2(0.5 * g1 - 2.0 * g2, g12) <= 1e-4
    assert util.rel_l2(g1, parts['d_params']) <= 1e-4


@pytest.mark.skipif(not ref_oracle.available(), reason='oracle/_ref not built')
def test_c5_batched_stroke_scenes_vs_oracle():
    """BASELINE configs[4]: 16 open one-segment cubic strokes per 64x64 scene, 2x2 spp, seed = scene index."""
    for b in (0, 1, 2, 3, 100, 511):
        topo, params = util.pack(scenes.batched_strokes(b))
        ref = oracle_check.render(topo, params, 64, 64, 2, 2, b)['image']
        got = util.gpu_render(topo, params, 64, 64, 2, 2, b)['image']
        assert np.abs(ref - got).max() <= 1e-5
        d_img = (2.0 * got / got.size).astype(np.float32)          # loss = mean of squares
        rb = oracle_check.render(topo, params, 64, 64, 2, 2, b, d_render_image=d_img)
        gb = util.gpu_render(topo, params, 64, 64, 2, 2, b, d_render_image=d_img)
        assert util.rel_l2(rb['d_params'], gb['d_params']) <= 1e-4


def test_skip_filter_grad_flag_only_drops_the_radius_entry():
    """DVG_BWD_SKIP_FILTER_GRAD (what pydiffvg passes when the filter radius takes no part in autograd): every other
    gradient is unchanged, the radius entry stays 0; sampled and prefiltered paths."""
    for pf in (False, True):
        topo, params = util.pack(scenes.zoo_prefilter() if pf else scenes.zoo(), filter_type=1, filter_radius=1.5)
        rng = np.random.RandomState(2)
        d_img = (rng.rand(128, 128, 4).astype(np.float32) - 0.5)
        full = util.gpu_render(topo, params, 128, 128, 2, 2, 4, d_render_image=d_img, use_prefiltering=pf)['d_params']
        skip = util.gpu_render(topo, params, 128, 128, 2, 2, 4, d_render_image=d_img, use_prefiltering=pf, extra_flags=4)['d_params']
        roff = int(topo[6])      # DVG_H_FILTER_RADIUS_OFF
        assert full[roff] != 0.0 and skip[roff] == 0.0
        keep = np.arange(full.shape[0]) != roff
        assert util.rel_l2(full[keep], skip[keep]) <= 2e-5


def test_gradient_stroke_colour_gradients_match_finite_differences():
    """Q4: the reference has no defined behaviour for the gradient of a Linear/RadialGradient STROKE colour (scene.cpp:866-889
    never assigns d_shape_groups[g].stroke_color and the backward pass faults).  Here it is accumulated like a fill's; the
    colour parameters move the image smoothly, so central finite differences are the check (sampled and prefiltered)."""
    from diffvg_b200 import pydiffvg
    pydiffvg.set_use_gpu(True)
    lin = dict(begin=torch.tensor([10.0, 12.0]), end=torch.tensor([52.0, 47.0]), offsets=torch.tensor([0.1, 0.5, 0.95]),
               stop_colors=torch.tensor([[0.9, 0.1, 0.2, 0.9], [0.2, 0.8, 0.3, 0.6], [0.1, 0.2, 0.9, 1.0]]))
    rad = dict(center=torch.tensor([40.0, 24.0]), radius=torch.tensor([20.0, 14.0]), offsets=torch.tensor([0.2, 0.8]),
               stop_colors=torch.tensor([[0.8, 0.7, 0.1, 1.0], [0.1, 0.5, 0.6, 0.5]]))
    leaves = {('lin', k): v.clone().requires_grad_(True) for k, v in lin.items()}
    leaves.update({('rad', k): v.clone().requires_grad_(True) for k, v in rad.items()})

    def scene(vals):
        shapes = [pydiffvg.Path(num_control_points=torch.tensor([2, 0]), points=torch.tensor([[8.0, 50.0], [20.0, 5.0], [40.0, 60.0], [55.0, 12.0], [58.0, 40.0]]),
                                is_closed=False, stroke_width=torch.tensor(5.0)),
                  pydiffvg.Circle(radius=torch.tensor(13.0), center=torch.tensor([40.0, 26.0]), stroke_width=torch.tensor(4.0))]
        groups = [pydiffvg.ShapeGroup(torch.tensor([0]), fill_color=None, stroke_color=pydiffvg.LinearGradient(
                      vals[('lin', 'begin')], vals[('lin', 'end')], vals[('lin', 'offsets')], vals[('lin', 'stop_colors')])),
                  pydiffvg.ShapeGroup(torch.tensor([1]), fill_color=torch.tensor([0.3, 0.3, 0.3, 0.4]), stroke_color=pydiffvg.RadialGradient(
                      vals[('rad', 'center')], vals[('rad', 'radius')], vals[('rad', 'offsets')], vals[('rad', 'stop_colors')]))]
        return shapes, groups

    wt = torch.rand(64, 64, 4, generator=torch.Generator().manual_seed(9)).double()
    for pf in (False, True):
        def loss_of(vals):
            shapes, groups = scene(vals)
            args = pydiffvg.RenderFunction.serialize_scene(64, 64, shapes, groups, use_prefiltering=pf)
            img = pydiffvg.RenderFunction.apply(64, 64, 2, 2, 5, None, *args)
            return (img.cpu().double() * wt).sum()
        for t in leaves.values():
            t.grad = None
        loss_of(leaves).backward()
        checked = 0
        for key, t in leaves.items():
            flat = t.detach().clone().reshape(-1)
            g = t.grad.reshape(-1)
            for i in range(flat.numel()):
                if key == ('rad', 'stop_colors') and i >= 4:
                    continue      # Q3: the reference's radial branch also adds d_color to the LAST stop (no `return`); reproduced
                eps = 0.005 if key[1] == 'offsets' else (0.02 if key[1] == 'stop_colors' else 0.25)
                vals = {k: v.detach() for k, v in leaves.items()}
                hi, lo = flat.clone(), flat.clone()
                hi[i] += eps; lo[i] -= eps
                vals[key] = hi.reshape(t.shape); lp = float(loss_of(vals))
                vals[key] = lo.reshape(t.shape); lm = float(loss_of(vals))
                fd = (lp - lm) / (2 * eps)
                # (the colour is piecewise linear in t: a central difference across the kinks at the stops is off by O(eps))
                tol = 0.08 if key[1] == 'offsets' else 0.03
                assert abs(fd - float(g[i])) <= tol * abs(fd) + 0.05, (pf, key, i, fd, float(g[i]))
                checked += 1 if abs(fd) > 0.05 else 0
        assert checked >= 15


def test_fast_stroke_accept_mode_differs_by_whole_samples_only():
    """dvg_set_fast_stroke_accept(1) (opt-in, DESIGN 'arithmetic contract'): samples the polyline bracket proves inside a
    curved stroke skip the reference's closest-point solve.  Geometrically right; the reference's solver misses a few
    near-tangent cases (Q21), so a handful of pixels may differ from the exact mode -- each by whole samples' weights,
    and only towards MORE coverage -- and the gradients stay within the usual tolerance."""
    from diffvg_b200 import _native as n
    topo, params = util.pack(scenes.painterly(512, 256))
    exact = util.gpu_render(topo, params, 256, 256, 4, 4, 0)['image']
    d_img = (np.random.RandomState(4).rand(256, 256, 4).astype(np.float32) - 0.5)
    g_exact = util.gpu_render(topo, params, 256, 256, 4, 4, 0, d_render_image=d_img)['d_params']
    assert n.lib.dvg_set_fast_stroke_accept(1) == 0
    try:
        fast = util.gpu_render(topo, params, 256, 256, 4, 4, 0)['image']
        g_fast = util.gpu_render(topo, params, 256, 256, 4, 4, 0, d_render_image=d_img)['d_params']
    finally:
        n.lib.dvg_set_fast_stroke_accept(0)
    diff = np.abs(fast - exact).max(axis=2)
    assert (diff > 1e-6).sum() <= 64                   # a handful of pixels
    assert diff.max() <= 3.0 / 16 + 1e-5              # a few samples of 16 at most
    assert util.rel_l2(g_exact, g_fast) <= 1e-3
    again = util.gpu_render(topo, params, 256, 256, 4, 4, 0)['image']
    assert np.array_equal(again, exact)                # the switch is off again


def test_row_costs_and_balanced_bands_reproduce_the_whole_render():
    """dvg_scene_row_costs: per tile row, the candidate entries of the whole-image bins; bands cut from it
    (sharded.balanced_row_partition) must render the same image and gradient as one call (prefiltered fill scene and
    sampled stroke scene)."""
    import ctypes
    from diffvg_b200 import _native as n, sharded
    for mk, pf, (W, H, ns) in ((scenes.blobs, True, (512, 512, 2)), (lambda: scenes.painterly(256, 256), False, (256, 256, 4))):
        topo, params = util.pack(mk())
        h = ctypes.c_void_p()
        t = np.ascontiguousarray(topo, np.int32)
        n.check(n.lib.dvg_scene_create(t.ctypes.data, t.shape[0], 0, ctypes.byref(h)))
        try:
            p = np.ascontiguousarray(params, np.float32)
            n.check(n.lib.dvg_scene_set_params(h, p.ctypes.data, p.shape[0], 0, None))
            costs = np.zeros(H + 1, np.float32)
            th = ctypes.c_int()
            assert n.lib.dvg_scene_row_costs(h, W, H, ns, ns, 1 if pf else 0, costs.ctypes.data, 3, ctypes.byref(th), None) != 0   # too small
            n.check(n.lib.dvg_scene_row_costs(h, W, H, ns, ns, 1 if pf else 0, costs.ctypes.data, H + 1, ctypes.byref(th), None))
        finally:
            n.lib.dvg_scene_destroy(h)
        assert th.value == sharded.tile_height(ns * ns)
        units = (H + th.value - 1) // th.value
        c = costs[:units]
        assert (c >= 0).all() and c.sum() > 0 and (costs[units:] == 0).all()
        bands = sharded.balanced_row_partition(c, H, 5, th.value, per_unit=4.0 * W / 8)
        d_img = (np.random.RandomState(6).rand(H, W, 4).astype(np.float32) - 0.5)
        whole = util.gpu_render(topo, params, W, H, ns, ns, 2, use_prefiltering=pf)['image']
        gw = util.gpu_render(topo, params, W, H, ns, ns, 2, use_prefiltering=pf, d_render_image=d_img)['d_params']
        parts = util.gpu_render_rows(topo, params, W, H, ns, ns, 2, bands, d_render_image=d_img, use_prefiltering=pf)
        assert np.abs(parts['image'] - whole).max() <= 1e-6
        assert util.rel_l2(gw, parts['d_params']) <= 1e-4


def test_prefilter_winding_prepass_equals_inline_and_backward_reuses_forward_words():
    """The prefiltered path answers the winding numbers of filled groups in a pre-pass (classify -> k_wave_solve_fill) and
    the render kernel reads them as words; dvg_debug_set_prefilter_inline(1) runs the same tests inline in the render
    kernel.  Both evaluate the same functions on the same inputs: images are identical, gradients equal up to the order
    of the atomic sums.  Then forward + backward on ONE scene object (pydiffvg): the backward pass re-uses the forward
    pass's words and must give the gradient of a fresh scene; a parameter change must not see stale words."""
    from diffvg_b200 import _native as n, pydiffvg
    cases = ((scenes.blobs(), 256, 256, 2, 2), (scenes.zoo(), 96, 96, 2, 2), (scenes.zoo(), 70, 50, 3, 1))
    for scene, W, H, nsx, nsy in cases:
        topo, params = util.pack(scene)
        d_img = np.random.RandomState(2).rand(H, W, 4).astype(np.float32) - 0.5
        a = util.gpu_render(topo, params, W, H, nsx, nsy, 0, use_prefiltering=True)['image']
        ga = util.gpu_render(topo, params, W, H, nsx, nsy, 0, use_prefiltering=True, d_render_image=d_img)['d_params']
        assert n.lib.dvg_debug_set_prefilter_inline(1) == 0
        try:
            b = util.gpu_render(topo, params, W, H, nsx, nsy, 0, use_prefiltering=True)['image']
            gb = util.gpu_render(topo, params, W, H, nsx, nsy, 0, use_prefiltering=True, d_render_image=d_img)['d_params']
        finally:
            n.lib.dvg_debug_set_prefilter_inline(0)
        if (nsx * nsy) & (nsx * nsy - 1):   # not a power of two: every sample splats with its own atomic, the sum order is free
            assert np.abs(a - b).max() <= 1e-6
        else:
            assert np.array_equal(a, b)
        assert util.rel_l2(gb, ga) <= 1e-4    # (atomic sums in a different order)
    # fragment cache: forward + backward on one scene object differentiate most samples from the records the forward kernel
    # left (at most 4 fragments per sample; the others go through the full kernel) -- against the full kernel for all
    import ctypes
    tiger = np.load(os.path.join(ROOT, 'tests', 'golden_svg', 'tiger.npz'))
    for topo, params, W, H in ((tiger['topo'], tiger['params'], 250, 256), util.pack(scenes.painterly(256, 128)) + (128, 128),
                               util.pack(scenes.zoo()) + (96, 96)):
        topo = np.ascontiguousarray(topo, np.int32); params = np.ascontiguousarray(params, np.float32)
        d_img = torch.from_numpy(np.random.RandomState(5).rand(H, W, 4).astype(np.float32) - 0.5).cuda()
        grads = []
        for mode in (0, 2):
            assert n.lib.dvg_debug_set_prefilter_inline(mode) == 0
            try:
                h = ctypes.c_void_p()
                n.check(n.lib.dvg_scene_create(topo.ctypes.data, topo.shape[0], 0, ctypes.byref(h)))
                st = torch.cuda.current_stream().cuda_stream
                n.check(n.lib.dvg_scene_set_params(h, params.ctypes.data, params.shape[0], 0, st))
                img = torch.empty(H, W, 4, device='cuda'); g = torch.empty(params.shape[0], device='cuda')
                n.check(n.lib.dvg_render_forward(h, None, img.data_ptr(), None, W, H, 2, 2, 0, 1, None, 0, st))
                n.check(n.lib.dvg_render_backward(h, None, d_img.data_ptr(), None, W, H, 2, 2, 0, 1, None, 0, g.data_ptr(), None, None, 0, st))
                torch.cuda.synchronize()
                grads.append(g.cpu().numpy())
                n.lib.dvg_scene_destroy(h)
            finally:
                n.lib.dvg_debug_set_prefilter_inline(0)
        assert np.linalg.norm(grads[1]) > 0
        assert util.rel_l2(grads[1], grads[0]) <= 1e-4
    # one scene object, forward then backward (words re-used), twice with different parameters
    pydiffvg.set_use_gpu(True)
    cw, ch, shapes, groups = scenes.blobs()
    packed, params0 = pydiffvg.RenderFunction.serialize_scene(cw, ch, shapes, groups, use_prefiltering=True)
    topo, params_np = util.pack((cw, ch, shapes, groups))
    W = H = 192
    ro = int(topo[scene_pack.H_FRAD_OFF])
    target = torch.rand(H, W, 4, generator=torch.Generator().manual_seed(3))
    for shift in (0.0, 0.25):
        p_np = params_np.copy()
        p_np += shift                     # moves every point (and tints every colour): the geometry changes
        p_np[ro] = params_np[ro]
        p = torch.from_numpy(p_np.copy()).requires_grad_(True)
        img = pydiffvg.RenderFunction.apply(W, H, 2, 2, 0, None, packed, p)
        loss = (img.cpu() - target).pow(2).mean()
        (g,) = torch.autograd.grad(loss, p)
        fresh = util.gpu_render(topo, p_np, W, H, 2, 2, 0, use_prefiltering=True)['image']
        assert np.array_equal(fresh, img.detach().cpu().numpy())
        d_img = (2.0 * (img.detach().cpu().numpy() - target.numpy()) / target.numel()).astype(np.float32)
        gf = util.gpu_render(topo, p_np, W, H, 2, 2, 0, use_prefiltering=True, d_render_image=d_img)['d_params']
        got = g.cpu().numpy().copy()
        got[ro] = gf[ro] = 0.0      # d_filter.radius is skipped unless the radius takes part in autograd
        assert util.rel_l2(gf, got) <= 1e-4
