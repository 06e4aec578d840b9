"""CPU tests of the SVG front end: path-data reader, `from_svg_path`, `svg_to_scene`,
`save_svg`, `imwrite`; and the oracle on the asset packs against the committed golden outputs.

Reference: pydiffvg/shape.py:63-172, parse_svg.py:17-586, save_svg.py:13-156, image.py:6-21; the
svgpathtools behaviour restated in diffvg_b200/pydiffvg/svg_path.py."""
import math
import os
import warnings

import numpy as np
import pytest
import torch

import util
from diffvg_b200 import pydiffvg
from diffvg_b200.pydiffvg import svg_path as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ASSETS = '/root/reference/apps/imgs'


def test_path_reader_commands_and_implicit_repeats():
    segs = sp.parse_path('M 10,20 L 30,20 30,40 H 5 V 7 h 2 v -3 Z')
    assert [s[0] for s in segs] == ['L'] * 7
    assert segs[0][1:] == (10 + 20j, 30 + 20j) and segs[1][2] == 30 + 40j
    assert segs[2][2] == 5 + 40j and segs[3][2] == 5 + 7j and segs[4][2] == 7 + 7j and segs[5][2] == 7 + 4j
    assert segs[6][1:] == (7 + 4j, 10 + 20j)                 # Z adds the closing line (current != start)
    # pairs after a moveto are linetos; relative moveto after Z starts from the sub-path start
    segs = sp.parse_path('m 1 1 2 0 0 2 z m 10 0 l 1 1')
    assert [(s[1], s[2]) for s in segs] == [(1 + 1j, 3 + 1j), (3 + 1j, 3 + 3j), (3 + 3j, 1 + 1j), (11 + 1j, 12 + 2j)]
    # numbers glued by signs / dots, exponents
    segs = sp.parse_path('M0.5-1.5l1e1-.5.5.25')
    assert segs[0][1:] == (0.5 - 1.5j, 10.5 - 2j) and segs[1][2] == 11 - 1.75j


def test_path_reader_curves_and_reflections():
    segs = sp.parse_path('M0,0 C1,2 3,4 5,6 S9,10 11,12 s1,1 2,2 Q1,1 2,0 T4,0 t2,0')
    assert segs[0] == ('C', 0j, 1 + 2j, 3 + 4j, 5 + 6j)
    assert segs[1] == ('C', 5 + 6j, 7 + 8j, 9 + 10j, 11 + 12j)          # reflection of (3,4) about (5,6)
    assert segs[2] == ('C', 11 + 12j, 13 + 14j, 12 + 13j, 13 + 14j)
    assert segs[3] == ('Q', 13 + 14j, 1 + 1j, 2 + 0j)
    assert segs[4] == ('Q', 2 + 0j, 3 - 1j, 4 + 0j) and segs[5] == ('Q', 4 + 0j, 5 + 1j, 6 + 0j)
    # S after a non-cubic uses the current point as first control point
    assert sp.parse_path('M0,0 L1,1 S2,2 3,3')[1] == ('C', 1 + 1j, 1 + 1j, 2 + 2j, 3 + 3j)


def test_arc_centre_parameterisation():
    a = sp.parse_path('M 10 0 A 10 10 0 0 1 0 10')[0][3]            # quarter circle about the origin, sweep
    assert abs(a['center']) < 1e-12 and abs(a['theta']) < 1e-9 and abs(a['delta'] - 90) < 1e-9
    a = sp.parse_path('M 10 0 A 10 10 0 1 0 0 10')[0][3]            # same circle, the other way round: 270 degrees
    assert abs(a['center']) < 1e-12 and abs(a['delta'] + 270) < 1e-9
    a = sp.parse_path('M 10 0 A 10 10 0 1 1 0 10')[0][3]            # large arc with sweep: the circle about (10, 10)
    assert abs(a['center'] - (10 + 10j)) < 1e-12 and abs(a['delta'] - 270) < 1e-9
    a = sp.parse_path('M 0 0 A 1 1 0 0 1 10 0')[0][3]               # radii too small: scaled up to fit
    assert abs(a['radius'] - (5 + 5j)) < 1e-12 and abs(a['center'] - 5) < 1e-12
    paths = pydiffvg.from_svg_path('M 10 0 A 10 10 0 0 1 -10 0')
    assert paths[0].num_control_points.tolist() == [2, 2] and paths[0].points.shape[0] == 7
    pts = paths[0].points.numpy()
    assert np.allclose(pts[3], [0, 10], atol=1e-5) and np.allclose(pts[6], [-10, 0], atol=1e-5)
    for t in np.linspace(0, 1, 9):                                    # the cubics stay on the circle
        b = ((1 - t) ** 3) * pts[0] + 3 * (1 - t) ** 2 * t * pts[1] + 3 * (1 - t) * t * t * pts[2] + t ** 3 * pts[3]
        assert abs(np.hypot(*b) - 10) < 5e-3


def test_from_svg_path_closing_rules_and_transform():
    # closed by coincidence of end and start: first point not repeated
    p = pydiffvg.from_svg_path('M 0,0 L 4,0 4,4 0,4 0,0 Z')[0]
    assert p.is_closed and p.num_control_points.tolist() == [0, 0, 0, 0] and p.points.tolist() == [[0, 0], [4, 0], [4, 4], [0, 4]]
    # Z with a gap adds the closing line
    p = pydiffvg.from_svg_path('M 0,0 L 4,0 4,4 Z')[0]
    assert p.is_closed and p.num_control_points.tolist() == [0, 0, 0] and p.points.shape[0] == 3
    # a closing line shorter than 1e-5 is dropped and the previous segment snapped onto the start
    p = pydiffvg.from_svg_path('M 0,0 C 1,0 2,1 2,2 C 2,3 1,3 0.000001,0 Z')[0]
    assert p.is_closed and p.num_control_points.tolist() == [2, 2] and p.points.shape[0] == 6
    # open path: kept open, or force-closed for fills
    p = pydiffvg.from_svg_path('M 0,0 C 1,0 2,1 2,2')[0]
    assert not p.is_closed and p.points.shape[0] == 4
    p = pydiffvg.from_svg_path('M 0,0 C 1,0 2,1 2,2', force_close=True)[0]
    assert p.is_closed and p.num_control_points.tolist() == [2, 0] and p.points.shape[0] == 4
    # two sub-paths, transform applied to the points
    m = torch.tensor([[1.0, 0.0, 5.0], [0.0, -1.0, 10.0], [0.0, 0.0, 1.0]])
    ps = pydiffvg.from_svg_path('M 0,0 L 1,0 1,1 z M 5,5 l 1,0', m)
    assert len(ps) == 2 and ps[0].points.tolist() == [[5, 10], [6, 10], [6, 9]] and ps[1].points.tolist() == [[10, 5], [11, 5]]
    assert pydiffvg.from_svg_path('') == []


def test_parse_transform_and_colors():
    m = pydiffvg.parse_transform('scale(1,-1) translate(0,-510) ')
    assert m.dtype == torch.float32 and m.tolist() == [[1, 0, 0], [0, -1, 510], [0, 0, 1]]
    m = pydiffvg.parse_transform('rotate(90 1 1)')
    assert np.allclose(m.numpy() @ np.array([2.0, 1.0, 1.0]), [1.0, 2.0, 1.0], atol=1e-6)
    m = pydiffvg.parse_transform('matrix(1 2 3 4 5 6)')
    assert m.tolist() == [[1, 3, 5], [2, 4, 6], [0, 0, 1]]
    assert pydiffvg.parse_color('#2C331E', {}).tolist() == pytest.approx([0x2c / 255, 0x33 / 255, 0x1e / 255, 1.0])
    assert pydiffvg.parse_color('#fa0', {}).tolist() == pytest.approx([1.0, 0xaa / 255, 0.0, 1.0])
    assert pydiffvg.parse_color('rgb(255,0,51)', {}).tolist() == pytest.approx([1.0, 0.0, 0.2, 1.0])
    assert pydiffvg.parse_color('none', {}) is None
    assert pydiffvg.parse_color('rebeccapurple', {}).tolist() == pytest.approx([0x66 / 255, 0x33 / 255, 0x99 / 255, 1.0])


SVG_DOC = '''<?xml version="1.0"?>
<svg xmlns="http://www.w3.org/2000/svg" xmlns:xlink="http://www.w3.org/1999/xlink" viewBox="0 0 64px 48px">
<style type="text/css"> .a{fill:#FF0000;} .b{fill:none;stroke:#00FF00;} </style>
<defs><linearGradient id="lg" x1="0" y1="0" x2="10" y2="0"><stop offset="0" stop-color="#000"/>
<stop offset="1" style="stop-color:#fff;stop-opacity:0.5"/></linearGradient></defs>
<g transform="translate(2,3)" fill="blue">
  <path class="a" d="M0,0 h10 v10 h-10 z" opacity="0.5"/>
  <path d="M20,0 l5,5" fill="none" stroke="rgb(0,0,255)" stroke-width="3px"/>
  <polygon points="0,0 4,0 4,4" fill="url(#lg)"/>
  <g transform="scale(2)"><circle cx="5" cy="6" r="2"/><line x1="0" y1="0" x2="3" y2="4" stroke="black"/>
  <rect x="7" y="8" width="3" height="4" style="fill:#00ff00;fill-rule:evenodd;stroke:red;stroke-width:1;opacity:0.25"/></g>
</g></svg>'''


def test_svg_to_scene_document(tmp_path):
    f = tmp_path / 'doc.svg'
    f.write_text(SVG_DOC)
    cw, ch, shapes, groups = pydiffvg.svg_to_scene(str(f))
    assert (cw, ch) == (64, 48) and len(shapes) == 6 and len(groups) == 6
    p0, g0 = shapes[0], groups[0]
    assert p0.is_closed and p0.points.tolist() == [[2, 3], [12, 3], [12, 13], [2, 13]]           # transform baked into points
    assert g0.fill_color.tolist() == pytest.approx([1, 0, 0, 0.5]) and g0.stroke_color is None
    assert torch.equal(g0.shape_to_canvas, torch.eye(3)) and g0.use_even_odd_rule is False
    assert float(p0.stroke_width) == 0.5                                                          # default stroke radius
    p1, g1 = shapes[1], groups[1]
    assert not p1.is_closed and g1.fill_color is None and g1.stroke_color.tolist() == [0, 0, 1, 1]
    assert float(p1.stroke_width) == 1.5                                                          # width 3px -> radius 1.5
    g2 = groups[2]
    assert isinstance(g2.fill_color, pydiffvg.LinearGradient) and g2.fill_color.stop_colors.tolist() == [[0, 0, 0, 1], [1, 1, 1, 0.5]]
    assert isinstance(shapes[2], pydiffvg.Polygon) and shapes[2].is_closed
    assert g2.shape_to_canvas.tolist() == [[1, 0, 2], [0, 1, 3], [0, 0, 1]]                       # non-path shapes keep the transform
    assert isinstance(shapes[3], pydiffvg.Circle) and groups[3].fill_color.tolist() == [0, 0, 1, 1]   # inherited from <g fill>
    assert groups[3].shape_to_canvas.tolist() == [[2, 0, 2], [0, 2, 3], [0, 0, 1]]
    assert isinstance(shapes[4], pydiffvg.Polygon) and not shapes[4].is_closed and groups[4].stroke_color.tolist() == [0, 0, 0, 1]
    r, g5 = shapes[5], groups[5]
    assert r.p_min.tolist() == [0, 0] and r.p_max.tolist() == [3, 4]      # reference quirk: x / y of <rect> are never read
    assert g5.use_even_odd_rule is True and g5.fill_color.tolist() == pytest.approx([0, 1, 0, 0.25])
    assert g5.stroke_color.tolist() == pytest.approx([1, 0, 0, 0.25]) and float(r.stroke_width) == 0.5
    # ... and the whole thing packs
    topo, params = util.pack((cw, ch, shapes, groups))
    assert topo[3] == 6 and np.isfinite(params).all()


@pytest.mark.skipif(not os.path.isdir(ASSETS), reason='the reference assets are only in the build container')
@pytest.mark.parametrize('name', ['tiger', 'flower'])
def test_asset_packs_are_what_the_loader_produces(name):
    """tests/golden_svg/*.npz (what the GPU box renders) == svg_to_scene + pack of the asset, bit for bit."""
    from golden.make_svg_golden import load_pack
    g = np.load(os.path.join(ROOT, 'tests', 'golden_svg', name + '.npz'))
    topo, params = load_pack(name)
    assert np.array_equal(topo, g['topo']) and np.array_equal(params, g['params'])
    if name == 'tiger':      # viewBox 495x510, 303 <path> elements
        assert (topo[1], topo[2], topo[3], topo[4]) == (495, 510, 303, 303)
    else:                    # viewBox 800x865, 1077 paths + 19 polygons
        assert (topo[1], topo[2], topo[4]) == (800, 865, 1096)


@pytest.mark.parametrize('name', ['tiger', 'flower'])
def test_oracle_reproduces_asset_golden_outputs(name):
    import oracle_check
    from golden.make_golden import d_image_for
    g = np.load(os.path.join(ROOT, 'tests', 'golden_svg', name + '.npz'))
    W, H, nsx, nsy, seed, pf = [int(v) for v in g['config']]
    img = oracle_check.render(g['topo'], g['params'], W, H, nsx, nsy, seed, use_prefiltering=bool(pf))['image']
    assert np.abs(img - g['image']).max() <= 1e-6
    b = oracle_check.render(g['topo'], g['params'], W, H, nsx, nsy, seed, use_prefiltering=bool(pf), d_render_image=d_image_for(name, H, W))
    assert util.rel_l2(g['d_params'], b['d_params']) <= 1e-4


def test_save_svg_round_trip(tmp_path):
    import scenes
    cw, ch, shapes, groups = scenes.zoo()
    # one shape per group and no per-point thickness: what save_svg can express (save_svg.py:78, 137); no
    # ellipse: the loader does not dispatch <ellipse> (parse_svg.py:391)
    keep = [i for i, g in enumerate(groups) if len(g.shape_ids) == 1 and shapes[int(g.shape_ids[0])].stroke_width.dim() == 0
            and not isinstance(g.fill_color, pydiffvg.RadialGradient) and not isinstance(shapes[int(g.shape_ids[0])], pydiffvg.Ellipse)]
    sh = [shapes[int(groups[i].shape_ids[0])] for i in keep]
    gr = [pydiffvg.ShapeGroup(torch.tensor([k]), groups[i].fill_color, groups[i].use_even_odd_rule, groups[i].stroke_color)
          for k, i in enumerate(keep)]
    f = tmp_path / 'out.svg'
    pydiffvg.save_svg(str(f), cw, ch, sh, gr)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        cw2, ch2, sh2, gr2 = pydiffvg.svg_to_scene(str(f))
    assert (cw2, ch2) == (cw, ch) and len(gr2) == len(gr)
    for a, b, ga, gb in zip(sh, sh2, gr, gr2):
        assert type(a).__name__ == type(b).__name__
        assert float(b.stroke_width) == pytest.approx(float(a.stroke_width))        # radius -> width -> radius
        if isinstance(a, pydiffvg.Path):
            assert a.num_control_points.tolist() == b.num_control_points.tolist()
            assert np.allclose(a.points.numpy(), b.points.numpy(), atol=1e-4)
        if isinstance(ga.stroke_color, torch.Tensor):
            assert np.allclose(ga.stroke_color.numpy(), gb.stroke_color.numpy(), atol=1 / 255 + 1e-6)


def test_imwrite_png(tmp_path):
    from PIL import Image
    img = torch.rand(7, 5, 4)
    f = tmp_path / 'sub' / 'a.png'
    pydiffvg.imwrite(img, str(f), gamma=2.2)
    got = np.asarray(Image.open(str(f)))
    ref = img.numpy().copy()
    ref[:, :, :3] = ref[:, :, :3] ** (1 / 2.2)
    assert got.shape == (7, 5, 4) and np.array_equal(got, (np.clip(ref, 0, 1) * 255).astype(np.uint8))
    # the built-in encoder (used where Pillow is absent) writes a valid PNG too
    from diffvg_b200.pydiffvg.image import _png_bytes
    a = (np.random.RandomState(0).rand(6, 9, 3) * 255).astype(np.uint8)
    (tmp_path / 'b.png').write_bytes(_png_bytes(a))
    assert np.array_equal(np.asarray(Image.open(str(tmp_path / 'b.png'))), a)
