"""CPU side of the batched front end (diffvg_b200/pydiffvg/batched.py): parameter layout of the vectorised stroke-scene
builder against the stock per-scene packing, topology checks of serialize_scenes, the lazy `scenes` list."""
import numpy as np
import pytest
import torch

import scenes
import util
from diffvg_b200 import pydiffvg, scene_pack


def _holders(pts, widths, rgba, ncp):
    shapes, groups = [], []
    for s in range(pts.shape[0]):
        shapes.append(pydiffvg.Path(num_control_points=torch.tensor(ncp, dtype=torch.int32), points=pts[s],
                                    stroke_width=widths[s], is_closed=False))
        groups.append(pydiffvg.ShapeGroup(shape_ids=torch.tensor([s]), fill_color=None, stroke_color=rgba[s]))
    return shapes, groups


def test_stroke_scene_args_layout_equals_per_scene_packing():
    g = torch.Generator().manual_seed(1)
    for npts, ncp in ((4, (2,)), (7, (2, 2)), (2, (0,)), (5, (0, 1, 0))):
        bs, ns, canvas = 3, 4, 48
        pts = torch.rand(bs, ns, npts, 2, generator=g) * canvas
        widths = torch.rand(bs, ns, generator=g) + 0.5
        rgba = torch.rand(bs, ns, 4, generator=g)
        packed, params = pydiffvg.stroke_scene_args(pts, widths, rgba, canvas, ncp)
        assert params.shape == (bs, packed.num_params)
        for k in range(bs):
            shapes, groups = _holders(pts[k], widths[k], rgba[k], ncp)
            topo, p = scene_pack.pack_scene_numpy(canvas, canvas, shapes, groups, 0, torch.tensor(0.5))
            assert np.array_equal(topo, packed.topo)
            assert np.array_equal(p, params[k].numpy())
    # differentiable with respect to the batch tensors
    pts.requires_grad_(True)
    _, params = pydiffvg.stroke_scene_args(pts, widths, rgba, canvas, ncp)
    params.sum().backward()
    assert float(pts.grad.sum()) == pts.numel()


def test_serialize_scenes_stacks_rows_and_rejects_mixed_topologies():
    a = scenes.batched_strokes(0)
    b = scenes.batched_strokes(1)
    packed, params = pydiffvg.serialize_scenes(64, 64, [(a[2], a[3]), (b[2], b[3])])
    assert params.shape == (2, packed.num_params)
    assert np.array_equal(params[1].numpy(), util.pack(b)[1]) and np.array_equal(packed.topo, util.pack(a)[0])
    c = scenes.batched_strokes(2, num_strokes=15)
    with pytest.raises(ValueError):
        pydiffvg.serialize_scenes(64, 64, [(a[2], a[3]), (c[2], c[3])])


def test_lazy_scenes_list():
    from diffvg_b200.pydiffvg.batched import _LazyScenes
    pts = torch.rand(3, 2, 4, 2)
    ls = _LazyScenes(pts, torch.rand(3, 2), torch.rand(3, 2, 4), 32, (2,))
    assert len(ls) == 3 and len(list(ls)) == 3 and len(ls[0:2]) == 2
    cw, ch, shapes, groups = ls[2]
    assert cw == 32 and len(shapes) == 2 and torch.equal(shapes[1].points, pts[2, 1]) and groups[1].fill_color is None
