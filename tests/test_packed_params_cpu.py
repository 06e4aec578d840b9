"""CPU tests of pydiffvg.PackedParams (the packed-parameter fast path of the pydiffvg face): same topology and
parameter layout as `serialize_scene`, holders re-pointed at views, a handful of leaves for the optimiser."""
import warnings

import numpy as np
import torch

import scenes
import util
from diffvg_b200 import pydiffvg, scene_pack


def test_layout_equals_serialize_scene_and_holders_are_views():
    for mk in (scenes.zoo, lambda: scenes.painterly(64, 128), scenes.zoo_prefilter):
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            scene = mk()
            topo0, params0 = util.pack(scene)
            cw, ch, shapes, groups = scene
            pp = pydiffvg.PackedParams(cw, ch, shapes, groups)
            packed, params = pp.scene_args()
            assert np.array_equal(packed.topo, topo0) and np.array_equal(params.detach().numpy(), params0)
            assert packed.num_params == params0.shape[0] and packed.needs_xform_grad is False
            # the stock walk over the re-pointed holders gathers the same array (the views alias the leaves)
            filt = pydiffvg.PixelFilter(type=pydiffvg.FilterType.box, radius=pp.filter.radius)
            packed2, params2 = pydiffvg.RenderFunction.serialize_scene(cw, ch, shapes, groups, filter=filt)
            assert np.array_equal(packed2.topo, topo0) and np.array_equal(params2.detach().numpy(), params0)
            # an optimiser step on the leaves is seen by every holder
            opt = torch.optim.SGD(pp.parameters(), lr=0.5)
            flat = pp.flat()
            (flat * torch.arange(flat.numel(), dtype=torch.float32)).sum().backward()
            opt.step()
            after = pp.flat().detach().numpy()
            assert not np.array_equal(after, params0)
            _, params3 = pydiffvg.RenderFunction.serialize_scene(cw, ch, shapes, groups, filter=filt)
            assert np.array_equal(params3.detach().numpy(), after)


def test_leaves_by_kind_and_in_place_clamp():
    cw, ch, shapes, groups = scenes.painterly(32, 128)
    pp = pydiffvg.PackedParams(cw, ch, shapes, groups)
    assert pp.points.numel() == sum(2 * s.points.shape[0] for s in shapes)
    assert pp.scalars.numel() == 32 and pp.colors.numel() == 4 * 32 and pp.transforms.numel() == 9
    assert pp.others.numel() == 0 and pp.filter_radius.numel() == 1      # the pixel-filter radius: a leaf of its own
    assert [l.requires_grad for l in pp.leaves] == [True, True, True, False, False, False]
    assert len(pp.parameters()) == 3
    shapes[3].stroke_width.data.clamp_(0.0, 0.25)      # painterly_rendering.py clamps through the holders
    assert float(pp.scalars.detach()[3]) == 0.25
    with torch.no_grad():
        pp.colors.clamp_(0.2, 0.3)
    assert float(groups[5].stroke_color.min()) >= 0.2 and float(groups[5].stroke_color.max()) <= 0.3 + 1e-6
    # gradients arrive per kind
    f = pp.flat()
    f.sum().backward()
    assert float(pp.points.grad.sum()) == pp.points.numel() and pp.transforms.grad is None
    # scene_args is cheap and stable: same topology key object, no scene walk
    a, b = pp.scene_args()[0], pp.scene_args(use_prefiltering=True)[0]
    assert a.topo_key is b.topo_key and b.use_prefiltering is True
