"""Bit-exact BVH / CDF / indexing leg of the parity contract (BASELINE.json north_star): every
selector of dvg_scene_dump, copied back from DEVICE memory, against the same dump of the compiled
reference's Scene (oracle/_ref via oracle/ref_capi.cpp dvgref_scene_dump).

Reference: scene.cpp:207-333 (lengths, CDFs, PMFs, point-id maps, sample ids), 431-494 (morton2D,
build_bvh), 496-684 (leaf boxes, the three BVH levels).  Trees compare word for word when their sort keys
are unique; std::sort leaves the order of equal keys to the implementation, so a tree with tied keys is
compared as leaf SETS plus its root box / radius."""
import numpy as np
import pytest
import torch

import oracle_check
import scenes
import util

pytestmark = pytest.mark.gpu


def kat_scene():
    """SURVEY 8c KAT: three open cubic strokes on a 256^2 canvas."""
    from diffvg_b200 import pydiffvg
    pts = [torch.tensor([[10., 10.], [20., 40.], [40., 20.], [60., 60.]]),
           torch.tensor([[100., 20.], [120., 30.], [110., 60.], [90., 80.], [70., 100.], [60., 120.], [80., 140.]]),
           torch.tensor([[200., 200.], [180., 220.], [220., 240.], [240., 210.]])]
    ncp = [torch.tensor([2]), torch.tensor([2, 2]), torch.tensor([2])]
    widths = [2.0, 1.5, 3.0]
    shapes = [pydiffvg.Path(num_control_points=ncp[i], points=pts[i], is_closed=False, stroke_width=torch.tensor(widths[i]))
              for i in range(3)]
    groups = [pydiffvg.ShapeGroup(torch.tensor([i]), fill_color=None, stroke_color=torch.tensor([0.1, 0.2, 0.3, 1.0]))
              for i in range(3)]
    return 256, 256, shapes, groups


def many_shape_groups():
    """Groups holding several shapes (group BVHs with more than one leaf, odd and even counts), a transform,
    per-point thickness, a closed polygon: exercises every leaf kind of the three levels."""
    from diffvg_b200 import pydiffvg
    g = torch.Generator().manual_seed(5)
    shapes, groups = [], []
    sid = 0
    for k, count in enumerate([1, 2, 3, 5, 8, 13]):
        ids = []
        for j in range(count):
            kind = (k + j) % 4
            c = torch.rand(2, generator=g) * 200 + 20
            if kind == 0:
                shapes.append(pydiffvg.Circle(radius=torch.rand(1, generator=g)[0] * 10 + 2, center=c, stroke_width=torch.tensor(1.0 + j)))
            elif kind == 1:
                shapes.append(pydiffvg.Rect(p_min=c - 5 - j, p_max=c + 7 + j, stroke_width=torch.tensor(0.5 + j)))
            elif kind == 2:
                nseg = 1 + (j % 4)
                pts = c + (torch.rand(3 * nseg + 1, 2, generator=g) - 0.5) * 60
                sw = torch.rand(3 * nseg + 1, generator=g) * 3 + 0.5 if j % 2 else torch.tensor(2.0)
                shapes.append(pydiffvg.Path(num_control_points=torch.zeros(nseg, dtype=torch.int32) + 2, points=pts,
                                            is_closed=False, stroke_width=sw))
            else:
                pts = c + (torch.rand(5, 2, generator=g) - 0.5) * 40
                shapes.append(pydiffvg.Polygon(points=pts, is_closed=True, stroke_width=torch.tensor(1.5)))
            ids.append(sid)
            sid += 1
        xf = torch.eye(3)
        if k % 2:
            xf = torch.tensor([[0.9, -0.2, 10.0 * k], [0.2, 0.9, -5.0 * k], [0.0, 0.0, 1.0]])
        groups.append(pydiffvg.ShapeGroup(torch.tensor(ids), fill_color=torch.rand(4, generator=g) if k % 3 else None,
                                          stroke_color=torch.rand(4, generator=g) if k % 3 != 1 else None, shape_to_canvas=xf))
    return 256, 256, shapes, groups


SCENES = [('kat', kat_scene), ('zoo', scenes.zoo), ('painterly64', lambda: scenes.painterly(64, 256)),
          ('blobs48', lambda: scenes.blobs(48, 256)), ('many_shape_groups', many_shape_groups),
          ('painterly2048', scenes.painterly)]


def _compare_tree(ref, got, what, index):
    ref = ref.reshape(-1, 7)
    got = got.reshape(-1, 7)
    assert ref.shape == got.shape, (what, index)
    n = (ref.shape[0] + 1) // 2
    if np.array_equal(ref, got):
        return 'exact'
    # differences are only legitimate among leaves whose sort keys tie: same leaf multiset, same root
    rl = sorted(map(tuple, ref[:n].tolist()))
    gl = sorted(map(tuple, got[:n].tolist()))
    assert rl == gl, 'leaf sets differ (what=%d index=%d)' % (what, index)
    assert np.array_equal(ref[-1, 2:], got[-1, 2:]), 'root box / radius differ (what=%d index=%d)' % (what, index)
    # ... and the keys must really tie: the leaf order differs only inside runs of equal keys
    boxes = ref[:n, 2:6].view(np.float32)
    if what == 2:
        key = 0.5 * (boxes[:, 1] + boxes[:, 3])
        assert len(np.unique(key)) < n, 'tree differs although all y-centre keys are unique'
    return 'ties'


@pytest.mark.parametrize('case', SCENES, ids=[c[0] for c in SCENES])
def test_scene_dump_matches_reference_bit_for_bit(case):
    name, mk = case
    topo, params = util.pack(mk())
    ns, ng = int(topo[3]), int(topo[4])
    from diffvg_b200 import scene_pack
    shape_recs = topo[topo[scene_pack.H_OFF_SHAPES]:topo[scene_pack.H_OFF_SHAPES] + ns * scene_pack.S_LEN].reshape(ns, scene_pack.S_LEN)
    paths = [s for s in range(ns) if shape_recs[s, 0] == scene_pack.SHAPE_PATH]
    cap_g, cap_p = (ng, paths) if ns <= 64 else (min(ng, 40), paths[:40])
    sel = [(0, 0)] + [(1, g) for g in range(cap_g)] + [(2, s) for s in cap_p]
    sel += [(3, 0), (4, 0), (5, 0), (9, 0), (10, 0)]
    for s in cap_p:
        sel += [(6, s), (7, s), (8, s)]
    got = util.gpu_scene_dump(topo, params, sel)
    verdicts = {}
    for (what, index), g in zip(sel, got):
        r = oracle_check.scene_dump(topo, params, what, index)
        if what <= 2:
            v = _compare_tree(r, g, what, index)
            verdicts[v] = verdicts.get(v, 0) + 1
        else:
            assert np.array_equal(r, g), 'selector %d index %d differs' % (what, index)
    assert verdicts.get('exact', 0) > 0
    if name == 'kat':   # the survey's known answers, straight from device memory
        f = lambda a: a.view(np.float32)
        nodes = got[0].reshape(-1, 7)
        assert nodes[:, :2].view(np.int32).tolist() == [[0, -1], [1, -1], [2, -1], [0, 1], [3, 2]]
        assert f(nodes[4, 2:]).tolist() == [10.0, 10.0, 240.0, 240.0, 3.0]
        assert verdicts == {'exact': len([s for s in sel if s[0] <= 2])}


def test_scene_dump_error_paths():
    topo, params = util.pack(scenes.zoo())
    with pytest.raises(RuntimeError):
        util.gpu_scene_dump(topo, params, [(11, 0)])
    with pytest.raises(RuntimeError):
        util.gpu_scene_dump(topo, params, [(2, 0)])   # shape 0 of the zoo is a circle: no path BVH
    with pytest.raises(RuntimeError):
        util.gpu_scene_dump(topo, params, [(1, 99)])
