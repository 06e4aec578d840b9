"""CPU-only tests (`-m "not gpu"`): the oracle against the reference's known answers and the
committed golden vectors, the host-side packing logic, and the C-ABI library's exported symbols.
No kernel is launched here."""
import ctypes
import glob
import os
import re

import numpy as np
import pytest
import torch

import emul
import ref_oracle
import scenes
import util
from diffvg_b200 import scene_pack

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = sorted(glob.glob(os.path.join(ROOT, 'tests', 'golden', '*.npz')))

needs_ref = pytest.mark.skipif(not ref_oracle.available(), reason='oracle/_ref not built (needs /root/reference)')

# (idx, seed) -> (state after init, first two floats), validated against pcg.h
PCG_KAT = [
    (0, 0, 0xf6e7b88658a69fc9, 0.452188373, 0.983064532),
    (1, 0, 0xa78ba0e0f1d19e25, 0.0703772306, 0.112021565),
    (0, 1, 0x4f39acb3a53c1ef6, 0.863092065, 0.755336404),
    (262143, 1, 0x400029050581209a, 0.628906608, 0.358397841),
    (4194303, 7, 0xc727c8286e921ba8, 0.997012258, 0.773138762),
]


def test_pcg_known_answers_product_header():
    """pcg.h:11-40 restated in dvg_common.cuh, compiled for the host by tests/host_emul."""
    for idx, seed, state, rx, ry in PCG_KAT:
        st, x, y = emul.pcg(idx, seed)
        assert st == state
        assert np.float32(x) == np.float32(rx) and np.float32(y) == np.float32(ry)


def test_golden_fixtures_exist():
    assert len(GOLDEN) >= 8


@needs_ref
@pytest.mark.parametrize('path', GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_reference_reproduces_golden(path):
    """The compiled reference (oracle/_ref) still produces the committed vectors: forward bit-exact
    (SURVEY Q20), gradients to atomic-order noise."""
    g = np.load(path)
    W, H, nsx, nsy, seed, ft = [int(v) for v in g['config']]
    bg = None
    if 'd_background' in g.files:
        from golden.make_golden import background_for
        bg = background_for(os.path.basename(path)[:-4], H, W)
    from golden.make_golden import d_image_for
    img = ref_oracle.render(g['topo'], g['params'], W, H, nsx, nsy, seed, background=bg)['image']
    if ft == 0:
        assert np.array_equal(img, g['image'])
    else:
        assert np.abs(img - g['image']).max() <= 1e-6
    d_img = d_image_for(os.path.basename(path)[:-4], H, W)
    bwd = ref_oracle.render(g['topo'], g['params'], W, H, nsx, nsy, seed, background=bg, d_render_image=d_img)
    assert util.rel_l2(g['d_params'], bwd['d_params']) <= 1e-5


@needs_ref
def test_reference_known_image_sums():
    """SURVEY 8c known answers captured from the reference's CPU build (256^2, 2x2 spp, seed 0)."""
    def s(scene, **kw):
        topo, params = util.pack(scene, kw.pop('ft', 0), kw.pop('fr', 0.5))
        return ref_oracle.render(topo, params, 256, 256, 2, 2, 0)['image']
    img = s(scenes.single_circle())
    assert abs(img.astype(np.float64).sum() - 11052.250240) < 1e-3
    assert np.allclose(img[128, 128], [.3, .6, .3, 1.0], atol=1e-6)
    with pytest.warns(Warning):
        img = s(scenes.single_stroke())
    assert abs(img.astype(np.float64).sum() - 4938.100155) < 1e-3
    img = s(scenes.single_stroke([10., 5., 4., 20.], fill=False))
    assert abs(img.astype(np.float64).sum() - 10697.300335) < 1e-3
    assert np.allclose(img[60, 135], [.6, .3, .6, .8], atol=1e-6)


@needs_ref
def test_reference_bvh_cdf_known_answers():
    """SURVEY 8c BVH / CDF KAT (3 open cubic stroke paths)."""
    from diffvg_b200 import pydiffvg
    P = [([[10, 10], [20, 40], [40, 20], [60, 60]], [2], 2.0),
         ([[100, 20], [120, 30], [110, 60], [90, 80], [70, 100], [60, 120], [80, 140]], [2, 2], 1.5),
         ([[200, 200], [180, 220], [220, 240], [240, 210]], [2], 3.0)]
    shapes, groups = [], []
    for i, (pts, ncp, w) in enumerate(P):
        shapes.append(pydiffvg.Path(torch.tensor(ncp), torch.tensor(pts, dtype=torch.float32), False, torch.tensor(w)))
        groups.append(pydiffvg.ShapeGroup(torch.tensor([i]), None, stroke_color=torch.tensor([0., 0., 0., 1.])))
    topo, params = util.pack((256, 256, shapes, groups))
    f = lambda w: w.view(np.float32)
    lengths = f(ref_oracle.scene_dump(topo, params, 3))
    assert np.allclose(lengths, [72.1463623, 137.820312, 67.0436325], rtol=1e-7)
    assert np.allclose(f(ref_oracle.scene_dump(topo, params, 4)), [0.260446489, 0.757974207, 1.0], rtol=1e-7)
    assert np.allclose(f(ref_oracle.scene_dump(topo, params, 5)), [0.260446489, 0.497527719, 0.242025763], rtol=1e-7)
    assert np.allclose(f(ref_oracle.scene_dump(topo, params, 6, 1)), [0.50051403, 1.0], rtol=1e-7)
    assert list(ref_oracle.scene_dump(topo, params, 8, 1).view(np.int32)) == [0, 3]
    nodes = ref_oracle.scene_dump(topo, params, 0).reshape(-1, 7)
    assert [tuple(n[:2].view(np.int32)) for n in nodes] == [(0, -1), (1, -1), (2, -1), (0, 1), (3, 2)]
    assert np.allclose(nodes[4][2:].view(np.float32), [10, 10, 240, 240, 3])
    pnodes = ref_oracle.scene_dump(topo, params, 2, 1).reshape(-1, 7)
    assert [tuple(n[:2].view(np.int32)) for n in pnodes] == [(0, -1), (1, -4), (0, 1)]
    # the product's host-compiled build functions produce the same tables, bit for bit
    for what, idx in ((3, 0), (4, 0), (5, 0), (6, 1), (7, 1), (8, 1)):
        assert np.array_equal(ref_oracle.scene_dump(topo, params, what, idx), emul.scene_dump(topo, params, what, idx))


@pytest.mark.parametrize('path', GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_product_arithmetic_on_host_matches_golden(path):
    """The product's predicate / tracer / boundary headers, compiled for the host
    (tests/host_emul), against the reference's golden vectors.  Forward tolerance 1e-5 abs
    (north_star); gradients per-tensor rel-L2 1e-4."""
    g = np.load(path)
    name = os.path.basename(path)[:-4]
    W, H, nsx, nsy, seed, ft = [int(v) for v in g['config']]
    from golden.make_golden import background_for, d_image_for
    bg = background_for(name, H, W) if 'd_background' in g.files else None
    img = emul.render(g['topo'], g['params'], W, H, nsx, nsy, seed, background=bg)['image']
    assert np.abs(img - g['image']).max() <= 1e-5
    bwd = emul.render(g['topo'], g['params'], W, H, nsx, nsy, seed, background=bg, d_render_image=d_image_for(name, H, W))
    assert util.rel_l2(g['d_params'], bwd['d_params']) <= 1e-4
    if bg is not None and nsx * nsy == 1:  # Q2: d_background is racy in the reference above 1 spp
        assert np.abs(bwd['d_background'] - g['d_background']).max() <= 1e-5


# ------------------------------------------------------------------ host logic: scene packing
def test_pack_layout_and_semantics():
    from diffvg_b200 import pydiffvg
    cw, ch, shapes, groups = scenes.zoo()
    topo, buckets = scene_pack.pack_scene(cw, ch, shapes, groups, 1, torch.tensor(1.5))
    params = scene_pack.concat_params(buckets)
    assert topo[scene_pack.H_MAGIC] == scene_pack.TOPO_MAGIC
    assert topo[scene_pack.H_NS] == len(shapes) and topo[scene_pack.H_NG] == len(groups)
    assert topo[scene_pack.H_NPARAMS] == params.numel()
    assert params[topo[scene_pack.H_FRAD_OFF]].item() == 1.5
    srec = topo[topo[scene_pack.H_OFF_SHAPES]:topo[scene_pack.H_OFF_GROUPS]].reshape(-1, scene_pack.S_LEN)
    # shape 4: per-point thickness -> no scalar width, thickness offset set (render_pytorch.py:67-72, 95-98)
    assert srec[4][0] == scene_pack.SHAPE_PATH and srec[4][2] == -1 and srec[4][3] >= 0
    th = params[srec[4][3]:srec[4][3] + srec[4][4]]
    assert torch.equal(th, shapes[4].stroke_width)
    # shape 5: closed polygon of 4 points -> 4 line segments; shape 7: open polygon -> 3 (render_pytorch.py:75-86)
    assert srec[5][5] == 4 and srec[7][5] == 3
    ncp = topo[topo[scene_pack.H_OFF_NCP] + srec[5][6]: topo[scene_pack.H_OFF_NCP] + srec[5][6] + 4]
    assert (ncp == 0).all()
    pts = params[srec[3][1]:srec[3][1] + 2 * srec[3][4]].reshape(-1, 2)
    assert torch.equal(pts, shapes[3].points)
    grec = topo[topo[scene_pack.H_OFF_GROUPS]:topo[scene_pack.H_OFF_NCP]].reshape(-1, scene_pack.G_LEN)
    assert grec[1][2] == scene_pack.COLOR_LINEAR and grec[1][4] == 3
    assert grec[4][2] == scene_pack.COLOR_NONE and grec[4][5] == scene_pack.COLOR_CONSTANT
    assert grec[5][1] == 2 and list(topo[topo[scene_pack.H_OFF_GSHAPES] + grec[5][0]:][:2]) == [5, 6]
    xf = params[grec[2][9]:grec[2][9] + 9].reshape(3, 3)
    assert torch.equal(xf, groups[2].shape_to_canvas)


def test_pack_gradients_flow_to_user_tensors():
    from diffvg_b200 import pydiffvg
    pts = torch.rand(4, 2, requires_grad=True)
    w = torch.tensor(2.0, requires_grad=True)
    col = torch.rand(4, requires_grad=True)
    shapes = [pydiffvg.Path(torch.tensor([2]), pts, False, w)]
    groups = [pydiffvg.ShapeGroup(torch.tensor([0]), None, stroke_color=col)]
    topo, buckets = scene_pack.pack_scene(64, 64, shapes, groups)
    params = scene_pack.concat_params(buckets)
    g = torch.arange(params.numel(), dtype=torch.float32)
    params.backward(g)
    srec = topo[topo[scene_pack.H_OFF_SHAPES]:][:scene_pack.S_LEN]
    assert torch.equal(pts.grad.reshape(-1), g[srec[1]:srec[1] + 8])
    assert w.grad.item() == g[srec[2]].item()
    grec = topo[topo[scene_pack.H_OFF_GROUPS]:][:scene_pack.G_LEN]
    assert torch.equal(col.grad, g[grec[6]:grec[6] + 4])


def _plain_concat(buckets, device):
    """The plain torch graph concat_params replaced (stack / cat per bucket): the reference for values and gradients."""
    parts = [f for f in (scene_pack._flatten_bucket(b, ts, device) for b, ts in enumerate(buckets)) if f is not None]
    return torch.cat(parts) if len(parts) > 1 else parts[0].clone()


def _leaves(buckets):
    seen, out = set(), []
    for b in buckets:
        for t in b:
            if id(t) not in seen and t.is_leaf and t.dtype.is_floating_point:
                seen.add(id(t))
                out.append(t)
    return out


@pytest.mark.parametrize('which', ['zoo', 'painterly', 'blobs'])
def test_concat_params_single_node_equals_plain_graph(which):
    scene = {'zoo': scenes.zoo, 'painterly': lambda: scenes.painterly(32, 64), 'blobs': lambda: scenes.blobs(16, 48)}[which]()
    cw, ch, shapes, groups = scene
    topo, buckets = scene_pack.pack_scene(cw, ch, shapes, groups, 0, torch.tensor(0.5, requires_grad=True))
    leaves = _leaves(buckets)
    was = [t.requires_grad for t in leaves]
    res = []
    try:
        for t in leaves:
            t.requires_grad_(True)
        for fn in (lambda: scene_pack.concat_params(buckets), lambda: _plain_concat(buckets, torch.device('cpu'))):
            for t in leaves:
                t.grad = None
            p = fn()
            assert p.dtype == torch.float32 and p.dim() == 1 and p.numel() == int(topo[scene_pack.H_NPARAMS])
            w = torch.linspace(-1.0, 2.0, p.numel())
            (p * w).sum().backward()
            res.append((p.detach().clone(), [t.grad.clone() for t in leaves]))
    finally:   # (the default identity transform is shared by every ShapeGroup of the process)
        for t, r in zip(leaves, was):
            t.grad = None
            t.requires_grad_(r)
    assert torch.equal(res[0][0], res[1][0])
    assert len(res[0][1]) == len(leaves) and all(g0.shape == t.shape for g0, t in zip(res[0][1], leaves))
    assert all(torch.equal(a, b) for a, b in zip(res[0][1], res[1][1]))


def test_concat_params_mixed_inputs():
    """Non-leaf inputs, float64 tensors, a tensor used twice, tensors that take no part in autograd, [1]-shaped scalars."""
    from diffvg_b200 import pydiffvg
    theta = torch.rand(3, 2, requires_grad=True)
    pts_a = theta * 20 + 5                                        # not a leaf
    pts_b = torch.rand(4, 2, dtype=torch.float64).mul(30).requires_grad_(True)
    wa = torch.tensor([2.0], requires_grad=True)                  # shape [1]
    wb = torch.tensor(1.5, requires_grad=True)                    # shape []
    col = torch.rand(4, requires_grad=True)                       # fill AND stroke colour of a group
    frozen = torch.rand(4)                                        # no gradient asked
    shapes = [pydiffvg.Path(torch.tensor([1]), pts_a, False, wa), pydiffvg.Path(torch.tensor([2]), pts_b, False, wb)]
    groups = [pydiffvg.ShapeGroup(torch.tensor([0]), col, stroke_color=col), pydiffvg.ShapeGroup(torch.tensor([1]), None, stroke_color=frozen)]
    topo, buckets = scene_pack.pack_scene(64, 64, shapes, groups)
    res = []
    for fn in (lambda: scene_pack.concat_params(buckets), lambda: _plain_concat(buckets, torch.device('cpu'))):
        for t in (theta, pts_b, wa, wb, col):
            t.grad = None
        p = fn()
        (p * torch.arange(1, p.numel() + 1, dtype=torch.float32)).sum().backward(retain_graph=True)   # (theta -> pts_a is shared by both runs)
        res.append((p.detach().clone(), [t.grad.clone() for t in (theta, pts_b, wa, wb, col)]))
        assert frozen.grad is None
    assert torch.equal(res[0][0], res[1][0])
    for a, b, t in zip(res[0][1], res[1][1], (theta, pts_b, wa, wb, col)):
        assert a.dtype == t.dtype and a.shape == t.shape and torch.equal(a, b)
    # nothing asks for a gradient: plain tensor out
    with torch.no_grad():
        assert not scene_pack.concat_params(buckets).requires_grad
    topo2, b2 = scene_pack.pack_scene(64, 64, [pydiffvg.Circle(torch.tensor(5.0), torch.tensor([8.0, 9.0]))],
                                      [pydiffvg.ShapeGroup(torch.tensor([0]), torch.tensor([0.1, 0.2, 0.3, 1.0]))])
    assert not scene_pack.concat_params(b2).requires_grad


def test_shared_transform_is_stored_once():
    cw, ch, shapes, groups = scenes.painterly(8, 64)
    eye = torch.eye(3)
    for g in groups:
        g.shape_to_canvas = eye
    topo, buckets = scene_pack.pack_scene(cw, ch, shapes, groups)
    assert len(buckets[scene_pack.B_MAT3]) == 1
    grec = topo[topo[scene_pack.H_OFF_GROUPS]:topo[scene_pack.H_OFF_NCP]].reshape(-1, scene_pack.G_LEN)
    assert len(set(grec[:, 9])) == 1


def test_pack_rejects_bad_input():
    from diffvg_b200 import pydiffvg
    c = pydiffvg.Circle(torch.tensor(3.0), torch.tensor([1.0, 2.0]))
    with pytest.raises(ValueError):
        scene_pack.pack_scene(8, 8, [c], [pydiffvg.ShapeGroup(torch.tensor([1]), torch.rand(4))])
    with pytest.raises(ValueError):
        scene_pack.pack_scene(8, 8, [c], [pydiffvg.ShapeGroup(torch.tensor([0]), torch.rand(3))])


# ------------------------------------------------------------------ C ABI surface
def _declared_functions():
    text = open(os.path.join(ROOT, 'include', 'diffvg_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(dvg_[a-z_0-9]+)\s*\(', text)))


def test_c_abi_library_exports_every_declared_symbol():
    lib_path = os.path.join(ROOT, 'diffvg_b200', 'libdiffvg_b200.so')
    if not os.path.exists(lib_path):
        import __graft_entry__
        __graft_entry__.build_library()
    lib = ctypes.CDLL(lib_path)
    names = _declared_functions()
    assert len(names) >= 10
    for n in names:
        assert hasattr(lib, n), 'missing export: ' + n
    lib.dvg_abi_version.restype = ctypes.c_int
    assert lib.dvg_abi_version() >= 1


def test_c_abi_rejects_malformed_topology_without_a_gpu():
    """Argument validation happens before any CUDA call, so it is testable on a CPU-only host."""
    from diffvg_b200 import _native as n
    h = ctypes.c_void_p()
    bad = np.zeros(4, np.int32)
    assert n.lib.dvg_scene_create(bad.ctypes.data, bad.shape[0], 0, ctypes.byref(h)) == 1
    assert b'header' in n.lib.dvg_last_error()
    topo, _ = util.pack(scenes.single_circle())
    topo = topo.copy()
    topo[scene_pack.H_MAGIC] = 0
    assert n.lib.dvg_scene_create(topo.ctypes.data, topo.shape[0], 0, ctypes.byref(h)) == 1
    from diffvg_b200 import pydiffvg
    e = pydiffvg.Ellipse(torch.tensor([3.0, 2.0]), torch.tensor([1.0, 2.0]), stroke_width=torch.tensor(1.0))
    topo, _ = util.pack((8, 8, [e], [pydiffvg.ShapeGroup(torch.tensor([0]), None, stroke_color=torch.rand(4))]))
    assert n.lib.dvg_scene_create(topo.ctypes.data, topo.shape[0], 0, ctypes.byref(h)) == n.DVG_ERR_UNSUPPORTED


def test_product_fails_loudly_without_cuda():
    from diffvg_b200 import pydiffvg
    if torch.cuda.is_available():
        pytest.skip('CUDA present')
    cw, ch, shapes, groups = scenes.single_circle()
    args = pydiffvg.RenderFunction.serialize_scene(cw, ch, shapes, groups)
    with pytest.raises(RuntimeError):
        pydiffvg.RenderFunction.apply(16, 16, 1, 1, 0, None, *args)


def test_pack_scene_memo_reuses_topology_and_tracks_changes():
    """The memoised front of pack_scene (scene_pack.py): same holders + new VALUES -> same topology
    object, tensors re-collected; any structural change -> a fresh full walk with a different blob."""
    import scenes
    cw, ch, shapes, groups = scenes.painterly(num_paths=16, canvas=64)
    fr = torch.tensor(0.5)
    topo0, b0 = scene_pack.pack_scene(cw, ch, shapes, groups, 0, fr)
    p0 = scene_pack.concat_params(b0)
    shapes[3].points = shapes[3].points + 1.0                 # new tensor object, same structure
    topo1, b1 = scene_pack.pack_scene(cw, ch, shapes, groups, 0, fr)
    assert topo1 is topo0
    p1 = scene_pack.concat_params(b1)
    assert p1.shape == p0.shape and not torch.equal(p0, p1)
    # against an un-memoised walk
    topo_full, bk, _ = scene_pack._pack_scene_full(cw, ch, shapes, groups, 0, fr)
    assert np.array_equal(topo_full, topo1)
    assert torch.equal(scene_pack.concat_params(bk.tensors), p1)
    # structural changes must miss the memo
    shapes[5].is_closed = True
    topo2, _ = scene_pack.pack_scene(cw, ch, shapes, groups, 0, fr)
    assert topo2 is not topo0 and not np.array_equal(topo2, topo0)
    shapes[5].is_closed = False
    groups[2].fill_color = torch.tensor([0.1, 0.2, 0.3, 1.0])
    with pytest.warns(Warning):
        topo3, b3 = scene_pack.pack_scene(cw, ch, shapes, groups, 0, fr)
    assert int(topo3[scene_pack.H_NPARAMS]) == int(topo0[scene_pack.H_NPARAMS]) + 4
    groups[2].fill_color = None
    shapes[7].stroke_width = torch.tensor([1.0, 2.0, 1.5, 0.5][:shapes[7].points.shape[0]] +
                                          [1.0] * max(0, shapes[7].points.shape[0] - 4))   # per-point thickness
    topo4, _ = scene_pack.pack_scene(cw, ch, shapes, groups, 0, fr)
    assert not np.array_equal(topo4, topo0)
    # a tensor whose element count no longer matches the recorded one falls back to the full walk (which raises)
    groups[1].stroke_color = torch.rand(3)
    with pytest.raises(ValueError):
        scene_pack.pack_scene(cw, ch, shapes, groups, 0, fr)


# ------------------------------------------------------------------ plain-C restatement (oracle/dvg_oracle.c)
def _c_oracle():
    import subprocess
    import c_oracle
    if not c_oracle.available():
        subprocess.check_call(['make', '-C', os.path.join(os.path.dirname(__file__), '..', 'oracle'), 'oracle'])
    return c_oracle


def test_c_restatement_pcg_known_answers():
    """(idx, seed) -> (state, rx, ry), validated against pcg.h."""
    c = _c_oracle()
    kats = [((0, 0), 0xf6e7b88658a69fc9, 0.452188373, 0.983064532),
            ((1, 0), 0xa78ba0e0f1d19e25, 0.0703772306, 0.112021565),
            ((0, 1), 0x4f39acb3a53c1ef6, 0.863092065, 0.755336404),
            ((262143, 1), 0x400029050581209a, 0.628906608, 0.358397841),
            ((4194303, 7), 0xc727c8286e921ba8, 0.997012258, 0.773138762)]
    for (idx, seed), state, rx, ry in kats:
        st, x, y = c.pcg(idx, seed)
        assert st == state and abs(x - rx) < 1e-8 and abs(y - ry) < 1e-8


@pytest.mark.parametrize('path', sorted(glob.glob(os.path.join(os.path.dirname(__file__), 'golden', '*.npz'))),
                         ids=lambda p: os.path.basename(p)[:-4])
def test_c_restatement_forward_matches_reference_golden(path):
    """The restatement is pinned against outputs of the compiled reference (tests/golden/*.npz, made by
    tests/golden/make_golden.py): forward images within the 1e-5 bar (bit-identical except where the pixel
    filter spreads a sample over several pixels and the summation order differs)."""
    c = _c_oracle()
    g = np.load(path)
    name = os.path.basename(path)[:-4]
    W, H, nsx, nsy, seed, ft = [int(v) for v in g['config']]
    from golden.make_golden import background_for
    bg = background_for(name, H, W) if 'd_background' in g.files else None
    img = c.render(g['topo'], g['params'], W, H, nsx, nsy, seed, background=bg)['image']
    assert np.abs(img - g['image']).max() <= 1e-5


def test_c_restatement_known_image_sums():
    """SURVEY 8c known answers from the reference CPU build (256^2, 2x2 spp, seed 0, float64 sums)."""
    c = _c_oracle()
    for scene, want in ((scenes.single_circle(), 11052.250240), (scenes.single_stroke(), 4938.100155)):
        topo, params = util.pack(scene)
        img = c.render(topo, params, 256, 256, 2, 2, 0)['image']
        assert abs(img.astype(np.float64).sum() - want) < 2e-3


def test_c_restatement_refuses_what_it_does_not_restate():
    c = _c_oracle()
    topo, params = util.pack(scenes.single_circle())
    with pytest.raises(RuntimeError):
        c.render(topo, params, 32, 32, 1, 1, 0, d_render_image=np.zeros((32, 32, 4), np.float32))


def test_tight_binning_never_drops_a_stroke_a_sample_could_touch():
    """dvg_buildfn.cuh bracket_reaches_tile (used by k_bin): a tile may drop a curved stroke only if NO point of the tile
    can be within the stroke.  Property check on random cubics and tiles with the product's own host-compiled code:
    whenever the tile test says "cannot reach", every sampled point of the tile is classified "certainly outside" by the
    polyline bracket and the exact closest-point test (the reference's algorithm) agrees."""
    import ctypes
    lib = emul._load()
    fp = ctypes.POINTER(ctypes.c_float)
    lib.emul_bracket_build.argtypes = [fp, ctypes.c_float, ctypes.c_float, fp]
    lib.emul_bracket_reaches_tile.argtypes = [fp] + [ctypes.c_float] * 4
    lib.emul_bracket_classify.argtypes = [fp, ctypes.c_float, ctypes.c_float]
    lib.emul_stroke_hit_cubic.argtypes = [fp, fp, ctypes.c_float, ctypes.c_float]
    rng = np.random.RandomState(12)
    dropped = kept = 0
    for _ in range(300):
        p0 = rng.rand(2) * 64
        pts = [p0]
        for _k in range(3):
            pts.append(pts[-1] + (rng.rand(2) - 0.5) * 25.6)
        pts = np.asarray(pts, np.float32).reshape(-1)
        r = np.float32(0.5 + 3.5 * rng.rand())
        cap = np.zeros(64, np.float32)
        lib.emul_bracket_build(pts.ctypes.data_as(fp), r, r, cap.ctypes.data_as(fp))
        rad = np.full(4, r, np.float32)
        for _t in range(40):
            tw, th = [(8, 2), (8, 8), (16, 8), (16, 16)][rng.randint(4)]
            x0 = np.float32(rng.randint(-8, 72)); y0 = np.float32(rng.randint(-8, 72))
            x1, y1 = np.float32(x0 + tw), np.float32(y0 + th)
            if lib.emul_bracket_reaches_tile(cap.ctypes.data_as(fp), x0, y0, x1, y1):
                kept += 1
                continue
            dropped += 1
            xs = np.concatenate([rng.rand(24) * tw + x0, [x0, x1, x0, x1]]).astype(np.float32)
            ys = np.concatenate([rng.rand(24) * th + y0, [y0, y0, y1, y1]]).astype(np.float32)
            for x, y in zip(xs, ys):
                assert lib.emul_bracket_classify(cap.ctypes.data_as(fp), x, y) < 0
                assert lib.emul_stroke_hit_cubic(pts.ctypes.data_as(fp), rad.ctypes.data_as(fp), x, y) == 0
    assert dropped > 1000 and kept > 200   # the test is exercised both ways


def test_per_primitive_quintic_split_is_bit_identical():
    """dvg_geom.cuh prim_quintic / quintic_of: the sample-independent part of the closest-point quintic formed once per
    primitive (what the exact-test kernel reads) gives the SAME five normalised coefficients, bit for bit, as
    cubic_quintic on every (cubic, sample) pair; the split points from the per-primitive isolator record agree with
    quintic_intervals' floats except where a root sits on a float rounding boundary (the two closed forms differ by
    ~1e-16 before the polish); the verdict through the kernel's bracket enumeration equals stroke_hit_cubic's."""
    import ctypes
    lib = emul._load()
    fp = ctypes.POINTER(ctypes.c_float)
    lib.emul_quintic_split_check.argtypes = [fp, fp, fp, ctypes.c_int, ctypes.POINTER(ctypes.c_longlong)]
    rng = np.random.RandomState(21)
    total = coeff = split = verdict = 0
    for c in range(400):
        p0 = rng.rand(2) * 512
        pts = [p0]
        for _k in range(3):
            pts.append(pts[-1] + (rng.rand(2) - 0.5) * (25.6 if c % 4 else 200.0))
        if c % 50 == 7:
            pts[3] = pts[0] + 3 * (pts[2] - pts[1])       # q3 == 0: a quadratic in disguise (A = 0)
        pts = np.asarray(pts, np.float32).reshape(-1)
        rad = (0.5 + 3.5 * rng.rand(4)).astype(np.float32)
        n = 2000
        lo, hi = pts.reshape(4, 2).min(0) - 6, pts.reshape(4, 2).max(0) + 6
        xy = (rng.rand(n, 2) * (hi - lo) + lo).astype(np.float32)
        out = (ctypes.c_longlong * 3)()
        lib.emul_quintic_split_check(pts.ctypes.data_as(fp), rad.ctypes.data_as(fp), xy.ctypes.data_as(fp), n, out)
        total += n; coeff += out[0]; split += out[1]; verdict += out[2]
    assert coeff == 0
    assert split <= total * 1e-4, (split, total)
    assert verdict == 0


def test_fast_cubic_winding_agrees_with_the_reference_sequence_whenever_it_answers():
    """dvg_geom.cuh cubic_winding_fast: roots from a polished float closed form, answered only when every decision (t in
    [0, 1], crossing right of the sample, crossing direction) is clear by a margin; otherwise the reference's operation
    sequence runs (cubic_winding_exact, with its correctly-rounded fall-back).  On random cubics, cubics with round
    coordinates sampled on a regular grid (crossings exactly at control points: the adversarial case of SVG assets under
    prefiltering), near-degenerate (almost quadratic / almost straight) ones and samples at the curve's own y-extrema:
    every answered pair must agree, and most pairs must be answered."""
    import ctypes
    lib = emul._load()
    fp = ctypes.POINTER(ctypes.c_float)
    lib.emul_winding_fast_check.argtypes = [fp, fp, ctypes.c_int, ctypes.POINTER(ctypes.c_longlong)]
    rng = np.random.RandomState(33)
    tot = ans = bad = 0
    common = [0, 0]      # random / round-coordinate / monotone cubics: what assets are made of
    cert = [0, 0]        # pairs answered by the classifier's per-primitive certificate / of them wrong
    for c in range(600):
        kind = c % 6
        if kind == 0:      # random
            pts = (rng.rand(4, 2) * 200).astype(np.float32)
        elif kind == 1:    # round coordinates, grid samples
            pts = rng.randint(0, 64, (4, 2)).astype(np.float32)
        elif kind == 2:    # almost quadratic in y
            pts = (rng.rand(4, 2) * 100).astype(np.float32)
            pts[3, 1] = pts[0, 1] + 3 * (pts[2, 1] - pts[1, 1]) + np.float32(rng.randn() * 1e-4)
        elif kind == 3:    # almost horizontal
            pts = (rng.rand(4, 2) * 100).astype(np.float32)
            pts[:, 1] = np.float32(40.0) + (rng.randn(4) * 1e-3).astype(np.float32)
        elif kind == 4:    # monotone, long
            pts = np.stack([rng.rand(4) * 500, np.sort(rng.rand(4) * 500)], axis=1).astype(np.float32)
        else:              # loop / cusp
            pts = np.asarray([[10, 10], [90, 80], [10, 80], [90, 10]], np.float32) + (rng.randn(4, 2) * 3).astype(np.float32)
        n = 3000
        if kind == 1:
            xs = rng.randint(0, 128, n) * 0.5 + 0.25
            ys = rng.randint(0, 128, n) * 0.5 + (0.25 if c % 12 == 1 else 0.0)     # on and off the control points' rows
        else:
            lo, hi = pts.min(0) - 5, pts.max(0) + 5
            xs = rng.rand(n) * (hi[0] - lo[0]) + lo[0]
            ys = rng.rand(n) * (hi[1] - lo[1]) + lo[1]
            # samples exactly at the y of control points and of the curve's end points
            ys[:8] = np.repeat(pts[:, 1], 2)
        xy = np.stack([xs, ys], axis=1).astype(np.float32)
        if kind == 4:      # monotone: a share of the samples left of the control polygon (what the classifier's certificate answers)
            xs[: n // 2] = pts[:, 0].min() - 1.0 - rng.rand(n // 2) * 40
            xy = np.stack([xs, ys], axis=1).astype(np.float32)
        out = (ctypes.c_longlong * 4)()
        lib.emul_winding_fast_check(np.ascontiguousarray(pts.reshape(-1)).ctypes.data_as(fp), np.ascontiguousarray(xy).ctypes.data_as(fp), n, out)
        tot += n; ans += out[0]; bad += out[1]
        cert[0] += out[2]; cert[1] += out[3]
        if kind in (0, 1, 4):
            common[0] += n; common[1] += out[0]
    assert bad == 0, (bad, ans, tot)
    assert cert[1] == 0 and cert[0] > 50000, cert           # the certificate: never wrong, and exercised
    assert ans >= 0.7 * tot, (ans, tot)                 # (the almost-quadratic kind never takes the fast form)
    assert common[1] >= 0.97 * common[0], common


def test_guided_cdf_search_equals_the_reference_bisection():
    """dvg_boundary.cuh cdf_sample_guided (2048-entry guide table over the shape cdf) picks the same shape as cdf.h's
    bisection for every u: random cdfs with runs of equal entries (zero-length shapes), tiny and large tables, u on bucket
    edges and on cdf values."""
    import ctypes
    lib = emul._load()
    fp = ctypes.POINTER(ctypes.c_float)
    lib.emul_cdf_guided_check.argtypes = [fp, ctypes.c_int, fp, ctypes.c_int]
    lib.emul_cdf_guided_check.restype = ctypes.c_longlong
    rng = np.random.RandomState(3)
    for n in (1, 2, 3, 17, 2048, 5000, 8192):
        for trial in range(6):
            lens = rng.rand(n).astype(np.float32) ** (1 + 3 * trial)
            lens[rng.rand(n) < 0.2 * (trial % 3)] = 0.0
            if lens.sum() == 0:
                lens[0] = 1.0
            c = np.zeros(n, np.float32)
            acc = np.float32(0)
            for i in range(n):                       # sequential float prefix sum, as the scene build
                acc = np.float32(lens[i]) if i == 0 else np.float32(lens[i] + acc)
                c[i] = acc
            c = (c / c[-1]).astype(np.float32)
            us = np.concatenate([rng.rand(20000).astype(np.float32), (np.arange(2049) / 2048.0).astype(np.float32)[:-1],
                                 c[:min(n, 3000)], np.nextafter(c[:min(n, 3000)], np.float32(0)),
                                 np.float32([0.0, np.nextafter(np.float32(1), np.float32(0))])]).astype(np.float32)
            us = us[(us >= 0) & (us < 1)]
            assert lib.emul_cdf_guided_check(np.ascontiguousarray(c).ctypes.data_as(fp), n, np.ascontiguousarray(us).ctypes.data_as(fp), len(us)) == 0
