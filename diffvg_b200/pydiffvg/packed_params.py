"""`PackedParams`: the scene's continuous parameters as a handful of leaf tensors the optimiser steps
directly, with the holders (`Path.points`, `ShapeGroup.stroke_color`, ...) re-pointed at views of them.

Why: `RenderFunction.serialize_scene` keeps the reference's calling convention -- thousands of small user
tensors, gathered by a differentiable concatenation every iteration and handed their gradients one
`AccumulateGrad` at a time.  At 2048 paths that is ~25 ms of host work per iteration around a ~7 ms render
(the reference pays 0.55 s there: render_pytorch.py:206-363 forward glue, 713-866 gradient read-back).
The renderer itself consumes ONE flat float32 array in the layout of include/dvg_scene_format.h and returns
ONE gradient array of the same layout, so the fast path is to keep the parameters in that layout:

    pp = pydiffvg.PackedParams(canvas_width, canvas_height, shapes, shape_groups)
    optim = torch.optim.Adam([{'params': [pp.points], 'lr': 1.0}, {'params': [pp.scalars], 'lr': 0.1},
                              {'params': [pp.colors], 'lr': 0.01}])
    for t in range(num_iter):
        optim.zero_grad()
        img = pydiffvg.RenderFunction.apply(w, h, 2, 2, t, None, *pp.scene_args())     # O(1), no scene walk
        loss(img).backward()                                                           # a handful of leaves, not 6144
        optim.step()
        pp.scalars.data.clamp_(1.0, max_width)     # the holders see every update: they are views

One leaf per KIND of parameter (all path points / all scalar stroke widths / all constant colours / the
transforms / everything else: per-point thickness, circle, ellipse, rect and gradient-colour parameters /
the pixel-filter radius; the last two kinds of leaves do not require a gradient unless asked to), so that the usual per-kind learning rates stay one optimiser group each.
The stock `serialize_scene` path is unchanged and accepts these holders too (it just gathers the views).
"""
import numpy as np
import torch

from .. import scene_pack
from . import pixel_filter as _pixel_filter
from .diffvg_enums import FilterType

__all__ = ['PackedParams']


class PackedParams:
    def __init__(self, canvas_width, canvas_height, shapes, shape_groups,
                 filter=None, device=None, requires_grad=True):
        """Packs the scene once (same walk and validation as `serialize_scene`), moves every parameter into one of
        six flat leaf tensors and replaces the holders' tensor attributes by views into them.  `device`: where the
        leaves live (default: where the first parameter tensor lives; CPU leaves are uploaded per iteration as one
        pinned copy).  The topology (shape counts, segment types, which colours are gradients, ...) is frozen."""
        if filter is None:
            filter = _pixel_filter.PixelFilter(type=FilterType.box, radius=torch.tensor(0.5))
        self.canvas_width, self.canvas_height = canvas_width, canvas_height
        self.shapes, self.shape_groups, self.filter = shapes, shape_groups, filter
        topo, bk, _ = scene_pack._pack_scene_full(canvas_width, canvas_height, shapes, shape_groups,
                                                  int(filter.type), filter.radius)
        topo.setflags(write=False)
        self.topo = topo
        if device is None:
            device = next((ts[0].device for ts in bk.tensors if ts), torch.device('cpu'))
        self.device = torch.device(device)
        self.leaves = []
        holders = (shapes, shape_groups)
        moved = {}   # id(original tensor) -> view, for tensor objects shared between holders (shape_to_canvas)
        with torch.no_grad():
            for b in range(scene_pack.NUM_BUCKETS):
                flat = scene_pack._flatten_bucket(b, bk.tensors[b], self.device)
                flat = torch.zeros(0, device=self.device) if flat is None else flat.detach().clone()
                self.leaves.append(flat)
                store = flat.detach()   # the holders' views alias the leaf's storage without being autograd views of it
                off = 0
                for t, (code, index, attr, sub, numel) in zip(bk.tensors[b], bk.sources[b]):
                    view = store[off:off + numel].view(t.shape if isinstance(t, torch.Tensor) and t.numel() == numel else (numel,))
                    off += numel
                    if code == scene_pack.SRC_FILTER:
                        filter.radius = view
                    else:
                        owner = holders[code][index]
                        if sub is not None:
                            owner = getattr(owner, attr)
                            setattr(owner, sub, view)
                        else:
                            setattr(owner, attr, view)
                    if isinstance(t, torch.Tensor):
                        moved[id(t)] = (t, view)
            for g in shape_groups:   # transforms stored once but referenced by many groups
                hit = moved.get(id(g.shape_to_canvas))
                if hit is not None and hit[0] is g.shape_to_canvas:
                    g.shape_to_canvas = hit[1]
        for leaf in self.leaves:
            if leaf.numel():
                leaf.requires_grad_(requires_grad)
        self.leaves[scene_pack.B_MAT3].requires_grad_(False)   # opt in: every boundary sample adds 9 terms to its group's transform
        self.leaves[scene_pack.B_FILTER].requires_grad_(False)   # opt in (apps/optimize_pixel_filter.py)
        self.num_params = int(topo[scene_pack.H_NPARAMS])
        assert sum(l.numel() for l in self.leaves) == self.num_params

    # one leaf per kind of parameter
    points = property(lambda self: self.leaves[scene_pack.B_POINTS], doc='all path points, flat [2 * total points]')
    scalars = property(lambda self: self.leaves[scene_pack.B_SCALAR], doc='scalar stroke widths')
    colors = property(lambda self: self.leaves[scene_pack.B_VEC4], doc='constant fill / stroke colours, flat [4 * n]')
    transforms = property(lambda self: self.leaves[scene_pack.B_MAT3], doc='shape_to_canvas matrices, flat [9 * n]; requires_grad off by default')
    others = property(lambda self: self.leaves[scene_pack.B_GENERIC],
                      doc='per-point thickness, circle / ellipse / rect parameters, gradient-colour parameters')
    filter_radius = property(lambda self: self.leaves[scene_pack.B_FILTER], doc='the pixel-filter radius [1]; requires_grad off by default')

    def parameters(self):
        """The non-empty leaves that require a gradient (what to hand to an optimiser)."""
        return [l for l in self.leaves if l.numel() and l.requires_grad]

    def flat(self):
        """`params` in the layout of include/dvg_scene_format.h (differentiable w.r.t. the leaves)."""
        parts = [l for l in self.leaves if l.numel()]
        return torch.cat(parts) if len(parts) > 1 else parts[0] * 1.0

    def scene_args(self, output_type=None, use_prefiltering=False, eval_positions=None):
        """What `RenderFunction.serialize_scene` returns, without walking the scene: splat into `apply`."""
        from .render_pytorch import OutputType, PackedScene
        if output_type is None:
            output_type = OutputType.color
        if eval_positions is None:
            eval_positions = torch.tensor([])
        packed = PackedScene(self.topo, self.canvas_width, self.canvas_height, output_type, use_prefiltering, eval_positions,
                             topo_key=self._topo_key())
        packed.needs_xform_grad = bool(self.transforms.requires_grad)
        packed.needs_filter_grad = bool(self.filter_radius.requires_grad)
        packed.filter_radius = float(self.filter.radius)
        packed.halo_rows = max(1, int(np.ceil(packed.filter_radius)))
        return [packed, self.flat()]

    def _topo_key(self):
        key = self.__dict__.get('_key')
        if key is None:
            key = self.__dict__['_key'] = self.topo.tobytes()
        return key
