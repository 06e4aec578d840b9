"""Batched scenes: many small scenes of ONE topology rendered and differentiated by one set of kernel launches.

The reference has no batched entry point: its batched front end (apps/generative_models/rendering.py:170-237
`line_render`, :239-307 `bezier_render`) loops over the batch in Python, builds `Path` / `ShapeGroup` holders per stroke,
constructs a `diffvg.Scene` per sample and calls `RenderFunction.apply` once per sample (two `diffvg.render` calls each
with the backward pass).  Here the scenes of a batch share one int32 topology blob and differ only in their rows of a
`[batch, num_params]` float tensor (layout of include/dvg_scene_format.h) and in their seeds:

    packed, params = pydiffvg.serialize_scenes(canvas_w, canvas_h, [(shapes_0, groups_0), (shapes_1, groups_1), ...])
    images = pydiffvg.render_batch(w, h, 2, 2, seeds, None, packed, params)        # [batch, h, w, 4]

`bezier_render` / `line_render` keep the reference's signatures and build `params` with a handful of tensor ops on the
whole batch (no per-stroke Python).  Scene b of a batch renders exactly what the single-scene path renders for the same
parameters and seed (`tests/test_batched_gpu.py`).
"""
import random

import numpy as np
import torch

from . import shape as _shape
from .. import scene_pack
from . import render_pytorch as _rp
from .render_pytorch import OutputType, PackedScene, _cuda_device, _native, _NativeScene, _scene_cache

__all__ = ['BatchRenderFunction', 'render_batch', 'serialize_scenes', 'stroke_scene_args', 'bezier_render', 'line_render']


def _get_batch_scene(packed, device_index, batch):
    key = (packed.topo_key, device_index, batch)
    ns = _scene_cache.get(key)
    if ns is None:
        ns = _NativeScene(packed.topo, device_index, batch)
        _scene_cache[key] = ns
        while len(_scene_cache) > _rp._SCENE_CACHE_MAX:
            _scene_cache.popitem(last=False)
    else:
        _scene_cache.move_to_end(key)
    return ns


def _seed_array(seeds, batch):
    if seeds is None:
        seeds = [random.randint(0, 1000000) for _ in range(batch)]      # rendering.py:16-17, once per sample
    if isinstance(seeds, torch.Tensor):
        seeds = seeds.detach().cpu().numpy()
    arr = np.ascontiguousarray(np.asarray(seeds, dtype=np.uint64).reshape(-1))
    if arr.shape[0] != batch:
        raise ValueError('need one seed per scene: got %d for a batch of %d' % (arr.shape[0], batch))
    return arr


class BatchRenderFunction(torch.autograd.Function):
    """`RenderFunction` over a batch: `apply(width, height, nsx, nsy, seeds, background_images, packed, params)` with
    `params` [batch, num_params] and `background_images` None or [batch, height, width, 4]; returns [batch, height,
    width, 4].  Colour output of the sampled path (what the batched apps use)."""

    @staticmethod
    def forward(ctx, width, height, num_samples_x, num_samples_y, seeds, background_images, packed, params):
        n = _native()
        dev = _cuda_device()
        if params.dim() != 2 or params.shape[1] != packed.num_params:
            raise ValueError('params must be [batch, %d]' % packed.num_params)
        if packed.output_type != OutputType.color or packed.use_prefiltering:
            raise NotImplementedError('batched scenes render colour images with the sampled path only')
        batch = params.shape[0]
        seeds = _seed_array(seeds, batch)
        ns = _get_batch_scene(packed, dev.index, batch)
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream().cuda_stream
            version = ns.set_params(params.reshape(-1), stream)
            images = torch.empty(batch, height, width, 4, device=dev, dtype=torch.float32)
            if background_images is not None:
                background_images = background_images.to(dev).float().contiguous()
                assert tuple(background_images.shape) == (batch, height, width, 4)
            n.check(n.lib.dvg_render_forward_batch(
                ns.handle, background_images.data_ptr() if background_images is not None else None, images.data_ptr(),
                width, height, num_samples_x, num_samples_y, seeds.ctypes.data, stream))
        ctx.native_scene, ctx.scene_version, ctx.packed, ctx.background_images = ns, version, packed, background_images
        ctx.dims = (width, height, num_samples_x, num_samples_y)
        ctx.seeds, ctx.device, ctx.params_device = seeds, dev, params.device
        ctx.save_for_backward(params)
        return images

    @staticmethod
    def backward(ctx, grad_images):
        n = _native()
        dev, ns = ctx.device, ctx.native_scene
        (params,) = ctx.saved_tensors
        width, height, nsx, nsy = ctx.dims
        grad_images = grad_images.to(dev).float().contiguous()
        bg = ctx.background_images
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream().cuda_stream
            if ns.version != ctx.scene_version:
                ctx.scene_version = ns.set_params(params.reshape(-1), stream)
            d_params = torch.empty_like(params, device=dev, dtype=torch.float32)
            d_bg = torch.empty_like(bg) if bg is not None else None
            n.check(n.lib.dvg_render_backward_batch(
                ns.handle, bg.data_ptr() if bg is not None else None, grad_images.data_ptr(), width, height, nsx, nsy,
                ctx.seeds.ctypes.data, d_params.data_ptr(), d_bg.data_ptr() if d_bg is not None else None,
                _rp.backward_flags(ctx.packed), stream))
        if d_params.device != ctx.params_device:
            d_params = d_params.to(ctx.params_device)
        return None, None, None, None, None, d_bg, None, d_params


render_batch = BatchRenderFunction.apply


def serialize_scenes(canvas_width, canvas_height, scenes, filter_type=0, filter_radius=None):
    """`scenes`: list of (shapes, shape_groups) with one common structure.  Returns [PackedScene, params[batch, N]]
    (differentiable with respect to the holders' tensors), to splat into `render_batch`."""
    topo0, rows, xform_grad, filter_grad = None, [], False, False
    for shapes, groups in scenes:
        topo, tensors = scene_pack.pack_scene(canvas_width, canvas_height, shapes, groups, filter_type, filter_radius)
        if topo0 is None:
            topo0 = topo
        elif topo.shape != topo0.shape or not np.array_equal(topo, topo0):
            raise ValueError('the scenes of a batch must share one topology (shape types, segment counts, groups, colour kinds)')
        xform_grad = xform_grad or any(t.requires_grad for t in tensors[scene_pack.B_MAT3])
        filter_grad = filter_grad or any(t.requires_grad for t in tensors[scene_pack.B_FILTER])
        rows.append(scene_pack.concat_params(tensors))
    packed = PackedScene(topo0, canvas_width, canvas_height, OutputType.color, False, torch.tensor([]))
    packed.needs_xform_grad, packed.needs_filter_grad = xform_grad, filter_grad
    return [packed, torch.stack(rows)]


_STROKE_TOPO = {}


def stroke_scene_args(all_points, all_widths, all_colors, canvas_size, num_control_points):
    """Scene arguments of `batch` scenes of `num_strokes` open stroked paths each (no fill, identity transform, box
    filter 0.5): `all_points` [batch, num_strokes, num_points, 2] in canvas units, `all_widths` [batch, num_strokes],
    `all_colors` [batch, num_strokes, 4] (RGBA), `num_control_points`: the per-segment control-point counts shared by
    every path (all 0: polyline strokes; all 2: cubic Beziers).  Everything stays a tensor op on the whole batch; the
    topology is packed once per (num_strokes, num_points, segment types, canvas size) from a template scene."""
    bs, num_strokes, num_pts, _ = all_points.shape
    ncp = tuple(int(c) for c in num_control_points)
    key = (num_strokes, num_pts, ncp, int(canvas_size))
    packed = _STROKE_TOPO.get(key)
    if packed is None:
        ncp_t = torch.tensor(ncp, dtype=torch.int32)
        shapes, groups = [], []
        for p in range(num_strokes):
            shapes.append(_shape.Path(num_control_points=ncp_t, points=torch.zeros(num_pts, 2), is_closed=False,
                                      stroke_width=torch.tensor(1.0)))
            groups.append(_shape.ShapeGroup(shape_ids=torch.tensor([p]), fill_color=None, stroke_color=torch.ones(4)))
        topo, bk, _ = scene_pack._pack_scene_full(canvas_size, canvas_size, shapes, groups, 0, torch.tensor(0.5))
        # the layout this function fills below: [points | widths | colours | eye(3) | filter radius]
        assert bk.sizes == [2 * num_pts * num_strokes, num_strokes, 4 * num_strokes, 9, 0, 1]
        topo.setflags(write=False)
        packed = PackedScene(topo, canvas_size, canvas_size, OutputType.color, False, torch.tensor([]))
        packed.needs_xform_grad = packed.needs_filter_grad = False
        _STROKE_TOPO[key] = packed
    dev, dt = all_points.device, torch.float32
    tail = torch.cat([torch.eye(3, device=dev, dtype=dt).reshape(-1), torch.full((1,), 0.5, device=dev, dtype=dt)])
    params = torch.cat([all_points.reshape(bs, -1).to(dt), all_widths.reshape(bs, -1).to(dt),
                        all_colors.reshape(bs, -1).to(dt), tail.expand(bs, -1)], dim=1)
    return [packed, params]


class _LazyScenes:
    """The `scenes` list the reference's front ends return, (canvas, canvas, shapes, shape_groups) per sample
    (rendering.py:213, 288): holders are only built for the samples somebody looks at (the apps save a few as SVG)."""

    def __init__(self, all_points, all_widths, all_colors, canvas_size, ncp):
        self.a = (all_points.detach(), all_widths.detach(), all_colors.detach(), canvas_size, ncp)

    def __len__(self):
        return self.a[0].shape[0]

    def __getitem__(self, k):
        pts, widths, colors, canvas_size, ncp = self.a
        if isinstance(k, slice):
            return [self[i] for i in range(*k.indices(len(self)))]
        shapes, groups = [], []
        for p in range(pts.shape[1]):
            shapes.append(_shape.Path(num_control_points=torch.tensor(ncp, dtype=torch.int32), points=pts[k, p].contiguous().cpu(),
                                      stroke_width=widths[k, p].cpu(), is_closed=False))
            groups.append(_shape.ShapeGroup(shape_ids=torch.tensor([p]), fill_color=None, stroke_color=colors[k, p].cpu()))
        return (canvas_size, canvas_size, shapes, groups)

    def __iter__(self):
        return (self[i] for i in range(len(self)))


def _stroke_render(all_points, all_widths, all_alphas, canvas_size, colors, ncp, samples, seeds):
    dev = all_points.device
    all_points = 0.5 * (all_points + 1.0) * canvas_size                  # rendering.py:182 / 251
    eps = 1e-4
    all_points = all_points + eps * torch.randn_like(all_points)         # rendering.py:184-185 / 253-254
    bs, num_strokes = all_points.shape[:2]
    rgb = colors if colors is not None else torch.ones(bs, num_strokes, 3, device=all_alphas.device, dtype=all_alphas.dtype)
    rgba = torch.cat([rgb.to(all_alphas.device), all_alphas.reshape(bs, num_strokes, 1)], dim=2)
    packed, params = stroke_scene_args(all_points, all_widths, rgba.to(all_points.device), canvas_size, ncp)
    raster = render_batch(canvas_size, canvas_size, samples, samples, _seed_array(seeds, bs), None, packed, params)
    raster = raster.permute(0, 3, 1, 2)                                    # [bs, 4, H, W]
    alpha = raster[:, 3:4]
    image = raster[:, :3] if colors is not None else raster[:, :1]
    output = (image * alpha).to(dev)                                       # alpha compositing, rendering.py:228-229
    return output, _LazyScenes(all_points, all_widths, rgba, canvas_size, ncp)


def line_render(all_points, all_widths, all_alphas, force_cpu=True, canvas_size=32, colors=None, seeds=None):
    """Reference: apps/generative_models/rendering.py:170-237.  `all_points` [bs, num_segments, 2, 2] in [-1, 1].
    `force_cpu` is accepted and ignored (it moved the reference's per-sample loop to the CPU renderer); `seeds`: one
    per sample (default: random per sample, as the reference's `render`)."""
    return _stroke_render(all_points, all_widths, all_alphas, canvas_size, colors, (0,), 2, seeds)


def bezier_render(all_points, all_widths, all_alphas, force_cpu=True, canvas_size=32, colors=None, seeds=None):
    """Reference: apps/generative_models/rendering.py:239-307.  `all_points` [bs, num_strokes, 3 k + 1, 2] in [-1, 1]:
    every stroke is a chain of k cubic segments."""
    num_segments = (all_points.shape[2] - 1) // 3
    return _stroke_render(all_points, all_widths, all_alphas, canvas_size, colors, (2,) * num_segments, 2, seeds)
