"""Pack a pydiffvg scene (lists of shape / group holders) into the two flat arrays the C ABI
takes (include/dvg_scene_format.h): an int32 topology blob and the list of float tensors
whose concatenation is `params`.

This replaces the per-object argument list built by the reference's
`RenderFunction.serialize_scene` (pydiffvg/render_pytorch.py:22-172) and the pybind11 object
graph rebuilt from it in `forward` (render_pytorch.py:206-363).  Semantics kept from there:
  * `stroke_width` with shape[0] > 1 is per-point thickness and the scalar stroke width
    becomes 0.0 (render_pytorch.py:67-72, 95-98);
  * a `Polygon` is a path of zero-control-point segments, one per point if closed and one
    fewer if open (render_pytorch.py:75-86);
  * fill / stroke colours may be None, a tensor[4], or a Linear/RadialGradient holder.
"""
import warnings

import numpy as np
import torch

from .pydiffvg import shape as _shape
from .pydiffvg import color as _color

# header / record indices: keep in sync with include/dvg_scene_format.h
TOPO_MAGIC = 0x44564731
H_LEN = 16
S_LEN = 8
G_LEN = 12
(H_MAGIC, H_CW, H_CH, H_NS, H_NG, H_FTYPE, H_FRAD_OFF, H_NPARAMS, H_TOTAL_SEGS, H_TOTAL_GSHAPES,
 H_OFF_SHAPES, H_OFF_GROUPS, H_OFF_NCP, H_OFF_GSHAPES, H_TOTAL_POINTS, H_RESERVED) = range(16)

SHAPE_CIRCLE, SHAPE_ELLIPSE, SHAPE_PATH, SHAPE_RECT = 0, 1, 2, 3
COLOR_NONE, COLOR_CONSTANT, COLOR_LINEAR, COLOR_RADIAL = -1, 0, 1, 2
SF_CLOSED, SF_DISTANCE_APPROX = 1, 2


class _ParamList:
    """Accumulates float tensors and hands out their offsets in the concatenation."""

    def __init__(self):
        self.tensors = []
        self.n = 0

    def add(self, t, expect=None):
        if not isinstance(t, torch.Tensor):
            t = torch.as_tensor(t, dtype=torch.float32)
        k = t.numel()
        if expect is not None and k != expect:
            raise ValueError('expected a tensor of %d elements, got shape %s' % (expect, tuple(t.shape)))
        off = self.n
        self.tensors.append(t)
        self.n += k
        return off


def _pack_color(color, params):
    """-> (type, params offset, num_stops)"""
    if color is None:
        return COLOR_NONE, 0, 0
    if isinstance(color, torch.Tensor):
        return COLOR_CONSTANT, params.add(color, 4), 0
    # duck-type so that the reference's own holder classes are accepted too
    if hasattr(color, 'begin') and hasattr(color, 'end'):
        n = color.offsets.shape[0]
        if color.stop_colors.shape[0] != n:
            raise ValueError('gradient offsets / stop_colors length mismatch')
        off = params.add(color.begin, 2)
        params.add(color.end, 2)
        params.add(color.offsets, n)
        params.add(color.stop_colors, 4 * n)
        return COLOR_LINEAR, off, n
    if hasattr(color, 'center') and hasattr(color, 'radius'):
        n = color.offsets.shape[0]
        if color.stop_colors.shape[0] != n:
            raise ValueError('gradient offsets / stop_colors length mismatch')
        off = params.add(color.center, 2)
        params.add(color.radius, 2)
        params.add(color.offsets, n)
        params.add(color.stop_colors, 4 * n)
        return COLOR_RADIAL, off, n
    raise TypeError('unsupported colour %r' % (color,))


def _kind(shape):
    # duck-typed so holders from the reference package work as well
    n = type(shape).__name__
    if n in ('Circle', 'Ellipse', 'Path', 'Polygon', 'Rect'):
        return n
    raise TypeError('unsupported shape %r' % (shape,))


def pack_scene(canvas_width, canvas_height, shapes, shape_groups, filter_type=0, filter_radius=None):
    """Returns (topo: np.int32[], tensors: list[Tensor]).  `torch.cat([t.reshape(-1) ...])`
    of `tensors` is `params`."""
    params = _ParamList()
    ns, ng = len(shapes), len(shape_groups)
    srec = np.zeros((ns, S_LEN), dtype=np.int32)
    ncp_chunks = []
    ncp_total = 0
    points_total = 0
    is_open_path = [False] * ns
    for i, shape in enumerate(shapes):
        kind = _kind(shape)
        r = srec[i]
        use_thickness = False
        r[3] = -1
        if kind == 'Circle':
            r[0] = SHAPE_CIRCLE
            r[1] = params.add(shape.radius, 1)
            params.add(shape.center, 2)
        elif kind == 'Ellipse':
            r[0] = SHAPE_ELLIPSE
            r[1] = params.add(shape.radius, 2)
            params.add(shape.center, 2)
        elif kind == 'Rect':
            r[0] = SHAPE_RECT
            r[1] = params.add(shape.p_min, 2)
            params.add(shape.p_max, 2)
        else:
            pts = shape.points
            if pts.dim() != 2 or pts.shape[1] != 2:
                raise ValueError('path points must be [N, 2]')
            npts = pts.shape[0]
            if kind == 'Path':
                ncp = shape.num_control_points
                ncp_np = ncp.detach().cpu().numpy().astype(np.int32, copy=False) if isinstance(ncp, torch.Tensor) \
                    else np.asarray(ncp, dtype=np.int32)
                sw = shape.stroke_width
                if isinstance(sw, torch.Tensor) and sw.dim() > 0 and sw.shape[0] > 1:
                    use_thickness = True
                flags = (SF_CLOSED if shape.is_closed else 0) | (SF_DISTANCE_APPROX if shape.use_distance_approx else 0)
            else:  # Polygon
                ncp_np = np.zeros(npts if shape.is_closed else npts - 1, dtype=np.int32)
                flags = SF_CLOSED if shape.is_closed else 0
            r[0] = SHAPE_PATH
            r[1] = params.add(pts, 2 * npts)
            if use_thickness:
                r[3] = params.add(shape.stroke_width, npts)
            r[4] = npts
            r[5] = ncp_np.shape[0]
            r[6] = ncp_total
            r[7] = flags
            ncp_chunks.append(ncp_np)
            ncp_total += ncp_np.shape[0]
            points_total += npts
            is_open_path[i] = not shape.is_closed
        if use_thickness:
            r[2] = -1
        else:
            r[2] = params.add(shape.stroke_width, 1)

    grec = np.zeros((ng, G_LEN), dtype=np.int32)
    gshape_chunks = []
    gshape_total = 0
    for g, group in enumerate(shape_groups):
        r = grec[g]
        ids = group.shape_ids
        ids_np = ids.detach().cpu().numpy().astype(np.int32, copy=False) if isinstance(ids, torch.Tensor) \
            else np.asarray(ids, dtype=np.int32)
        ids_np = ids_np.reshape(-1)
        if ids_np.size == 0:
            raise ValueError('shape group %d has no shapes' % g)
        if ids_np.min() < 0 or ids_np.max() >= ns:
            raise ValueError('shape group %d references a shape id out of range' % g)
        r[0] = gshape_total
        r[1] = ids_np.shape[0]
        gshape_chunks.append(ids_np)
        gshape_total += ids_np.shape[0]
        r[2], r[3], r[4] = _pack_color(group.fill_color, params)
        if group.fill_color is not None:
            # render_pytorch.py:131-136
            for sid in ids_np:
                if is_open_path[sid]:
                    warnings.warn('Detected non-closed paths with fill color. This might causes unexpected results.',
                                  Warning)
        r[5], r[6], r[7] = _pack_color(group.stroke_color, params)
        r[8] = 1 if group.use_even_odd_rule else 0
        r[9] = params.add(group.shape_to_canvas, 9)

    frad_off = params.add(filter_radius if filter_radius is not None else torch.tensor(0.5), 1)

    off_shapes = H_LEN
    off_groups = off_shapes + ns * S_LEN
    off_ncp = off_groups + ng * G_LEN
    off_gshapes = off_ncp + ncp_total
    topo = np.zeros(off_gshapes + gshape_total, dtype=np.int32)
    topo[H_MAGIC] = TOPO_MAGIC
    topo[H_CW] = int(canvas_width)
    topo[H_CH] = int(canvas_height)
    topo[H_NS] = ns
    topo[H_NG] = ng
    topo[H_FTYPE] = int(filter_type)
    topo[H_FRAD_OFF] = frad_off
    topo[H_NPARAMS] = params.n
    topo[H_TOTAL_SEGS] = ncp_total
    topo[H_TOTAL_GSHAPES] = gshape_total
    topo[H_OFF_SHAPES] = off_shapes
    topo[H_OFF_GROUPS] = off_groups
    topo[H_OFF_NCP] = off_ncp
    topo[H_OFF_GSHAPES] = off_gshapes
    topo[H_TOTAL_POINTS] = points_total
    topo[off_shapes:off_groups] = srec.reshape(-1)
    topo[off_groups:off_ncp] = grec.reshape(-1)
    if ncp_total:
        topo[off_ncp:off_gshapes] = np.concatenate(ncp_chunks)
    if gshape_total:
        topo[off_gshapes:] = np.concatenate(gshape_chunks)
    return topo, params.tensors


def concat_params(tensors, device=None):
    """Differentiable concatenation of the parameter tensors into the flat `params`.

    Gradients flow back to the user's tensors through autograd's CatBackward in C++ instead
    of the reference's O(#shapes) Python read-back loop (render_pytorch.py:713-866)."""
    flat = []
    devs = set()
    for t in tensors:
        if t.dtype != torch.float32:
            t = t.to(torch.float32)
        flat.append(t.reshape(-1))
        devs.add(t.device)
    if len(devs) > 1:
        target = device if device is not None else torch.device('cpu')
        flat = [t.to(target) for t in flat]
    return torch.cat(flat) if len(flat) > 1 else flat[0].clone()


def pack_scene_numpy(canvas_width, canvas_height, shapes, shape_groups, filter_type=0, filter_radius=None):
    """(topo, params) as numpy arrays -- what the oracle entry points take."""
    topo, tensors = pack_scene(canvas_width, canvas_height, shapes, shape_groups, filter_type, filter_radius)
    params = concat_params([t.detach().cpu() for t in tensors]).numpy().astype(np.float32, copy=False)
    return topo, np.ascontiguousarray(params)
