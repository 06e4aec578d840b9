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
_OPEN_FILL_WARNING = 'Detected non-closed paths with fill color. This might causes unexpected results.'


# Parameter tensors are collected into a few "buckets" so that the flat `params` can be built
# with a handful of torch ops (one cat for all path points, one stack for all stroke widths,
# one for all constant colours, ...) instead of one reshape per tensor: at 2048 paths the
# per-tensor version costs ~130 ms of pure Python per iteration, the bucketed one a few ms.
# A tensor OBJECT that occurs several times (typically the default `shape_to_canvas`
# torch.eye(3) shared by every ShapeGroup) is stored once; autograd then sums the gradients
# of all its uses, which is exactly what sharing a tensor means.
B_POINTS, B_SCALAR, B_VEC4, B_MAT3, B_GENERIC, B_FILTER, NUM_BUCKETS = 0, 1, 2, 3, 4, 5, 6
SRC_SHAPE, SRC_GROUP, SRC_FILTER = 0, 1, 2


class _Buckets:
    def __init__(self):
        self.tensors = [[] for _ in range(NUM_BUCKETS)]
        self.sizes = [0] * NUM_BUCKETS
        self.seen = {}
        # where every entry came from: (holder list code, holder index, attribute, sub-attribute, numel);
        # replayed by the memoised fast path of pack_scene
        self.sources = [[] for _ in range(NUM_BUCKETS)]
        self.src = (SRC_FILTER, 0)

    def at(self, code, index):
        self.src = (code, index)

    def add(self, bucket, t, numel, attr=None, sub=None):
        off = self.sizes[bucket]
        self.tensors[bucket].append(t)
        self.sources[bucket].append((self.src[0], self.src[1], attr, sub, numel))
        self.sizes[bucket] = off + numel
        return off

    def add_generic(self, t, expect, attr=None, sub=None):
        if not isinstance(t, torch.Tensor):
            t = torch.as_tensor(t, dtype=torch.float32)
        if t.numel() != expect:
            raise ValueError('expected a tensor of %d elements, got shape %s' % (expect, tuple(t.shape)))
        return self.add(B_GENERIC, t, expect, attr, sub)

    def add_shared(self, bucket, t, numel, attr=None):
        key = id(t)
        hit = self.seen.get(key)
        if hit is not None and hit[0] is t:
            return hit[1]
        off = self.add(bucket, t, numel, attr)
        self.seen[key] = (t, off)
        return off


def _np_cached(holder, attr, value):
    """int32 numpy copy of a small index tensor, cached on the holder object and invalidated
    by the tensor's identity / in-place version counter."""
    if not isinstance(value, torch.Tensor):
        return np.asarray(value, dtype=np.int32).reshape(-1)
    cache = holder.__dict__.get('_dvg_' + attr)
    if cache is not None and cache[0] is value and cache[1] == value._version:
        return cache[2]
    arr = value.detach().cpu().numpy().astype(np.int32, copy=True).reshape(-1)
    holder.__dict__['_dvg_' + attr] = (value, value._version, arr)
    return arr


def _pack_color(color, bk, attr):
    """-> (type, bucket, offset in bucket, num_stops)"""
    if color is None:
        return COLOR_NONE, B_GENERIC, 0, 0
    if isinstance(color, torch.Tensor):
        if color.dim() != 1 or color.shape[0] != 4:
            raise ValueError('a constant colour must be a tensor of 4 elements')
        return COLOR_CONSTANT, B_VEC4, bk.add(B_VEC4, color, 4, attr), 0
    # duck-typed so that the reference's own holder classes are accepted too
    if hasattr(color, 'begin') and hasattr(color, 'end'):
        n = color.offsets.shape[0]
        if color.stop_colors.shape[0] != n:
            raise ValueError('gradient offsets / stop_colors length mismatch')
        off = bk.add_generic(color.begin, 2, attr, 'begin')
        bk.add_generic(color.end, 2, attr, 'end')
        bk.add_generic(color.offsets, n, attr, 'offsets')
        bk.add_generic(color.stop_colors, 4 * n, attr, 'stop_colors')
        return COLOR_LINEAR, B_GENERIC, off, n
    if hasattr(color, 'center') and hasattr(color, 'radius'):
        n = color.offsets.shape[0]
        if color.stop_colors.shape[0] != n:
            raise ValueError('gradient offsets / stop_colors length mismatch')
        off = bk.add_generic(color.center, 2, attr, 'center')
        bk.add_generic(color.radius, 2, attr, 'radius')
        bk.add_generic(color.offsets, n, attr, 'offsets')
        bk.add_generic(color.stop_colors, 4 * n, attr, 'stop_colors')
        return COLOR_RADIAL, B_GENERIC, off, n
    raise TypeError('unsupported colour %r' % (color,))


def _add_width(bk, sw, attr='stroke_width'):
    if isinstance(sw, torch.Tensor) and sw.dim() == 0:
        return B_SCALAR, bk.add(B_SCALAR, sw, 1, attr)
    return B_GENERIC, bk.add_generic(sw, 1, attr)


def _pack_scene_full(canvas_width, canvas_height, shapes, shape_groups, filter_type, filter_radius):
    """The complete walk: validates everything, returns (topo, _Buckets, warned)."""
    bk = _Buckets()
    warned = False
    ns, ng = len(shapes), len(shape_groups)
    srec = []      # rows of DVG_SHAPE_REC_LEN ints (offsets still bucket-relative)
    sbucket = []   # (bucket of PARAM_OFF, bucket of WIDTH_OFF, bucket of THICK_OFF)
    ncp_chunks = []
    ncp_total = 0
    points_total = 0
    is_open_path = [False] * ns
    for i, shape in enumerate(shapes):
        bk.at(SRC_SHAPE, i)
        kind = type(shape).__name__
        if kind == 'Path' or kind == 'Polygon':
            pts = shape.points
            if pts.dim() != 2 or pts.shape[1] != 2:
                raise ValueError('path points must be [N, 2]')
            npts = pts.shape[0]
            thick_off = -1
            use_thickness = False
            if kind == 'Path':
                ncp_np = _np_cached(shape, 'ncp', shape.num_control_points)
                sw = shape.stroke_width
                if isinstance(sw, torch.Tensor) and sw.dim() > 0 and sw.shape[0] > 1:
                    use_thickness = True
                flags = (SF_CLOSED if shape.is_closed else 0) | (SF_DISTANCE_APPROX if shape.use_distance_approx else 0)
            else:
                ncp_np = np.zeros(npts if shape.is_closed else npts - 1, dtype=np.int32)
                flags = SF_CLOSED if shape.is_closed else 0
            poff = bk.add(B_POINTS, pts, 2 * npts, 'points')
            if use_thickness:
                thick_off = bk.add_generic(shape.stroke_width, npts, 'stroke_width')
                wb, woff = B_GENERIC, -1
            else:
                wb, woff = _add_width(bk, shape.stroke_width)
            nseg = ncp_np.shape[0]
            srec.append((SHAPE_PATH, poff, woff, thick_off, npts, nseg, ncp_total, flags))
            sbucket.append((B_POINTS, wb, B_GENERIC))
            ncp_chunks.append(ncp_np)
            ncp_total += nseg
            points_total += npts
            is_open_path[i] = not shape.is_closed
        else:
            if kind == 'Circle':
                t = SHAPE_CIRCLE
                poff = bk.add_generic(shape.radius, 1, 'radius')
                bk.add_generic(shape.center, 2, 'center')
            elif kind == 'Ellipse':
                t = SHAPE_ELLIPSE
                poff = bk.add_generic(shape.radius, 2, 'radius')
                bk.add_generic(shape.center, 2, 'center')
            elif kind == 'Rect':
                t = SHAPE_RECT
                poff = bk.add_generic(shape.p_min, 2, 'p_min')
                bk.add_generic(shape.p_max, 2, 'p_max')
            else:
                raise TypeError('unsupported shape %r' % (shape,))
            wb, woff = _add_width(bk, shape.stroke_width)
            srec.append((t, poff, woff, -1, 0, 0, 0, 0))
            sbucket.append((B_GENERIC, wb, B_GENERIC))

    grec = []
    gbucket = []
    gshape_chunks = []
    gshape_total = 0
    any_open = any(is_open_path)
    for g, group in enumerate(shape_groups):
        bk.at(SRC_GROUP, g)
        ids_np = _np_cached(group, 'ids', group.shape_ids)
        k = ids_np.shape[0]
        if k == 0:
            raise ValueError('shape group %d has no shapes' % g)
        ft, fb, foff, fstops = _pack_color(group.fill_color, bk, 'fill_color')
        if ft != COLOR_NONE and any_open:
            # render_pytorch.py:131-136
            for sid in ids_np:
                if 0 <= sid < ns and is_open_path[sid]:
                    warned = True
                    warnings.warn(_OPEN_FILL_WARNING, Warning)
        st, sb, soff, sstops = _pack_color(group.stroke_color, bk, 'stroke_color')
        xf = group.shape_to_canvas
        if xf.dim() != 2 or xf.shape[0] != 3 or xf.shape[1] != 3:
            raise ValueError('shape_to_canvas must be [3, 3]')
        xoff = bk.add_shared(B_MAT3, xf, 9, 'shape_to_canvas')
        grec.append((gshape_total, k, ft, foff, fstops, st, soff, sstops, 1 if group.use_even_odd_rule else 0, xoff, 0, 0))
        gbucket.append((fb, sb))
        gshape_chunks.append(ids_np)
        gshape_total += k

    fr = filter_radius if filter_radius is not None else torch.tensor(0.5)
    bk.at(SRC_FILTER, 0)
    # a bucket of its own (the last float of `params`): an optimiser stepping the stroke widths or the other leaves of a
    # PackedParams must not move the pixel filter along
    frb, froff = B_FILTER, bk.add(B_FILTER, fr if isinstance(fr, torch.Tensor) else torch.as_tensor(fr, dtype=torch.float32), 1, None)

    # bucket bases in the final concatenation order
    base = [0] * NUM_BUCKETS
    acc = 0
    for b in range(NUM_BUCKETS):
        base[b] = acc
        acc += bk.sizes[b]
    num_params = acc
    base_np = np.asarray(base, dtype=np.int64)

    srec = np.asarray(srec, dtype=np.int64).reshape(ns, S_LEN)
    sb_np = np.asarray(sbucket, dtype=np.int64).reshape(ns, 3)
    srec[:, 1] += base_np[sb_np[:, 0]]
    srec[:, 2] = np.where(srec[:, 2] >= 0, srec[:, 2] + base_np[sb_np[:, 1]], -1)
    srec[:, 3] = np.where(srec[:, 3] >= 0, srec[:, 3] + base_np[sb_np[:, 2]], -1)
    grec = np.asarray(grec, dtype=np.int64).reshape(ng, G_LEN)
    gb_np = np.asarray(gbucket, dtype=np.int64).reshape(ng, 2)
    grec[:, 3] += base_np[gb_np[:, 0]]
    grec[:, 6] += base_np[gb_np[:, 1]]
    grec[:, 9] += base[B_MAT3]

    gshapes = np.concatenate(gshape_chunks) if gshape_total else np.zeros(0, np.int32)
    if gshape_total and (gshapes.min() < 0 or gshapes.max() >= ns):
        raise ValueError('a shape group references a shape id out of range')

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
    topo[H_FRAD_OFF] = froff + base[frb]
    topo[H_NPARAMS] = num_params
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
        topo[off_gshapes:] = gshapes
    return topo, bk, warned


# ---------------------------------------------------------------------------------------------
# Memoised front: an optimisation loop calls serialize_scene every iteration with the same
# holders and only the VALUES of their tensors changed.  The structural signature below (types,
# point counts, flags, identity + version of the integer tensors, identity of the transform
# tensors that decide sharing) is cheap to compute; when it matches a previous call, the topology
# blob is reused and the parameter tensors are re-collected by replaying the recorded sources
# (each checked against its recorded element count; any surprise falls back to the full walk).
_MEMO = []
_MEMO_MAX = 8
_T = torch.Tensor


def _sig_tensor(ap, t):
    """Appends the structural facts of one parameter tensor as plain scalars (tuples and torch.Size
    objects are avoided on purpose: thousands of live containers per call make the cyclic GC run)."""
    if isinstance(t, _T):
        n = t.dim()
        ap(n)
        ap(t.shape[0] if n else -1)
    else:
        ap(None)
        ap(type(t))


def _sig_color(ap, c):
    if c is None:
        ap(0)
    elif isinstance(c, _T):
        ap(1)
        _sig_tensor(ap, c)
    else:
        ap(type(c))
        _sig_tensor(ap, getattr(c, 'offsets', None))
        _sig_tensor(ap, getattr(c, 'stop_colors', None))


def _signature(canvas_width, canvas_height, shapes, shape_groups, filter_type, filter_radius):
    sig = [int(canvas_width), int(canvas_height), int(filter_type), len(shapes), len(shape_groups)]
    ap = sig.append
    _sig_tensor(ap, filter_radius)
    keep = []
    kp = keep.append
    for s in shapes:   # hot loop: inlined on purpose
        d = s.__dict__
        pts = d.get('points')
        ap(type(s))
        if pts is not None:
            ncp = d.get('num_control_points')
            kp(ncp)
            ap(pts.shape[0])
            ap(d.get('is_closed'))
            ap(d.get('use_distance_approx'))
            ap(id(ncp))
            ap(ncp._version if ncp is not None else None)
        sw = d.get('stroke_width')
        if isinstance(sw, _T):
            n = sw.dim()
            ap(n)
            if n:
                ap(sw.shape[0])
        else:
            ap(type(sw))
    for g in shape_groups:
        d = g.__dict__
        ids, xf = d['shape_ids'], d['shape_to_canvas']
        kp(ids)
        kp(xf)
        ap(id(ids))
        ap(ids._version)
        ap(id(xf))
        ap(d['use_even_odd_rule'])
        c = d['fill_color']
        if c is None:
            ap(0)
        elif isinstance(c, _T):
            ap(c.dim())
            ap(c.shape[0])
        else:
            _sig_color(ap, c)
        c = d['stroke_color']
        if c is None:
            ap(0)
        elif isinstance(c, _T):
            ap(c.dim())
            ap(c.shape[0])
        else:
            _sig_color(ap, c)
    return sig, keep


def _replay(sources, shapes, shape_groups, filter_radius):
    holders = (shapes, shape_groups)
    out = []
    for bucket_sources in sources:
        ts = []
        for code, index, attr, sub, numel in bucket_sources:
            if code == SRC_FILTER:
                t = filter_radius
            else:
                t = getattr(holders[code][index], attr)
                if sub is not None:
                    t = getattr(t, sub)
            if not isinstance(t, _T):
                t = torch.as_tensor(t, dtype=torch.float32)
            if t.numel() != numel:
                return None
            ts.append(t)
        out.append(ts)
    return out


def pack_scene(canvas_width, canvas_height, shapes, shape_groups, filter_type=0, filter_radius=None):
    """Returns (topo: np.int32[], buckets: list[list[Tensor]]).  `concat_params(buckets)` is `params`."""
    if filter_radius is None:
        filter_radius = torch.tensor(0.5)
    sig = keep = None
    try:
        sig, keep = _signature(canvas_width, canvas_height, shapes, shape_groups, filter_type, filter_radius)
    except Exception:
        sig = None   # malformed holders: the full walk raises the proper error
    if sig is not None:
        for k, m in enumerate(_MEMO):
            if m[0] == sig:
                tensors = _replay(m[3], shapes, shape_groups, filter_radius)
                if tensors is None:
                    break
                if m[4]:
                    warnings.warn(_OPEN_FILL_WARNING, Warning)
                if k:
                    _MEMO.insert(0, _MEMO.pop(k))
                return m[2], tensors
    topo, bk, warned = _pack_scene_full(canvas_width, canvas_height, shapes, shape_groups, filter_type, filter_radius)
    if sig is not None:
        topo.setflags(write=False)
        _MEMO.insert(0, (sig, keep, topo, bk.sources, warned))
        del _MEMO[_MEMO_MAX:]
    return topo, bk.tensors


def _flatten_bucket(bucket, tensors, device):
    """One flat float32 tensor for a bucket, using a single cat/stack when the tensors agree
    on device and dtype (the common case) and a per-tensor path otherwise."""
    if not tensors:
        return None
    try:
        if bucket == B_POINTS:
            return torch.cat(tensors, dim=0).reshape(-1).to(device=device, dtype=torch.float32)
        if bucket != B_GENERIC:
            return torch.stack([t.reshape(()) if bucket in (B_SCALAR, B_FILTER) else t for t in tensors]).reshape(-1).to(device=device, dtype=torch.float32)
    except (RuntimeError, TypeError):
        pass
    return torch.cat([t.to(device=device, dtype=torch.float32).reshape(-1) for t in tensors])


def _bucket_plan(bucket, tensors):
    """How one bucket is flattened: ('stack', shape, dtype, device) -- tensors of one shape; ('cat', trailing shape, dtype,
    device, rows) -- path points [n_i, 2] with different n_i; None -- mixed devices / dtypes / shapes: per tensor."""
    t0 = tensors[0]
    shp, dt, dv = t0.shape, t0.dtype, t0.device
    if not dt.is_floating_point:
        return None
    same = True
    for t in tensors:
        if t.dtype != dt or t.device != dv:
            return None
        if t.shape != shp:
            same = False
    if same:
        return ('stack', shp, dt, dv)
    if bucket == B_POINTS and len(shp) == 2 and all(t.dim() == 2 and t.shape[1] == shp[1] for t in tensors):
        return ('cat', shp[1:], dt, dv, [t.shape[0] for t in tensors])
    return None


class _PackParams(torch.autograd.Function):
    """`params` = the buckets' tensors, flattened and concatenated, as ONE autograd node whose backward hands every tensor a
    view of its slice of `d_params`.  The plain `torch.stack / cat` graph (a view node per scalar, Stack/CatBackward per
    bucket) spent a third more time in the autograd engine at 6 144 leaf tensors; values and gradients are the same."""

    @staticmethod
    def forward(ctx, sizes, device, *tensors):
        parts, plans, i = [], [], 0
        for b, n in enumerate(sizes):
            ts = tensors[i:i + n]
            i += n
            if not n:
                continue
            plan = _bucket_plan(b, ts)
            if plan is None:
                parts.append(torch.cat([t.to(device=device, dtype=torch.float32).reshape(-1) for t in ts]))
                plan = ('each', [(t.shape, t.dtype, t.device) for t in ts])
            elif plan[0] == 'stack':
                parts.append(torch.stack(ts).reshape(-1).to(device=device, dtype=torch.float32))
            else:
                parts.append(torch.cat(ts, dim=0).reshape(-1).to(device=device, dtype=torch.float32))
            plans.append((n, plan))
        ctx.plans = plans
        return torch.cat(parts) if len(parts) > 1 else parts[0].clone()

    @staticmethod
    def backward(ctx, g):
        grads, off = [None, None], 0
        for n, plan in ctx.plans:
            if plan[0] == 'stack':
                k = n * int(np.prod(plan[1], dtype=np.int64))
                gb = g[off:off + k].to(device=plan[3], dtype=plan[2])       # (one copy per bucket when the user's tensors live elsewhere)
                grads.extend(gb.view((n,) + tuple(plan[1])).unbind(0))
            elif plan[0] == 'cat':
                k = sum(plan[4]) * int(np.prod(plan[1], dtype=np.int64))
                gb = g[off:off + k].to(device=plan[3], dtype=plan[2])
                grads.extend(gb.view((-1,) + tuple(plan[1])).split(plan[4]))
            else:
                k = 0
                for shp, dt, dv in plan[1]:
                    m = int(np.prod(shp, dtype=np.int64))
                    grads.append(g[off + k:off + k + m].reshape(shp).to(device=dv, dtype=dt))
                    k += m
            off += k
        return tuple(grads)


def concat_params(buckets, device=None):
    """Differentiable concatenation of the parameter tensors into the flat `params`.

    Gradients flow back to the user's tensors through ONE autograd node (`_PackParams`) that slices `d_params`
    into views, instead of the reference's O(#shapes) Python read-back loop (render_pytorch.py:713-866).
    `device`: where to build `params` (default: the device of the first path-point tensor,
    i.e. wherever the user keeps the scene)."""
    if device is None:
        for b in buckets:
            if b:
                device = b[0].device
                break
    device = torch.device(device) if device is not None else torch.device('cpu')
    return _PackParams.apply([len(b) for b in buckets], device, *[t for b in buckets for t in b])


def pack_scene_numpy(canvas_width, canvas_height, shapes, shape_groups, filter_type=0, filter_radius=None):
    """(topo, params) as numpy arrays -- what the oracle entry points take."""
    topo, buckets = pack_scene(canvas_width, canvas_height, shapes, shape_groups, filter_type, filter_radius)
    with torch.no_grad():
        params = concat_params(buckets, torch.device('cpu')).numpy().astype(np.float32, copy=False)
    return topo, np.ascontiguousarray(params)
