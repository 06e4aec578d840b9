"""Plain holders for the vector primitives; constructor signatures follow the reference
(pydiffvg/shape.py:5-61), and `from_svg_path` (shape.py:63-172) on top of svg_path.py instead of the
third-party svgpathtools package (absent here)."""
import torch


class Circle:
    def __init__(self, radius, center, stroke_width=torch.tensor(1.0), id=''):
        self.radius = radius
        self.center = center
        self.stroke_width = stroke_width
        self.id = id


class Ellipse:
    def __init__(self, radius, center, stroke_width=torch.tensor(1.0), id=''):
        self.radius = radius
        self.center = center
        self.stroke_width = stroke_width
        self.id = id


class Path:
    def __init__(self, num_control_points, points, is_closed,
                 stroke_width=torch.tensor(1.0), id='', use_distance_approx=False):
        self.num_control_points = num_control_points
        self.points = points
        self.is_closed = is_closed
        self.stroke_width = stroke_width
        self.id = id
        self.use_distance_approx = use_distance_approx


class Polygon:
    def __init__(self, points, is_closed, stroke_width=torch.tensor(1.0), id=''):
        self.points = points
        self.is_closed = is_closed
        self.stroke_width = stroke_width
        self.id = id


class Rect:
    def __init__(self, p_min, p_max, stroke_width=torch.tensor(1.0), id=''):
        self.p_min = p_min
        self.p_max = p_max
        self.stroke_width = stroke_width
        self.id = id


class ShapeGroup:
    def __init__(self, shape_ids, fill_color, use_even_odd_rule=True, stroke_color=None,
                 shape_to_canvas=torch.eye(3), id=''):
        self.shape_ids = shape_ids
        self.fill_color = fill_color
        self.use_even_odd_rule = use_even_odd_rule
        self.stroke_color = stroke_color
        self.shape_to_canvas = shape_to_canvas
        self.id = id


def from_svg_path(path_str, shape_to_canvas=torch.eye(3), force_close=False):
    """SVG path data -> list of `Path` holders, one per continuous sub-path, points pre-multiplied by
    `shape_to_canvas` (reference pydiffvg/shape.py:63-172; the svgpathtools part is restated in svg_path.py).

    Kept from the reference: a sub-path counts as closed when its end equals its start (or lies within 1e-5 of
    it, or `force_close` adds the closing line); a closing line shorter than 1e-5 is dropped and the previous
    segment snapped onto the start; a closed sub-path does not repeat its first point; arcs become cubics of
    at most a quarter turn each (shape.py:107-157, including its use of `phi` -- already in radians -- as
    degrees)."""
    import math
    from . import svg_path as sp
    segs = sp.parse_path(path_str)
    if len(segs) == 0:
        return []
    ret_paths = []
    for sub in sp.continuous_subpaths(segs):
        closed = sp.seg_start(sub[0]) == sp.seg_end(sub[-1])
        if closed:
            tail = sub[-1]
            if len(sub) > 1 and tail[0] == 'L' and abs(tail[2] - tail[1]) < 1e-5:
                sub.pop()
                sub[-1] = sp.with_end(sub[-1], sp.seg_start(sub[0]))
        else:
            beg, end = sp.seg_start(sub[0]), sp.seg_end(sub[-1])
            if abs(end - beg) < 1e-5:
                sub[-1] = sp.with_end(sub[-1], beg)
                closed = True
            elif force_close:
                sub.append(('L', end, beg))
                closed = True
        num_control_points = []
        points = []
        for i, e in enumerate(sub):
            s0 = sp.seg_start(e)
            if i == 0:
                points.append((s0.real, s0.imag))
            else:
                assert s0.real == points[-1][0] and s0.imag == points[-1][1]
            kind = e[0]
            if kind == 'L':
                num_control_points.append(0)
            elif kind == 'Q':
                num_control_points.append(1)
                points.append((e[2].real, e[2].imag))
            elif kind == 'C':
                num_control_points.append(2)
                points.append((e[2].real, e[2].imag))
                points.append((e[3].real, e[3].imag))
            else:
                a = e[3]
                start = a['theta'] * math.pi / 180.0
                stop = (a['theta'] + a['delta']) * math.pi / 180.0
                sign = -1.0 if stop < start else 1.0
                epsilon = 0.00001
                rx, ry = a['radius'].real, a['radius'].imag
                cx, cy = a['center'].real, a['center'].imag
                rot = a['phi'] * math.pi / 180.0
                cos_rot, sin_rot = math.cos(rot), math.sin(rot)
                while sign * (stop - start) > epsilon:
                    step = stop - start
                    step = min(step, 0.5 * math.pi) if step > 0.0 else max(step, -0.5 * math.pi)
                    alpha = step / 2.0
                    cos_alpha, sin_alpha = math.cos(alpha), math.sin(alpha)
                    cot_alpha = 1.0 / math.tan(alpha)
                    phi = start + alpha
                    cos_phi, sin_phi = math.cos(phi), math.sin(phi)
                    lambda_ = (4.0 - cos_alpha) / 3.0
                    mu = sin_alpha + (cos_alpha - lambda_) * cot_alpha
                    last = sign * (stop - (start + step)) <= epsilon
                    num_control_points.append(2)
                    for sgn in (1.0, -1.0):
                        x = lambda_ * cos_phi + sgn * mu * sin_phi
                        y = lambda_ * sin_phi - sgn * mu * cos_phi
                        points.append((cx + rx * (x * cos_rot - y * sin_rot), cy + ry * (x * sin_rot + y * cos_rot)))
                    if not last:
                        points.append((cx + rx * math.cos(rot + start + step), cy + ry * math.sin(rot + start + step)))
                    start += step
            e1 = sp.seg_end(e)
            if i != len(sub) - 1:
                points.append((e1.real, e1.imag))
            elif closed:
                assert e1.real == points[0][0] and e1.imag == points[0][1]
            else:
                points.append((e1.real, e1.imag))
        points = torch.tensor(points, dtype=torch.float)
        points = torch.cat((points, torch.ones([points.shape[0], 1])), dim=1) @ torch.transpose(shape_to_canvas, 0, 1)
        points = points / points[:, 2:3]
        points = points[:, :2].contiguous()
        ret_paths.append(Path(torch.tensor(num_control_points), points, closed))
    return ret_paths
