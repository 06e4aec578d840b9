"""Deterministic synthetic scenes shared by the tests, bench.py and the golden-vector generator.
They restate the scene shapes of the reference apps the configs in BASELINE.md name
(apps/single_circle.py, single_stroke.py, painterly_rendering.py:39-106, the blob generator
painterly_rendering.py:45-74, generative_models/rendering.py:239-307)."""
import random

import torch

from diffvg_b200 import pydiffvg


def single_circle(radius=40.0, center=(128.0, 128.0), color=(0.3, 0.6, 0.3, 1.0)):
    c = pydiffvg.Circle(radius=torch.tensor(radius), center=torch.tensor(center))
    g = pydiffvg.ShapeGroup(shape_ids=torch.tensor([0]), fill_color=torch.tensor(color))
    return 256, 256, [c], [g]


def single_stroke(thickness=None, fill=True):
    pts = torch.tensor([[120., 30.], [150., 60.], [90., 198.], [60., 218.]])
    sw = torch.tensor(thickness) if thickness is not None else torch.tensor(5.0)
    p = pydiffvg.Path(num_control_points=torch.tensor([2]), points=pts, is_closed=False, stroke_width=sw)
    g = pydiffvg.ShapeGroup(shape_ids=torch.tensor([0]),
                            fill_color=torch.tensor([0., 0., 0., 0.]) if fill else None,
                            stroke_color=torch.tensor([0.6, 0.3, 0.6, 0.8]))
    return 256, 256, [p], [g]


def painterly(num_paths=2048, canvas=512, seed=1234):
    """the exact call order reproduces the oracle sums in SURVEY 8c."""
    random.seed(seed)
    torch.manual_seed(seed)
    shapes, groups = [], []
    for i in range(num_paths):
        num_segments = random.randint(1, 3)
        num_control_points = torch.zeros(num_segments, dtype=torch.int32) + 2
        points = []
        p0 = (random.random(), random.random())
        points.append(p0)
        for j in range(num_segments):
            radius = 0.05
            p1 = (p0[0] + radius * (random.random() - 0.5), p0[1] + radius * (random.random() - 0.5))
            p2 = (p1[0] + radius * (random.random() - 0.5), p1[1] + radius * (random.random() - 0.5))
            p3 = (p2[0] + radius * (random.random() - 0.5), p2[1] + radius * (random.random() - 0.5))
            points.append(p1)
            points.append(p2)
            points.append(p3)
            p0 = p3
        points = torch.tensor(points)
        points[:, 0] *= canvas
        points[:, 1] *= canvas
        path = pydiffvg.Path(num_control_points=num_control_points, points=points,
                             stroke_width=torch.tensor(1.0 + 3.0 * random.random()), is_closed=False)
        shapes.append(path)
        groups.append(pydiffvg.ShapeGroup(
            shape_ids=torch.tensor([len(shapes) - 1]), fill_color=None,
            stroke_color=torch.tensor([random.random(), random.random(), random.random(), random.random()])))
    return canvas, canvas, shapes, groups


def blobs(num_paths=1024, canvas=512, seed=1234):
    """Closed filled cubic blobs (painterly_rendering.py:45-74 `use_blob`)."""
    random.seed(seed)
    torch.manual_seed(seed)
    shapes, groups = [], []
    for i in range(num_paths):
        num_segments = random.randint(3, 5)
        num_control_points = torch.zeros(num_segments, dtype=torch.int32) + 2
        points = []
        p0 = (random.random(), random.random())
        points.append(p0)
        for j in range(num_segments):
            radius = 0.05
            p1 = (p0[0] + radius * (random.random() - 0.5), p0[1] + radius * (random.random() - 0.5))
            p2 = (p1[0] + radius * (random.random() - 0.5), p1[1] + radius * (random.random() - 0.5))
            p3 = (p2[0] + radius * (random.random() - 0.5), p2[1] + radius * (random.random() - 0.5))
            points.append(p1)
            points.append(p2)
            if j < num_segments - 1:
                points.append(p3)
                p0 = p3
        points = torch.tensor(points)
        points[:, 0] *= canvas
        points[:, 1] *= canvas
        path = pydiffvg.Path(num_control_points=num_control_points, points=points,
                             stroke_width=torch.tensor(1.0), is_closed=True)
        shapes.append(path)
        groups.append(pydiffvg.ShapeGroup(
            shape_ids=torch.tensor([len(shapes) - 1]),
            fill_color=torch.tensor([random.random(), random.random(), random.random(), random.random()])))
    return canvas, canvas, shapes, groups


def zoo(canvas=128, seed=7):
    """A small scene touching every primitive / colour / transform feature at once."""
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.rand(*s, generator=g)
    shapes, groups = [], []
    # 0: filled + stroked circle
    shapes.append(pydiffvg.Circle(radius=torch.tensor(18.0), center=torch.tensor([40.0, 36.0]), stroke_width=torch.tensor(2.5)))
    groups.append(pydiffvg.ShapeGroup(torch.tensor([0]), fill_color=torch.tensor([0.9, 0.2, 0.1, 0.7]),
                                      stroke_color=torch.tensor([0.1, 0.1, 0.8, 1.0])))
    # 1: filled ellipse with a linear gradient
    shapes.append(pydiffvg.Ellipse(radius=torch.tensor([22.0, 12.0]), center=torch.tensor([84.0, 40.0])))
    groups.append(pydiffvg.ShapeGroup(torch.tensor([1]), fill_color=pydiffvg.LinearGradient(
        begin=torch.tensor([60.0, 30.0]), end=torch.tensor([108.0, 52.0]), offsets=torch.tensor([0.0, 0.5, 1.0]),
        stop_colors=torch.tensor([[0.2, 0.5, 0.7, 1.0], [0.7, 0.2, 0.5, 0.8], [0.1, 0.9, 0.3, 1.0]]))))
    # 2: rect, fill + stroke, translated (the reference asserts on rotated rects in the backward
    #    pass: accumulate_boundary_gradient compares the normal with the axis directions exactly)
    shapes.append(pydiffvg.Rect(p_min=torch.tensor([-14.0, -9.0]), p_max=torch.tensor([14.0, 9.0]), stroke_width=torch.tensor(1.5)))
    groups.append(pydiffvg.ShapeGroup(torch.tensor([2]), fill_color=torch.tensor([0.2, 0.8, 0.8, 0.6]),
                                      stroke_color=torch.tensor([0.0, 0.0, 0.0, 0.9]),
                                      shape_to_canvas=torch.tensor([[1.0, 0.0, 40.0], [0.0, 1.0, 92.0], [0.0, 0.0, 1.0]])))
    # 3: closed path mixing line / quadratic / cubic, radial gradient, non-zero rule, rotated + scaled
    shapes.append(pydiffvg.Path(num_control_points=torch.tensor([0, 1, 2, 0]),
                                points=torch.tensor([[-25.0, -25.0], [15.0, -23.0], [27.0, -5.0], [17.0, 15.0],
                                                     [5.0, 30.0], [-15.0, 25.0], [-27.0, 17.0]]),
                                is_closed=True, stroke_width=torch.tensor(2.0)))
    c, s = 0.8660254 * 0.9, 0.5 * 0.9
    groups.append(pydiffvg.ShapeGroup(torch.tensor([3]), use_even_odd_rule=False, fill_color=pydiffvg.RadialGradient(
        center=torch.tensor([95.0, 95.0]), radius=torch.tensor([30.0, 24.0]), offsets=torch.tensor([0.1, 0.9]),
        stop_colors=torch.tensor([[0.9, 0.9, 0.1, 1.0], [0.3, 0.1, 0.6, 0.5]])),
        stroke_color=torch.tensor([0.4, 0.2, 0.1, 1.0]),
        shape_to_canvas=torch.tensor([[c, -s, 95.0], [s, c, 95.0], [0.0, 0.0, 1.0]])))
    # 4: open quadratic + cubic stroke with per-point thickness
    shapes.append(pydiffvg.Path(num_control_points=torch.tensor([1, 2]),
                                points=torch.tensor([[10.0, 70.0], [30.0, 50.0], [50.0, 75.0], [60.0, 95.0], [30.0, 100.0], [15.0, 118.0]]),
                                is_closed=False, stroke_width=torch.tensor([3.0, 1.5, 2.0, 4.0, 1.0, 2.5])))
    groups.append(pydiffvg.ShapeGroup(torch.tensor([4]), fill_color=None, stroke_color=torch.tensor([0.5, 0.7, 0.2, 0.75])))
    # 5: a group of two shapes (polygon + circle), even-odd, translated
    shapes.append(pydiffvg.Polygon(points=torch.tensor([[0.0, 0.0], [30.0, 4.0], [26.0, 28.0], [4.0, 22.0]]), is_closed=True,
                                   stroke_width=torch.tensor(1.0)))
    shapes.append(pydiffvg.Circle(radius=torch.tensor(9.0), center=torch.tensor([14.0, 13.0]), stroke_width=torch.tensor(1.0)))
    groups.append(pydiffvg.ShapeGroup(torch.tensor([5, 6]), fill_color=torch.tensor([0.3, 0.3, 0.9, 0.85]),
                                      stroke_color=torch.tensor([0.9, 0.6, 0.1, 1.0]),
                                      shape_to_canvas=torch.tensor([[1.0, 0.0, 88.0], [0.0, 1.0, 4.0], [0.0, 0.0, 1.0]])))
    # 6: open polyline stroke
    shapes.append(pydiffvg.Polygon(points=torch.tensor([[5.0, 5.0], [25.0, 12.0], [12.0, 28.0], [30.0, 30.0]]), is_closed=False,
                                   stroke_width=torch.tensor(1.25)))
    groups.append(pydiffvg.ShapeGroup(torch.tensor([7]), fill_color=None, stroke_color=torch.tensor([0.0, 0.5, 0.5, 1.0])))
    return canvas, canvas, shapes, groups


def batched_strokes(scene_index, canvas=64, num_strokes=16):
    """one scene of the batched config (generative_models/rendering.py:239-307)."""
    g = torch.Generator().manual_seed(1000 + scene_index)
    shapes, groups = [], []
    for i in range(num_strokes):
        pts = torch.rand(4, 2, generator=g) * canvas
        w = 0.5 + 2.0 * torch.rand(1, generator=g)[0]
        a = torch.rand(1, generator=g)[0]
        shapes.append(pydiffvg.Path(num_control_points=torch.tensor([2]), points=pts, is_closed=False, stroke_width=w))
        groups.append(pydiffvg.ShapeGroup(torch.tensor([i]), fill_color=None,
                                          stroke_color=torch.stack([torch.tensor(1.0), torch.tensor(1.0), torch.tensor(1.0), a])))
    return canvas, canvas, shapes, groups


def zoo_prefilter(canvas=128):
    """Shapes the reference's SDF-prefiltering path supports (closest_point asserts on ellipses and never
    finds circles, compute_distance.h:18-24, 354-372): rect, mixed path, thick open path, polygon +
    filled circle group, open cubic stroke over a closed blob."""
    shapes, groups = [], []
    shapes.append(pydiffvg.Rect(p_min=torch.tensor([-14.0, -9.0]), p_max=torch.tensor([14.0, 9.0]), stroke_width=torch.tensor(1.5)))
    groups.append(pydiffvg.ShapeGroup(torch.tensor([0]), fill_color=torch.tensor([0.2, 0.8, 0.8, 0.6]),
                                      stroke_color=torch.tensor([0.0, 0.0, 0.0, 0.9]),
                                      shape_to_canvas=torch.tensor([[1.0, 0.0, 40.0], [0.0, 1.0, 92.0], [0.0, 0.0, 1.0]])))
    shapes.append(pydiffvg.Path(num_control_points=torch.tensor([0, 1, 2, 0]),
                                points=torch.tensor([[-25.0, -25.0], [15.0, -23.0], [27.0, -5.0], [17.0, 15.0],
                                                     [5.0, 30.0], [-15.0, 25.0], [-27.0, 17.0]]),
                                is_closed=True, stroke_width=torch.tensor(2.0)))
    c, s = 0.8660254 * 0.9, 0.5 * 0.9
    groups.append(pydiffvg.ShapeGroup(torch.tensor([1]), use_even_odd_rule=False, fill_color=pydiffvg.RadialGradient(
        center=torch.tensor([95.0, 95.0]), radius=torch.tensor([30.0, 24.0]), offsets=torch.tensor([0.1, 0.9]),
        stop_colors=torch.tensor([[0.9, 0.9, 0.1, 1.0], [0.3, 0.1, 0.6, 0.5]])),
        stroke_color=torch.tensor([0.4, 0.2, 0.1, 1.0]),
        shape_to_canvas=torch.tensor([[c, -s, 95.0], [s, c, 95.0], [0.0, 0.0, 1.0]])))
    shapes.append(pydiffvg.Path(num_control_points=torch.tensor([1, 2]),
                                points=torch.tensor([[10.0, 70.0], [30.0, 50.0], [50.0, 75.0], [60.0, 95.0], [30.0, 100.0], [15.0, 118.0]]),
                                is_closed=False, stroke_width=torch.tensor(2.25)))
    groups.append(pydiffvg.ShapeGroup(torch.tensor([2]), fill_color=None, stroke_color=torch.tensor([0.5, 0.7, 0.2, 0.75])))
    shapes.append(pydiffvg.Polygon(points=torch.tensor([[0.0, 0.0], [30.0, 4.0], [26.0, 28.0], [4.0, 22.0]]), is_closed=True,
                                   stroke_width=torch.tensor(1.0)))
    shapes.append(pydiffvg.Circle(radius=torch.tensor(9.0), center=torch.tensor([14.0, 13.0]), stroke_width=torch.tensor(1.0)))
    groups.append(pydiffvg.ShapeGroup(torch.tensor([3, 4]), fill_color=torch.tensor([0.3, 0.3, 0.9, 0.85]),
                                      stroke_color=torch.tensor([0.9, 0.6, 0.1, 1.0]),
                                      shape_to_canvas=torch.tensor([[1.0, 0.0, 88.0], [0.0, 1.0, 4.0], [0.0, 0.0, 1.0]])))
    shapes.append(pydiffvg.Path(num_control_points=torch.tensor([2, 2, 2]),
                                points=torch.tensor([[20.0, 20.0], [35.0, 8.0], [52.0, 12.0], [60.0, 28.0], [66.0, 42.0],
                                                     [48.0, 56.0], [34.0, 50.0], [22.0, 46.0], [12.0, 34.0]]),
                                is_closed=True, stroke_width=torch.tensor(1.0)))
    groups.append(pydiffvg.ShapeGroup(torch.tensor([5]), fill_color=pydiffvg.LinearGradient(
        begin=torch.tensor([20.0, 10.0]), end=torch.tensor([60.0, 50.0]), offsets=torch.tensor([0.0, 1.0]),
        stop_colors=torch.tensor([[0.9, 0.2, 0.1, 0.9], [0.1, 0.3, 0.9, 0.7]]))))
    shapes.append(pydiffvg.Path(num_control_points=torch.tensor([2]),
                                points=torch.tensor([[15.0, 15.0], [50.0, 20.0], [30.0, 60.0], [70.0, 55.0]]),
                                is_closed=False, stroke_width=torch.tensor(3.0)))
    groups.append(pydiffvg.ShapeGroup(torch.tensor([6]), fill_color=None, stroke_color=torch.tensor([0.1, 0.1, 0.1, 0.8])))
    return canvas, canvas, shapes, groups
