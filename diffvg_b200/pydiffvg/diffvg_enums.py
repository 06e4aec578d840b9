"""The enum surface apps use from the reference's pybind11 module `diffvg`
(diffvg.cpp:1651-1792): diffvg.FilterType.box, diffvg.ShapeType.path, ...
Numeric values match the C ABI (include/dvg_scene_format.h)."""
from enum import IntEnum


class ShapeType(IntEnum):
    circle = 0
    ellipse = 1
    path = 2
    rect = 3


class ColorType(IntEnum):
    constant = 0
    linear_gradient = 1
    radial_gradient = 2


class FilterType(IntEnum):
    box = 0
    tent = 1
    parabolic = 2
    hann = 3
