"""diffvg_b200: B200-native differentiable vector-graphics rasteriser (hot path of
BachiLi/diffvg behind the unchanged pydiffvg API).  See DESIGN.md."""
__version__ = '0.1.0'
