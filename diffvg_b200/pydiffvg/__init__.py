"""Drop-in mirror of the reference `pydiffvg` package surface (reference pydiffvg/__init__.py:1-9).
`optimize_svg` (an SVG-level optimiser front end) is out of scope (apps and optimisers)."""
from .device import *  # noqa: F401,F403
from .shape import *  # noqa: F401,F403
from .pixel_filter import *  # noqa: F401,F403
from .diffvg_enums import FilterType, ShapeType, ColorType  # noqa: F401  (the reference keeps these in its pybind module `diffvg`)
from .render_pytorch import *  # noqa: F401,F403
from .image import *  # noqa: F401,F403
from .parse_svg import *  # noqa: F401,F403
from .color import *  # noqa: F401,F403
from .save_svg import *  # noqa: F401,F403
from .packed_params import *  # noqa: F401,F403
from .batched import *  # noqa: F401,F403
