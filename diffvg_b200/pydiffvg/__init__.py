"""Drop-in mirror of the reference `pydiffvg` package surface that sits on the hot path
(reference pydiffvg/__init__.py:1-9; parse_svg / save_svg / optimize_svg are out of scope,
SURVEY section 2 rows 18-20)."""
from .device import *  # noqa: F401,F403
from .shape import *  # noqa: F401,F403
from .pixel_filter import *  # noqa: F401,F403
from .color import *  # noqa: F401,F403
from .render_pytorch import *  # noqa: F401,F403
