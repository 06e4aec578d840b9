"""TEST INFRASTRUCTURE (oracle) -- single front door for the checkers.

`render` / `scene_dump` dispatch to oracle/_ref (the UNMODIFIED reference renderer compiled from
/root/reference by oracle/Makefile; ref_oracle.py) when it has been built, else to the plain-C
restatement (oracle/liboracle.so; c_oracle.py), which covers the forward colour path only.  What neither
can answer raises OracleUnavailable: under pytest that is a *skip* of the rest of the test (whatever was
compared before it still counts).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this; the product package never does.
"""
import os

import ref_oracle
import c_oracle

try:   # under pytest "the oracle cannot answer this" skips instead of failing
    import pytest as _pytest
    _Base = _pytest.skip.Exception
except Exception:   # pragma: no cover
    _Base = RuntimeError


class OracleUnavailable(_Base):
    pass


def _use_ref():
    return ref_oracle.available() and os.environ.get('DVG_ORACLE', '') != 'port'


def kind():
    """'reference' when the compiled reference is available, else 'port'."""
    return 'reference' if _use_ref() else 'port'


def render(*args, **kwargs):
    if _use_ref():
        return ref_oracle.render(*args, **kwargs)
    forward_colour = (kwargs.get('d_render_image') is None and kwargs.get('d_render_sdf') is None and
                      not kwargs.get('use_prefiltering') and not kwargs.get('want_sdf') and kwargs.get('want_image', True))
    if not forward_colour:
        raise OracleUnavailable('oracle/_ref (the compiled reference) is not built here and oracle/dvg_oracle.c restates the '
                                'forward colour path only')
    kwargs.pop('variant', None)
    topo, _params, width, height, nsx, nsy = args[:6]
    if float(width) * height * nsx * nsy * int(topo[4]) > 5e9:   # samples x shape groups: single-threaded, no culling
        raise OracleUnavailable('too large for the single-threaded C restatement; build oracle/_ref')
    return c_oracle.render(*args, **kwargs)


def scene_dump(*args, **kwargs):
    if _use_ref():
        return ref_oracle.scene_dump(*args, **kwargs)
    raise OracleUnavailable('scene dumps need oracle/_ref')
