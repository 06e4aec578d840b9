"""TEST INFRASTRUCTURE (oracle) -- single front door for the checkers.

`render` / `scene_dump` dispatch to oracle/_ref (the UNMODIFIED reference renderer compiled from
/root/reference by oracle/Makefile; ref_oracle.py) when it has been built, else to the plain-C
restatement (oracle/liboracle.so; c_oracle.py).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this; the product package never does.
"""
import ref_oracle
import c_oracle


def kind():
    """'reference' when the compiled reference is available, else 'port'."""
    return 'reference' if ref_oracle.available() else 'port'


def render(*args, **kwargs):
    if ref_oracle.available():
        return ref_oracle.render(*args, **kwargs)
    return c_oracle.render(*args, **kwargs)


def scene_dump(*args, **kwargs):
    if ref_oracle.available():
        return ref_oracle.scene_dump(*args, **kwargs)
    return c_oracle.scene_dump(*args, **kwargs)
