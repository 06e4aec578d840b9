import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


def pytest_report_header(config):
    try:
        import oracle_check
        return 'oracle: %s (reference = oracle/_ref, the compiled unmodified reference; port = oracle/dvg_oracle.c, forward colour path only)' % oracle_check.kind()
    except Exception as e:   # pragma: no cover
        return 'oracle: unavailable (%r)' % (e,)


def pytest_collection_modifyitems(config, items):
    """The GPU parity suite is only meaningful against the compiled reference: without oracle/_ref every backward,
    prefilter and SDF comparison would silently turn into a skip.  Fail loudly instead (DVG_ALLOW_PORT_ORACLE=1 opts out)."""
    import pytest
    if os.environ.get('DVG_ALLOW_PORT_ORACLE') == '1':
        return
    gpu_items = [it for it in items if it.get_closest_marker('gpu') is not None]
    if not gpu_items:
        return
    markexpr = config.getoption('-m') or ''
    if 'not gpu' in markexpr:
        return
    try:
        import torch
        if not torch.cuda.is_available():
            return      # no GPU here: the gpu tests fail on their own, for the right reason
        import oracle_check
        kind = oracle_check.kind()
    except Exception:
        kind = 'unavailable'
    if kind != 'reference':
        def _fail():
            pytest.fail("GPU parity tests need oracle/_ref (the compiled reference): found oracle kind %r. Build it with "
                        "`make -C oracle ref` where /root/reference exists, or set DVG_ALLOW_PORT_ORACLE=1." % kind)
        gpu_items[0].runtest = _fail   # one loud failure (with -x it also ends the run) instead of a row of silent skips
