"""SVG path-data reader behind `pydiffvg.from_svg_path` (reference pydiffvg/shape.py:63-172).

The reference delegates this to the third-party package `svgpathtools` (setup.py:94 `install_requires =
["svgpathtools"]`, unpinned; README.md:28), which is not vendored under the reference tree and is absent
from this image.  What the reference needs from it is restated here from svgpathtools' published
behaviour (release 1.4.x, current at the reference's commit):

  * `parse_path`: tokens = command letters and floats matched by `[-+]?[0-9]*\\.?[0-9]+(?:[eE][-+]?[0-9]+)?`;
    commands M m L l H h V v C c S s Q q T t A a Z z with implicit repetition (coordinates after a moveto
    are linetos); coordinates are Python complex numbers (float64); `Z` appends a closing Line only when the
    current point differs from the sub-path start;
  * `Path.continuous_subpaths()`: the segment list is cut wherever one segment's end is not exactly the next
    one's start; `isclosed()` is `start == end` of such a run (the Z flag itself is not consulted);
  * `Arc`: endpoint -> centre parameterisation of the SVG implementation notes (F.6.5), radii scaled up when
    no ellipse fits; `theta` / `delta` in degrees, `phi` in radians.

Segments are plain tuples: ('L', start, end), ('Q', start, control, end), ('C', start, c1, c2, end),
('A', start, end, dict(radius, center, theta, delta, phi)).
"""
import cmath
import math
import re

_COMMANDS = set('MmZzLlHhVvCcSsQqTtAa')
_COMMAND_RE = re.compile(r'([MmZzLlHhVvCcSsQqTtAa])')
_FLOAT_RE = re.compile(r'[-+]?[0-9]*\.?[0-9]+(?:[eE][-+]?[0-9]+)?')


def _tokens(d):
    for chunk in _COMMAND_RE.split(d):
        if chunk in _COMMANDS:
            yield chunk
        else:
            for tok in _FLOAT_RE.findall(chunk):
                yield tok


def _clip(v, lo, hi):
    return lo if v < lo else (hi if v > hi else v)


def _arc(start, radius, rotation, large_arc, sweep, end):
    """Centre parameterisation of one elliptical arc; returns the fields from_svg_path reads."""
    rx, ry = abs(radius.real), abs(radius.imag)
    large_arc, sweep = bool(large_arc), bool(sweep)
    phi = math.radians(rotation)
    rot = cmath.exp(1j * phi)
    zp1 = (1 / rot) * (start - end) / 2
    x1p, y1p = zp1.real, zp1.imag
    check = (x1p * x1p) / (rx * rx) + (y1p * y1p) / (ry * ry)
    if check > 1:   # no ellipse of these radii passes through both end points: scale the radii up
        rx *= math.sqrt(check)
        ry *= math.sqrt(check)
    rx2, ry2 = rx * rx, ry * ry
    tmp = rx2 * y1p * y1p + ry2 * x1p * x1p
    radicand = (rx2 * ry2 - tmp) / tmp
    radical = math.sqrt(radicand) if radicand >= 0 else 0.0
    cp = radical * (rx * y1p / ry - 1j * ry * x1p / rx)
    if large_arc == sweep:
        cp = -cp
    center = rot * cp + (start + end) / 2
    u1 = complex(_clip((x1p - cp.real) / rx, -1, 1), _clip((y1p - cp.imag) / ry, -1, 1))
    u2 = complex(_clip((-x1p - cp.real) / rx, -1, 1), _clip((-y1p - cp.imag) / ry, -1, 1))
    if u1.imag > 0:
        theta = math.degrees(math.acos(u1.real))
    elif u1.imag < 0:
        theta = -math.degrees(math.acos(u1.real))
    else:
        theta = 0 if u1.real > 0 else 180
    det = u1.real * u2.imag - u1.imag * u2.real
    dot = u1.real * u2.real + u1.imag * u2.imag
    if dot > 1 or dot < -1:
        dot = round(dot)
    if det > 0:
        delta = math.degrees(math.acos(dot))
    elif det < 0:
        delta = -math.degrees(math.acos(dot))
    else:
        delta = 0 if dot > 0 else 180
    if not sweep and delta >= 0:
        delta -= 360
    elif large_arc and delta <= 0:
        delta += 360
    return dict(radius=complex(rx, ry), center=center, theta=theta, delta=delta, phi=phi)


def parse_path(d):
    """Path data string -> flat list of segments in drawing order."""
    toks = list(_tokens(d))
    toks.reverse()
    pop = toks.pop

    def num():
        return float(pop())

    def pt():
        x = float(pop())
        y = float(pop())
        return complex(x, y)

    segs = []
    cur = 0j
    start = None
    cmd = None
    last = None
    absolute = True
    while toks:
        if toks[-1] in _COMMANDS:
            last = cmd
            c = pop()
            absolute = c.isupper()
            cmd = c.upper()
        else:
            if cmd is None:
                raise ValueError('path data: coordinates without a command in %r' % d)
            last = cmd
        if cmd == 'M':
            p = pt()
            cur = p if absolute else cur + p
            start = cur
            cmd = 'L'   # further pairs are implicit linetos
        elif cmd == 'Z':
            if not (cur == start):
                segs.append(('L', cur, start))
            cur = start
            cmd = None
        elif cmd == 'L':
            p = pt()
            if not absolute:
                p += cur
            segs.append(('L', cur, p))
            cur = p
        elif cmd == 'H':
            p = complex(num(), cur.imag)
            if not absolute:
                p += cur.real
            segs.append(('L', cur, p))
            cur = p
        elif cmd == 'V':
            p = complex(cur.real, num())
            if not absolute:
                p += cur.imag * 1j
            segs.append(('L', cur, p))
            cur = p
        elif cmd == 'C':
            c1, c2, e = pt(), pt(), pt()
            if not absolute:
                c1 += cur
                c2 += cur
                e += cur
            segs.append(('C', cur, c1, c2, e))
            cur = e
        elif cmd == 'S':
            # first control point: reflection of the previous cubic's second one, else the current point
            c1 = cur + cur - segs[-1][3] if (last is not None and last in 'CS') else cur
            c2, e = pt(), pt()
            if not absolute:
                c2 += cur
                e += cur
            segs.append(('C', cur, c1, c2, e))
            cur = e
        elif cmd == 'Q':
            c1, e = pt(), pt()
            if not absolute:
                c1 += cur
                e += cur
            segs.append(('Q', cur, c1, e))
            cur = e
        elif cmd == 'T':
            c1 = cur + cur - segs[-1][2] if (last is not None and last in 'QT') else cur
            e = pt()
            if not absolute:
                e += cur
            segs.append(('Q', cur, c1, e))
            cur = e
        elif cmd == 'A':
            radius = pt()
            rotation, large_arc, sweep = num(), num(), num()
            e = pt()
            if not absolute:
                e += cur
            segs.append(('A', cur, e, _arc(cur, radius, rotation, large_arc, sweep, e)))
            cur = e
    return segs


def seg_start(s):
    return s[1]


def seg_end(s):
    return s[2] if s[0] in ('L', 'A') else s[-1]


def with_end(s, e):
    if s[0] == 'L':
        return ('L', s[1], e)
    if s[0] == 'A':
        return ('A', s[1], e, s[3])
    return s[:-1] + (e,)


def continuous_subpaths(segs):
    """Runs of segments that join exactly end-to-start."""
    if not segs:
        return []
    out, first = [], 0
    for i in range(len(segs) - 1):
        if seg_end(segs[i]) != seg_start(segs[i + 1]):
            out.append(list(segs[first:i + 1]))
            first = i + 1
    out.append(list(segs[first:]))
    return out
