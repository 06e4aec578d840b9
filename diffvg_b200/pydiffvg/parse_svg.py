"""`svg_to_scene`: SVG file -> (canvas_width, canvas_height, shapes, shape_groups), the loader every
reference app that starts from an asset goes through (reference pydiffvg/parse_svg.py:392-586; used by
apps/render_svg.py, refine_svg.py, finite_difference_comp.py on apps/imgs/tiger.svg / flower.svg).

The reference leans on three third-party packages that are not part of its tree and are absent here:
svgpathtools (path data: see svg_path.py), cssutils (the `<style>` class rules) and matplotlib.colors
(named colours).  The last two are replaced by a rule splitter and the CSS colour-keyword table below.

Behaviour kept on purpose (drop-in parity with what the reference builds from the same file):
  * stroke-width is halved into a stroke RADIUS, default radius 0.5 (parse_svg.py:301, 327-331);
  * `<path>` points are multiplied by the accumulated transform and the group keeps the identity;
    every other element keeps its own coordinates and passes the transform as `shape_to_canvas`;
  * a fill inherited from a `<g>` / a style class is the SAME tensor object for every shape that inherits
    it, and its alpha is overwritten in place by each shape's opacity (parse_svg.py:311-317);
  * `<rect>`: `x` / `y` are never read (the reference tests `0.0 in node.attrib`) and p_max is
    (x + width, x + height) (parse_svg.py:496-505); `<ellipse>` is not dispatched (is_shape, :391);
  * fill defaults to opaque black; filled paths are force-closed; default fill rule is non-zero.
"""
import os
import re
import warnings
import xml.etree.ElementTree as etree

import numpy as np
import torch

from .color import LinearGradient, RadialGradient
from .shape import Circle, Polygon, Rect, ShapeGroup, from_svg_path

__all__ = ['svg_to_scene', 'parse_scene', 'parse_transform', 'parse_color', 'parse_style']


def remove_namespaces(s):
    return re.sub('{.*}', '', s)


# CSS colour keywords (CSS Color Module Level 4, the table matplotlib.colors.to_rgba resolves names with)
_CSS_COLORS = dict(
    aliceblue='f0f8ff', antiquewhite='faebd7', aqua='00ffff', aquamarine='7fffd4', azure='f0ffff', beige='f5f5dc',
    bisque='ffe4c4', black='000000', blanchedalmond='ffebcd', blue='0000ff', blueviolet='8a2be2', brown='a52a2a',
    burlywood='deb887', cadetblue='5f9ea0', chartreuse='7fff00', chocolate='d2691e', coral='ff7f50',
    cornflowerblue='6495ed', cornsilk='fff8dc', crimson='dc143c', cyan='00ffff', darkblue='00008b',
    darkcyan='008b8b', darkgoldenrod='b8860b', darkgray='a9a9a9', darkgreen='006400', darkgrey='a9a9a9',
    darkkhaki='bdb76b', darkmagenta='8b008b', darkolivegreen='556b2f', darkorange='ff8c00', darkorchid='9932cc',
    darkred='8b0000', darksalmon='e9967a', darkseagreen='8fbc8f', darkslateblue='483d8b', darkslategray='2f4f4f',
    darkslategrey='2f4f4f', darkturquoise='00ced1', darkviolet='9400d3', deeppink='ff1493', deepskyblue='00bfff',
    dimgray='696969', dimgrey='696969', dodgerblue='1e90ff', firebrick='b22222', floralwhite='fffaf0',
    forestgreen='228b22', fuchsia='ff00ff', gainsboro='dcdcdc', ghostwhite='f8f8ff', gold='ffd700',
    goldenrod='daa520', gray='808080', green='008000', greenyellow='adff2f', grey='808080', honeydew='f0fff0',
    hotpink='ff69b4', indianred='cd5c5c', indigo='4b0082', ivory='fffff0', khaki='f0e68c', lavender='e6e6fa',
    lavenderblush='fff0f5', lawngreen='7cfc00', lemonchiffon='fffacd', lightblue='add8e6', lightcoral='f08080',
    lightcyan='e0ffff', lightgoldenrodyellow='fafad2', lightgray='d3d3d3', lightgreen='90ee90', lightgrey='d3d3d3',
    lightpink='ffb6c1', lightsalmon='ffa07a', lightseagreen='20b2aa', lightskyblue='87cefa', lightslategray='778899',
    lightslategrey='778899', lightsteelblue='b0c4de', lightyellow='ffffe0', lime='00ff00', limegreen='32cd32',
    linen='faf0e6', magenta='ff00ff', maroon='800000', mediumaquamarine='66cdaa', mediumblue='0000cd',
    mediumorchid='ba55d3', mediumpurple='9370db', mediumseagreen='3cb371', mediumslateblue='7b68ee',
    mediumspringgreen='00fa9a', mediumturquoise='48d1cc', mediumvioletred='c71585', midnightblue='191970',
    mintcream='f5fffa', mistyrose='ffe4e1', moccasin='ffe4b5', navajowhite='ffdead', navy='000080',
    oldlace='fdf5e6', olive='808000', olivedrab='6b8e23', orange='ffa500', orangered='ff4500', orchid='da70d6',
    palegoldenrod='eee8aa', palegreen='98fb98', paleturquoise='afeeee', palevioletred='db7093', papayawhip='ffefd5',
    peachpuff='ffdab9', peru='cd853f', pink='ffc0cb', plum='dda0dd', powderblue='b0e0e6', purple='800080',
    rebeccapurple='663399', red='ff0000', rosybrown='bc8f8f', royalblue='4169e1', saddlebrown='8b4513',
    salmon='fa8072', sandybrown='f4a460', seagreen='2e8b57', seashell='fff5ee', sienna='a0522d', silver='c0c0c0',
    skyblue='87ceeb', slateblue='6a5acd', slategray='708090', slategrey='708090', snow='fffafa',
    springgreen='00ff7f', steelblue='4682b4', tan='d2b48c', teal='008080', thistle='d8bfd8', tomato='ff6347',
    turquoise='40e0d0', violet='ee82ee', wheat='f5deb3', white='ffffff', whitesmoke='f5f5f5', yellow='ffff00',
    yellowgreen='9acd32')


def _hex_rgb(s):
    s = s.lstrip('#')
    if len(s) == 3:
        s = ''.join(ch + ch for ch in s)
    return [int(s[i:i + 2], 16) / 255.0 for i in (0, 2, 4)]


def parse_color(s, defs):
    """Colour attribute -> tensor[4] (alpha 1), a gradient holder from `defs` (url(#id)), or None."""
    if s is None or isinstance(s, torch.Tensor):
        return s
    s = s.lstrip(' ')
    if s == 'none':
        return None
    if s[0] == '#':
        return torch.tensor(_hex_rgb(s) + [1.0])
    if s[:3] == 'url':
        return defs[s[4:-1].lstrip('#')]
    if s[:4] == 'rgb(':
        r, g, b = s[4:-1].split(',')[:3]
        return torch.tensor([int(r) / 255.0, int(g) / 255.0, int(b) / 255.0, 1.0])
    name = s.strip().lower()
    if name in _CSS_COLORS:
        return torch.tensor(_hex_rgb(_CSS_COLORS[name]) + [1.0])
    warnings.warn('Unknown color command ' + s)
    return torch.tensor([0.0, 0.0, 0.0, 1.0])


def parse_style(s, defs):
    """'key:value;key:value' -> dict; fill / stroke values become colours at once so that shapes of one class
    share one tensor."""
    out = {}
    for decl in s.split(';'):
        kv = decl.split(':')
        if len(kv) == 2:
            key, value = kv[0].strip(), kv[1].strip()
            out[key] = parse_color(value, defs) if key in ('fill', 'stroke') else value
    return out


def _one_transform(item):
    kind, values = item.split('(')
    v = [float(x) for x in values.replace(',', ' ').split(' ') if x]
    m = np.identity(3)
    if 'matrix' in kind:
        m[0:2, 0:3] = np.array([v[0:6:2], v[1:6:2]])
    elif 'translate' in item:
        m[0, 2] = v[0]
        if len(v) > 1:
            m[1, 2] = v[1]
    elif 'scale' in item:
        m[0, 0] = v[0]
        m[1, 1] = v[1] if len(v) > 1 else v[0]
    elif 'rotate' in item:
        a = v[0] * np.pi / 180.0
        ox, oy = (v[1], v[2]) if len(v) == 3 else (0.0, 0.0)
        to = np.identity(3)
        to[0, 2], to[1, 2] = ox, oy
        back = np.identity(3)
        back[0, 2], back[1, 2] = -ox, -oy
        r = np.identity(3)
        r[0:2, 0:2] = np.array([[np.cos(a), -np.sin(a)], [np.sin(a), np.cos(a)]])
        m = to.dot(r).dot(back)
    elif 'skewX' in item:
        m[0, 1] = np.tan(v[0] * np.pi / 180.0)
    elif 'skewY' in item:
        m[1, 0] = np.tan(v[0] * np.pi / 180.0)
    else:
        warnings.warn('Unknown SVG transform type: {0}'.format(kind))
    return m


def parse_transform(transform_str):
    """SVG transform list -> float32 3x3 (float64 product, left to right); identity for an empty string."""
    if not transform_str:
        return np.identity(3)
    if not isinstance(transform_str, str):
        raise TypeError('Must provide a string to parse')
    total = np.identity(3)
    for item in transform_str.split(')')[:-1]:
        total = total.dot(_one_transform(item))
    return torch.from_numpy(total).type(torch.float32)


def _apply(transform, p):
    q = transform @ torch.cat((p, torch.ones([1])))
    return (q / q[2])[:2]


def _gradient_stops(node, defs, offsets, stop_colors):
    for child in node:
        if remove_namespaces(child.tag) != 'stop':
            continue
        color = [0.0, 0.0, 0.0, 1.0]
        sources = [child.attrib]
        if 'style' in child.attrib:
            sources.append(parse_style(child.attrib['style'], defs))
        for src in sources:
            if 'stop-color' in src:
                c = parse_color(src['stop-color'], defs)
                color[:3] = [c[0], c[1], c[2]]
            if 'stop-opacity' in src:
                color[3] = float(src['stop-opacity'])
        offsets.append(float(child.attrib['offset']))
        stop_colors.append(color)
    if isinstance(offsets, list):
        offsets = torch.tensor(offsets)
    if isinstance(stop_colors, list):
        stop_colors = torch.tensor(stop_colors)
    return offsets, stop_colors


def _inherit_gradient(node, defs):
    begin, end, offsets, stop_colors = torch.tensor([0.0, 0.0]), torch.tensor([0.0, 0.0]), [], []
    for key in node.attrib:
        if remove_namespaces(key) == 'href':
            parent = defs[node.attrib[key].lstrip('#')]
            begin, end = parent.begin, parent.end
            offsets, stop_colors = parent.offsets, parent.stop_colors
    return begin, end, offsets, stop_colors


def parse_linear_gradient(node, transform, defs):
    begin, end, offsets, stop_colors = _inherit_gradient(node, defs)
    for attrib in node.attrib:
        name = remove_namespaces(attrib)
        if name == 'x1':
            begin[0] = float(node.attrib['x1'])
        elif name == 'y1':
            begin[1] = float(node.attrib['y1'])
        elif name == 'x2':
            end[0] = float(node.attrib['x2'])
        elif name == 'y2':
            end[1] = float(node.attrib['y2'])
        elif name == 'gradientTransform':
            transform = transform @ parse_transform(node.attrib['gradientTransform'])
    begin, end = _apply(transform, begin), _apply(transform, end)
    offsets, stop_colors = _gradient_stops(node, defs, offsets, stop_colors)
    return LinearGradient(begin, end, offsets, stop_colors)


def parse_radial_gradient(node, transform, defs):
    # the reference hands (begin, end) of the inherited gradient -- zeros without an href -- to the
    # RadialGradient constructor as (center, radius) (parse_svg.py:200-262, "TODO: this is incorrect" there)
    begin, end, offsets, stop_colors = _inherit_gradient(node, defs)
    for attrib in node.attrib:
        if remove_namespaces(attrib) == 'gradientTransform':
            transform = transform @ parse_transform(node.attrib['gradientTransform'])
    offsets, stop_colors = _gradient_stops(node, defs, offsets, stop_colors)
    return RadialGradient(begin, end, offsets, stop_colors)


_RULE_RE = re.compile(r'([^{}]+)\{([^{}]*)\}')


def parse_stylesheet(node, transform, defs):
    """`.name { key: value; ... }` rules of a <style> element -> defs[name] = style dict."""
    text = re.sub(r'/\*.*?\*/', '', node.text or '', flags=re.S)
    for selector, body in _RULE_RE.findall(text):
        name = selector.strip()
        if len(name) >= 2 and name[0] == '.':
            defs[name[1:]] = parse_style(body, defs)
    return defs


def parse_defs(node, transform, defs):
    for child in node:
        tag = remove_namespaces(child.tag)
        if tag == 'linearGradient' and 'id' in child.attrib:
            defs[child.attrib['id']] = parse_linear_gradient(child, transform, defs)
        elif tag == 'radialGradient' and 'id' in child.attrib:
            defs[child.attrib['id']] = parse_radial_gradient(child, transform, defs)
        elif tag == 'style':
            defs = parse_stylesheet(child, transform, defs)
    return defs


def _radius_from_width(text):
    if text[-2:] == 'px':
        text = text[:-2]
    return torch.tensor(float(text) / 2.0)


def _fill_rule(value, current):
    if value == 'evenodd':
        return True
    if value == 'nonzero':
        return False
    warnings.warn('Unknown fill-rule: {}'.format(value))
    return current


def parse_common_attrib(node, transform, fill_color, defs):
    """-> (transform, fill colour, stroke colour, stroke radius, even-odd flag) of one element."""
    attribs = {}
    if 'class' in node.attrib:
        attribs.update(defs[node.attrib['class']])
    attribs.update(node.attrib)
    name = node.attrib.get('id', '')
    stroke_color = None
    stroke_width = torch.tensor(0.5)
    use_even_odd_rule = False
    new_transform = transform
    if 'transform' in attribs:
        new_transform = transform @ parse_transform(attribs['transform'])
    if 'fill' in attribs:
        fill_color = parse_color(attribs['fill'], defs)
    fill_opacity = 1.0
    if 'fill-opacity' in attribs:
        fill_opacity *= float(attribs['fill-opacity'])
    if 'opacity' in attribs:
        fill_opacity *= float(attribs['opacity'])
    if isinstance(fill_color, torch.Tensor):   # gradients ignore opacity
        fill_color[3] = fill_opacity
    if 'fill-rule' in attribs:
        use_even_odd_rule = _fill_rule(attribs['fill-rule'], use_even_odd_rule)
    if 'stroke' in attribs:
        stroke_color = parse_color(attribs['stroke'], defs)
    if 'stroke-width' in attribs:
        stroke_width = _radius_from_width(attribs['stroke-width'])
    if 'stroke-opacity' in attribs:
        stroke_color[3] = torch.tensor(float(attribs['stroke-opacity']))
    if 'style' in attribs:
        style = parse_style(attribs['style'], defs)
        if 'fill' in style:
            fill_color = parse_color(style['fill'], defs)
        fill_opacity = 1.0
        if 'fill-opacity' in style:
            fill_opacity *= float(style['fill-opacity'])
        if 'opacity' in style:
            fill_opacity *= float(style['opacity'])
        if 'fill-rule' in style:
            use_even_odd_rule = _fill_rule(style['fill-rule'], use_even_odd_rule)
        if isinstance(fill_color, torch.Tensor):
            fill_color[3] = fill_opacity
        if 'stroke' in style:   # already a colour / None here, so the reference's "!= 'none'" guard always passes
            stroke_color = parse_color(style['stroke'], defs)
            if isinstance(stroke_color, torch.Tensor):
                if 'stroke-opacity' in style:
                    stroke_color[3] = float(style['stroke-opacity'])
                if 'opacity' in style:
                    stroke_color[3] *= float(style['opacity'])
            if 'stroke-width' in style:
                stroke_width = _radius_from_width(style['stroke-width'])
        for c in (fill_color, stroke_color):
            if isinstance(c, LinearGradient):
                c.begin, c.end = _apply(new_transform, c.begin), _apply(new_transform, c.end)
        if 'filter' in style:
            print('*** WARNING ***: Ignoring filter for path with id "{}"'.format(name))
    return new_transform, fill_color, stroke_color, stroke_width, use_even_odd_rule


def is_shape(tag):
    return tag in ('path', 'polygon', 'line', 'circle', 'rect')


def parse_shape(node, transform, fill_color, shapes, shape_groups, defs):
    tag = remove_namespaces(node.tag)
    new_transform, fill, stroke, stroke_width, even_odd = parse_common_attrib(node, transform, fill_color, defs)
    name = node.attrib.get('id', '')
    if tag == 'path':
        paths = from_svg_path(node.attrib['d'], new_transform, fill is not None)
        for idx, path in enumerate(paths):
            assert path.points.shape[1] == 2
            path.stroke_width = stroke_width
            path.source_id = name
            path.id = '{}-{}'.format(name, idx) if len(paths) > 1 else name
        first = len(shapes)
        shapes = shapes + paths
        shape_groups.append(ShapeGroup(shape_ids=torch.tensor(list(range(first, len(shapes)))), fill_color=fill,
                                       stroke_color=stroke, use_even_odd_rule=even_odd, id=name))
        return shapes, shape_groups
    if tag == 'polygon':
        pts = [[float(y) for y in re.split(',| ', x)] for x in node.attrib['points'].strip().split(' ') if x]
        shape = Polygon(torch.tensor(pts, dtype=torch.float32).view(-1, 2), fill is not None)
    elif tag == 'line':
        p1 = torch.tensor([float(node.attrib['x1']), float(node.attrib['y1'])])
        p2 = torch.tensor([float(node.attrib['x2']), float(node.attrib['y2'])])
        shape = Polygon(torch.stack((p1, p2)), False)
    elif tag == 'circle':
        shape = Circle(radius=torch.tensor(float(node.attrib['r'])),
                       center=torch.tensor([float(node.attrib['cx']), float(node.attrib['cy'])]))
    elif tag == 'rect':
        x = y = 0.0   # never read from the element in the reference (see the module docstring)
        w, h = float(node.attrib['width']), float(node.attrib['height'])
        shape = Rect(p_min=torch.tensor([x, y]), p_max=torch.tensor([x + w, x + h]))
    else:
        return shapes, shape_groups
    shape.stroke_width = stroke_width
    shape_ids = torch.tensor([len(shapes)])
    shapes.append(shape)
    kwargs = dict(shape_ids=shape_ids, fill_color=fill, stroke_color=stroke, use_even_odd_rule=even_odd,
                  shape_to_canvas=new_transform)
    if tag == 'polygon':
        kwargs['id'] = name
    shape_groups.append(ShapeGroup(**kwargs))
    return shapes, shape_groups


def parse_group(node, transform, fill_color, shapes, shape_groups, defs):
    if 'transform' in node.attrib:
        transform = transform @ parse_transform(node.attrib['transform'])
    if 'fill' in node.attrib:
        fill_color = parse_color(node.attrib['fill'], defs)
    for child in node:
        tag = remove_namespaces(child.tag)
        if is_shape(tag):
            shapes, shape_groups = parse_shape(child, transform, fill_color, shapes, shape_groups, defs)
        elif tag == 'g':
            shapes, shape_groups = parse_group(child, transform, fill_color, shapes, shape_groups, defs)
    return shapes, shape_groups


def _leading_int(s):
    return int(float(''.join(ch for ch in s if not ch.isalpha())))


def parse_scene(node):
    canvas_width = canvas_height = -1
    defs = {}
    shapes, shape_groups = [], []
    fill_color = torch.tensor([0.0, 0.0, 0.0, 1.0])
    transform = torch.eye(3)
    if 'viewBox' in node.attrib:
        box = node.attrib['viewBox'].split()
        canvas_width, canvas_height = _leading_int(box[2]), _leading_int(box[3])
    else:
        if 'width' in node.attrib:
            canvas_width = _leading_int(node.attrib['width'])
        else:
            print('Warning: Can\'t find canvas width.')
        if 'height' in node.attrib:
            canvas_height = _leading_int(node.attrib['height'])
        else:
            print('Warning: Can\'t find canvas height.')
    for child in node:
        tag = remove_namespaces(child.tag)
        if tag == 'defs':
            defs = parse_defs(child, transform, defs)
        elif tag == 'style':
            defs = parse_stylesheet(child, transform, defs)
        elif tag == 'linearGradient' and 'id' in child.attrib:
            defs[child.attrib['id']] = parse_linear_gradient(child, transform, defs)
        elif tag == 'radialGradient' and 'id' in child.attrib:
            defs[child.attrib['id']] = parse_radial_gradient(child, transform, defs)
        elif is_shape(tag):
            shapes, shape_groups = parse_shape(child, transform, fill_color, shapes, shape_groups, defs)
        elif tag == 'g':
            shapes, shape_groups = parse_group(child, transform, fill_color, shapes, shape_groups, defs)
    return canvas_width, canvas_height, shapes, shape_groups


def svg_to_scene(filename):
    """Load an SVG file and convert it to PyTorch tensors (reference parse_svg.py:574-586)."""
    root = etree.parse(filename).getroot()
    cwd = os.getcwd()
    if os.path.dirname(filename) != '':
        os.chdir(os.path.dirname(filename))
    try:
        return parse_scene(root)
    finally:
        os.chdir(cwd)
