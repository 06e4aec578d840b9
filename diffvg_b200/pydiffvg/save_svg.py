"""`save_svg` (reference pydiffvg/save_svg.py:13-156): scene holders -> SVG 1.1 text.

One element per shape group, built from the group's FIRST shape (the reference does the same:
save_svg.py:78), stroke-width written back as a diameter (2 x the stroke radius), constant colours as
`rgb(r, g, b)` + opacity, linear gradients as <linearGradient> defs, optional display-gamma filter."""
import xml.etree.ElementTree as etree
from xml.dom import minidom

from .color import LinearGradient
from .shape import Circle, Ellipse, Path, Polygon, Rect

__all__ = ['save_svg', 'prettify']


def prettify(elem):
    return minidom.parseString(etree.tostring(elem, 'utf-8')).toprettyxml(indent='  ')


def _rgb(c):
    return 'rgb({}, {}, {})'.format(int(255 * c[0]), int(255 * c[1]), int(255 * c[2]))


def _path_data(shape):
    ncp = shape.num_control_points.data.cpu().numpy()
    pts = shape.points.data.cpu().numpy()
    n = pts.shape[0]
    d = 'M {} {}'.format(pts[0, 0], pts[0, 1])
    k = 1
    for c in ncp:
        if c == 0:
            p = k % n
            d += ' L {} {}'.format(pts[p, 0], pts[p, 1])
            k += 1
        elif c == 1:
            p = (k + 1) % n
            d += ' Q {} {} {} {}'.format(pts[k, 0], pts[k, 1], pts[p, 0], pts[p, 1])
            k += 2
        elif c == 2:
            p = (k + 2) % n
            d += ' C {} {} {} {} {} {}'.format(pts[k, 0], pts[k, 1], pts[k + 1, 0], pts[k + 1, 1], pts[p, 0], pts[p, 1])
            k += 3
    return d


def save_svg(filename, width, height, shapes, shape_groups, use_gamma=False):
    root = etree.Element('svg')
    for k, v in (('version', '1.1'), ('xmlns', 'http://www.w3.org/2000/svg'), ('width', str(width)), ('height', str(height))):
        root.set(k, v)
    defs = etree.SubElement(root, 'defs')
    g = etree.SubElement(root, 'g')
    if use_gamma:
        f = etree.SubElement(defs, 'filter')
        for k, v in (('id', 'gamma'), ('x', '0'), ('y', '0'), ('width', '100%'), ('height', '100%')):
            f.set(k, v)
        transfer = etree.SubElement(f, 'feComponentTransfer')
        transfer.set('color-interpolation-filters', 'sRGB')
        for ch in 'RGBA':
            func = etree.SubElement(transfer, 'feFunc' + ch)
            func.set('type', 'gamma')
            func.set('amplitude', str(1))
            func.set('exponent', str(1 / 2.2))
        g.set('style', 'filter:url(#gamma)')

    for i, group in enumerate(shape_groups):
        for color, name in ((group.fill_color, 'shape_{}_fill'.format(i)), (group.stroke_color, 'shape_{}_stroke'.format(i))):
            if isinstance(color, LinearGradient):
                node = etree.SubElement(defs, 'linearGradient')
                node.set('id', name)
                node.set('x1', str(color.begin[0].item()))
                node.set('y1', str(color.begin[1].item()))
                node.set('x2', str(color.end[0].item()))
                node.set('y2', str(color.end[1].item()))
                offsets = color.offsets.data.cpu().numpy()
                for j in range(offsets.shape[0]):
                    stop = etree.SubElement(node, 'stop')
                    c = color.stop_colors[j, :]
                    stop.set('offset', str(offsets[j]))
                    stop.set('stop-color', _rgb(c))
                    stop.set('stop-opacity', '{}'.format(c[3]))

    for i, group in enumerate(shape_groups):
        shape = shapes[group.shape_ids[0]]
        if isinstance(shape, Circle):
            node = etree.SubElement(g, 'circle')
            node.set('r', str(shape.radius.item()))
            node.set('cx', str(shape.center[0].item()))
            node.set('cy', str(shape.center[1].item()))
        elif isinstance(shape, Polygon):
            node = etree.SubElement(g, 'polygon')
            pts = shape.points.data.cpu().numpy()
            node.set('points', ' '.join('{} {}'.format(pts[j, 0], pts[j, 1]) for j in range(pts.shape[0])))
        elif isinstance(shape, Path):
            node = etree.SubElement(g, 'path')
            node.set('d', _path_data(shape))
        elif isinstance(shape, Rect):
            node = etree.SubElement(g, 'rect')
            node.set('x', str(shape.p_min[0].item()))
            node.set('y', str(shape.p_min[1].item()))
            node.set('width', str(shape.p_max[0].item() - shape.p_min[0].item()))
            node.set('height', str(shape.p_max[1].item() - shape.p_min[1].item()))
        elif isinstance(shape, Ellipse):
            node = etree.SubElement(g, 'ellipse')
            node.set('cx', str(shape.center[0].item()))
            node.set('cy', str(shape.center[1].item()))
            node.set('rx', str(shape.radius[0].item()))
            node.set('ry', str(shape.radius[1].item()))
        else:
            raise TypeError('unsupported shape %r' % (shape,))
        node.set('stroke-width', str(2 * shape.stroke_width.data.cpu().item()))
        if group.fill_color is None:
            node.set('fill', 'none')
        elif isinstance(group.fill_color, LinearGradient):
            node.set('fill', 'url(#shape_{}_fill)'.format(i))
        else:
            c = group.fill_color.data.cpu().numpy()
            node.set('fill', _rgb(c))
            node.set('opacity', str(c[3]))
        if group.stroke_color is not None:
            if isinstance(group.stroke_color, LinearGradient):
                node.set('stroke', 'url(#shape_{}_stroke)'.format(i))
            else:
                c = group.stroke_color.data.cpu().numpy()
                node.set('stroke', _rgb(c))
                node.set('stroke-opacity', str(c[3]))
            node.set('stroke-linecap', 'round')
            node.set('stroke-linejoin', 'round')
    with open(filename, 'w') as f:
        f.write(prettify(root))
