/*
 * dvg_scene_format.h -- the packed scene exchanged across the C ABI.
 *
 * A scene is two flat host (or device) arrays:
 *
 *   topo   : int32[]  -- structure (counts, types, flags, offsets).  Built once per
 *                        scene topology by the host language (Python here).
 *   params : float32[] -- every continuous parameter of the scene (points, thickness,
 *                        radii, stroke widths, colours, gradient stops, transforms,
 *                        filter radius).  `d_params` returned by the backward pass has
 *                        exactly the same layout, so the host side can hand it back to
 *                        autograd as ONE tensor and NCCL can all-reduce it in one call.
 *
 * This replaces the per-object pybind11 surface of the reference
 * (diffvg.cpp:1651-1792: Circle/Ellipse/Path/Rect/Shape/ShapeGroup/Constant/
 * LinearGradient/RadialGradient/Filter objects holding raw pointers into torch
 * tensors, shape.h:9-169, color.h:7-63) with plain arrays.
 *
 * topo layout (all int32):
 *   header  [DVG_TOPO_HEADER_LEN]
 *   shapes  [num_shapes][DVG_SHAPE_REC_LEN]      at header[DVG_H_OFF_SHAPES]
 *   groups  [num_groups][DVG_GROUP_REC_LEN]      at header[DVG_H_OFF_GROUPS]
 *   ncp     [total_segments]                     at header[DVG_H_OFF_NCP]
 *           (num_control_points per path segment: 0 line, 1 quadratic, 2 cubic)
 *   gshapes [total_group_shapes]                 at header[DVG_H_OFF_GSHAPES]
 *           (concatenated ShapeGroup.shape_ids)
 */
#ifndef DVG_SCENE_FORMAT_H
#define DVG_SCENE_FORMAT_H

#include <stdint.h>

#define DVG_TOPO_MAGIC 0x44564731 /* 'DVG1' */

enum {
    DVG_H_MAGIC = 0,
    DVG_H_CANVAS_W = 1,
    DVG_H_CANVAS_H = 2,
    DVG_H_NUM_SHAPES = 3,
    DVG_H_NUM_GROUPS = 4,
    DVG_H_FILTER_TYPE = 5,       /* DvgFilterType (filter.h:6-11) */
    DVG_H_FILTER_RADIUS_OFF = 6, /* params offset of the filter radius scalar */
    DVG_H_NUM_PARAMS = 7,        /* length of params */
    DVG_H_TOTAL_SEGS = 8,        /* sum of num_base_points over path shapes */
    DVG_H_TOTAL_GSHAPES = 9,     /* sum of num_shapes over groups (num_total_shapes, scene.cpp:935-939) */
    DVG_H_OFF_SHAPES = 10,
    DVG_H_OFF_GROUPS = 11,
    DVG_H_OFF_NCP = 12,
    DVG_H_OFF_GSHAPES = 13,
    DVG_H_TOTAL_POINTS = 14,     /* sum of num_points over path shapes */
    DVG_H_RESERVED = 15,
    DVG_TOPO_HEADER_LEN = 16
};

/* shape.h:9-14 ShapeType, same numeric order */
enum DvgShapeType { DVG_SHAPE_CIRCLE = 0, DVG_SHAPE_ELLIPSE = 1, DVG_SHAPE_PATH = 2, DVG_SHAPE_RECT = 3 };
/* color.h:7-11 ColorType; -1 = colour absent (None in pydiffvg) */
enum DvgColorType { DVG_COLOR_NONE = -1, DVG_COLOR_CONSTANT = 0, DVG_COLOR_LINEAR = 1, DVG_COLOR_RADIAL = 2 };
/* filter.h:6-11 FilterType */
enum DvgFilterType { DVG_FILTER_BOX = 0, DVG_FILTER_TENT = 1, DVG_FILTER_PARABOLIC = 2, DVG_FILTER_HANN = 3 };

enum {
    DVG_S_TYPE = 0,
    /* params offset of the shape's own floats:
     *   circle : radius, center.x, center.y                (3)
     *   ellipse: radius.x, radius.y, center.x, center.y    (4)
     *   rect   : p_min.x, p_min.y, p_max.x, p_max.y        (4)
     *   path   : points[2*num_points]                                         */
    DVG_S_PARAM_OFF = 1,
    DVG_S_WIDTH_OFF = 2,  /* params offset of the scalar stroke_width; -1 => 0.0
                             (paths with per-point thickness, render_pytorch.py:95-98) */
    DVG_S_THICK_OFF = 3,  /* path only: params offset of thickness[num_points], or -1 */
    DVG_S_NUM_POINTS = 4, /* path only */
    DVG_S_NUM_SEGS = 5,   /* path only: num_base_points */
    DVG_S_NCP_OFF = 6,    /* path only: index into the ncp array */
    DVG_S_FLAGS = 7,      /* bit0 is_closed, bit1 use_distance_approx */
    DVG_SHAPE_REC_LEN = 8
};
#define DVG_SF_CLOSED 1
#define DVG_SF_DISTANCE_APPROX 2

enum {
    DVG_G_SHAPES_OFF = 0, /* index into the gshapes array */
    DVG_G_NUM_SHAPES = 1,
    DVG_G_FILL_TYPE = 2,  /* DvgColorType */
    /* params offset of the colour record:
     *   constant: r,g,b,a                                                  (4)
     *   linear  : begin.xy, end.xy, offsets[n], stop_colors[4n]            (4+5n)
     *   radial  : center.xy, radius.xy, offsets[n], stop_colors[4n]        (4+5n) */
    DVG_G_FILL_OFF = 3,
    DVG_G_FILL_STOPS = 4,
    DVG_G_STROKE_TYPE = 5,
    DVG_G_STROKE_OFF = 6,
    DVG_G_STROKE_STOPS = 7,
    DVG_G_EVEN_ODD = 8,
    DVG_G_XFORM_OFF = 9,  /* params offset of shape_to_canvas, 9 floats row-major */
    DVG_G_RESERVED0 = 10,
    DVG_G_RESERVED1 = 11,
    DVG_GROUP_REC_LEN = 12
};

#endif /* DVG_SCENE_FORMAT_H */
