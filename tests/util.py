"""Shared helpers for the parity tests."""
import numpy as np
import torch

from diffvg_b200 import scene_pack


def pack(scene, filter_type=0, filter_radius=0.5):
    cw, ch, shapes, groups = scene
    return scene_pack.pack_scene_numpy(cw, ch, shapes, groups, filter_type, torch.tensor(filter_radius))


def rel_l2(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(a), 1e-30))


def image_report(ref, got, spp):
    """max-abs error, number of pixels off by more than 1e-5, and how many of those look like
    whole-sample flips (error close to a multiple of 1/spp in some channel)."""
    d = np.abs(ref.astype(np.float64) - got.astype(np.float64))
    bad = d.max(axis=2) > 1e-5
    return dict(max_abs=float(d.max()), bad_pixels=int(bad.sum()))


def gpu_render(topo, params, width, height, nsx, nsy, seed, background=None, d_render_image=None,
               skip_xform_grad=False, use_prefiltering=False, want_image=True, want_sdf=False, eval_positions=None,
               d_render_sdf=None, want_d_translation=False, extra_flags=0):
    """Drive the product C ABI directly (ctypes) with device buffers managed through torch.  Same
    keyword surface as oracle/ref_oracle.render."""
    import ctypes
    from diffvg_b200 import _native as n
    dev = torch.device('cuda', 0)
    h = ctypes.c_void_p()
    topo = np.ascontiguousarray(topo, np.int32)
    n.check(n.lib.dvg_scene_create(topo.ctypes.data, topo.shape[0], 0, ctypes.byref(h)))
    try:
        stream = torch.cuda.current_stream().cuda_stream
        p = np.ascontiguousarray(params, np.float32)
        n.check(n.lib.dvg_scene_set_params(h, p.ctypes.data, p.shape[0], 0, stream))
        bg = torch.from_numpy(background).to(dev).contiguous() if background is not None else None
        ep = torch.from_numpy(np.ascontiguousarray(eval_positions, np.float32)).to(dev) if eval_positions is not None else None
        n_eval = 0 if ep is None else ep.shape[0]
        ptr = lambda t: t.data_ptr() if t is not None else None
        pf = 1 if use_prefiltering else 0
        out = {}
        if d_render_image is None and d_render_sdf is None:
            img = torch.empty(height, width, 4, device=dev) if want_image else None
            sdf = None
            if want_sdf:
                sdf = torch.empty((n_eval, 1) if n_eval else (height, width, 1), device=dev)
            n.check(n.lib.dvg_render_forward(h, ptr(bg), ptr(img), ptr(sdf), width, height, nsx, nsy, int(seed), pf,
                                             ptr(ep), n_eval, stream))
            out['image'] = img.cpu().numpy() if img is not None else None
            out['sdf'] = sdf.cpu().numpy() if sdf is not None else None
        else:
            dimg = torch.from_numpy(np.ascontiguousarray(d_render_image, np.float32)).to(dev) if d_render_image is not None else None
            dsdf = torch.from_numpy(np.ascontiguousarray(d_render_sdf, np.float32)).to(dev) if d_render_sdf is not None else None
            dpar = torch.empty(p.shape[0], device=dev)
            dbg = torch.empty_like(bg) if bg is not None else None
            dtr = torch.empty(height, width, 2, device=dev) if want_d_translation else None
            n.check(n.lib.dvg_render_backward(h, ptr(bg), ptr(dimg), ptr(dsdf), width, height, nsx, nsy, int(seed), pf,
                                              ptr(ep), n_eval, dpar.data_ptr(), ptr(dbg), ptr(dtr),
                                              (1 if skip_xform_grad else 0) | extra_flags, stream))
            out['d_params'] = dpar.cpu().numpy()
            out['d_background'] = dbg.cpu().numpy() if dbg is not None else None
            out['d_translation'] = dtr.cpu().numpy() if dtr is not None else None
        torch.cuda.synchronize()
        return out
    finally:
        n.lib.dvg_scene_destroy(h)


def gpu_render_rows(topo, params, width, height, nsx, nsy, seed, row_ranges, d_render_image=None, use_prefiltering=False):
    """Render the image (and gradient) as independent row shards through the `_rows` entry points,
    the way one-process-per-GPU sharding does, and combine: image rows are disjoint, gradients add."""
    import ctypes
    from diffvg_b200 import _native as n
    dev = torch.device('cuda', 0)
    h = ctypes.c_void_p()
    topo = np.ascontiguousarray(topo, np.int32)
    n.check(n.lib.dvg_scene_create(topo.ctypes.data, topo.shape[0], 0, ctypes.byref(h)))
    try:
        stream = torch.cuda.current_stream().cuda_stream
        p = np.ascontiguousarray(params, np.float32)
        n.check(n.lib.dvg_scene_set_params(h, p.ctypes.data, p.shape[0], 0, stream))
        img = torch.full((height, width, 4), float('nan'), device=dev)
        for (r0, r1) in row_ranges:
            n.check(n.lib.dvg_render_forward_rows(h, None, img.data_ptr(), width, height, nsx, nsy, int(seed), 1 if use_prefiltering else 0, r0, r1, stream))
        out = {'image': img.cpu().numpy()}
        if d_render_image is not None:
            dimg = torch.from_numpy(np.ascontiguousarray(d_render_image, np.float32)).to(dev)
            total = np.zeros(p.shape[0], np.float64)
            for (r0, r1) in row_ranges:
                dpar = torch.empty(p.shape[0], device=dev)
                n.check(n.lib.dvg_render_backward_rows(h, None, dimg.data_ptr(), width, height, nsx, nsy, int(seed),
                                                       1 if use_prefiltering else 0, r0, r1, dpar.data_ptr(), None, 0, stream))
                total += dpar.cpu().numpy().astype(np.float64)
            out['d_params'] = total
        torch.cuda.synchronize()
        return out
    finally:
        n.lib.dvg_scene_destroy(h)


def gpu_scene_dump(topo, params, selectors):
    """dvg_scene_dump of the product library for a list of (what, index) selectors -> list of uint32 arrays.
    Everything comes back from DEVICE memory (the tables the kernels read, the trees of dvg_bvh.cu)."""
    import ctypes
    from diffvg_b200 import _native as n
    h = ctypes.c_void_p()
    topo = np.ascontiguousarray(topo, np.int32)
    n.check(n.lib.dvg_scene_create(topo.ctypes.data, topo.shape[0], 0, ctypes.byref(h)))
    try:
        stream = torch.cuda.current_stream().cuda_stream
        p = np.ascontiguousarray(params, np.float32)
        n.check(n.lib.dvg_scene_set_params(h, p.ctypes.data, p.shape[0], 0, stream))
        out = []
        cap = 1 << 22
        buf = np.zeros(cap, np.uint32)
        for what, index in selectors:
            cnt = n.lib.dvg_scene_dump(h, what, index, buf.ctypes.data, cap, stream)
            if cnt < 0:
                raise RuntimeError(n.lib.dvg_last_error().decode())
            out.append(buf[:cnt].copy())
        return out
    finally:
        n.lib.dvg_scene_destroy(h)


def gpu_render_batch(topo, params_rows, width, height, nsx, nsy, seeds, backgrounds=None, d_render_images=None,
                     skip_xform_grad=False):
    """`batch` scenes of one topology through dvg_scene_create_batch / dvg_render_*_batch.  params_rows [B, N],
    seeds [B]; images [B, H, W, 4]."""
    import ctypes
    from diffvg_b200 import _native as n
    dev = torch.device('cuda', 0)
    h = ctypes.c_void_p()
    topo = np.ascontiguousarray(topo, np.int32)
    p = np.ascontiguousarray(params_rows, np.float32)
    B = p.shape[0]
    n.check(n.lib.dvg_scene_create_batch(topo.ctypes.data, topo.shape[0], 0, B, ctypes.byref(h)))
    try:
        stream = torch.cuda.current_stream().cuda_stream
        n.check(n.lib.dvg_scene_set_params(h, p.ctypes.data, p.size, 0, stream))
        sd = np.ascontiguousarray(np.asarray(seeds, np.uint64))
        bg = torch.from_numpy(np.ascontiguousarray(backgrounds, np.float32)).to(dev) if backgrounds is not None else None
        ptr = lambda t: t.data_ptr() if t is not None else None
        out = {}
        if d_render_images is None:
            img = torch.empty(B, height, width, 4, device=dev)
            n.check(n.lib.dvg_render_forward_batch(h, ptr(bg), img.data_ptr(), width, height, nsx, nsy, sd.ctypes.data, stream))
            out['image'] = img.cpu().numpy()
        else:
            dimg = torch.from_numpy(np.ascontiguousarray(d_render_images, np.float32)).to(dev)
            dpar = torch.empty(B, p.shape[1], device=dev)
            dbg = torch.empty_like(bg) if bg is not None else None
            n.check(n.lib.dvg_render_backward_batch(h, ptr(bg), dimg.data_ptr(), width, height, nsx, nsy, sd.ctypes.data,
                                                    dpar.data_ptr(), ptr(dbg), 1 if skip_xform_grad else 0, stream))
            out['d_params'] = dpar.cpu().numpy()
            out['d_background'] = dbg.cpu().numpy() if dbg is not None else None
        torch.cuda.synchronize()
        return out
    finally:
        n.lib.dvg_scene_destroy(h)
