"""One-process-per-GPU sharding of the render path.

The path shards with no data-path exchange: every pixel sample and every boundary sample is
independent given the replicated scene, and its RNG stream is a function of the GLOBAL sample
index only (pcg.h:32-40), so any partition of the index space reproduces the single-GPU sample
set.  Two partitions are offered:

  * one large render (C3 / C4): rank r owns a contiguous band of pixel rows, aligned to the
    tile height, and the boundary-sample indices of the same rows (`row_partition`,
    `ShardedRenderFunction`).  Forward: the bands are assembled with one all-gather (a band is
    rows*W*4 floats).  Backward: every rank holds the full d_image (the loss is computed on the
    assembled image), runs `dvg_render_backward_rows` on its band, and the per-rank gradient
    buffers `float[num_params]` are summed with ONE all-reduce.
  * batched scenes (C5): scenes are dealt to ranks by index (`batch_partition`); per-scene
    gradients need no exchange, shared upstream parameters are all-reduced by the caller (DDP).

`torch.distributed` is plumbing only (NCCL on the GPUs, gloo in the CPU tests); the kernels are
in libdiffvg_b200.so.  The collectives operate on whatever device the tensors live on, which
is what lets tests/test_sharding_cpu.py exercise this file with world_size 2 on gloo.
"""
import torch
import torch.distributed as dist


def row_partition(height, world, align=1):
    """[(row_begin, row_end)] * world: contiguous bands covering [0, height), every interior
    boundary a multiple of `align` (the tile height, so that no tile is split between ranks).
    Ranks may receive empty bands when height / align < world."""
    assert height >= 0 and world >= 1 and align >= 1
    units = (height + align - 1) // align
    out = []
    for r in range(world):
        b = min(height, ((units * r) // world) * align)
        e = min(height, ((units * (r + 1)) // world) * align)
        out.append((b, e))
    return out


def balanced_row_partition(costs, height, world, align, per_unit=0.0):
    """Bands of equal COST instead of equal height: `costs[k]` = cost of pixel rows [k * align, (k + 1) * align) (the tile
    rows of dvg_scene_row_costs), `per_unit` = cost every tile row carries whatever its content (sample generation, splat,
    the boundary samples of its pixels).  Cuts lie on multiples of `align`; every rank gets at least one unit while there
    are enough of them.  Deterministic, so ranks that see the same costs agree on the bands."""
    import numpy as np
    c = np.asarray(costs, np.float64) + float(per_unit)
    units = c.shape[0]
    assert units == (height + align - 1) // align and world >= 1
    if units <= world:
        return row_partition(height, world, align)
    cum = np.concatenate([[0.0], np.cumsum(c)])
    total = cum[-1]
    cuts = [0]
    for r in range(1, world):
        k = int(np.searchsorted(cum, total * r / world, side='left'))
        if k > 0 and abs(cum[k - 1] - total * r / world) <= abs(cum[min(k, units)] - total * r / world):
            k -= 1
        k = max(k, cuts[-1] + 1)                 # at least one unit per rank ...
        k = min(k, units - (world - r))          # ... and enough left for the ranks after this one
        cuts.append(k)
    cuts.append(units)
    return [(min(height, cuts[r] * align), min(height, cuts[r + 1] * align)) for r in range(world)]


def rebalance_bands(bands, times, height, align, damping=1.0):
    """Bands re-cut from MEASURED per-rank times: rank r took `times[r]` for rows `bands[r]`, so its rows cost
    times[r] / rows each; the new cuts give every rank the same share of that piecewise-constant cost profile (cuts on
    multiples of `align`, at least one unit per rank).  Two or three rounds during warm-up settle within a few per cent
    of even: the bin-count estimate of `balanced_bands` cannot see how expensive a tile's candidates are.  `damping` < 1
    moves only part of the way (noisy timings).  Deterministic in its inputs: gather the times, then call it on every rank."""
    import numpy as np
    world = len(bands)
    units = (height + align - 1) // align
    if units <= world:
        return list(bands)
    dens = np.zeros(units, np.float64)
    for (b, e), t in zip(bands, times):
        u0, u1 = b // align, (e + align - 1) // align
        if u1 > u0:
            dens[u0:u1] = max(float(t), 1e-9) / (u1 - u0)
    new = balanced_row_partition(dens, height, world, align)
    if damping < 1.0:
        cuts = [int(round(((1 - damping) * ob + damping * nb) / align)) * align for (ob, _), (nb, _) in zip(bands, new)]
        cuts = cuts + [height]
        for r in range(1, world):                     # keep the cuts increasing by at least one unit
            cuts[r] = max(cuts[r], cuts[r - 1] + align)
        for r in range(world - 1, 0, -1):
            cuts[r] = min(cuts[r], cuts[r + 1] - align if r + 1 < world else ((height - 1) // align) * align)
        new = [(cuts[r], cuts[r + 1]) for r in range(world)]
    return new


def balanced_bands(packed, params, width, height, num_samples_x, num_samples_y, world, per_tile=4.0, group=None):
    """Row bands of equal estimated cost for the scene's CURRENT parameters (one whole-image binning + a read-back:
    call it every few dozen iterations, not every step -- the content of an optimisation moves slowly -- and hand the
    result to `ShardedRenderFunction.apply(..., bands=...)`).  `per_tile`: fixed cost of a tile in candidate units.
    Every rank must call it (rank 0's answer is broadcast so that all ranks cut at the same rows)."""
    import ctypes
    import numpy as np
    from .pydiffvg import render_pytorch as rp
    n = rp._native()
    dev = rp._cuda_device()
    ns = rp._get_native_scene(packed, dev.index)
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream().cuda_stream
        ns.set_params(params, stream)
        cap = height + 1
        costs = np.zeros(cap, np.float32)
        th = ctypes.c_int()
        n.check(n.lib.dvg_scene_row_costs(ns.handle, width, height, num_samples_x, num_samples_y,
                                          1 if packed.use_prefiltering else 0, costs.ctypes.data, cap, ctypes.byref(th), stream))
    align = th.value
    units = (height + align - 1) // align
    bands = balanced_row_partition(costs[:units], height, world, align, per_unit=per_tile * ((width + 7) // 8))
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        t = torch.tensor(bands, dtype=torch.int64, device=dev if dist.get_backend(group) == 'nccl' else 'cpu')
        dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        bands = [tuple(int(v) for v in row) for row in t.cpu().tolist()]
    return bands


def stripe_partition(height, world, stripe=16):
    """Round-robin stripes of `stripe` rows: [[(b, e), ...]] * world.  Balances non-uniform content
    (SURVEY 8e 'efficiency risks'); each stripe is one *_rows call."""
    out = [[] for _ in range(world)]
    for k, b in enumerate(range(0, height, stripe)):
        out[k % world].append((b, min(height, b + stripe)))
    return out


def batch_partition(num_scenes, world):
    """[range] * world: scene b goes to rank b % world (seeds differ per scene, so the deal is
    cost-balanced in expectation)."""
    return [range(r, num_scenes, world) for r in range(world)]


def sample_range(rows, width, spp):
    """Global sample-index range [begin, end) of a row band: idx = ((y*W + x)*nsy + sy)*nsx + sx
    (diffvg.cpp:1168-1176) is row-major in y, so a band of rows is one contiguous range; the same
    range of boundary-sample indices goes with it (diffvg.cpp:1332-1336)."""
    b, e = rows
    return b * width * spp, e * width * spp


def allgather_rows(band, bands, group=None):
    """Assemble the full image from per-rank row bands (band: [rows_r, W, C])."""
    world = dist.get_world_size(group)
    if world == 1:
        return band
    assert len(bands) == world
    w, c = band.shape[1], band.shape[2]
    rmax = max(e - b for b, e in bands)   # bands differ by at most one alignment unit: pad to equal size
    mine = band.contiguous()
    if mine.shape[0] < rmax:
        mine = torch.cat([mine, mine.new_zeros(rmax - mine.shape[0], w, c)], dim=0)
    buf = torch.empty(world, rmax, w, c, dtype=band.dtype, device=band.device)
    dist.all_gather_into_tensor(buf, mine, group=group) if band.is_cuda else \
        dist.all_gather(list(buf.unbind(0)), mine, group=group)
    return torch.cat([buf[r, :e - b] for r, (b, e) in enumerate(bands)], dim=0)


# Own render time of a rank (CUDA events around its dvg_render_*_rows calls): what `rebalance_bands` needs.  The step time
# of a rank is useless for that -- the collectives make every rank wait for the slowest one.
_TIMING = {'on': False, 'events': []}


def time_compute(on=True):
    """Start / stop recording the device time of this rank's own render calls inside `ShardedRenderFunction`."""
    _TIMING['on'] = bool(on)
    _TIMING['events'] = []


def compute_ms():
    """Device milliseconds this rank spent in its own render calls since `time_compute(True)` (synchronises)."""
    torch.cuda.synchronize()
    total = sum(a.elapsed_time(b) for a, b in _TIMING['events'])
    _TIMING['events'] = []
    return total


class _timed:
    def __enter__(self):
        if _TIMING['on']:
            self.a, self.b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            self.a.record()
        return self

    def __exit__(self, *exc):
        if _TIMING['on']:
            self.b.record()
            _TIMING['events'].append((self.a, self.b))
        return False


def band_gradient_mode(packed, bands):
    """What a rank needs of the other ranks' `d_image` rows in the backward pass of a band loss (`gather=False`):
    'own'    -- nothing: prefiltered path without the gradient of the pixel-filter radius (a sample reads d_image at its
                own pixel only);
    'halo'   -- `halo_rows` = ceil(filter radius) rows either side of the band: prefiltered path with `d_filter.radius`,
                which reads d_image over the whole (2 ceil(r) + 1)^2 footprint of a sample (diffvg.cpp:1250-1268);
    'gather' -- every row: the boundary samples of the sampled path land anywhere in the image."""
    if not packed.use_prefiltering:
        return 'gather'
    if not getattr(packed, 'needs_filter_grad', True) and float(getattr(packed, 'filter_radius', 0.5)) <= 0.5:
        return 'own'
    return 'halo' if min(e - b for b, e in bands) >= int(getattr(packed, 'halo_rows', 1)) else 'gather'


def band_gradient_image(grad_band, bands, rank, height, width, mode, halo=1, group=None):
    """Full-size `d_image [height, width, 4]` for the backward call of rank `rank` from its band's gradient
    `[rows, width, 4]` (see `band_gradient_mode`).  Rows the mode does not fill are left uninitialised: the backward
    pass does not read them."""
    world = len(bands)
    rb, re = bands[rank]
    if mode == 'gather':
        return allgather_rows(grad_band, bands, group)
    full = torch.empty(height, width, 4, device=grad_band.device, dtype=grad_band.dtype)
    full[rb:re] = grad_band
    if mode == 'own':
        return full
    hl = halo
    edge = torch.cat([grad_band[:hl], grad_band[-hl:]], dim=0).contiguous()
    buf = torch.empty((world,) + tuple(edge.shape), dtype=edge.dtype, device=edge.device)
    dist.all_gather_into_tensor(buf, edge, group=group) if edge.is_cuda else \
        dist.all_gather(list(buf.unbind(0)), edge, group=group)
    full[max(0, rb - hl):rb].zero_()            # halo rows: the neighbours' rows below, zero at the image edge
    full[re:min(height, re + hl)].zero_()
    if rank > 0:
        full[rb - hl:rb] = buf[rank - 1, hl:]
    if rank < world - 1:
        full[re:re + hl] = buf[rank + 1, :hl]
    return full


def allreduce_gradients(d_params, group=None):
    """Sum the per-rank gradient buffers in place (the only collective on the backward path:
    num_params floats, 155 KB at the painterly config -- latency-bound on NVSwitch)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(d_params, op=dist.ReduceOp.SUM, group=group)
    return d_params


class ShardedRenderFunction(torch.autograd.Function):
    """`RenderFunction.apply` for one large render split by pixel rows over the ranks of `group`.
    Every rank calls it with the same scene / seed and gets the full image; gradients w.r.t. the packed
    parameters are complete (summed over ranks) on every rank.  Colour output only.

    `gather=False` returns only the rank's own band `[rows, W, 4]` (rows = `row_partition(...)[rank]`) for a loss
    that is itself computed per band: no image exchange in the forward pass.  The backward pass then needs the
    other ranks' d_image rows only for the boundary samples of the sampled path (they land anywhere in the image):
    with `use_prefiltering` only `halo_rows` rows either side of the band are exchanged, and only when the gradient of
    the pixel-filter radius is asked for (`packed.needs_filter_grad`); otherwise the d_image bands are all-gathered."""

    @staticmethod
    def forward(ctx, width, height, num_samples_x, num_samples_y, seed, background_image, packed, params, group=None,
                gather=True, bands=None):
        from .pydiffvg import render_pytorch as rp
        n = rp._native()
        dev = rp._cuda_device()
        assert packed.output_type == rp.OutputType.color, 'the SDF output is not row-sharded'
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        ns = rp._get_native_scene(packed, dev.index)
        tile_h = tile_height(num_samples_x * num_samples_y)
        if bands is None:
            bands = row_partition(height, world, tile_h)
        else:   # e.g. balanced_bands(...): cuts must lie on tile rows
            bands = [(int(b), int(e)) for b, e in bands]
            assert len(bands) == world and bands[0][0] == 0 and bands[-1][1] == height
            assert all(b % tile_h == 0 for b, _ in bands) and all(bands[k][1] == bands[k + 1][0] for k in range(world - 1))
        rb, re = bands[rank]
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream().cuda_stream
            version = ns.set_params(params, stream)
            if background_image is not None:
                background_image = background_image.to(dev).contiguous().float()
                assert background_image.shape == (height, width, 4)
            wide = float(getattr(packed, 'filter_radius', 0.5)) > 0.5
            # the call zeroes and fills the rows of its band; only the all-reduce of wide filters reads the other rows
            full = (torch.zeros if (wide and world > 1) else torch.empty)(height, width, 4, device=dev, dtype=torch.float32)
            with _timed():
                n.check(n.lib.dvg_render_forward_rows(
                    ns.handle, background_image.data_ptr() if background_image is not None else None, full.data_ptr(),
                    width, height, num_samples_x, num_samples_y, int(seed), 1 if packed.use_prefiltering else 0,
                    rb, re, stream))
            if not gather:
                assert not wide or world == 1, 'gather=False needs a pixel filter of radius <= 0.5 (samples splat across band edges)'
                img = full[rb:re]
            elif world > 1 and wide:
                # a sample splats onto pixels up to ceil(radius) rows outside its band: every rank's buffer holds the
                # contributions of ITS samples to the whole image, the image is their sum
                dist.all_reduce(full, op=dist.ReduceOp.SUM, group=group)
                img = full
            else:
                img = allgather_rows(full[rb:re], bands, group) if world > 1 else full
        ctx.native_scene, ctx.scene_version, ctx.packed = ns, version, packed
        ctx.background_image = background_image
        ctx.geom = (width, height, num_samples_x, num_samples_y, seed, rb, re)
        ctx.gather, ctx.bands, ctx.world = gather, bands, world
        ctx.device, ctx.params_device, ctx.group = dev, params.device, group
        ctx.save_for_backward(params)
        return img

    @staticmethod
    def backward(ctx, grad_img):
        from .pydiffvg import render_pytorch as rp
        n = rp._native()
        dev, ns = ctx.device, ctx.native_scene
        (params,) = ctx.saved_tensors
        width, height, nsx, nsy, seed, rb, re = ctx.geom
        bg = ctx.background_image
        grad_img = grad_img.to(dev).float().contiguous()
        with torch.cuda.device(dev):
            if not ctx.gather and ctx.world > 1:   # band-shaped gradient -> full-size d_image
                grad_img = band_gradient_image(grad_img, ctx.bands, dist.get_rank(ctx.group), height, width,
                                               band_gradient_mode(ctx.packed, ctx.bands), int(getattr(ctx.packed, 'halo_rows', 1)), ctx.group)
            stream = torch.cuda.current_stream().cuda_stream
            if ns.version != ctx.scene_version:
                ctx.scene_version = ns.set_params(params, stream)
            d_params = torch.empty(ctx.packed.num_params, device=dev, dtype=torch.float32)
            d_bg = torch.zeros_like(bg) if bg is not None else None
            with _timed():
                n.check(n.lib.dvg_render_backward_rows(
                    ns.handle, bg.data_ptr() if bg is not None else None, grad_img.data_ptr(),
                    width, height, nsx, nsy, int(seed), 1 if ctx.packed.use_prefiltering else 0, rb, re,
                    d_params.data_ptr(), d_bg.data_ptr() if d_bg is not None else None,
                    rp.backward_flags(ctx.packed), stream))
            allreduce_gradients(d_params, ctx.group)
            if d_bg is not None:
                allreduce_gradients(d_bg, ctx.group)
        if d_params.device != ctx.params_device:
            d_params = d_params.to(ctx.params_device)
        return None, None, None, None, None, d_bg, None, d_params, None, None, None


def tile_height(spp):
    """Tile height the library bins with for `spp` samples per pixel (csrc/dvg_capi.cu choose_tile):
    row bands are aligned to it so that a tile never straddles two ranks."""
    if spp >= 16:
        return 2
    if spp >= 2:
        return 8
    return 16
