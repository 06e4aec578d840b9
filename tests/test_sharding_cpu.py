"""Host-side logic of the multi-GPU path (diffvg_b200/sharded.py), world_size 2 on gloo / CPU.
The kernels themselves are covered by the -m gpu tests (test_gpu_parity.py::test_row_sharding_*);
here: partitions tile the index space, band assembly and the gradient all-reduce are correct, and
the row split reproduces the whole render when applied to the ORACLE (pixel rows are independent
because the RNG stream is a function of the global sample index, pcg.h:32-40)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from diffvg_b200 import sharded


def test_row_partition_covers_and_aligns():
    for height in (1, 2, 7, 64, 510, 512, 2048):
        for world in (1, 2, 3, 4, 8):
            for align in (1, 2, 8, 16):
                bands = sharded.row_partition(height, world, align)
                assert len(bands) == world
                assert bands[0][0] == 0 and bands[-1][1] == height
                for (b0, e0), (b1, e1) in zip(bands, bands[1:]):
                    assert e0 == b1 and b0 <= e0
                for b, e in bands[:-1]:
                    assert e % align == 0 or e == height
                sizes = [e - b for b, e in bands]
                assert max(sizes) - min(sizes) <= 2 * align   # balanced to within the alignment


def test_stripe_and_batch_partitions():
    st = sharded.stripe_partition(100, 3, 16)
    rows = sorted(r for part in st for b, e in part for r in range(b, e))
    assert rows == list(range(100))
    bp = sharded.batch_partition(512, 8)
    assert sorted(i for r in bp for i in r) == list(range(512))
    assert all(len(r) == 64 for r in bp)


def test_sample_range_matches_index_layout():
    w, nsx, nsy = 13, 2, 3
    b, e = sharded.sample_range((4, 9), w, nsx * nsy)
    idx = [((y * w + x) * nsy + sy) * nsx + sx for y in range(4, 9) for x in range(w) for sy in range(nsy) for sx in range(nsx)]
    assert b == min(idx) and e == max(idx) + 1 and e - b == len(idx)


def test_tile_height_matches_library_choice():
    # csrc/dvg_capi.cu choose_tile
    assert [sharded.tile_height(s) for s in (1, 2, 4, 9, 16, 64)] == [16, 8, 8, 8, 2, 2]


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, height, width, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        full = torch.arange(height * width * 4, dtype=torch.float32).reshape(height, width, 4)
        bands = sharded.row_partition(height, world, 2)
        b, e = bands[rank]
        img = sharded.allgather_rows(full[b:e].clone(), bands)
        ok_img = bool(torch.equal(img, full))
        # gradient sum: rank r contributes (r + 1) * base; every rank must end with the same total
        base = torch.linspace(-1, 1, 1001)
        g = sharded.allreduce_gradients(base * (rank + 1))
        ok_grad = bool(torch.allclose(g, base * sum(range(1, world + 1))))
        q.put((rank, ok_img, ok_grad))
    finally:
        dist.destroy_process_group()


def test_allgather_rows_and_gradient_allreduce_world2():
    world = 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 10, 3, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(r[1] and r[2] for r in res)


def _worker_band_gradient(rank, world, port, height, width, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        full = torch.arange(height * width * 4, dtype=torch.float32).reshape(height, width, 4) + 1.0
        bands = sharded.row_partition(height, world, 2)
        b, e = bands[rank]
        band = full[b:e].clone()
        ok = {}
        g = sharded.band_gradient_image(band, bands, rank, height, width, 'gather')
        ok['gather'] = bool(torch.equal(g, full))
        g = sharded.band_gradient_image(band, bands, rank, height, width, 'own')
        ok['own'] = g.shape == full.shape and bool(torch.equal(g[b:e], band))
        for hl in (1, 2):
            g = sharded.band_gradient_image(band, bands, rank, height, width, 'halo', hl)
            lo, hi = max(0, b - hl), min(height, e + hl)
            # own rows and the neighbours' halo rows are the true d_image; nothing else is promised
            ok['halo%d' % hl] = bool(torch.equal(g[lo:hi], full[lo:hi]))
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_band_gradient_image_world2():
    """Backward pass of a band loss: what each rank assembles of d_image in the three exchange modes."""
    world = 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_band_gradient, args=(r, world, port, 12, 3, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in res) == [0, 1]
    for _, ok in res:
        assert all(ok.values()), ok


def test_band_gradient_mode_follows_the_scene():
    class P:
        pass
    bands = [(0, 8), (8, 16)]
    p = P(); p.use_prefiltering = False
    assert sharded.band_gradient_mode(p, bands) == 'gather'            # boundary samples land anywhere
    p.use_prefiltering = True; p.needs_filter_grad = True; p.filter_radius = 0.5; p.halo_rows = 1
    assert sharded.band_gradient_mode(p, bands) == 'halo'              # d_filter.radius reads the 3x3 footprint
    p.needs_filter_grad = False
    assert sharded.band_gradient_mode(p, bands) == 'own'
    p.filter_radius = 1.5; p.halo_rows = 2
    assert sharded.band_gradient_mode(p, bands) == 'halo'              # (a wide filter: samples read beyond their pixel)
    p.halo_rows = 9
    assert sharded.band_gradient_mode(p, bands) == 'gather'            # bands thinner than the halo


def test_row_split_reproduces_whole_render_on_the_oracle():
    """The property the sharding relies on, checked on the checker itself: rendering with the canvas
    height and image height scaled to a band is NOT what *_rows does -- *_rows keeps the global sample
    indices.  The oracle has no row entry point, so the property is checked through linearity of the
    backward pass instead: the gradient of a d_image that is zero outside a band of rows equals that
    band's share, and the shares of a partition sum to the whole (interior and boundary samples both
    scatter with weights read from d_image at their own pixel only when the filter is the 0.5 box)."""
    import oracle_check
    import scenes
    import util
    topo, params = util.pack(scenes.painterly(num_paths=24, canvas=32))
    w = h = 32
    rng = np.random.RandomState(0)
    d_img = rng.randn(h, w, 4).astype(np.float32)
    whole = oracle_check.render(topo, params, w, h, 2, 2, 5, d_render_image=d_img)['d_params'].astype(np.float64)
    acc = np.zeros_like(whole)
    for b, e in sharded.row_partition(h, 2, 8):
        part = np.zeros_like(d_img)
        part[b:e] = d_img[b:e]
        acc += oracle_check.render(topo, params, w, h, 2, 2, 5, d_render_image=part)['d_params'].astype(np.float64)
    assert np.linalg.norm(acc - whole) <= 1e-4 * np.linalg.norm(whole)


def test_balanced_row_partition_equalises_cost():
    import numpy as np
    rng = np.random.RandomState(0)
    for height, align, world in ((2048, 8, 8), (2048, 2, 8), (512, 8, 4), (510, 8, 3), (64, 8, 8), (40, 8, 8)):
        units = (height + align - 1) // align
        x = np.linspace(-1, 1, units)
        costs = 50.0 * np.exp(-6 * x * x) + rng.rand(units)         # dense in the middle, like flower.svg
        bands = sharded.balanced_row_partition(costs, height, world, align, per_unit=1.0)
        assert len(bands) == world and bands[0][0] == 0 and bands[-1][1] == height
        for k, (b, e) in enumerate(bands):
            assert b % align == 0 and b <= e
            if k:
                assert bands[k - 1][1] == b
        if units > world:
            assert all(e > b for b, e in bands)
            c = costs + 1.0
            share = [c[b // align:(e + align - 1) // align].sum() for b, e in bands]
            even = [c[b // align:(e + align - 1) // align].sum() for b, e in sharded.row_partition(height, world, align)]
            assert max(share) <= max(even) + 1e-9
            if units >= 16 * world:
                assert max(share) <= 1.15 * c.sum() / world
        else:
            assert bands == sharded.row_partition(height, world, align)
    # uniform cost: the plain partition (up to one unit)
    bands = sharded.balanced_row_partition(np.ones(256), 2048, 8, 8)
    assert bands == sharded.row_partition(2048, 8, 8)


def test_rebalance_bands_converges_on_a_hidden_cost_profile():
    import numpy as np
    height, align, world = 2048, 8, 8
    units = height // align
    x = np.linspace(-1, 1, units)
    true = 1.0 + 6.0 * np.exp(-5 * (x - 0.2) ** 2)                 # cost per tile row, unknown to the partitioner
    fixed = 0.05 * true.sum() / world                              # plus a constant per rank

    def times(bands):
        return [true[b // align:e // align].sum() + fixed for b, e in bands]
    bands = sharded.row_partition(height, world, align)
    first = max(times(bands))
    for _ in range(3):
        bands = sharded.rebalance_bands(bands, times(bands), height, align)
        assert bands[0][0] == 0 and bands[-1][1] == height and all(b % align == 0 and e > b for b, e in bands)
        assert all(bands[k][1] == bands[k + 1][0] for k in range(world - 1))
    t = times(bands)
    assert max(t) < 0.8 * first and max(t) <= 1.04 * (sum(t) / world)
    damped = sharded.rebalance_bands(sharded.row_partition(height, world, align), times(sharded.row_partition(height, world, align)), height, align, damping=0.5)
    assert damped[0][0] == 0 and damped[-1][1] == height and all(e > b for b, e in damped)
