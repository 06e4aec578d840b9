#!/usr/bin/env python
"""Multi-GPU check of the row-sharded render (diffvg_b200/sharded.py), run under torchrun on one box:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/multigpu_check.py

Every rank renders its band of pixel rows of the SAME scene / seed (forward: all-gather of the bands;
backward: boundary samples of the same rows, NCCL all-reduce of the gradient buffer) and compares the
assembled image and the summed gradients with an unsharded render of its own.  Prints one line per rank."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    dist.init_process_group('nccl', device_id=dev)
    from diffvg_b200 import pydiffvg, sharded
    import scenes
    pydiffvg.set_use_gpu(True)
    pydiffvg.set_device(dev)
    ok = True
    for name, scene, (w, h, nsx, nsy), pf in (('painterly256', scenes.painterly(num_paths=256, canvas=128), (128, 128, 4, 4), False),
                                              ('blobs64', scenes.blobs(num_paths=64, canvas=96), (96, 96, 2, 2), False),
                                              ('blobs64-pf', scenes.blobs(num_paths=64, canvas=96), (192, 192, 2, 2), True),
                                              ('zoo-hann1.5', scenes.zoo(), (128, 128, 2, 2), 'hann')):
        cw, ch, shapes, groups = scene
        if pf == 'hann':   # wide pixel filter: samples splat across band edges
            args = pydiffvg.RenderFunction.serialize_scene(cw, ch, shapes, groups, filter=pydiffvg.PixelFilter(
                type=pydiffvg.FilterType.hann, radius=torch.tensor(1.5)))
        else:
            args = pydiffvg.RenderFunction.serialize_scene(cw, ch, shapes, groups, use_prefiltering=pf)
        packed, params = args
        params = params.detach().to(dev).requires_grad_(True)
        target = torch.rand(h, w, 4, generator=torch.Generator().manual_seed(3)).to(dev)
        img1 = pydiffvg.RenderFunction.apply(w, h, nsx, nsy, 7, None, packed, params)
        (g1,) = torch.autograd.grad((img1 - target).pow(2).mean(), params)
        img2 = sharded.ShardedRenderFunction.apply(w, h, nsx, nsy, 7, None, packed, params)
        (g2,) = torch.autograd.grad((img2 - target).pow(2).mean(), params)
        d_img = float((img1 - img2).detach().abs().max())
        rel = float((g1 - g2).norm() / g1.norm().clamp_min(1e-30))
        # per-band loss, no image exchange in the forward pass (gather=False; needs a filter radius <= 0.5)
        rb, re = sharded.row_partition(h, world, sharded.tile_height(nsx * nsy))[rank]
        if pf == 'hann':
            img3, g3 = img1[rb:re], g1
        else:
            img3 = sharded.ShardedRenderFunction.apply(w, h, nsx, nsy, 7, None, packed, params, None, False)
            (g3,) = torch.autograd.grad((img3 - target[rb:re]).pow(2).sum() / target.numel(), params)
        rel3 = float((g1 - g3).norm() / g1.norm().clamp_min(1e-30))
        worst = int((g1 - g3).abs().argmax())
        if rel3 > 1e-4 or rel > 1e-4:
            print('   rank %d %s: gather rel %.3g, band-loss rel %.3g, worst entry %d of %d: %g vs %g (filter radius at %d)' % (
                rank, name, rel, rel3, worst, g1.numel(), float(g1[worst]), float(g3[worst]), int(packed.topo[6])), flush=True)
        if pf is True:   # the same with the gradient of the pixel-filter radius asked for: the backward pass exchanges halo rows
            packed.needs_filter_grad = True
            (g1f,) = torch.autograd.grad((pydiffvg.RenderFunction.apply(w, h, nsx, nsy, 7, None, packed, params) - target).pow(2).mean(), params)
            img4 = sharded.ShardedRenderFunction.apply(w, h, nsx, nsy, 7, None, packed, params, None, False)
            (g4,) = torch.autograd.grad((img4 - target[rb:re]).pow(2).sum() / target.numel(), params)
            packed.needs_filter_grad = False
            rel3 = max(rel3, float((g1f - g4).norm() / g1f.norm().clamp_min(1e-30)))
            assert float((g1f - g1).abs().max()) > 0, 'the filter-radius gradient should differ from the masked one'
        rel = max(rel, rel3)
        d_img = max(d_img, float((img1[rb:re] - img3).detach().abs().max()))
        good = d_img <= 1e-6 and rel <= 1e-4
        ok = ok and good
        print('rank %d/%d %-12s bands %s image max-abs %.3g grad rel-L2 %.3g %s' % (
            rank, world, name, sharded.row_partition(h, world, sharded.tile_height(nsx * nsy)), d_img, rel, 'OK' if good else 'MISMATCH'),
            flush=True)
    flag = torch.tensor([0 if ok else 1], device=dev)
    dist.all_reduce(flag)
    dist.destroy_process_group()
    sys.exit(int(flag.item() != 0))


if __name__ == '__main__':
    main()
