#!/usr/bin/env python
"""Strong scaling of ONE large render split by pixel rows (BASELINE configs[3] shape: 2048x2048, 2x2 spp,
use_prefiltering, fill-heavy scene), one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29513 \
        tools/scale_c4.py [--strokes]

Each step: set_params + scene build (every rank, replicated) -> forward of the rank's row band -> all-gather of the
bands -> loss gradient on the assembled image -> backward of the band -> NCCL all-reduce of the gradient buffer.
Timed with CUDA events per rank, max over ranks; rank 0 prints one JSON line.  --strokes: the painterly scene at
2048^2, 2x2 spp without prefiltering (boundary pass included) instead."""
import json
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
    import warnings
    warnings.simplefilter('ignore')
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        os.environ.pop('NCCL_DEBUG', None)
        dist.init_process_group('nccl', device_id=dev)
    from diffvg_b200 import pydiffvg, sharded
    import scenes
    pydiffvg.set_use_gpu(True)
    pydiffvg.set_device(dev)
    strokes = '--strokes' in sys.argv
    cw, ch, shapes, groups = scenes.painterly() if strokes else scenes.blobs()
    W = H = 2048
    packed, params = pydiffvg.RenderFunction.serialize_scene(cw, ch, shapes, groups, use_prefiltering=not strokes)
    params = params.detach().to(dev).requires_grad_(True)
    target = torch.rand(H, W, 4, generator=torch.Generator().manual_seed(1)).to(dev)

    band_loss = '--band-loss' in sys.argv   # loss computed per band: no image exchange (see sharded.py)
    rb, re = sharded.row_partition(H, world, sharded.tile_height(4))[rank]

    def step(seed):
        params.grad = None
        if world > 1 and band_loss:
            img = sharded.ShardedRenderFunction.apply(W, H, 2, 2, seed, None, packed, params, None, False)
            ((img - target[rb:re]).pow(2).sum() / target.numel()).backward()
            return
        if world > 1:
            img = sharded.ShardedRenderFunction.apply(W, H, 2, 2, seed, None, packed, params)
        else:
            img = pydiffvg.RenderFunction.apply(W, H, 2, 2, seed, None, packed, params)
        (img - target).pow(2).mean().backward()

    for i in range(3):
        step(i)
    steps = 8
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(steps):
        step(3 + i)
    b.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / steps
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    if rank == 0:
        print(json.dumps({'workload': ('painterly strokes' if strokes else 'blobs1024 prefiltered') + ' 2048x2048 2x2 spp, row-sharded, fwd+bwd',
                          'n_gpus': world, 'band_loss': '--band-loss' in sys.argv, 'ms_per_step': ms, 'it_per_s': 1e3 / ms, 'grad_norm': float(params.grad.norm())}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
