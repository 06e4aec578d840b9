#!/usr/bin/env python
"""bench.py -- forward+backward iterations/s on the painterly config (BASELINE.json configs[2], the one
its metric is quoted on): 2048 random open cubic strokes, 512x512, 4x4 spp, L2 loss against a
synthetic target; one step = scene build + forward + loss gradient + backward.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Own arm (default): the CUDA path through the C ABI.
  value      device-resident step (params / target already in HBM), CUDA-event timed per step, L2 flushed
             between steps, max over ranks.
  e2e        the same step through the pydiffvg face with HOST tensors: pydiffvg.PackedParams leaves on the
             CPU -> H2D of the packed params -> RenderFunction.apply -> loss -> backward -> D2H of the gradient
             and the loss.  `e2e_stock_api` is the same through the per-tensor `serialize_scene` convention of
             the reference (6144 leaf tensors at this config).
  N > 1      one process per GPU (torchrun).  Headline line: rank r renders seed = step*N + r of the C3 scene
             (N independent sample sets per step -- the batch partition of SURVEY 8e) and the per-GPU gradient
             buffers are summed with one NCCL all-reduce; value = N*K / max-over-ranks time ("weak").
             `strong`: ONE 2048x2048 render split by pixel rows over the N ranks through
             diffvg_b200.sharded.ShardedRenderFunction (BASELINE.json configs[3]: flower.svg, use_prefiltering,
             2x2 spp; and the painterly scene at 2048^2 through the sampled path), timed next to the same render
             on rank 0 alone in the same run: ms/step at N and at 1, speed-up, efficiency.
Reference arm (--impl reference): the reference's own CPU implementation (oracle/_ref, the unmodified
sources compiled by oracle/Makefile, stock flags) on the host cores, on a bounded sample.
"""
import argparse
import glob
import json
import os
import subprocess
import sys
import tempfile
import time
import warnings

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

W = H = 512
NSX = NSY = 4
NUM_PATHS = 2048
N_SAMPLES = W * H * NSX * NSY
EVALS = {'fwd': 1.0, 'interior': 1.0, 'edge': 1.918}   # colour evaluations per pixel sample (edge: 2 sides x 96 % valid)

# Algorithmic flops per colour evaluation at this config, split by pipe (SURVEY 8d event model, F_alg, per pixel
# sample): E2 group-leaf visits 75 + E9 fixed 40 | E4 set-up 4.93 x 200 (90 of them FP64) + E5 bracket evaluations
# 13.5 x 41 (FP64) | E6 Newton 29.7 x 43 (FP64) + E7 accepted roots 7 x 50 | E8 fragments 1.52 x 15 + E9 40.
# name -> (FP32 flop, FP64 flop, passes the kernel serves in one step; the interior backward pass re-uses the forward
# pass's classification and exact tests)
KERNEL_MODEL = {
    'k_wave_classify_px': (115.0, 0.0, ('fwd',)),
    'k_wave_classify_edge': (115.0, 0.0, ('edge',)),
    'k_wave_stroke_solve': (542.0 + 350.0, 997.0 + 1277.0, ('fwd', 'edge')),   # E4 + E5 (set-up, bracket evaluations) and E6 + E7 (Newton, accepted roots)
    'k_wave_composite_px<false>': (63.0, 0.0, ('fwd',)),
    'k_wave_composite_px<true>': (63.0, 0.0, ('interior',)),
    'k_wave_composite_edge': (63.0, 0.0, ('edge',)),
}
STEP_FP32 = (115.0 + 542.0 + 350.0 + 63.0) * (EVALS['fwd'] + EVALS['edge']) + 63.0 * EVALS['interior']
STEP_FP64 = (997.0 + 1277.0) * (EVALS['fwd'] + EVALS['edge'])

METRIC = 'fwd+bwd iters/s'
UNIT = 'it/s'
WORKLOAD = 'painterly: 2048 open cubic strokes (1-3 segments, width 1-4), 512x512, 4x4 spp, L2 loss, fwd+bwd'


def config_dict(n_gpus):
    return {'workload': WORKLOAD, 'num_paths': NUM_PATHS, 'width': W, 'height': H, 'spp': NSX * NSY,
            'samples_per_iter': 2 * N_SAMPLES, 'cache': 'L2 flushed between timed steps (256 MiB write)',
            'sharding': 'none' if n_gpus == 1 else 'headline: per-rank seeds (batch partition) + NCCL all-reduce of the gradient '
                        'buffer; `strong`: one 2048^2 render split by pixel rows'}


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.proc = None
        self.path = None
        try:
            fd, self.path = tempfile.mkstemp(suffix='.csv')
            os.close(fd)
            self.f = open(self.path, 'w')
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(gpu_index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '20'], stdout=self.f,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': []}
        if self.proc is None:
            return out
        try:
            self.proc.terminate()
            self.proc.wait(timeout=5)
            self.f.close()
            rows = [l.strip().split(', ') for l in open(self.path) if l.strip()]
            sm = [float(r[0]) for r in rows if len(r) >= 6]
            if sm:
                out['sm_mhz'] = float(np.median(sm))
                out['sm_max_mhz'] = float(rows[0][1])
                names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
                out['reasons'] = [n for k, n in enumerate(names) if any(r[2 + k].strip() == 'Active' for r in rows if len(r) >= 6)]
                out['samples'] = len(sm)
        except Exception:
            pass
        finally:
            try:
                os.unlink(self.path)
            except Exception:
                pass
        return out


# ------------------------------------------------------------------------------------------ reference arm
CPU_NOTE = ('unmodified reference sources, stock flags (oracle/_ref/diffvg*.so, not the zero-initialising parity build), driven '
            'through oracle/ref_capi.cpp: the ~0.55 s/iteration of pydiffvg/pybind glue the reference spends in Python '
            '(SURVEY 3.3) is NOT included, which favours the reference')


def reference_arm(args, rank):
    """The reference's own CPU implementation, all host threads, bounded sample per step."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import oracle_check
    import scenes
    import util
    warnings.simplefilter('ignore')
    topo, params = util.pack(scenes.painterly())
    cores = os.cpu_count()
    target_full = torch.rand(H, W, 4, generator=torch.Generator().manual_seed(1234)).numpy()

    def one(rows, seed):
        """fwd + bwd on a `rows`-row render of the same scene (the canvas is squeezed vertically into
        fewer pixel rows: same scene and per-sample work, rows/512 of the samples)."""
        t0 = time.perf_counter()
        img = oracle_check.render(topo, params, W, rows, NSX, NSY, seed, variant='plain')['image']
        d_img = (2.0 * (img - target_full[:rows]) / img.size).astype(np.float32)
        oracle_check.render(topo, params, W, rows, NSX, NSY, seed, d_render_image=d_img, variant='plain')
        return time.perf_counter() - t0

    if oracle_check.kind() != 'reference':   # oracle/_ref absent: the C restatement has no backward pass (and no threads)
        print(json.dumps({'impl': 'reference', 'unavailable': 'oracle/_ref (the compiled reference) is not built here; '
                          'oracle/dvg_oracle.c restates the forward colour path only'}), flush=True)
        return
    t_probe = one(32, 0)
    t_full_est = t_probe * (H / 32.0)
    budget = 150.0
    frac = min(1.0, budget / max((args.steps + args.warmup) * t_full_est, 1e-9))
    rows = int(max(16, min(H, (int(H * frac) // 16) * 16)))
    for i in range(args.warmup):
        one(rows, i)
    t = 0.0
    for i in range(args.steps):
        t += one(rows, args.warmup + i)
    ms_sample = 1e3 * t / args.steps
    ms_iter = ms_sample * (H / rows)       # one full 512-row iteration at the measured sample rate
    value = 1e3 / ms_iter
    sample = '%d of %d pixel rows per step (%.3f of the samples of one iteration), fwd+bwd, Scene rebuilt per call; %s' % (
        rows, H, rows / H, CPU_NOTE)
    line = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': ms_iter, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32+f64', 'data': 'synthetic', 'config': config_dict(1),
            'msamples_per_s': 2 * N_SAMPLES / (ms_iter * 1e-3) / 1e6,
            'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': oracle_check.kind(), 'sample': sample},
            'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ strong scaling (N > 1)
def strong_scaling(dev, rank, world, barrier, max_over_ranks, steps=6, warmup=3):
    """One 2048x2048 render split by pixel rows over all ranks, against the same render on rank 0 alone."""
    import torch.distributed as dist
    import scenes
    from diffvg_b200 import pydiffvg, sharded
    from diffvg_b200.pydiffvg.render_pytorch import OutputType, PackedScene
    SW = SH = 2048
    out = []
    flower = np.load(os.path.join(ROOT, 'tests', 'golden_svg', 'flower.npz'))
    cases = [('C4 flower.svg (1096 groups, 10.5 k cubics), 2048x2048, use_prefiltering, 2x2 spp, band loss, fwd+bwd', 'flower', True),
             ('painterly 2048 strokes at 2048x2048, 2x2 spp, sampled path (boundary pass), band loss, fwd+bwd', 'painterly', False)]
    for label, which, pf in cases:
        if which == 'flower':
            topo, params_np = flower['topo'], flower['params']
            packed = PackedScene(np.ascontiguousarray(topo), int(topo[1]), int(topo[2]), OutputType.color, True, torch.tensor([]))
            packed.needs_xform_grad = False
            packed.needs_filter_grad = False      # the pixel-filter radius is a constant, as in every reference app (both arms)
            packed.filter_radius, packed.halo_rows = 0.5, 1
            params = torch.from_numpy(params_np).to(dev).requires_grad_(True)
        else:
            cw, ch, shapes, groups = scenes.painterly()
            packed, params = pydiffvg.RenderFunction.serialize_scene(cw, ch, shapes, groups)
            params = params.detach().to(dev).requires_grad_(True)
        target = torch.rand(SH, SW, 4, generator=torch.Generator().manual_seed(1)).to(dev)
        # bands of equal estimated cost (whole-image bin counts of the start parameters), cut once outside the timed region
        bands = sharded.balanced_bands(packed, params.detach(), SW, SH, 2, 2, world)
        rb, re = bands[rank]

        def step_sharded(seed):
            params.grad = None
            # loss per band: no image exchange in the forward pass; backward: halo rows of d_image only on the prefiltered
            # path, an all-gather of the d_image bands on the sampled path (its boundary samples land anywhere)
            img = sharded.ShardedRenderFunction.apply(SW, SH, 2, 2, seed, None, packed, params, None, False, bands)
            ((img - target[rb:re]).pow(2).sum() / target.numel()).backward()

        def step_single(seed):
            params.grad = None
            img = pydiffvg.RenderFunction.apply(SW, SH, 2, 2, seed, None, packed, params)
            (img - target).pow(2).mean().backward()

        def timed(fn, only_rank0):
            if only_rank0 and rank != 0:
                barrier()
                barrier()
                return 0.0
            for i in range(warmup):
                fn(i)
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for i in range(steps):
                fn(warmup + i)
            b.record()
            barrier()
            return a.elapsed_time(b) / steps

        # bands re-cut from measured per-rank times (three rounds, outside the timed region): the bin-count estimate cannot
        # see how expensive a tile's candidates are
        for rnd in range(3):
            step_sharded(100 + rnd)
            barrier()
            sharded.time_compute(True)      # the rank's own render calls only: the collectives make every rank wait for the slowest
            for i in range(2):
                step_sharded(110 + 2 * rnd + i)
            mine = torch.tensor([sharded.compute_ms()], dtype=torch.float64, device=dev)
            sharded.time_compute(False)
            allt = [torch.zeros_like(mine) for _ in range(world)]
            dist.all_gather(allt, mine)
            bands = sharded.rebalance_bands(bands, [float(t.item()) for t in allt], SH, sharded.tile_height(4))
            rb, re = bands[rank]
        ms1 = max_over_ranks(timed(step_single, True))
        g1 = float(params.grad.norm()) if rank == 0 else 0.0
        msn = max_over_ranks(timed(step_sharded, False))
        gn = float(params.grad.norm())
        out.append({'workload': label, 'ms_per_step_1gpu': ms1, 'ms_per_step': msn, 'n_gpus': world, 'speedup': ms1 / msn,
                    'efficiency': ms1 / msn / world, 'it_per_s': 1e3 / msn, 'grad_norm_1gpu': g1, 'grad_norm': gn,
                    'bands': 'rows balanced by whole-image bin counts (sharded.balanced_bands), then re-cut three times from measured per-rank times during warm-up (sharded.rebalance_bands): %s' % (bands,),
                    'timing': 'CUDA events around %d steps after %d warm-up steps, max over ranks; 1-GPU figure on rank 0 of the same run' % (steps, warmup)})
        del target, params
    return out


# ------------------------------------------------------------------------------------------ own arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='own', choices=['own', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-strong', action='store_true', help='N > 1: skip the 2048^2 strong-scaling sub-records')
    ap.add_argument('--quick', action='store_true', help='development: device-timed value and per-kernel times only (no e2e legs)')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'own' else args.warmup

    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if args.impl == 'reference':
        reference_arm(args, rank)
        return

    import ctypes
    import torch.distributed as dist
    from diffvg_b200 import _native as n
    from diffvg_b200 import pydiffvg
    import scenes
    import util
    warnings.simplefilter('ignore')

    assert torch.cuda.is_available(), 'bench.py needs a CUDA device (there is no CPU fallback)'
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        if os.environ.get('NCCL_DEBUG', '').upper() in ('VERSION', 'WARN'):
            os.environ.pop('NCCL_DEBUG')           # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
        dist.init_process_group('nccl', device_id=dev)
        # one process per GPU on a shared host: keep ATen's intra-op pool from oversubscribing the cores
        torch.set_num_threads(max(1, (os.cpu_count() or 8) // world))
    n_gpus = world

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- inputs, resident in HBM before the timed region
    scene = scenes.painterly()
    topo, params_np = util.pack(scene)
    params_dev = torch.from_numpy(params_np).to(dev)
    target = torch.rand(H, W, 4, generator=torch.Generator().manual_seed(1234)).to(dev)
    h = ctypes.c_void_p()
    n.check(n.lib.dvg_scene_create(topo.ctypes.data, topo.shape[0], local_rank, ctypes.byref(h)))
    stream = torch.cuda.current_stream().cuda_stream
    img = torch.empty(H, W, 4, device=dev)
    d_params = torch.empty(params_np.shape[0], device=dev)
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    inv_numel = 2.0 / img.numel()

    def device_step(i, collective=True):
        seed = i * world + rank
        n.check(n.lib.dvg_scene_set_params(h, params_dev.data_ptr(), params_dev.numel(), 1, stream))
        n.check(n.lib.dvg_render_forward(h, None, img.data_ptr(), None, W, H, NSX, NSY, seed, 0, None, 0, stream))
        d_img = (img - target) * inv_numel          # d/d img of mean((img - target)^2)
        n.check(n.lib.dvg_render_backward(h, None, d_img.data_ptr(), None, W, H, NSX, NSY, seed, 0, None, 0,
                                          d_params.data_ptr(), None, None, 0, stream))
        if world > 1 and collective:
            dist.all_reduce(d_params)

    def timed(step_fn, steps, first):
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        barrier()
        t0 = time.perf_counter()
        for k in range(steps):
            flush_buf.fill_(k & 0xff)               # evict L2 between timed steps (outside the event pair)
            evs[k][0].record()
            step_fn(first + k)
            evs[k][1].record()
        barrier()
        wall = time.perf_counter() - t0
        dev_ms = sum(a.elapsed_time(b) for a, b in evs)
        return dev_ms, wall

    for i in range(args.warmup):
        device_step(i)
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = n.launch_count()
    dev_ms, wall = timed(device_step, args.steps, args.warmup)
    launches = n.launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    ms_per_step = max_over_ranks(dev_ms / args.steps)
    value = world * 1e3 / ms_per_step

    # ---- e2e: pydiffvg face, host tensors in / host gradients + loss out
    pydiffvg.set_use_gpu(True)
    pydiffvg.set_device(dev)
    cw, ch, shapes, groups = scene

    def time_e2e(step_fn, steps):
        for i in range(3):
            step_fn(i)
        barrier()
        t0 = time.perf_counter()
        for k in range(steps):
            step_fn(args.warmup + k)
        barrier()
        return max_over_ranks(1e3 * (time.perf_counter() - t0) / steps)

    # (a) the per-tensor convention of the reference: 2048 x (points, stroke_width) + 2048 colours as separate CPU leaves
    leaves = []
    for s in shapes:
        s.points.requires_grad_(True)
        s.stroke_width.requires_grad_(True)
        leaves += [s.points, s.stroke_width]
    for g in groups:
        g.stroke_color.requires_grad_(True)
        leaves.append(g.stroke_color)
    e2e_bytes = {}

    def e2e_stock_step(i):
        seed = i * world + rank
        for t in leaves:
            t.grad = None
        sargs = pydiffvg.RenderFunction.serialize_scene(cw, ch, shapes, groups)
        out = pydiffvg.RenderFunction.apply(W, H, NSX, NSY, seed, None, *sargs)
        loss = (out - target).pow(2).mean()
        loss.backward()                              # gradients land on the CPU leaves (D2H inside)
        if world > 1:
            flat = torch.cat([t.grad.reshape(-1) for t in leaves]).to(dev)
            dist.all_reduce(flat)
        return loss.item()

    e2e_stock_ms = time_e2e(e2e_stock_step, max(3, args.steps // 2)) if not args.quick else float('nan')
    for t in leaves:
        t.requires_grad_(False)
        t.grad = None

    # (b) pydiffvg.PackedParams: the same parameters as five CPU leaves in the renderer's own layout
    pp = pydiffvg.PackedParams(cw, ch, shapes, groups, device=torch.device('cpu'))

    def e2e_step(i):
        seed = i * world + rank
        for t in pp.leaves:
            t.grad = None
        sargs = pp.scene_args()                      # [PackedScene, params (CPU)]
        out = pydiffvg.RenderFunction.apply(W, H, NSX, NSY, seed, None, *sargs)   # H2D of params inside
        loss = (out - target).pow(2).mean()
        loss.backward()                              # D2H of the gradient inside; five AccumulateGrad
        if world > 1:
            flat = torch.cat([t.grad for t in pp.parameters()]).to(dev)
            dist.all_reduce(flat)
        e2e_bytes['h2d'] = sargs[1].numel() * 4
        e2e_bytes['d2h'] = sargs[1].numel() * 4 + 4
        return loss.item()                           # D2H of the loss

    e2e_ms = time_e2e(e2e_step, args.steps if not args.quick else 1)
    e2e_value = world * 1e3 / e2e_ms

    # ---- N > 1: strong scaling of one 2048^2 render split by rows
    strong = None
    if world > 1 and not args.no_strong:
        try:
            strong = strong_scaling(dev, rank, world, barrier, max_over_ranks)
        except Exception as e:   # keep the headline line alive
            strong = [{'error': repr(e)}]

    # ---- per-kernel CUDA-event times over the same steps (separate pass: events between kernels)
    roofline = None
    kernels = None
    if rank == 0:
        n.profile_enable(True)
        for k in range(args.steps):
            flush_buf.fill_(k & 0xff)
            device_step(args.warmup + k, collective=False)   # rank 0 alone: no collective in this pass
        torch.cuda.synchronize(dev)
        rep = n.profile_report()
        n.profile_enable(False)
        kernels = {}
        for k, (c, ms) in rep.items():
            k = k[:-len('<false>')] if k.startswith('k_wave_classify') and k.endswith('<false>') else k   # (the <true> forms are the overflow retries)
            e = kernels.setdefault(k, {'launches': 0, 'ms_per_step': 0.0})
            e['launches'] += c
            e['ms_per_step'] += ms / args.steps
        total_k = sum(v['ms_per_step'] for v in kernels.values())
        top = max(kernels, key=lambda k: kernels[k]['ms_per_step'])
        fp32_peak = n.measure_peak(0, local_rank)
        fp64_peak = n.measure_peak(1, local_rank)
        f32, f64, passes = KERNEL_MODEL.get(top, (STEP_FP32, STEP_FP64, ()))
        evals = sum(EVALS[p] for p in passes) if passes else 1.0
        top_ms = kernels[top]['ms_per_step']
        g32, g64 = f32 * evals * N_SAMPLES, f64 * evals * N_SAMPLES       # algorithmic flops of the kernel per STEP, by pipe
        bound = 'fp64' if g64 / max(fp64_peak, 1e-9) >= g32 / max(fp32_peak, 1e-9) else 'fp32'
        achieved = (g64 if bound == 'fp64' else g32) / (top_ms * 1e-3) / 1e12
        peak = fp64_peak if bound == 'fp64' else fp32_peak
        floor_ms = (g64 / (fp64_peak * 1e12) + g32 / (fp32_peak * 1e12)) * 1e3
        step_floor_ms = (STEP_FP64 * N_SAMPLES / (fp64_peak * 1e12) + STEP_FP32 * N_SAMPLES / (fp32_peak * 1e12)) * 1e3
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        except Exception:
            pass
        # compulsory DRAM traffic of the whole step: images + scene tables
        alg_bytes = W * H * 4 * 4 * 2 + W * H * 4 + 2 * params_np.nbytes * 8
        roofline = {'bound': bound, 'kernel': top, 'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s',
                    'frac': achieved / peak if peak else None,
                    'frac_definition': 'algorithmic %s flops of the dominant kernel / its time / the %s FMA-chain peak (flops of the other pipe are not counted)' % (bound.upper(), bound.upper()),
                    'kernel_two_pipe_frac': floor_ms / top_ms, 'step_two_pipe_frac': step_floor_ms / ms_per_step,
                    'two_pipe_definition': 'time floor = FP64 flops / FP64 peak + FP32 flops / FP32 peak (pipes not overlapped), divided by the measured time',
                    'peak_source': 'FMA-chain probes measured live in this run (MEASURED_PEAKS.json has no CUDA-core figure): '
                                   'FP32 %.2f, FP64 %.2f TFLOP/s' % (fp32_peak, fp64_peak),
                    'kernel_ms_per_step': top_ms, 'kernel_launches_per_step': kernels[top]['launches'] / args.steps,
                    'kernel_share_of_step': top_ms / total_k,
                    'algorithmic_flops_per_step': {'fp32': g32, 'fp64': g64},
                    'model': 'SURVEY 8d event model: %.0f FP32 + %.0f FP64 flop per colour evaluation for this kernel (%s), '
                             '%.3f evaluations per pixel sample' % (f32, f64, '+'.join(passes), evals),
                    'step_achieved_tflops': (STEP_FP32 + STEP_FP64) * N_SAMPLES / (ms_per_step * 1e-3) / 1e12,
                    'hbm': {'algorithmic_bytes': alg_bytes, 'achieved_gbs': alg_bytes / (ms_per_step * 1e-3) / 1e9,
                            'peak_gbs': peaks.get('hbm_gbs', 6650.0),
                            'peak_source': 'measured' if 'hbm_gbs' in peaks else 'fallback'},
                    'traffic': None, 'fp64_pipe_pct': None}
        try:   # DRAM bytes and FP64-pipe utilisation of this kernel's longest launch, from the newest committed ncu --set full capture
            cands = sorted(glob.glob(os.path.join(ROOT, 'profiles', 'r*_ncu_traffic.json')))
            tr = json.load(open(cands[-1]))['kernels'].get(top)
            if tr:
                roofline['traffic'] = tr['dram_bytes']
                roofline['fp64_pipe_pct'] = tr.get('fp64_pipe_pct')
                roofline['traffic_note'] = 'dram__bytes_read+write of the longest launch of %s (%.3f ms) in %s' % (
                    top, tr['ms'], os.path.relpath(cands[-1], ROOT))
        except Exception:
            pass

    # ---- CPU baseline + the reference's own CUDA build (rank 0, N = 1)
    cpu_baseline = None
    reference_cuda = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sys.path.insert(0, os.path.join(ROOT, 'oracle'))
        import oracle_check
        import ref_oracle
        try:
            if oracle_check.kind() != 'reference':
                raise oracle_check.OracleUnavailable('oracle/_ref is not built here')
            t0 = time.perf_counter()
            ref_img = oracle_check.render(topo, params_np, W, H, NSX, NSY, 0, variant='plain')['image']
            d_img_np = (2.0 * (ref_img - target.cpu().numpy()) / ref_img.size).astype(np.float32)
            oracle_check.render(topo, params_np, W, H, NSX, NSY, 0, d_render_image=d_img_np, variant='plain')
            t_cpu = time.perf_counter() - t0
            cpu_baseline = {'value': 1.0 / t_cpu, 'unit': UNIT, 'cores': os.cpu_count(), 'kind': oracle_check.kind(),
                            'sample': '1 full fwd+bwd iteration of the same workload (seed 0), all host threads, '
                                      'Scene rebuilt per call as the reference does; ' + CPU_NOTE}
            if ref_oracle.cuda_available():
                try:
                    ms, cimg, _ = ref_oracle.cuda_bench(topo, params_np, W, H, NSX, NSY, 0, d_img_np, 2, 5)
                    reference_cuda = {'ms_per_step': ms, 'value': 1e3 / ms, 'unit': UNIT,
                                      'max_abs_image_diff_vs_cpu_reference': float(np.abs(cimg - ref_img).max()),
                                      'what': "the reference's own CUDA path (diffvg.cpp / scene.cpp through nvcc for sm_100a, oracle/Makefile "
                                              'ref_cuda) on this GPU: per step Scene() + forward render() + Scene() + backward render(), wall '
                                              'clock, images in device memory, no Python glue; informational (SURVEY 8d)'}
                except Exception as e:
                    reference_cuda = {'unavailable': repr(e)}
            else:
                reference_cuda = {'unavailable': 'oracle/_ref/diffvg_cuda.so not built'}
        except oracle_check.OracleUnavailable as e:
            cpu_baseline = {'value': None, 'unit': UNIT, 'cores': os.cpu_count(), 'kind': 'port', 'sample': 'unavailable: ' + str(e)}

    if rank == 0:
        line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': n_gpus, 'steps': args.steps,
                'warmup': args.warmup, 'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak',
                'vs_baseline': None, 'dtype': 'f32+f64', 'data': 'synthetic', 'config': config_dict(n_gpus),
                'msamples_per_s': world * 2 * N_SAMPLES / (ms_per_step * 1e-3) / 1e6,
                'wall_ms_per_step': 1e3 * wall / args.steps,
                'e2e': {'value': e2e_value, 'unit': UNIT, 'ms_per_step': e2e_ms,
                        'h2d_bytes_per_step': e2e_bytes.get('h2d', 0), 'd2h_bytes_per_step': e2e_bytes.get('d2h', 0),
                        'api': 'pydiffvg.PackedParams (CPU leaves) -> RenderFunction.apply -> loss.backward -> loss.item()'},
                'e2e_stock_api': {'value': world * 1e3 / e2e_stock_ms, 'unit': UNIT, 'ms_per_step': e2e_stock_ms,
                                  'api': 'RenderFunction.serialize_scene over 6144 CPU leaf tensors (the reference\'s calling convention)'},
                'gpu_launches': launches, 'clocks': clocks, 'roofline': roofline, 'cpu_baseline': cpu_baseline,
                'reference_cuda': reference_cuda, 'strong': strong, 'kernels': kernels}
        print(json.dumps(line), flush=True)
    n.lib.dvg_scene_destroy(h)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
