#!/usr/bin/env python
"""bench.py -- forward+backward iterations/s on the painterly config (BASELINE.json configs[2],
the one its metric is quoted on): 2048 random open cubic strokes, 512x512, 4x4 spp, L2 loss
against a synthetic target; one step = scene build + forward + loss gradient + backward.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Own arm (default): the CUDA path through the C ABI.
  value  : device-resident step (params / target already in HBM), CUDA-event timed per step,
           L2 flushed between steps, max over ranks.
  e2e    : the same step through pydiffvg.RenderFunction with HOST tensors (serialize_scene on
           CPU tensors -> H2D of the packed params -> render -> backward -> D2H of the gradient
           and the loss).
  N > 1  : one process per GPU (torchrun); rank r renders seed = step*N + r of the same scene
           (N independent sample sets per step) and the per-GPU gradient buffers are summed with
           one NCCL all-reduce; value = N*K / max-over-ranks time ("weak" scaling).
Reference arm (--impl reference): the reference's own CPU implementation (oracle/_ref, the
unmodified sources compiled by oracle/Makefile) on the host cores, on a bounded sample.
"""
import argparse
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
# Algorithmic flops per colour evaluation at this config (SURVEY 8d event model, F_alg):
FALG_PER_EVAL = 3.3e3
EVALS_FWD = 1.0          # colour evaluations per pixel sample, forward
EVALS_INTERIOR = 1.0     # interior backward (forward recompute)
EVALS_BOUNDARY = 1.918   # boundary pass (two sides x 96% valid samples)

# Share of the event model per kernel: flop per colour evaluation, the passes the kernel serves in one step
# (the interior backward pass re-uses the forward pass's classification and exact tests), and the pipe it is
# bound by.  E2 group-leaf visits 75 + E9 fixed 40 | E4 set-up 986 + E5 bracket evaluations 553 |
# E6 Newton 1277 + E7 accepted roots 350 | E8 fragments 23 (per pixel sample at this config).
KERNEL_MODEL = {
    'k_wave_classify_px': (115.0, ('fwd',), 'fp32'),
    'k_wave_classify_edge': (115.0, ('edge',), 'fp32'),
    'k_wave_stroke_setup': (1539.0, ('fwd', 'edge'), 'fp64'),
    'k_wave_stroke_newton': (1627.0, ('fwd', 'edge'), 'fp64'),
    'k_wave_composite_px<false>': (63.0, ('fwd',), 'fp32'),
    'k_wave_composite_px<true>': (63.0, ('interior',), 'fp32'),
    'k_wave_composite_edge': (63.0, ('edge',), 'fp32'),
    'k_edge': (3304.0, ('edge',), 'fp32'),
    'k_render<true>': (3304.0, ('interior',), 'fp32'),
    'k_render<false>': (3304.0, ('fwd',), 'fp32'),
}

METRIC = 'fwd+bwd iters/s'
UNIT = 'it/s'
WORKLOAD = 'painterly: 2048 open cubic strokes (1-3 segments, width 1-4), 512x512, 4x4 spp, L2 loss, fwd+bwd'


def config_dict(n_gpus):
    return {'workload': WORKLOAD, 'num_paths': NUM_PATHS, 'width': W, 'height': H, 'spp': NSX * NSY,
            'samples_per_iter': 2 * N_SAMPLES, 'cache': 'L2 flushed between timed steps (256 MiB write)',
            'sharding': 'none' if n_gpus == 1 else 'per-rank seeds, NCCL all-reduce of the gradient buffer'}


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
        img = oracle_check.render(topo, params, W, rows, NSX, NSY, seed)['image']
        d_img = (2.0 * (img - target_full[:rows]) / img.size).astype(np.float32)
        oracle_check.render(topo, params, W, rows, NSX, NSY, seed, d_render_image=d_img)
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
    sample = '%d of %d pixel rows per step (%.3f of the samples of one iteration), fwd+bwd, Scene rebuilt per call' % (
        rows, H, rows / H)
    line = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': ms_iter, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32+f64', 'data': 'synthetic', 'config': config_dict(1),
            'msamples_per_s': 2 * N_SAMPLES / (ms_iter * 1e-3) / 1e6,
            'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': oracle_check.kind(), 'sample': sample},
            'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ own arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='own', choices=['own', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
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

    # ---- e2e: pydiffvg API, host tensors in / host gradients + loss out
    pydiffvg.set_use_gpu(True)
    pydiffvg.set_device(dev)
    cw, ch, shapes, groups = scene
    leaves = []
    for s in shapes:
        s.points.requires_grad_(True)
        s.stroke_width.requires_grad_(True)
        leaves += [s.points, s.stroke_width]
    for g in groups:
        g.stroke_color.requires_grad_(True)
        leaves.append(g.stroke_color)
    e2e_bytes = {}

    def e2e_step(i):
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
        e2e_bytes['h2d'] = sargs[1].numel() * 4
        e2e_bytes['d2h'] = sargs[1].numel() * 4 + 4
        return loss.item()

    for i in range(3):
        e2e_step(i)
    barrier()
    t0 = time.perf_counter()
    e2e_steps = args.steps
    for k in range(e2e_steps):
        e2e_step(args.warmup + k)
    barrier()
    e2e_ms = max_over_ranks(1e3 * (time.perf_counter() - t0) / e2e_steps)
    e2e_value = world * 1e3 / e2e_ms

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
        kernels = {k: {'launches': c, 'ms_per_step': ms / args.steps} for k, (c, ms) in rep.items()}
        total_k = sum(v['ms_per_step'] for v in kernels.values())
        top = max(kernels, key=lambda k: kernels[k]['ms_per_step'])
        fp32_peak = n.measure_peak(0, local_rank)
        fp64_peak = n.measure_peak(1, local_rank)
        # algorithmic flops of the kernel per STEP: its share of the event model (KERNEL_MODEL) times the colour
        # evaluations of the passes it serves; a kernel launched by several passes is timed over all its launches
        flops_per_eval, passes, bound = KERNEL_MODEL.get(top, (FALG_PER_EVAL, ('fwd', 'edge'), 'fp32'))
        evals = sum({'fwd': EVALS_FWD, 'interior': EVALS_INTERIOR, 'edge': EVALS_BOUNDARY}[p] for p in passes)
        flops = flops_per_eval * evals * N_SAMPLES
        top_ms = kernels[top]['ms_per_step']
        achieved = flops / (top_ms * 1e-3) / 1e12
        peak = fp64_peak if bound == 'fp64' else fp32_peak
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        except Exception:
            pass
        # compulsory DRAM traffic of the whole step: queues + result words + images + scene tables
        alg_bytes = W * H * 4 * 4 * 2 + W * H * 4 + 2 * params_np.nbytes * 8
        roofline = {'bound': bound, 'kernel': top, 'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s',
                    'frac': achieved / peak if peak else None,
                    'peak_source': 'FMA-chain probes measured live in this run (MEASURED_PEAKS.json has no CUDA-core figure): '
                                   'FP32 %.2f, FP64 %.2f TFLOP/s' % (fp32_peak, fp64_peak),
                    'kernel_ms_per_step': top_ms, 'kernel_launches_per_step': kernels[top]['launches'] / args.steps,
                    'kernel_share_of_step': top_ms / total_k,
                    'algorithmic_flops_per_step': flops, 'model': 'SURVEY 8d event model: %.0f flop per colour evaluation '
                    'for this kernel (%s), %.3f evaluations per pixel sample' % (flops_per_eval, '+'.join(passes), evals),
                    'step_achieved_tflops': FALG_PER_EVAL * (EVALS_FWD + EVALS_INTERIOR + EVALS_BOUNDARY) * N_SAMPLES / (ms_per_step * 1e-3) / 1e12,
                    'hbm': {'algorithmic_bytes': alg_bytes, 'achieved_gbs': alg_bytes / (ms_per_step * 1e-3) / 1e9,
                            'peak_gbs': peaks.get('hbm_gbs', 6650.0),
                            'peak_source': 'measured' if 'hbm_gbs' in peaks else 'fallback'},
                    'traffic': None}
        try:   # DRAM bytes of this kernel's longest launch, from the committed ncu --set full capture
            tr = json.load(open(os.path.join(ROOT, 'profiles', 'r1_ncu_traffic.json')))['kernels'].get(top)
            if tr:
                roofline['traffic'] = tr['dram_bytes']
                roofline['traffic_note'] = 'dram__bytes_read+write of the longest launch of %s (%.3f ms) in profiles/r1_ncu_traffic.json' % (top, tr['ms'])
        except Exception:
            pass

    # ---- CPU baseline (rank 0, N = 1): the reference's CPU path on a bounded sample
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sys.path.insert(0, os.path.join(ROOT, 'oracle'))
        import oracle_check
        try:
            if oracle_check.kind() != 'reference':
                raise oracle_check.OracleUnavailable('oracle/_ref is not built here')
            t0 = time.perf_counter()
            ref_img = oracle_check.render(topo, params_np, W, H, NSX, NSY, 0)['image']
            d_img_np = (2.0 * (ref_img - target.cpu().numpy()) / ref_img.size).astype(np.float32)
            oracle_check.render(topo, params_np, W, H, NSX, NSY, 0, d_render_image=d_img_np)
            t_cpu = time.perf_counter() - t0
            cpu_baseline = {'value': 1.0 / t_cpu, 'unit': UNIT, 'cores': os.cpu_count(), 'kind': oracle_check.kind(),
                            'sample': '1 full fwd+bwd iteration of the same workload (seed 0), all host threads, '
                                      'Scene rebuilt per call as the reference does'}
        except oracle_check.OracleUnavailable as e:
            cpu_baseline = {'value': None, 'unit': UNIT, 'cores': os.cpu_count(), 'kind': 'port', 'sample': 'unavailable: ' + str(e)}

    if rank == 0:
        line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': n_gpus, 'steps': args.steps,
                'warmup': args.warmup, 'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak',
                'vs_baseline': None, 'dtype': 'f32+f64', 'data': 'synthetic', 'config': config_dict(n_gpus),
                'msamples_per_s': world * 2 * N_SAMPLES / (ms_per_step * 1e-3) / 1e6,
                'wall_ms_per_step': 1e3 * wall / args.steps,
                'e2e': {'value': e2e_value, 'unit': UNIT, 'ms_per_step': e2e_ms,
                        'h2d_bytes_per_step': e2e_bytes.get('h2d', 0), 'd2h_bytes_per_step': e2e_bytes.get('d2h', 0)},
                'gpu_launches': launches, 'clocks': clocks, 'roofline': roofline, 'cpu_baseline': cpu_baseline,
                'kernels': kernels}
        print(json.dumps(line), flush=True)
    n.lib.dvg_scene_destroy(h)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
