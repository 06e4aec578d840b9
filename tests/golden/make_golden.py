"""Generates tests/golden/*.npz from the UNMODIFIED reference renderer (oracle/_ref, built from
/root/reference by oracle/Makefile).  Run in the build container only:

    make -C oracle ref && python tests/golden/make_golden.py

Each fixture holds the packed scene (topo, params), the render configuration and the reference's
outputs: forward image and, for a fixed pseudo-random d_render_image, the flat parameter gradient
(+ d_background where a background is used).  Forward images of the reference are run-to-run
deterministic for box filters; gradients are reproducible to ~1e-6 rel-L2 (atomic order).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)

import ref_oracle  # noqa: E402
import scenes  # noqa: E402
import util  # noqa: E402

# name -> (scene, W, H, nsx, nsy, seed, filter_type, filter_radius, use_background)
CASES = {
    'circle':        (lambda: scenes.single_circle(), 64, 64, 2, 2, 0, 0, 0.5, False),
    'stroke':        (lambda: scenes.single_stroke(), 64, 64, 2, 2, 0, 0, 0.5, False),
    'stroke_thick':  (lambda: scenes.single_stroke([10., 5., 4., 20.], fill=False), 64, 64, 2, 2, 1, 0, 0.5, False),
    'zoo':           (lambda: scenes.zoo(), 96, 96, 2, 2, 3, 0, 0.5, False),
    'zoo_bg_1spp':   (lambda: scenes.zoo(), 64, 64, 1, 1, 3, 0, 0.5, True),
    'zoo_tent':      (lambda: scenes.zoo(), 80, 48, 3, 3, 5, 1, 1.5, False),
    'zoo_hann':      (lambda: scenes.zoo(), 48, 48, 2, 2, 9, 3, 2.0, True),
    'painterly64':   (lambda: scenes.painterly(64, 128), 64, 64, 4, 4, 0, 0, 0.5, False),
    'blobs48':       (lambda: scenes.blobs(48, 128), 64, 64, 2, 2, 0, 0, 0.5, False),
    'batched0':      (lambda: scenes.batched_strokes(0), 64, 64, 2, 2, 0, 0, 0.5, False),
}


def d_image_for(name, H, W):
    seed = sum(ord(c) for c in name)
    return (np.random.RandomState(seed).rand(H, W, 4).astype(np.float32) - 0.5)


def background_for(name, H, W):
    seed = 17 + sum(ord(c) for c in name)
    return np.random.RandomState(seed).rand(H, W, 4).astype(np.float32)


# Other render modes of the path: SDF prefiltering ('pf'), SDF output ('sdf', with or without
# eval_positions) and the per-pixel translation gradient of render_grad ('dtrans'), stored under
# tests/golden_modes/.  name -> (mode, scene, W, H, nsx, nsy, seed, filter_type, filter_radius, use_background)
MODE_CASES = {
    'pf_stroke':     ('pf', lambda: scenes.single_stroke(), 64, 64, 1, 1, 0, 0, 0.5, False),
    'pf_zoo':        ('pf', lambda: scenes.zoo_prefilter(), 96, 96, 2, 2, 5, 0, 0.5, False),
    'pf_zoo_bg':     ('pf', lambda: scenes.zoo_prefilter(), 80, 48, 1, 1, 2, 0, 0.5, True),
    'pf_zoo_hann':   ('pf', lambda: scenes.zoo_prefilter(), 64, 64, 2, 2, 5, 3, 1.5, False),
    'pf_blobs':      ('pf', lambda: scenes.blobs(48, 128), 64, 64, 2, 2, 1, 0, 0.5, False),
    'sdf_stroke':    ('sdf', lambda: scenes.single_stroke(), 48, 48, 2, 2, 0, 0, 0.5, False),
    'sdf_zoo':       ('sdf', lambda: scenes.zoo_prefilter(), 64, 64, 2, 2, 5, 0, 0.5, False),
    'sdf_zoo_eval':  ('sdf_eval', lambda: scenes.zoo_prefilter(), 128, 128, 1, 1, 0, 0, 0.5, False),
    'sdf_circle':    ('sdf', lambda: scenes.single_circle(), 32, 32, 1, 1, 0, 0, 0.5, False),
    'dtrans_zoo':    ('dtrans', lambda: scenes.zoo(), 64, 64, 2, 2, 3, 0, 0.5, False),
}
MODES_DIR = os.path.join(os.path.dirname(HERE), 'golden_modes')


def eval_positions_for(name, H, W, n=300):
    seed = 5 + sum(ord(c) for c in name)
    return (np.random.RandomState(seed).rand(n, 2) * np.asarray([W, H])).astype(np.float32)


def d_sdf_for(name, shape):
    seed = 9 + sum(ord(c) for c in name)
    return (np.random.RandomState(seed).rand(*shape).astype(np.float32) - 0.5)


def run_mode(render, name, mode, topo, params, W, H, nsx, nsy, seed, use_bg):
    """One mode case through `render` (ref_oracle.render, or an emulation / GPU front end with the same
    keyword surface).  Returns the dict of arrays a fixture stores."""
    bg = background_for(name, H, W) if use_bg else None
    out = {}
    if mode == 'pf':
        out['image'] = render(topo, params, W, H, nsx, nsy, seed, background=bg, use_prefiltering=True)['image']
        b = render(topo, params, W, H, nsx, nsy, seed, background=bg, use_prefiltering=True,
                   d_render_image=d_image_for(name, H, W), want_d_translation=True)
        out['d_params'] = b['d_params']
        out['d_translation'] = b['d_translation']
        if use_bg:
            out['d_background'] = b['d_background']
    elif mode in ('sdf', 'sdf_eval'):
        ep = eval_positions_for(name, H, W) if mode == 'sdf_eval' else None
        out['sdf'] = render(topo, params, W, H, nsx, nsy, seed, want_image=False, want_sdf=True, eval_positions=ep)['sdf']
        b = render(topo, params, W, H, nsx, nsy, seed, eval_positions=ep, d_render_sdf=d_sdf_for(name, out['sdf'].shape),
                   want_d_translation=True)
        out['d_params'] = b['d_params']
        out['d_translation'] = b['d_translation']
    else:  # dtrans: plain colour render, boundary term writes the translation gradient
        b = render(topo, params, W, H, nsx, nsy, seed, d_render_image=d_image_for(name, H, W), want_d_translation=True)
        out['d_params'] = b['d_params']
        out['d_translation'] = b['d_translation']
    return out


def main_modes():
    os.makedirs(MODES_DIR, exist_ok=True)
    for name, (mode, mk, W, H, nsx, nsy, seed, ft, fr, use_bg) in MODE_CASES.items():
        topo, params = util.pack(mk(), ft, fr)
        out = run_mode(ref_oracle.render, name, mode, topo, params, W, H, nsx, nsy, seed, use_bg)
        out.update(topo=topo, params=params, config=np.asarray([W, H, nsx, nsy, seed, ft], np.int64), filter_radius=np.float32(fr))
        np.savez_compressed(os.path.join(MODES_DIR, name + '.npz'), **out)
        first = out.get('image', out.get('sdf', out['d_translation']))
        print('%-14s %-8s sum %.6f |d_params| %.6g' % (name, mode, first.astype(np.float64).sum(),
                                                       np.linalg.norm(out['d_params'].astype(np.float64))))


def main():
    for name, (mk, W, H, nsx, nsy, seed, ft, fr, use_bg) in CASES.items():
        topo, params = util.pack(mk(), ft, fr)
        bg = background_for(name, H, W) if use_bg else None
        d_img = d_image_for(name, H, W)
        fwd = ref_oracle.render(topo, params, W, H, nsx, nsy, seed, background=bg)
        bwd = ref_oracle.render(topo, params, W, H, nsx, nsy, seed, background=bg, d_render_image=d_img)
        out = dict(topo=topo, params=params, config=np.asarray([W, H, nsx, nsy, seed, ft], np.int64),
                   filter_radius=np.float32(fr), image=fwd['image'], d_params=bwd['d_params'])
        if use_bg:
            out['d_background'] = bwd['d_background']
        np.savez_compressed(os.path.join(HERE, name + '.npz'), **out)
        print('%-14s image sum %.6f |d_params| %.6g' % (name, fwd['image'].astype(np.float64).sum(),
                                                        np.linalg.norm(bwd['d_params'].astype(np.float64))))


if __name__ == '__main__':
    if len(sys.argv) < 2 or sys.argv[1] != 'modes':
        main()
    if len(sys.argv) < 2 or sys.argv[1] == 'modes':
        main_modes()
