"""Tuning sweep of the intersection kernel on config E (forward only): pixels/thread x chunk x filter mode.
Prints the kernel's CUDA-event time (measured inside the library) and its fraction of FP32-FMA peak."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch                     # noqa: E402
import scene_io                  # noqa: E402
import surf_renderer_b200        # noqa: E402
from surf_renderer_b200 import scenes as synth   # noqa: E402
from surf_renderer_b200._lib import lib          # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--splats', type=int, default=100000)
ap.add_argument('--size', type=int, default=1024)
ap.add_argument('--reps', type=int, default=3)
ap.add_argument('--ppt', default='4,8,16')
ap.add_argument('--chunk', default='512,1024,2048')
ap.add_argument('--mode', default='0')
ap.add_argument('--instr', type=float, default=10.0, help='FMA-pipe lane-instr per test used for the peak fraction')
args = ap.parse_args()

L = lib()
print('fma peak lane-instr/s: scalar %.4e  packed %.4e  (theoretical %.4e)' % (
    L.surf_fma_peak(0, 8192, None), L.surf_fma_peak(1, 8192, None), 148 * 128 * 1.965e9))
scene = scene_io.clone_scene(synth.config_e(m=args.splats, width=args.size, height=args.size), device='cuda')
tests = float(args.splats) * args.size * args.size
L.surf_set_kernel_timing(1)
ref = None
rows = []
for mode in [int(x) for x in args.mode.split(',')]:
    for ppt in [int(x) for x in args.ppt.split(',')]:
        for chunk in [int(x) for x in args.chunk.split(',')]:
            best = 1e9
            try:
                for _ in range(args.reps):
                    with torch.no_grad():
                        res = surf_renderer_b200.render(scene, _pixels_per_thread=ppt, _chunk_prims=chunk, _math_mode=mode)
                    best = min(best, L.surf_last_kernel_ms(0))
            except ValueError as e:
                print('mode %d ppt %d chunk %d: skipped (%s)' % (mode, ppt, chunk, e))
                continue
            if ref is None:
                ref = res['nearest'].clone()
            same = bool(torch.equal(ref, res['nearest']))
            frac = tests * args.instr / (best * 1e-3) / (148 * 128 * 1.965e9)
            rows.append({'mode': mode, 'ppt': ppt, 'chunk': chunk, 'ms': best, 'frac_fp32_peak': frac, 'same_nearest': same})
            print('mode %d ppt %d chunk %4d : %8.3f ms  %.1f%% of FP32 FMA peak  %.3e tests/s  same=%s' % (
                mode, ppt, chunk, best, 100 * frac, tests / (best * 1e-3), same), flush=True)
os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
json.dump(rows, open(os.path.join(ROOT, 'gpurun_out', 'sweep_intersect.json'), 'w'), indent=1)
