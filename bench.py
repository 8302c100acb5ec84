#!/usr/bin/env python
"""bench.py - headline benchmark of the render hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W                 # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...       # the reference's own CPU implementation, host cores
    python bench.py --workload config_d ...                       # BASELINE configs[3]: 64 scenes x 5000 splats at 128x128

Workload config_e (default; config.workload = "config_e"): BASELINE.json configs[4], the configuration the north-star
target is quoted on - the inverse-rendering step of 100 000 synthetic disk splats at 1024x1024.
One step = render (forward) -> mean((image-target)^2) -> backward to splat positions, normals, albedo and light
positions -> Adam step (the loop body of the reference's test_optimization.py:100-125).

  value   ray-primitive tests per second of the whole job = H*W*M per step / step time, inputs resident in HBM,
          through surf_renderer_b200.MSEStep (one surf_step_mse call per step) + torch's fused Adam.
  e2e     the same render -> loss -> backward through the C ABI with HOST buffers (surf_step_host_begin / _end): every
          step copies the scene and the target from pinned host memory, and the gradients and the loss back to host
          memory, inside the timed region.  The optimizer of a host-side caller runs on the host and is NOT part of
          this leg (config.e2e_step says so).
  N > 1   the frame is sharded into row bands, one per GPU (strong scaling of one frame): no collective in the
          intersection data path, band-local loss, ONE in-place NCCL all-reduce over the packed [gradients | loss] buffer.

Workload config_d: a stacked batch of 64 scenes (5000 splats each, own camera) at 128x128, scenes sharded over the
ranks; step = forward, all-gather of the images, a weighted-sum loss on the gathered batch (stand-in for the GAN
discriminator the reference feeds these frames to, gan.py:326-377), backward; gradients of the light rig shared by all
scenes all-reduced.
"""
from __future__ import annotations

import argparse
import ctypes as C
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FMA_INSTR_PER_DISK_TEST = 10      # n.d 3, t 1, rel 3, |rel|^2 3 (FFMA/FMUL lane-instructions) - the plane filter
FMA_INSTR_DENSE_TEST = 4          # k_intersect_batch (stacked batches, frames <= 256x256): (oc . d)^2 - c0 >= 0 from the staged plane records
FMA_INSTR_SPHERE_TEST = 3         # default path of large disk frames: |oc' . d| >= 1 (k_filter_const<8, 1>), 3 FFMA2/FMUL2 lane-instr
# math_mode -> (kernel name, FMA-pipe lane-instr per ray-disk test, description)
INTERSECT_MODES = {
    0: ('k_filter_const', FMA_INSTR_SPHERE_TEST, 'bounding-sphere filter through the constant bank / uniform registers (3 lane-instr per test), '
        'plane filter + exact test on the candidates in k_narrow_queue'),
    6: ('k_filter_const', FMA_INSTR_PER_DISK_TEST, 'plane filter (SURVEY 8(d) formulation, 10 lane-instr per test + 1 MUFU.RCP) through the constant bank / uniform registers'),
    5: ('k_intersect', FMA_INSTR_PER_DISK_TEST, 'plane filter (10 lane-instr per test), records staged in shared memory by TMA'),
}
N_SM, FP32_LANES = 148, 128
L2_FLUSH_BYTES = 144 * 1024 * 1024     # > the 126 MB L2


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='surf', choices=['surf', 'reference'])
    ap.add_argument('--workload', default='config_e', choices=['config_e', 'config_d'])
    ap.add_argument('--splats', type=int, default=None)
    ap.add_argument('--size', type=int, default=None)
    ap.add_argument('--ppt', type=int, default=0, help='pixels per thread of the intersection kernel (0 = default)')
    ap.add_argument('--chunk', type=int, default=0)
    ap.add_argument('--math', type=int, default=0, help='intersection kernel: 0 = default (sphere + plane filter through the constant bank), 6 = plane filter through the constant bank, '
                    '5 = plane filter staged by TMA (round 1 kernel), 3 = screen-space mode')
    ap.add_argument('--graph', action='store_true', help='replay the step from a CUDA graph')
    ap.add_argument('--torch-adam', action='store_true', help="torch.optim.Adam(fused=True) instead of the library's packed Adam kernel")
    ap.add_argument('--ref-size', type=int, default=0, help='reference arm: viewport edge of the bounded sample (0 = auto)')
    ap.add_argument('--ref-budget', type=float, default=150.0, help='reference arm: seconds of CPU work')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-fast', action='store_true')
    return ap.parse_args()


def measured_peaks():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            return json.load(f), 'measured'
    except Exception:
        return {'hbm_gbs': 6650.0, 'sm_max_mhz': 1965.0}, 'fallback'


def source_hash():
    """sha256 over the library sources: ties an ncu capture under profiles/ to the build it was taken on"""
    h = hashlib.sha256()
    csrc = os.path.join(ROOT, 'surf_renderer_b200', 'csrc')
    files = sorted(f for f in os.listdir(csrc) if f.endswith(('.cu', '.cuh', '.h')))
    for f in files:
        with open(os.path.join(csrc, f), 'rb') as fh:
            h.update(f.encode() + b'\0' + fh.read())
    with open(os.path.join(ROOT, 'include', 'surf_b200.h'), 'rb') as fh:
        h.update(fh.read())
    return h.hexdigest()[:16]


def ncu_traffic(kernel, workload, world):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the ncu --set full capture recorded in
    profiles/ncu_traffic.json - only when that capture was taken on THIS build (same source hash), workload and N."""
    try:
        with open(os.path.join(ROOT, 'profiles', 'ncu_traffic.json')) as f:
            rec = json.load(f)
        if rec.get('source_hash') != source_hash() or rec.get('workload') != workload or rec.get('n_gpus', 1) != world:
            return None
        return rec['kernels'].get(kernel)
    except Exception:
        return None


# ----------------------------------------------------------------------------------------------------------
# clocks sampling during the timed region
# ----------------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU during the timed region.  NVML in-process (cheap calls from a
    thread); falls back to an `nvidia-smi -lms 200` child when pynvml is unavailable."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index, period_s=0.05):
        self.index, self.period = index, period_s
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        self.proc, self.rows = None, []
        self.mode = None

    def _nvml_loop(self):
        import pynvml as nv
        h = self.handle
        bits = {'hw_slowdown': nv.nvmlClocksThrottleReasonHwSlowdown,
                'hw_thermal_slowdown': nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                'sw_thermal_slowdown': nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                'sw_power_cap': nv.nvmlClocksThrottleReasonSwPowerCap}
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for name, bit in bits.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(self.period)

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            # NVML indexes physical devices; honour CUDA_VISIBLE_DEVICES when it is a plain index list
            vis = os.environ.get('CUDA_VISIBLE_DEVICES')
            phys = self.index
            if vis:
                try:
                    phys = int(vis.split(',')[self.index])
                except Exception:
                    phys = self.index
            self.handle = nv.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(self.handle, nv.NVML_CLOCK_SM))
            self.mode = 'nvml'
            self._thread = threading.Thread(target=self._nvml_loop, daemon=True)
            self._thread.start()
            return
        except Exception:
            pass
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '200'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.mode = 'nvidia-smi'
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        import numpy as np
        if self.mode == 'nvml':
            self._stop.set()
            self._thread.join(timeout=1.0)
            return {'sm_mhz': float(np.median(self.samples)) if self.samples else None, 'sm_max_mhz': self.max_mhz,
                    'reasons': sorted(self.reasons), 'samples': len(self.samples), 'source': 'nvml'}
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['clock sampling unavailable']}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith('active'):
                        reasons.add(nm)
            except Exception:
                pass
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm), 'source': 'nvidia-smi'}


# ----------------------------------------------------------------------------------------------------------
# workload descriptions shared by both arms
# ----------------------------------------------------------------------------------------------------------
def workload_config(args):
    if args.workload == 'config_e':
        M, S = args.splats or 100_000, args.size or 1024
        return {'workload': 'config_e', 'splats': M, 'width': S, 'height': S, 'lights': 3,
                'step': 'render fwd + mse(image,target) + bwd(pos,normal,albedo,light_pos) + Adam (%s)' % ('torch fused' if args.torch_adam else 'surf_adam_step'),
                'e2e_step': 'H2D(scene, target) + render fwd + mse + bwd + grad all-reduce + D2H(gradients, loss); the host-side optimizer is not part of this leg',
                'sharding': 'row-bands x%d' % max(1, args.gpus),
                'l2': 'flushed between timed steps (%d MiB fill enqueued on the stream, inside the timed region)' % (L2_FLUSH_BYTES >> 20)}, float(M) * S * S
    M, S, B = args.splats or 5000, args.size or 128, 64
    return {'workload': 'config_d', 'scenes': B, 'splats': M, 'width': S, 'height': S, 'lights': 7,
            'step': 'render_batch fwd (stacked scenes) + all-gather(image) + weighted-sum loss + bwd(pos,normal,light_pos) + all-reduce(shared light grads)',
            'e2e_step': 'H2D(splat positions, normals, camera eyes of the batch) + the same step + D2H(loss)',
            'sharding': 'scene-blocks x%d' % max(1, args.gpus),
            'l2': 'flushed between timed steps (%d MiB fill enqueued on the stream, inside the timed region)' % (L2_FLUSH_BYTES >> 20)}, float(B) * M * S * S


# ----------------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation (oracle/_ref = the unmodified files; else the oracle port)
# ----------------------------------------------------------------------------------------------------------
def run_reference_arm(args):
    os.environ['CUDA_VISIBLE_DEVICES'] = ''          # the reference picks CUDA tensors whenever torch sees a GPU
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return 0
    import numpy as np
    import torch
    from oracle import ref_runner, torch_oracle
    from surf_renderer_b200 import scenes as synth
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    config, tests_per_step = workload_config(args)
    if ref_runner.available():
        render, _ = ref_runner.load()
        kind = 'reference'
        conv = ref_runner.scene_for_reference
        what = 'UNMODIFIED reference diffrend.torch.renderer.render (oracle/_ref) on CPU tensors'
    else:
        render, kind, conv = torch_oracle.render, 'port', synth.clone_scene
        what = 'torch-CPU op-for-op port of the reference (oracle/torch_oracle.py; oracle/_ref absent)'

    if args.workload == 'config_e':
        M = config['splats']
        full = synth.config_e(m=M, width=config['width'], height=config['height'])

        def make(size):
            sc = synth.config_e(m=M, width=size, height=size)
            tg = conv(synth.config_e_target_scene(sc))
            with torch.no_grad():
                target = render(tg, tile_size=512)['image']
            return sc, target

        def one_step(sc, target):
            s = conv(sc)
            leaves = [s['objects']['disk']['pos'], s['objects']['disk']['normal'], s['materials']['albedo'], s['lights']['pos']]
            for t in leaves:
                t.requires_grad_(True)
            t0 = time.perf_counter()
            res = render(s, tile_size=512)
            loss = ((res['image'] - target) ** 2).mean()
            t1 = time.perf_counter()
            loss.backward()
            return t1 - t0, time.perf_counter() - t1

        # the same scene (all M splats) on a reduced viewport of the same camera: the reference materialises [M, N]
        # tensors and keeps them for autograd, so the full 1024x1024 frame needs ~2 h and > 1 TB on the CPU
        size = args.ref_size
        if not size:
            sc, tg = make(16)
            tf, tb = one_step(sc, tg)
            per_px = (tf + tb) / 256.0
            want = args.ref_budget / max(1, args.steps + min(args.warmup, 1)) / max(per_px, 1e-9)
            size = int(max(16, min(48, np.sqrt(want) // 8 * 8)))
        sc, tg = make(size)
        n_px = size * size
        sample = '%dx%d viewport of the same camera x all %d splats, fwd + mse + bwd' % (size, size, M)
        unit_tests = float(M) * n_px
        del full
    else:
        B = config['scenes']
        scenes = [synth.config_d_scene(i, m=config['splats'], width=config['width'], height=config['height']) for i in range(2)]
        g = torch.Generator().manual_seed(1)
        ws = [torch.rand(config['height'], config['width'], 3, generator=g) for _ in scenes]

        def one_step(scs, _):
            tf = tb = 0.0
            for sc, w in zip(scs, ws):
                s = conv(sc)
                for t in (s['objects']['disk']['pos'], s['objects']['disk']['normal'], s['lights']['pos']):
                    t.requires_grad_(True)
                t0 = time.perf_counter()
                res = render(s, tile_size=512, double_sided=True)
                loss = (res['image'] * w).sum()
                t1 = time.perf_counter()
                loss.backward()
                tf += t1 - t0
                tb += time.perf_counter() - t1
            return tf, tb
        sc, tg = scenes, None
        sample = '2 of the %d scenes (%d splats at %dx%d each), fwd + weighted-sum loss + bwd, one render() per scene like gan.py:326-377' % (
            B, config['splats'], config['width'], config['height'])
        unit_tests = 2.0 * config['splats'] * config['width'] * config['height']

    times = []
    t_start = time.perf_counter()
    for i in range(min(args.warmup, 1) + max(1, args.steps)):
        tf, tb = one_step(sc, tg)
        if i >= min(args.warmup, 1):
            times.append((tf, tb))
        if time.perf_counter() - t_start > args.ref_budget and times:
            break
    tf = float(np.mean([t[0] for t in times])); tb = float(np.mean([t[1] for t in times]))
    value = unit_tests / (tf + tb)
    cb = {'value': value, 'unit': 'tests/s', 'cores': threads, 'kind': kind,
          'sample': '%s; %s; %d timed step(s), fwd %.2fs + bwd %.2fs each, tile_size=512, %d torch threads' % (sample, what, len(times), tf, tb, threads),
          'fwd_tests_per_s': unit_tests / tf, 'ms_per_sample_step': 1e3 * (tf + tb)}
    line = {'impl': 'reference', 'metric': 'ray-primitive tests/s (fwd+bwd inverse-rendering step)' if args.workload == 'config_e'
            else 'ray-primitive tests/s (fwd+bwd over a batch of scenes)',
            'value': value, 'unit': 'tests/s', 'n_gpus': 0, 'steps': len(times), 'warmup': min(args.warmup, 1),
            'ms_per_step': 1e3 * (tf + tb), 'ms_per_step_is': 'one bounded SAMPLE step (see cpu_baseline.sample), not a full-size step',
            'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': config, 'cpu_baseline': cb,
            'full_step_s_extrapolated': tests_per_step / value,
            'e2e': {'value': value, 'unit': 'tests/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    if args.workload == 'config_e':
        try:
            line['cpu_baseline_numpy'] = run_numpy_baseline(synth.config_e(m=config['splats'], width=config['width'], height=config['height']))
        except Exception as e:      # the numpy twin is a side number; never fail the arm on it
            line['cpu_baseline_numpy'] = {'error': repr(e)}
    print(json.dumps(line))
    return 0


def run_numpy_baseline(scene, n_pix=96, reps=2):
    """The reference's second CPU renderer (diffrend/numpy/renderer.py, restated in oracle/numpy_oracle.py): forward
    only (it has no gradients), float64, Lambertian, and it materialises the whole [M, N, 4] tensor - so the sample is
    a small pixel subset of the same frame and the same splats."""
    import torch
    from oracle import numpy_oracle
    hs = numpy_oracle.homogeneous_scene(scene)
    m = int(hs['objects']['disk']['pos'].shape[0])
    vp = scene['camera']['viewport']
    n = (vp[2] - vp[0]) * (vp[3] - vp[1])
    g = torch.Generator().manual_seed(123)
    sub = torch.randperm(n, generator=g)[:n_pix].sort().values.numpy()
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter()
        numpy_oracle.render(hs, pixel_subset=sub)
        best = min(best, time.perf_counter() - t0)
    return {'value': float(m) * len(sub) / best, 'unit': 'tests/s (forward only)', 'cores': 1, 'kind': 'port',
            'sample': '%d random pixels of the frame x %d splats, forward, best of %d (%.2fs); numpy float64 port of '
                      'the reference numpy twin (no tiling: [M,N,4] temporaries; elementwise numpy is single-threaded)'
                      % (len(sub), m, reps, best)}


def cpu_baseline_subprocess(args, budget_s):
    """The reference arm in a child process with the GPUs hidden (the reference chooses its device at import)."""
    cmd = [sys.executable, os.path.abspath(__file__), '--impl', 'reference', '--workload', args.workload, '--steps', '3',
           '--warmup', '1', '--ref-budget', str(budget_s)]
    if args.splats:
        cmd += ['--splats', str(args.splats)]
    if args.size:
        cmd += ['--size', str(args.size)]
    env = dict(os.environ, CUDA_VISIBLE_DEVICES='')
    for k in ('RANK', 'WORLD_SIZE', 'LOCAL_RANK', 'MASTER_ADDR', 'MASTER_PORT'):
        env.pop(k, None)
    try:
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=budget_s * 4 + 240, env=env)
        for ln in reversed(out.stdout.strip().splitlines()):
            if ln.startswith('{'):
                rec = json.loads(ln)
                return rec.get('cpu_baseline'), rec.get('cpu_baseline_numpy')
        return {'error': (out.stderr or out.stdout)[-400:]}, None
    except Exception as e:
        return {'error': repr(e)}, None


# ----------------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------------
def timed_region(step, steps, warmup, flush, barrier, world, dev, sampler_rank0):
    """W warm-up steps, then exactly K steps between barriers, timed with CUDA events on the current stream, max over
    ranks.  The L2 is flushed between steps by a fill larger than the L2, enqueued on the same stream."""
    import torch
    import torch.distributed as dist
    for _ in range(max(3, warmup)):
        step()
    barrier()
    if sampler_rank0 is not None:
        sampler_rank0.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0 = time.perf_counter()
    ev0.record()
    out = None
    for i in range(steps):
        flush.fill_(i & 0xff)
        out = step()
    ev1.record()
    barrier()
    wall = time.perf_counter() - t0
    total_ms = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([total_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    clocks = sampler_rank0.stop() if sampler_rank0 is not None else None
    return total_ms / steps, wall, clocks, out


class _DevBlock:
    """a raw device allocation of the C library as a CUDA-array-interface object (for torch.as_tensor)"""

    def __init__(self, ptr, n_floats):
        self.__cuda_array_interface__ = {'shape': (int(n_floats),), 'typestr': '<f4', 'data': (int(ptr), False), 'version': 2}


def run_config_e(args, rank, world, local_rank):
    import numpy as np   # noqa: F401
    import torch
    import torch.distributed as dist
    import surf_renderer_b200
    from surf_renderer_b200 import scenes as synth
    from surf_renderer_b200.scenes import clone_scene
    from surf_renderer_b200 import dist as sdist
    from surf_renderer_b200._lib import check, lib
    from surf_renderer_b200.marshal import Marshalled, make_options
    dev = torch.device('cuda', local_rank)
    config, tests_per_step = workload_config(args)
    M, H, W = config['splats'], config['height'], config['width']
    scene = synth.config_e(m=M, width=W, height=H)
    target_scene = synth.config_e_target_scene(scene)
    peaks, peak_src = measured_peaks()
    params = {'_pixels_per_thread': args.ppt, '_chunk_prims': args.chunk, '_math_mode': args.math}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.no_grad():
        tgt = surf_renderer_b200.render(clone_scene(target_scene, device=dev), **params)['image']
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)

    def build(step_params):
        sc = clone_scene(scene, device=dev)
        leaves = [sc['objects']['disk']['pos'], sc['objects']['disk']['normal'], sc['materials']['albedo'], sc['lights']['pos']]
        for t in leaves:
            t.requires_grad_(True)
        plan = surf_renderer_b200.MSEStep(sc, tgt, group=(True if world > 1 else None), **step_params)
        if args.torch_adam:
            opt = torch.optim.Adam(leaves, lr=1e-4, fused=True, capturable=args.graph)
        else:
            opt = surf_renderer_b200.PackedAdam(plan, lr=1e-4)          # one kernel over the packed gradient buffer

        def step():
            loss = plan()
            opt.step()
            return loss
        if args.graph:
            graphed = surf_renderer_b200.GraphedStep(step, warmup=3, capture_error_mode='thread_local' if world > 1 else 'global')
            return plan, graphed
        return plan, step

    plan, step = build(params)
    lib().surf_set_kernel_timing(0 if args.graph else 1)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    for _ in range(2):
        step()
    barrier()
    lib().surf_set_kernel_timing(0 if args.graph else 1)      # reset the library's per-kernel event ring
    ms_per_step, wall, clocks, loss = timed_region(step, args.steps, args.warmup, flush, barrier, world, dev, sampler)
    value = tests_per_step / (ms_per_step * 1e-3)
    n_timed = C.c_int32()
    k_mean = {k: lib().surf_mean_kernel_ms(k, C.byref(n_timed)) for k in (0, 1, 2)}
    lib().surf_set_kernel_timing(0)

    # ---- roofline of the dominant kernel (k_intersect), timed live with CUDA events inside the library
    isect_ms = k_mean[0] if k_mean[0] > 0 else None
    f_clk = float(peaks.get('sm_max_mhz', 1965.0)) * 1e6
    peak_lane = N_SM * FP32_LANES * f_clk                    # lane-instr/s
    roofline = None
    if isect_ms:
        tests_launch = tests_per_step / world
        kname, instr_per_test, kdesc = INTERSECT_MODES.get(args.math, ('k_intersect', FMA_INSTR_PER_DISK_TEST, 'math_mode %d' % args.math))
        ach_lane = tests_launch * instr_per_test / (isect_ms * 1e-3)
        fma_meas = max(lib().surf_fma_peak(0, 8192, None), lib().surf_fma_peak(1, 8192, None))
        roofline = {'bound': 'fp32_fma', 'kernel': kname, 'achieved': ach_lane * 2 / 1e12, 'peak': peak_lane * 2 / 1e12,
                    'unit': 'TFLOP/s', 'frac': ach_lane / peak_lane, 'formulation': kdesc,
                    'timed': 'the whole intersection stage of a step on the launching stream (CUDA events inside the library): for '
                             'k_filter_const that is every launch of the kernel over the constant-bank loads of the frame plus '
                             'k_sphere_records, k_narrow_queue, k_const_fallback, k_inside_disks',
                    # dram bytes of the dominant kernel per launch from the ncu --set full capture of THIS build, else null
                    'traffic': ncu_traffic(kname, 'config_e', world), 'source_hash': source_hash(),
                    'peak_source': 'theoretical 148 SM x 128 lanes x sm_max_mhz (%s MEASURED_PEAKS.json has no fp32 entry)' % peak_src,
                    'peak_measured': fma_meas * 2 / 1e12, 'frac_of_measured': ach_lane / fma_meas if fma_meas > 0 else None,
                    'algorithmic': '%d FMA-pipe lane-instr per ray-disk test x %.3g tests per step and rank' % (instr_per_test, tests_launch),
                    'kernel_ms': isect_ms, 'kernel_share_of_step': isect_ms / ms_per_step,
                    'step_minus_kernel_ms': ms_per_step - isect_ms,
                    'shade_ms': k_mean[1], 'backward_ms': k_mean[2], 'launches_timed': int(n_timed.value)}
        # the two HBM-side kernels (north star: achieved GB/s against the measured copy bandwidth)
        n_loc = H * W / world
        hbm = float(peaks.get('hbm_gbs', 6650.0))
        # k_shade in the step: rays 12 + key 8 + target 12 read; image 12 + depth 4 + nearest 8 + d(image) 12 written
        shade_bytes = n_loc * (12 + 8 + 12 + 12 + 4 + 8 + 12)
        # k_backward: rays 12, nearest 8, depth 4, d(image) 12 read (+ ~24 B per hit pixel of atomics, not counted)
        bwd_bytes = n_loc * (12 + 8 + 4 + 12)
        roofline['hbm_kernels'] = {
            'peak_gbs': hbm, 'peak_source': peak_src + ' MEASURED_PEAKS.json hbm_gbs',
            'k_shade': {'ms': k_mean[1], 'algorithmic_bytes': shade_bytes, 'achieved_gbs': shade_bytes / (k_mean[1] * 1e-3) / 1e9,
                        'frac': shade_bytes / (k_mean[1] * 1e-3) / 1e9 / hbm, 'traffic': ncu_traffic('k_shade', 'config_e', world)} if k_mean[1] > 0 else None,
            'k_backward': {'ms': k_mean[2], 'algorithmic_bytes': bwd_bytes, 'achieved_gbs': bwd_bytes / (k_mean[2] * 1e-3) / 1e9,
                           'frac': bwd_bytes / (k_mean[2] * 1e-3) / 1e9 / hbm, 'traffic': ncu_traffic('k_backward', 'config_e', world)} if k_mean[2] > 0 else None}
    launches_per_step = int(plan.launches) + (0 if args.torch_adam else 2)      # + k_adam_advance, k_adam_packed

    # ---- extra: the same step with the other intersection kernels (identical results): the plane filter of SURVEY 8(d)
    # through the constant bank (math_mode 6) and staged by TMA (math_mode 5, the round-1 kernel), each with its own
    # fraction of the FP32-FMA peak at 10 lane-instr per test
    other_modes = {}
    if args.math == 0 and not args.no_fast and roofline is not None:
        for mode in (6, 5):
            plan_m, step_m = build(dict(params, _math_mode=mode))
            lib().surf_set_kernel_timing(1)
            for _ in range(2):
                step_m()
            barrier()
            lib().surf_set_kernel_timing(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n_m = 3
            e0.record()
            for i in range(n_m):
                flush.fill_(i & 0xff)
                step_m()
            e1.record()
            barrier()
            k_ms = lib().surf_mean_kernel_ms(0, None)
            lib().surf_set_kernel_timing(0)
            other_modes['math_mode_%d' % mode] = {
                'kernel': INTERSECT_MODES[mode][0], 'formulation': INTERSECT_MODES[mode][2], 'ms_per_step': e0.elapsed_time(e1) / n_m,
                'kernel_ms': k_ms, 'frac': (tests_per_step / world) * INTERSECT_MODES[mode][1] / (k_ms * 1e-3) / peak_lane if k_ms > 0 else None,
                'note': 'per-rank time, not max over ranks'}
            del plan_m, step_m
        roofline['other_kernels'] = other_modes

    # ---- extra: the opt-in screen-space intersection kernel (math_mode 3), same step, same results
    fast = None
    if args.math == 0 and not args.no_fast:
        plan_f, step_f = build(dict(params, _math_mode=3, _pixels_per_thread=0))
        lib().surf_set_kernel_timing(1)
        for _ in range(3):
            step_f()
        barrier()
        lib().surf_set_kernel_timing(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n_fast = max(3, min(args.steps, 10))
        e0.record()
        for i in range(n_fast):
            flush.fill_(i & 0xff)
            step_f()
        e1.record()
        barrier()
        ms_fast = e0.elapsed_time(e1) / n_fast
        fast = {'math_mode': 3, 'ms_per_step': ms_fast, 'tests_per_s': tests_per_step / (ms_fast * 1e-3),
                'frames_per_s': 1e3 / ms_fast, 'intersect_kernel_ms': lib().surf_mean_kernel_ms(0, None),
                'note': 'per-pair screen-space bounding-circle level-1 test (k_intersect_screen); bit-identical outputs; '
                        'per-rank time, not max over ranks'}
        lib().surf_set_kernel_timing(0)
        del plan_f, step_f

    # ---- e2e: the C ABI with HOST buffers (pinned), H2D + D2H inside the timed region
    e2e = None
    if not args.no_e2e:
        hs = clone_scene(scene)
        m = Marshalled(hs, 'cpu')
        m.floats = [t.pin_memory() for t in m.floats]
        m.ints = {k: v.pin_memory() for k, v in m.ints.items()}
        m.cam_vecs = {k: v.pin_memory() for k, v in m.cam_vecs.items()}
        n_total = H * W
        p0, p1 = sdist.band_range(n_total, rank, world)
        target_host = tgt.detach().reshape(-1, 3)[p0:p1].cpu().pin_memory()
        grads = [torch.zeros_like(t).pin_memory() for t in m.floats]
        csc, ccam = m.c_scene(), m.c_camera()
        copt = make_options(params, (p0, p1))
        csg = m.c_grads(grads)
        ctx = lib().surf_context_create(local_rank)
        loss_c = C.c_float()
        ext = torch.cuda.ExternalStream(lib().surf_context_stream(ctx), device=dev)
        scale = 1.0 / (3.0 * n_total)
        blk, cnt = C.c_void_p(), C.c_size_t()

        def e2e_step():
            check(lib().surf_step_host_begin(ctx, C.byref(csc), C.byref(ccam), C.byref(copt), target_host.data_ptr(), scale))
            if world > 1:      # sum the per-band partial gradients (and losses) across ranks on the device block
                check(lib().surf_context_device_grads(ctx, C.byref(blk), C.byref(cnt)))
                block = torch.as_tensor(_DevBlock(blk.value, cnt.value), device=dev)
                with torch.cuda.stream(ext):
                    dist.all_reduce(block, op=dist.ReduceOp.SUM)
            check(lib().surf_step_host_end(ctx, C.byref(csg), C.byref(loss_c)))
        for _ in range(3):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        n_e2e = max(3, min(args.steps, 10))
        for _ in range(n_e2e):
            e2e_step()
        barrier()
        dt = (time.perf_counter() - t0) / n_e2e
        if world > 1:
            t = torch.tensor([dt], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        h2d, d2h = C.c_uint64(), C.c_uint64()
        lib().surf_context_last_transfer(ctx, C.byref(h2d), C.byref(d2h))
        lib().surf_context_destroy(ctx)
        e2e = {'value': tests_per_step / dt, 'unit': 'tests/s', 'h2d_bytes_per_step': int(h2d.value) * world,
               'd2h_bytes_per_step': int(d2h.value) * world, 'ms_per_step': dt * 1e3, 'frames_per_s': 1.0 / dt,
               'loss': float(loss_c.value), 'timed_with': 'host wall clock around %d steps between barriers, max over ranks' % n_e2e,
               'api': 'surf_step_host_begin / surf_step_host_end (C ABI, pinned host buffers)'}

    if rank != 0:
        return None
    cpu_baseline = cpu_baseline_numpy = None
    if not args.no_cpu_baseline and world == 1:
        cpu_baseline, cpu_baseline_numpy = cpu_baseline_subprocess(args, budget_s=20.0)
    return {'metric': 'ray-primitive tests/s (fwd+bwd inverse-rendering step)', 'value': value, 'unit': 'tests/s',
            'n_gpus': world, 'steps': args.steps, 'warmup': max(3, args.warmup), 'ms_per_step': ms_per_step,
            'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': dict(config, graph=bool(args.graph)), 'frames_per_s': 1e3 / ms_per_step, 'loss': float(loss.detach()),
            'gpu_launches': launches_per_step * args.steps, 'gpu_launches_per_step': launches_per_step,
            'clocks': clocks, 'roofline': roofline, 'cpu_baseline': cpu_baseline, 'cpu_baseline_numpy': cpu_baseline_numpy, 'e2e': e2e,
            'fast_mode': fast, 'wall_s_timed_region': wall}


def run_config_d(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    import surf_renderer_b200
    from surf_renderer_b200 import dist as sdist, scenes as synth
    from surf_renderer_b200._lib import lib
    dev = torch.device('cuda', local_rank)
    config, tests_per_step = workload_config(args)
    B, M, S = config['scenes'], config['splats'], config['width']
    peaks, peak_src = measured_peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    host = synth.config_d_batch(B, m=M, width=S, height=S, pin=True)      # batched tensors in pinned host memory
    plan = sdist.ShardedBatchStep(host, device=dev, group=(True if world > 1 else None), double_sided=True)
    w = torch.rand(B, S, S, 3, device=dev, generator=torch.Generator(device=dev).manual_seed(1))
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)

    def eager_step():
        return plan.step(lambda image: (image * w).sum())

    step = eager_step
    if args.graph:      # the whole step incl. its two NCCL collectives replays from one CUDA graph
        step = surf_renderer_b200.GraphedStep(eager_step, warmup=3, capture_error_mode='thread_local' if world > 1 else 'global')

    lib().surf_set_kernel_timing(0 if args.graph else 1)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    for _ in range(2):
        step()
    barrier()
    lib().surf_set_kernel_timing(0 if args.graph else 1)
    ms_per_step, wall, clocks, loss = timed_region(step, args.steps, args.warmup, flush, barrier, world, dev, sampler)
    n_timed = C.c_int32()
    k_mean = {k: lib().surf_mean_kernel_ms(k, C.byref(n_timed)) for k in (0, 1, 2)}
    if args.graph:      # kernel durations from a few eager steps after the timed region (events cannot be read inside a graph)
        lib().surf_set_kernel_timing(1)
        for _ in range(3):
            eager_step()
        barrier()
        k_mean = {k: lib().surf_mean_kernel_ms(k, C.byref(n_timed)) for k in (0, 1, 2)}
    lib().surf_set_kernel_timing(0)
    value = tests_per_step / (ms_per_step * 1e-3)
    f_clk = float(peaks.get('sm_max_mhz', 1965.0)) * 1e6
    peak_lane = N_SM * FP32_LANES * f_clk
    roofline = None
    if k_mean[0] > 0:
        tests_launch = tests_per_step / world
        ach_lane = tests_launch * FMA_INSTR_DENSE_TEST / (k_mean[0] * 1e-3)
        roofline = {'bound': 'fp32_fma', 'kernel': 'k_intersect_batch', 'achieved': ach_lane * 2 / 1e12, 'peak': peak_lane * 2 / 1e12,
                    'unit': 'TFLOP/s', 'frac': ach_lane / peak_lane, 'traffic': ncu_traffic('k_intersect_batch', 'config_d', world),
                    'source_hash': source_hash(),
                    'peak_source': 'theoretical 148 SM x 128 lanes x sm_max_mhz (%s)' % peak_src,
                    'algorithmic': '%d FMA-pipe lane-instr per ray-disk test (bounding-sphere test from the staged plane records; the plane filter runs on what passes) x %.3g tests per launch' % (FMA_INSTR_DENSE_TEST, tests_launch),
                    'kernel_ms': k_mean[0], 'kernel_share_of_step': k_mean[0] / ms_per_step, 'step_minus_kernel_ms': ms_per_step - k_mean[0],
                    'shade_ms': k_mean[1], 'backward_ms': k_mean[2], 'launches_timed': int(n_timed.value)}
    # ---- extra: the opt-in screen-space intersection kernel (math_mode 3), same step, same results
    fast = None
    if not args.no_fast:
        plan_f = sdist.ShardedBatchStep(host, device=dev, group=(True if world > 1 else None), double_sided=True, _math_mode=3)
        for _ in range(3):
            plan_f.step(lambda image: (image * w).sum())
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n_fast = max(3, min(args.steps, 10))
        e0.record()
        for i in range(n_fast):
            flush.fill_(i & 0xff)
            loss_f = plan_f.step(lambda image: (image * w).sum())
        e1.record()
        barrier()
        ms_fast = e0.elapsed_time(e1) / n_fast
        fast = {'math_mode': 3, 'ms_per_step': ms_fast, 'tests_per_s': tests_per_step / (ms_fast * 1e-3), 'loss': float(loss_f),
                'note': 'k_intersect_screen per scene over the internal stream pool (no fused batch kernel for this mode); '
                        'bit-identical outputs; per-rank time, not max over ranks'}
        del plan_f

    # e2e: the same step with the per-step H2D of the batch inputs (pinned host) and the D2H read of the loss
    e2e = None
    if not args.no_e2e:
        def e2e_step():
            plan.upload()
            loss = step()
            return float(loss)          # D2H read of the result
        for _ in range(3):
            e2e_step()
        barrier()
        n_e2e = max(3, min(args.steps, 10))
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            e2e_step()
        barrier()
        dt = (time.perf_counter() - t0) / n_e2e
        if world > 1:
            t = torch.tensor([dt], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {'value': tests_per_step / dt, 'unit': 'tests/s', 'h2d_bytes_per_step': int(plan.upload_bytes) * world,
               'd2h_bytes_per_step': 4 * world, 'ms_per_step': dt * 1e3,
               'timed_with': 'host wall clock around %d steps between barriers, max over ranks' % n_e2e,
               'api': 'surf_renderer_b200.dist.ShardedBatchStep (upload from pinned host arrays + step + loss read-back)'}
    if rank != 0:
        return None
    cpu_baseline = None
    if not args.no_cpu_baseline and world == 1:
        cpu_baseline, _ = cpu_baseline_subprocess(args, budget_s=20.0)
    return {'metric': 'ray-primitive tests/s (fwd+bwd over a batch of scenes)', 'value': value, 'unit': 'tests/s',
            'n_gpus': world, 'steps': args.steps, 'warmup': max(3, args.warmup), 'ms_per_step': ms_per_step,
            'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': dict(config, graph=bool(args.graph)), 'batches_per_s': 1e3 / ms_per_step, 'loss': float(loss.detach()),
            'gpu_launches': int(plan.launches) * args.steps, 'gpu_launches_per_step': int(plan.launches),
            'clocks': clocks, 'roofline': roofline, 'cpu_baseline': cpu_baseline, 'e2e': e2e, 'fast_mode': fast,
            'wall_s_timed_region': wall}


def main():
    args = parse_args()
    if args.impl == 'reference':
        return run_reference_arm(args)
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    import torch
    import torch.distributed as dist
    assert torch.cuda.is_available(), 'bench.py needs a CUDA device; there is no CPU fallback'
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    line = (run_config_e if args.workload == 'config_e' else run_config_d)(args, rank, world, local_rank)
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == '__main__':
    sys.exit(main())
