#!/usr/bin/env python
"""bench.py - headline benchmark of the render hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host CPU cores

Workload (config.workload = "config_e"): BASELINE.json configs[4], the configuration the north-star target is
quoted on - the inverse-rendering step of 100 000 synthetic disk splats at 1024x1024.
One step = render (forward) -> mean((image-target)^2) -> backward to splat positions, normals, albedo and light
positions -> Adam step.  `value` = ray-primitive tests per second of the whole job = H*W*M per step / step time
with inputs resident in HBM; `e2e` = the same through the C-ABI host-pointer call (pinned HOST buffers in,
gradients + loss back in host memory, H2D/D2H inside the timed region).  N>1: the frame is sharded into row
bands, one per GPU (strong scaling of one frame), image bands all-gathered, packed gradients all-reduced (NCCL).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np   # noqa: E402
import torch         # noqa: E402

FMA_INSTR_PER_DISK_TEST = 10      # n.d 3, t 1, rel 3, |rel|^2 3 (FFMA/FMUL lane-instructions)
N_SM, FP32_LANES = 148, 128


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='surf', choices=['surf', 'reference'])
    ap.add_argument('--splats', type=int, default=100_000)
    ap.add_argument('--size', type=int, default=1024)
    ap.add_argument('--ppt', type=int, default=0, help='pixels per thread of the intersection kernel (0 = default)')
    ap.add_argument('--chunk', type=int, default=0)
    ap.add_argument('--math', type=int, default=0, help='intersection kernel: 0 = default ray-plane FFMA2 filter (10 instr/test), 3 = screen-space fast mode')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    return ap.parse_args()


def measured_peaks():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            return json.load(f), 'measured'
    except Exception:
        return {'hbm_gbs': 6650.0, 'sm_max_mhz': 1965.0}, 'fallback'


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
# reference arm / cpu baseline: the oracle (op-for-op restatement of the reference's torch path) on host cores
# ----------------------------------------------------------------------------------------------------------
def cpu_reference_step(scene, target_scene, subset, threads):
    """One bounded sample of the workload on the CPU: fwd + loss + bwd over `subset` pixels of the frame."""
    from oracle import torch_oracle
    from surf_renderer_b200.scenes import clone_scene
    torch.set_num_threads(threads)
    sc = clone_scene(scene, requires_grad=False)
    leaves = [sc['objects']['disk']['pos'], sc['objects']['disk']['normal'], sc['materials']['albedo'], sc['lights']['pos']]
    for t in leaves:
        t.requires_grad_(True)
    t0 = time.perf_counter()
    with torch.no_grad():
        tgt = torch_oracle.render(target_scene, pixel_subset=subset, tile_size=512)['image']
    t_target = time.perf_counter() - t0
    t0 = time.perf_counter()
    res = torch_oracle.render(sc, pixel_subset=subset, tile_size=512)
    loss = ((res['image'] - tgt) ** 2).mean()
    t_fwd = time.perf_counter() - t0
    t0 = time.perf_counter()
    loss.backward()
    t_bwd = time.perf_counter() - t0
    return t_fwd, t_bwd, t_target


def cpu_sample(scene, n_pix_sample, seed=123):
    vp = scene['camera']['viewport']
    n = (vp[2] - vp[0]) * (vp[3] - vp[1])
    g = torch.Generator().manual_seed(seed)
    return torch.randperm(n, generator=g)[:n_pix_sample].sort().values


def run_cpu_baseline(scene, target_scene, budget_s, threads, steps=1, warmup=0):
    m = int(scene['objects']['disk']['pos'].shape[0])
    # calibrate on 128 pixels, then size the sample for the time budget
    sub = cpu_sample(scene, 128)
    tf, tb, _ = cpu_reference_step(scene, target_scene, sub, threads)
    per_pix = (tf + tb) / 128
    vp = scene['camera']['viewport']
    n_total = (vp[2] - vp[0]) * (vp[3] - vp[1])
    n_pix = int(max(128, min(8192, n_total, budget_s / max(per_pix, 1e-9) / max(1, steps + warmup))))
    sub = cpu_sample(scene, n_pix)
    n_pix = int(sub.numel())
    times = []
    for i in range(warmup + steps):
        tf, tb, _ = cpu_reference_step(scene, target_scene, sub, threads)
        if i >= warmup:
            times.append((tf, tb))
    tf = float(np.mean([t[0] for t in times])); tb = float(np.mean([t[1] for t in times]))
    tests = float(m) * n_pix
    return {'value': tests / (tf + tb), 'unit': 'tests/s', 'cores': threads, 'kind': 'port',
            'sample': '%d random pixels of the %dx%d frame x %d splats, fwd+bwd (fwd %.2fs, bwd %.2fs), torch-CPU '
                      'op-for-op port of the reference, tile_size=512' % (n_pix, scene['camera']['viewport'][2],
                                                                          scene['camera']['viewport'][3], m, tf, tb),
            'fwd_tests_per_s': tests / tf, 'ms_per_sample_step': 1e3 * (tf + tb)}, n_pix, tf + tb


def run_numpy_baseline(scene, n_pix=96, reps=2):
    """The reference's second CPU renderer (diffrend/numpy/renderer.py, restated in oracle/numpy_oracle.py): forward
    only (it has no gradients), float64, Lambertian, and it materialises the whole [M, N, 4] tensor - so the sample is
    a small pixel subset of the same frame and the same splats."""
    from oracle import numpy_oracle
    hs = numpy_oracle.homogeneous_scene(scene)
    m = int(hs['objects']['disk']['pos'].shape[0])
    sub = cpu_sample(scene, n_pix).numpy()
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter()
        numpy_oracle.render(hs, pixel_subset=sub)
        best = min(best, time.perf_counter() - t0)
    return {'value': float(m) * len(sub) / best, 'unit': 'tests/s (forward only)', 'cores': 1, 'kind': 'port',
            'sample': '%d random pixels of the frame x %d splats, forward, best of %d (%.2fs); numpy float64 port of '
                      'the reference numpy twin (no tiling: [M,N,4] temporaries; elementwise numpy is single-threaded)'
                      % (len(sub), m, reps, best)}


# ----------------------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    from surf_renderer_b200 import scenes as synth
    scene = synth.config_e(m=args.splats, width=args.size, height=args.size)
    target_scene = synth.config_e_target_scene(scene)
    M, H, W = args.splats, args.size, args.size
    tests_per_step = float(M) * H * W
    config = {'workload': 'config_e', 'splats': M, 'width': W, 'height': H, 'lights': 3,
              'step': 'render fwd + mse(image,target) + bwd(pos,normal,albedo,light_pos) + Adam',
              'sharding': 'row-bands x%d' % max(1, args.gpus),
              'l2': 'flushed between timed steps (256 MiB fill enqueued on the stream, inside the timed region)'}

    if args.impl == 'reference':
        if rank != 0:
            return 0
        threads = os.cpu_count() or 1
        cb, n_pix, step_s = run_cpu_baseline(scene, target_scene, budget_s=150.0, threads=threads,
                                             steps=max(1, args.steps), warmup=max(0, min(args.warmup, 1)))
        line = {'impl': 'reference', 'metric': 'ray-primitive tests/s (fwd+bwd inverse-rendering step)',
                'value': cb['value'], 'unit': 'tests/s', 'n_gpus': 0, 'steps': args.steps, 'warmup': args.warmup,
                'ms_per_step': 1e3 * step_s, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
                'dtype': 'f32', 'data': 'synthetic', 'config': config, 'cpu_baseline': cb,
                'frames_per_s_extrapolated': cb['value'] / tests_per_step,
                'cpu_baseline_numpy': run_numpy_baseline(scene),
                'e2e': {'value': cb['value'], 'unit': 'tests/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ our arm
    import torch.distributed as dist
    import surf_renderer_b200
    from surf_renderer_b200.scenes import clone_scene
    from surf_renderer_b200 import _abi, dist as sdist
    from surf_renderer_b200._lib import check, lib
    from surf_renderer_b200.marshal import Marshalled, make_options
    assert torch.cuda.is_available(), 'bench.py needs a CUDA device; there is no CPU fallback'
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    peaks, peak_src = measured_peaks()
    params = {'_pixels_per_thread': args.ppt, '_chunk_prims': args.chunk, '_math_mode': args.math}

    sc = clone_scene(scene, device=dev)
    leaves = [sc['objects']['disk']['pos'], sc['objects']['disk']['normal'], sc['materials']['albedo'], sc['lights']['pos']]
    for t in leaves:
        t.requires_grad_(True)
    opt = torch.optim.Adam(leaves, lr=1e-4, fused=True)
    with torch.no_grad():
        tgt = surf_renderer_b200.render(clone_scene(target_scene, device=dev), **params)['image']
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    lib().surf_set_kernel_timing(1)
    launches = []

    def step():
        opt.zero_grad(set_to_none=True)
        if world > 1:
            res = sdist.render_bands(sc, gather=('image',), **params)
        else:
            res = surf_renderer_b200.render(sc, **params)
        n_launch = lib().surf_last_launch_count()
        loss = ((res['image'] - tgt) ** 2).mean()
        loss.backward()
        n_launch += lib().surf_last_launch_count()
        if world > 1:
            sdist.allreduce_gradients(leaves)
        opt.step()
        return loss, n_launch

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    lib().surf_set_kernel_timing(1)                   # reset the library's per-kernel event ring
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_wall0 = time.perf_counter()
    ev0.record()
    for i in range(args.steps):
        flush.fill_(i & 0xff)                      # L2 flush between steps (stream-ordered, inside the timed region)
        loss, n_launch = step()
        launches.append(n_launch)
    ev1.record()
    barrier()
    wall = time.perf_counter() - t_wall0
    total_ms = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([total_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = total_ms / args.steps
    value = tests_per_step / (ms_per_step * 1e-3)
    n_timed = C.c_int32()
    k_mean = {k: lib().surf_mean_kernel_ms(k, C.byref(n_timed)) for k in (0, 1, 2)}

    # ---- roofline of the dominant kernel (k_intersect), timed live with CUDA events inside the library
    isect_ms = k_mean[0] if k_mean[0] > 0 else None
    f_clk = float(peaks.get('sm_max_mhz', 1965.0)) * 1e6
    peak_lane = N_SM * FP32_LANES * f_clk                    # lane-instr/s
    roofline = None
    if isect_ms:
        tests_launch = tests_per_step / world
        ach_lane = tests_launch * FMA_INSTR_PER_DISK_TEST / (isect_ms * 1e-3)
        fma_meas = max(lib().surf_fma_peak(0, 8192, None), lib().surf_fma_peak(1, 8192, None))
        roofline = {'bound': 'fp32_fma', 'kernel': 'k_intersect', 'achieved': ach_lane * 2 / 1e12, 'peak': peak_lane * 2 / 1e12,
                    'unit': 'TFLOP/s', 'frac': ach_lane / peak_lane,
                    # dram__bytes_read+write of one k_intersect launch, ncu --set full (profiles/r1_ncu_full_step_kernels.json);
                    # only meaningful for the default single-GPU config it was captured on
                    'traffic': 22395904 if (world == 1 and M == 100_000 and H == 1024 and args.math == 0) else None,
                    'peak_source': 'theoretical 148 SM x 128 lanes x sm_max_mhz (%s MEASURED_PEAKS.json has no fp32 entry)' % peak_src,
                    'peak_measured': fma_meas * 2 / 1e12, 'frac_of_measured': ach_lane / fma_meas if fma_meas > 0 else None,
                    'algorithmic': '%d FMA-pipe lane-instr per ray-disk test x %.3g tests per launch' % (FMA_INSTR_PER_DISK_TEST, tests_launch),
                    'kernel_ms': isect_ms, 'kernel_share_of_step': isect_ms / ms_per_step,
                    'shade_ms': k_mean[1], 'backward_ms': k_mean[2], 'launches_timed': int(n_timed.value)}
        # the two HBM-side kernels (north star: achieved GB/s against the measured copy bandwidth)
        n_loc = H * W / world
        hbm = float(peaks.get('hbm_gbs', 6650.0))
        shade_bytes = n_loc * (12 + 8 + 48)            # rays + z-buffer key read; image/depth/normal/pos/nearest written
        bwd_bytes = n_loc * (12 + 8 + 4 + 12)          # rays, nearest, depth, d(image) read (+ ~28 B per hit pixel of atomics)
        roofline['hbm_kernels'] = {
            'peak_gbs': hbm, 'peak_source': peak_src + ' MEASURED_PEAKS.json hbm_gbs',
            'k_shade': {'ms': k_mean[1], 'algorithmic_bytes': shade_bytes, 'achieved_gbs': shade_bytes / (k_mean[1] * 1e-3) / 1e9,
                        'frac': shade_bytes / (k_mean[1] * 1e-3) / 1e9 / hbm} if k_mean[1] > 0 else None,
            'k_backward': {'ms': k_mean[2], 'algorithmic_bytes': bwd_bytes, 'achieved_gbs': bwd_bytes / (k_mean[2] * 1e-3) / 1e9,
                           'frac': bwd_bytes / (k_mean[2] * 1e-3) / 1e9 / hbm} if k_mean[2] > 0 else None}

    # ---- extra: the opt-in screen-space intersection kernel (math_mode 3), same step, same results
    fast = None
    if args.math == 0:
        params_fast = dict(params, _math_mode=3, _pixels_per_thread=0)
        params_keep = dict(params)
        params.clear(); params.update(params_fast)
        for _ in range(3):
            step()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n_fast = max(3, min(args.steps, 10))
        e0.record()
        for _ in range(n_fast):
            step()
        e1.record()
        barrier()
        ms_fast = e0.elapsed_time(e1) / n_fast
        fast = {'math_mode': 3, 'ms_per_step': ms_fast, 'tests_per_s': tests_per_step / (ms_fast * 1e-3),
                'frames_per_s': 1e3 / ms_fast, 'intersect_kernel_ms': lib().surf_last_kernel_ms(0),
                'note': 'per-pair screen-space bounding-circle level-1 test (2.25 FMA-pipe lane-instr/test); bit-identical outputs'}
        params.clear(); params.update(params_keep)

    # ---- e2e: C-ABI host-pointer call, pinned host buffers, H2D + D2H inside the timed region
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

        want = [i for i, nme in enumerate(m.names) if nme in ('objects/disk/pos', 'objects/disk/normal',
                                                                'materials/albedo', 'lights/pos')]
        flat_host = torch.empty(sum(grads[i].numel() for i in want)).pin_memory()

        def e2e_step():
            check(lib().surf_render_backward_host(ctx, C.byref(csc), C.byref(ccam), C.byref(copt), None, None,
                                                  target_host.data_ptr(), C.byref(loss_c), C.byref(csg)))
            if world > 1:      # sum the per-band partial gradients across ranks (packed buffer, one all-reduce)
                torch.cat([grads[i].reshape(-1) for i in want], out=flat_host)
                flat_dev = flat_host.to(dev, non_blocking=True)
                dist.all_reduce(flat_dev, op=dist.ReduceOp.SUM)
                flat_host.copy_(flat_dev, non_blocking=True)
                torch.cuda.synchronize()
        for _ in range(2):
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
               'api': 'surf_render_backward_host (C ABI, pinned host buffers)'}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    cpu_baseline = cpu_baseline_numpy = None
    if not args.no_cpu_baseline and world == 1:
        cpu_baseline, _, _ = run_cpu_baseline(scene, target_scene, budget_s=20.0, threads=os.cpu_count() or 1)
        cpu_baseline_numpy = run_numpy_baseline(scene)

    line = {'metric': 'ray-primitive tests/s (fwd+bwd inverse-rendering step)', 'value': value, 'unit': 'tests/s',
            'n_gpus': world, 'steps': args.steps, 'warmup': max(3, args.warmup), 'ms_per_step': ms_per_step,
            'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': config, 'frames_per_s': 1e3 / ms_per_step, 'loss': float(loss.detach()),
            'gpu_launches': int(sum(launches)), 'gpu_launches_per_step': int(launches[-1]) if launches else 0,
            'clocks': clocks, 'roofline': roofline, 'cpu_baseline': cpu_baseline, 'cpu_baseline_numpy': cpu_baseline_numpy, 'e2e': e2e,
            'fast_mode': fast,
            'wall_s_timed_region': wall}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == '__main__':
    sys.exit(main())
