"""Every BASELINE.json config on one B200: forward and fwd+bwd time (CUDA events), ray-primitive tests/s, and the
intersection kernel's share and FP32-FMA fraction.  Algorithmic FMA-pipe lane-instr per test of the filter that runs:
disk 3 (frames above 256x256: bounding-sphere test through the constant bank, k_filter_const) or 4 (dense / batch kernel:
bounding-sphere test from the staged plane records), sphere 10, triangle 16, plane 4 (the plane filter of a disk,
math_mode 5 / 6, is 10).  A  basic.json 64x64 - B  bunny.splat 256x256 - C  torus_1K.obj 512x512 - D  64 x 5000 splats
128x128 (stacked batch) - E  100K splats 1024x1024.  Scenes A-C come from the reference-generated fixtures in
tests/golden (same primitives, the viewport set to the config's size)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
import scene_io, surf_renderer_b200
from surf_renderer_b200 import scenes as synth
from surf_renderer_b200._lib import lib
from surf_renderer_b200.renderer import _stack_scenes

INSTR = {'disk': 4, 'sphere': 10, 'triangle': 16, 'plane': 4}
INSTR_DISK_CONST = 3
PEAK = 148 * 128 * 1.965e9          # FP32 lane-instr/s at the measured max clock


def fixture(name, size, grad_fields):
    scene, params, _, _, _ = scene_io.load_case(os.path.join(ROOT, 'tests', 'golden', name + '.npz'))
    scene['camera']['viewport'] = [0, 0, size, size]
    sc = scene_io.clone_scene(scene, device='cuda')
    for kind, f in grad_fields:
        sc['objects'][kind][f].requires_grad_(True)
    sc['materials']['albedo'].requires_grad_(True)
    return sc, params


def timed(fn, reps, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def measure(name, scene, params, n_scenes=1, reps=20, batched=False):
    objs = scene['objects']
    H = W = scene['camera']['viewport'][2]
    counts = {k: int((v['face'] if k == 'triangle' else v['pos']).shape[-3 if k == 'triangle' else -2]) for k, v in objs.items()}
    tests = float(sum(counts.values())) * H * W * n_scenes
    const_path = not batched and H * W > 256 * 256 and 'triangle' not in counts and params.get('_math_mode', 0) == 0
    lane_instr = float(sum((INSTR_DISK_CONST if (k == 'disk' and const_path and c >= 256) else INSTR[k]) * c for k, c in counts.items())) * H * W * n_scenes
    call = (lambda: surf_renderer_b200.render_batch(scene, **params)) if batched else (lambda: surf_renderer_b200.render(scene, **params))

    def fwd():
        with torch.no_grad():
            call()

    def fb():
        call()['image'].sum().backward()
    t_f = timed(fwd, reps)
    t_fb = timed(fb, reps)
    # the same fwd + loss + bwd as ONE library call (MSEStep), eager and replayed from a CUDA graph
    t_step = t_graph = None
    if not batched:
        with torch.no_grad():
            target = torch.rand_like(call()['image'])
        plan = surf_renderer_b200.MSEStep(scene, target, **params)
        t_step = timed(plan, reps)
        graphed = surf_renderer_b200.GraphedStep(plan, warmup=3)
        t_graph = timed(graphed, reps)
        for t in plan.leaves:
            t.grad = None
    L = lib()
    L.surf_set_kernel_timing(1)
    fwd(); fwd(); fwd()
    torch.cuda.synchronize()
    k_ms = L.surf_mean_kernel_ms(0, None)
    L.surf_set_kernel_timing(0)
    row = {'config': name, 'primitives': counts, 'scenes': n_scenes, 'size': [H, W], 'tests_per_frame': tests,
           'forward_ms': t_f, 'fwd_bwd_ms': t_fb, 'mse_step_ms': t_step, 'mse_step_graph_ms': t_graph, 'forward_tests_per_s': tests / (t_f * 1e-3),
           'fwd_bwd_tests_per_s': tests / (t_fb * 1e-3), 'fwd_bwd_frames_per_s': n_scenes * 1e3 / t_fb,
           'k_intersect_ms': k_ms, 'k_intersect_frac_fp32_peak': lane_instr / (k_ms * 1e-3) / PEAK,
           'k_intersect_share_of_forward': k_ms / t_f}
    print(json.dumps(row), flush=True)
    return row


rows = []
sc, p = fixture('a_basic_json_64', 64, [('triangle', 'face')])
rows.append(measure('A basic.json 64x64', sc, p))
sc, _ = fixture('a_basic_mixed_64', 64, [('disk', 'pos')])
rows.append(measure("A' plane+sphere+disk 64x64", sc, {}))
sc, p = fixture('b_bunny_48', 256, [('disk', 'pos'), ('disk', 'normal')])
rows.append(measure('B bunny.splat 256x256', sc, {}))
rows.append(measure('B bunny.splat 256x256, math_mode 4 (dense)', sc, {'_math_mode': 4}))
sc, p = fixture('c_torus_64', 512, [('triangle', 'face'), ('triangle', 'normal')])
rows.append(measure('C torus_1K.obj 512x512', sc, {'double_sided': True}))
st = _stack_scenes([scene_io.clone_scene(synth.config_d_scene(i), device='cuda') for i in range(64)])
for f in ('pos', 'normal'):
    st['objects']['disk'][f].requires_grad_(True)
rows.append(measure('D 64 scenes x 5000 splats 128x128 (stacked batch)', st, {'double_sided': True}, n_scenes=64, batched=True))
sc = scene_io.clone_scene(synth.config_e(), device='cuda')
for f in ('pos', 'normal'):
    sc['objects']['disk'][f].requires_grad_(True)
rows.append(measure('E 100K splats 1024x1024', sc, {}, reps=5))
os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
json.dump(rows, open(os.path.join(ROOT, 'gpurun_out', 'bench_configs.json'), 'w'), indent=1)
