"""Scenes-over-GPUs (BASELINE configs[3]: 64 scenes x 5000 splats at 128x128, fwd+bwd): run under torchrun, one rank
per GPU.  Checks render_batch_sharded against a single-GPU render_batch of the whole batch and times the step with
CUDA events (max over ranks).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_batch_check.py
"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
import torch.distributed as dist
import scene_io, surf_renderer_b200
from surf_renderer_b200 import dist as sdist, scenes as synth
from surf_renderer_b200.renderer import _stack_scenes

rank, world, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
B = 64


def build():
    st = _stack_scenes([scene_io.clone_scene(synth.config_d_scene(i), device='cuda') for i in range(B)])
    lights = st['lights']['pos'][0].detach().clone().requires_grad_(True)      # one light rig shared by all scenes
    st['lights']['pos'] = lights
    for f in ('pos', 'normal'):
        st['objects']['disk'][f] = st['objects']['disk'][f].detach().requires_grad_(True)
    return st, lights


st, lights = build()
w = torch.rand(B, 128, 128, 3, device='cuda', generator=torch.Generator(device='cuda').manual_seed(1))


def step():
    for t in (st['objects']['disk']['pos'], st['objects']['disk']['normal'], lights):
        t.grad = None
    res = sdist.render_batch_sharded(st, double_sided=True)
    (res['image'] * w).sum().backward()
    sdist.allreduce_gradients([st['objects']['disk']['pos'], st['objects']['disk']['normal'], lights])
    return res


for _ in range(3):
    res = step()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 10
e0.record()
for _ in range(reps):
    res = step()
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / reps], device='cuda')
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)

# parity against the whole batch on one GPU
st1, lights1 = build()
ref = surf_renderer_b200.render_batch(st1, double_sided=True)
(ref['image'] * w).sum().backward()
ok = torch.equal(res['image'], ref['image']) and torch.equal(res['nearest'], ref['nearest'])
ok = ok and torch.allclose(st['objects']['disk']['pos'].grad, st1['objects']['disk']['pos'].grad, rtol=1e-4, atol=1e-6)
ok = ok and torch.allclose(lights.grad, lights1.grad, rtol=1e-3, atol=1e-4)
flag = torch.tensor([1 if ok else 0], device='cuda')
if world > 1:
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    out = {'n_gpus': world, 'scenes': B, 'splats': 5000, 'size': 128, 'fwd_bwd_ms': float(ms), 'parity_all_ranks': bool(flag.item()),
           'tests_per_s': B * 5000 * 128 * 128 / (float(ms) * 1e-3)}
    print(json.dumps(out))
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    open(os.path.join(ROOT, 'gpurun_out', 'dist_batch_n%d.json' % world), 'w').write(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
sys.exit(0 if flag.item() else 1)
