"""Scenes-over-GPUs (BASELINE configs[3]: 64 scenes x 5000 splats at 128x128, fwd+bwd): run under torchrun, one rank
per GPU.  Checks dist.ShardedBatchStep at full size against a single-GPU render_batch of the whole batch (images
bit-identical, block gradients and the all-reduced light gradients within atomics tolerance) and times the step with
CUDA events (max over ranks), eager and replayed from a CUDA graph.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_batch_check.py
"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
import torch.distributed as dist
import scene_io, surf_renderer_b200
from surf_renderer_b200 import dist as sdist, scenes as synth

rank, world, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
if world > 1:
    dist.init_process_group('nccl', device_id=dev)
B = 64
batch = synth.config_d_batch(B)
w = torch.rand(B, 128, 128, 3, device=dev, generator=torch.Generator(device=dev).manual_seed(1))
plan = sdist.ShardedBatchStep(batch, device=dev, group=(True if world > 1 else None), double_sided=True)
loss_fn = lambda im: (im * w).sum()      # noqa: E731


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms)


ms_eager = timed(lambda: plan.step(loss_fn))
# parity against the whole batch on one GPU
plan.step(loss_fn)
ref_sc = scene_io.clone_scene(batch, device=dev)
for t in (ref_sc['objects']['disk']['pos'], ref_sc['objects']['disk']['normal'], ref_sc['lights']['pos']):
    t.requires_grad_(True)
ref = surf_renderer_b200.render_batch(ref_sc, double_sided=True)
(ref['image'] * w).sum().backward()
b0, b1 = plan.block
ok = torch.equal(plan.image_full.view(B, 128, 128, 3), ref['image'])
gp, gr = plan.leaves['objects/disk/pos'].grad, ref_sc['objects']['disk']['pos'].grad[b0:b1]
ok = ok and torch.allclose(gp, gr, rtol=1e-4, atol=2e-6 * float(gr.abs().max()))
gl, glr = plan.leaves['lights/pos'].grad, ref_sc['lights']['pos'].grad
ok = ok and torch.allclose(gl, glr, rtol=1e-3, atol=1e-5 * float(glr.abs().max()))
graphed = surf_renderer_b200.GraphedStep(lambda: plan.step(loss_fn), warmup=3, capture_error_mode='thread_local' if world > 1 else 'global')
ms_graph = timed(graphed)
flag = torch.tensor([1 if ok else 0], device=dev)
if world > 1:
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    out = {'n_gpus': world, 'scenes': B, 'splats': 5000, 'size': 128, 'fwd_bwd_ms_eager': ms_eager, 'fwd_bwd_ms_graph': ms_graph,
           'parity_all_ranks': bool(flag.item()), 'tests_per_s_graph': B * 5000 * 128 * 128 / (ms_graph * 1e-3)}
    print(json.dumps(out))
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    open(os.path.join(ROOT, 'gpurun_out', 'dist_batch_n%d.json' % world), 'w').write(json.dumps(out))
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
sys.stdout.flush()
os._exit(0 if flag.item() else 1)
