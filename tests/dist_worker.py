"""Worker of tests/test_gpu_multi.py: run under torchrun with one rank per GPU (NCCL).  Checks, on every rank, that the
multi-GPU paths reproduce the single-GPU results (pixels / scenes are independent: renderer.py:170-198, gan.py:326-377):

  * MSEStep over row bands: loss and all-reduced gradients equal the single-GPU step (atomics tolerance), the gathered
    image is bit-identical;
  * render_bands: gathered image / depth / nearest bit-identical to the full-frame render;
  * ShardedBatchStep: gathered images bit-identical to render_batch of the whole batch, block gradients and the
    all-reduced shared light gradients equal (atomics tolerance);
  * (reported, not asserted) whether the band step incl. its NCCL all-reduce replays from a CUDA graph.

Prints one JSON line from rank 0 and exits non-zero on any rank that saw a mismatch.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch                         # noqa: E402
import torch.distributed as dist     # noqa: E402
import scene_io                      # noqa: E402
import surf_renderer_b200            # noqa: E402
from surf_renderer_b200 import dist as sdist, scenes as synth   # noqa: E402

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
report, ok = {}, True


def close(a, b, rtol=1e-4, atol_scale=2e-6):
    scale = float(b.abs().max()) if b.numel() else 0.0
    return bool(torch.allclose(a, b, rtol=rtol, atol=atol_scale * max(scale, 1e-30)))


# ---- 1. MSEStep over row bands vs the single-GPU step -------------------------------------------------------------
# bands of more than 256 x 256 pixels: the constant-bank intersection path (two streams) runs next to the NCCL stream
scene = synth.config_e(m=6000, width=512, height=384, radius=0.02)
target = surf_renderer_b200.render(scene_io.clone_scene(synth.config_e_target_scene(scene, jitter=0.01), device=dev))['image'].detach()


def leaves_of(sc):
    ls = [sc['objects']['disk']['pos'], sc['objects']['disk']['normal'], sc['materials']['albedo'], sc['lights']['pos']]
    for t in ls:
        t.requires_grad_(True)
    return ls


one = scene_io.clone_scene(scene, device=dev)
l1 = leaves_of(one)
p1 = surf_renderer_b200.MSEStep(one, target)
loss1 = p1()
band = scene_io.clone_scene(scene, device=dev)
lb = leaves_of(band)
pb = surf_renderer_b200.MSEStep(band, target, group=True)
lossb = pb()
torch.cuda.synchronize()
report['band_step_loss'] = [float(loss1), float(lossb)]
ok &= abs(float(loss1) - float(lossb)) <= 2e-6 * abs(float(loss1))
for a, b, name in zip(lb, l1, ('pos', 'normal', 'albedo', 'light_pos')):
    good = close(a.grad, b.grad)
    report['band_step_grad_' + name] = good
    ok &= good
img = pb.gather_image()
ok &= bool(torch.equal(img, p1.image.view_as(img)))
report['band_step_image_bit_identical'] = bool(torch.equal(img, p1.image.view_as(img)))

# ---- 2. render_bands (autograd path) vs the full frame ------------------------------------------------------------
full = surf_renderer_b200.render(scene_io.clone_scene(scene, device=dev))
rb = sdist.render_bands(scene_io.clone_scene(scene, device=dev), gather=('image', 'depth', 'nearest'))
same = all(bool(torch.equal(rb[k], full[k])) for k in ('image', 'depth', 'nearest'))
report['render_bands_bit_identical'] = same
ok &= same

# ---- 3. ShardedBatchStep vs render_batch of the whole batch on one GPU ---------------------------------------------
B = 2 * world
batch = synth.config_d_batch(B, m=800, width=48, height=40, radius=0.05)
w = torch.rand(B, 40, 48, 3, device=dev, generator=torch.Generator(device=dev).manual_seed(4))
plan = sdist.ShardedBatchStep(batch, device=dev, group=True, double_sided=True)
loss_s = plan.step(lambda im: (im * w).sum())
ref_sc = scene_io.clone_scene(batch, device=dev)
for t in (ref_sc['objects']['disk']['pos'], ref_sc['objects']['disk']['normal'], ref_sc['lights']['pos']):
    t.requires_grad_(True)
ref = surf_renderer_b200.render_batch(ref_sc, double_sided=True)
loss_r = (ref['image'] * w).sum()
loss_r.backward()
torch.cuda.synchronize()
b0, b1 = plan.block
same = bool(torch.equal(plan.image_full.view(B, 40, 48, 3), ref['image']))
report['sharded_batch_images_bit_identical'] = same
ok &= same
g_pos = close(plan.leaves['objects/disk/pos'].grad, ref_sc['objects']['disk']['pos'].grad[b0:b1])
g_nrm = close(plan.leaves['objects/disk/normal'].grad, ref_sc['objects']['disk']['normal'].grad[b0:b1])
g_lgt = close(plan.leaves['lights/pos'].grad, ref_sc['lights']['pos'].grad, rtol=1e-3, atol_scale=1e-5)
report['sharded_batch_grads'] = [g_pos, g_nrm, g_lgt]
ok &= g_pos and g_nrm and g_lgt
ok &= abs(float(loss_s) - float(loss_r)) <= 1e-5 * abs(float(loss_r))

# ---- 4. the band step incl. NCCL replayed from a CUDA graph (reported) ---------------------------------------------
try:
    gsc = scene_io.clone_scene(scene, device=dev)
    gl = leaves_of(gsc)
    gp = surf_renderer_b200.MSEStep(gsc, target, group=True)
    opt = torch.optim.Adam(gl, lr=1e-4, capturable=True)

    def gstep():
        loss = gp()
        opt.step()
        return loss
    graphed = surf_renderer_b200.GraphedStep(gstep, warmup=3, capture_error_mode='thread_local')
    vals = [float(graphed()) for _ in range(3)]
    torch.cuda.synchronize()
    report['graph_with_nccl'] = {'captured': True, 'losses': vals}
    del graphed
except Exception as e:      # noqa: BLE001
    report['graph_with_nccl'] = {'captured': False, 'error': repr(e)[:300]}
    torch.cuda.synchronize()

flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    report['ok_all_ranks'] = bool(flag.item())
    report['world'] = world
    print(json.dumps(report))
    out_dir = os.path.join(ROOT, 'gpurun_out')
    os.makedirs(out_dir, exist_ok=True)
    with open(os.path.join(out_dir, 'dist_worker_n%d.json' % world), 'w') as f:
        f.write(json.dumps(report))
code = 0 if flag.item() else 1
torch.cuda.synchronize()
dist.barrier()
sys.stdout.flush()
# A process group whose collectives were captured into a CUDA graph can hang in destroy_process_group(); every rank has
# passed the barrier above, so leave without the teardown.
os._exit(code)
