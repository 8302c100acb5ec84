"""Forward renders of one BASELINE config for profiling: python tools/run_config.py B|C|D|E [reps] [math_mode]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch                      # noqa: E402
import scene_io                   # noqa: E402
import surf_renderer_b200         # noqa: E402
from surf_renderer_b200 import scenes as synth   # noqa: E402
from surf_renderer_b200._lib import lib           # noqa: E402

which = sys.argv[1]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
extra = {'_math_mode': int(sys.argv[3])} if len(sys.argv) > 3 else {}
if len(sys.argv) > 4:
    extra['_chunk_prims'] = int(sys.argv[4])
if len(sys.argv) > 5:
    extra['_pixels_per_thread'] = int(sys.argv[5])


def fixture(name, size):
    scene, params, _, _, _ = scene_io.load_case(os.path.join(ROOT, 'tests', 'golden', name + '.npz'))
    scene['camera']['viewport'] = [0, 0, size, size]
    return scene_io.clone_scene(scene, device='cuda')


if which == 'B':
    sc, call = fixture('b_bunny_48', 256), lambda s: surf_renderer_b200.render(s, **extra)
elif which == 'C':
    sc, call = fixture('c_torus_64', 512), lambda s: surf_renderer_b200.render(s, double_sided=True, **extra)
elif which == 'D':
    sc = scene_io.clone_scene(synth.config_d_batch(64), device='cuda')
    call = lambda s: surf_renderer_b200.render_batch(s, double_sided=True, **extra)      # noqa: E731
else:
    sc, call = scene_io.clone_scene(synth.config_e(), device='cuda'), lambda s: surf_renderer_b200.render(s, **extra)
if os.environ.get('RUN_GRAPH'):
    # back-to-back replays from a CUDA graph: the GPU never idles between launches (clocks stay up)
    with torch.no_grad():
        for _ in range(3):
            call(sc)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(10):
                call(sc)
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
    print(which, extra, 'graph replay: %.4f ms per forward (all kernels)' % (e0.elapsed_time(e1) / (10 * reps)))
    sys.exit(0)
lib().surf_set_kernel_timing(1)
with torch.no_grad():
    for _ in range(reps):
        call(sc)
torch.cuda.synchronize()
print(which, extra, 'intersect kernel ms (mean of %d): %.4f' % (reps, lib().surf_mean_kernel_ms(0, None)))
