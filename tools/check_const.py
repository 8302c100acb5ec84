"""k_filter_const path (math_mode 0 on large disk frames) against the staged TMA kernel (math_mode 5): every output of
every pixel must be identical; timings of both.  python tools/check_const.py [E|bunny1024]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch                      # noqa: E402
import scene_io                   # noqa: E402
import surf_renderer_b200         # noqa: E402
from surf_renderer_b200 import scenes as synth   # noqa: E402
from surf_renderer_b200._lib import lib           # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else 'E'
if which.startswith('E'):
    scene = synth.config_e()
    if len(which) > 1:        # E128: a 1024 x 128 frame, the size of one of eight row bands
        scene['camera']['viewport'] = [0, 0, 1024, int(which[1:])]
    sc = scene_io.clone_scene(scene, device='cuda')
else:
    scene, _, _, _, _ = scene_io.load_case(os.path.join(ROOT, 'tests', 'golden', 'b_bunny_48.npz'))
    scene['camera']['viewport'] = [0, 0, 1024, 1024]
    sc = scene_io.clone_scene(scene, device='cuda')
with torch.no_grad():
    a = surf_renderer_b200.render(sc)
    b = surf_renderer_b200.render(sc, _math_mode=5)
    torch.cuda.synchronize()
    for k in ('nearest', 'depth', 'image', 'pos', 'normal'):
        same = torch.equal(a[k], b[k])
        print(k, 'identical' if same else 'DIFFERENT: %d pixels' % int((a[k] != b[k]).reshape(a[k].shape[0] * a[k].shape[1], -1).any(1).sum()))
    print('hit pixels', int((a['depth'] < 1000).sum()))
    L = lib()
    for mode in (0, 5):
        L.surf_set_kernel_timing(1)
        for _ in range(4):
            surf_renderer_b200.render(sc, _math_mode=mode)
        torch.cuda.synchronize()
        print('math_mode', mode, 'intersection stage ms (mean of 4): %.4f' % L.surf_mean_kernel_ms(0, None))
        L.surf_set_kernel_timing(0)
