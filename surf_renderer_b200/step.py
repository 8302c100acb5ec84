"""The reference's inverse-rendering step as ONE library call (reference: diffrend/torch/test_optimization.py:100-125).

The reference optimises a scene with

    res = render(scene)                      # test_optimization.py:100
    loss = mean((res['image'] - target)**2)  # :104 (criterion = MSELoss)
    loss.backward()                          # :121
    optimizer.step()                         # :122

``MSEStep`` is the same step through ``surf_step_mse`` (include/surf_b200.h): forward, the loss and d(loss)/d(image)
fused into the shading epilogue, and the backward, enqueued by one host call - no autograd graph, no intermediate
full-frame tensors, nothing between the caller's leaves and the kernels.  The plan is built once per scene (the
marshalling, the workspace, the output buffers and ONE packed gradient buffer are reused by every step), so a step
costs one memset + one C call, and - on several GPUs - one NCCL all-reduce:

  * tiles over GPUs (BASELINE configs[4]): with a process group, rank r renders the flat pixel band
    ``band_range(H*W, r, G)`` against all primitives and evaluates the loss of ITS band with the weight 1/(3 H W) of the
    full-frame mean; the bands' partial losses and partial gradients add up, so one in-place ``all_reduce(SUM)`` over
    the packed buffer ``[gradients of every leaf | loss]`` completes the step.  The image is never gathered: the loss
    does not need it (``gather_image()`` fetches it on demand).
  * the leaves' ``.grad`` are views into the packed buffer, so an optimizer (torch.optim.Adam(..., capturable=True) for
    graph capture) steps them directly and nothing is copied before or after the collective.

Everything is enqueued on the current stream without host synchronisation: ``GraphedStep(lambda: (plan(), opt.step()))``
captures the whole step into one CUDA graph.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist

from . import _abi
from ._lib import check, lib
from .marshal import Marshalled, make_options
from .renderer import _resolve_device, _stream_ptr, get_param_value


def _band_range(n_pixels, rank, world):
    base, rem = divmod(n_pixels, world)
    p0 = rank * base + min(rank, rem)
    return p0, p0 + base + (1 if rank < rem else 0)


class MSEStep:
    """plan = MSEStep(scene, target, group=None, **params);  loss = plan()  ->  0-d device tensor, leaves' .grad set.

    `scene`: the reference's scene dict; its float tensors with requires_grad=True are the leaves (they must already be
    contiguous fp32 tensors on the CUDA device, so that the kernels read the very tensors the optimizer updates).
    `target`: [H, W, 3] image.  `params`: render()'s kwargs (double_sided, use_quartic, shadow, ...).
    `group`: a torch.distributed process group (or True for the default group): shard the frame into row bands."""

    def __init__(self, scene, target, group=None, keep_outputs=('image',), **params):
        if get_param_value('vis_stat', params, False):
            raise RuntimeError('Removed Support for vis_stat')
        dev = _resolve_device(scene)
        self.device = dev
        self.m = m = Marshalled(scene, dev)
        self.params = dict(params)
        self.group = None if group in (None, True) else group
        self.world = dist.get_world_size(self.group) if (group is not None and dist.is_initialized()) else 1
        self.rank = dist.get_rank(self.group) if self.world > 1 else 0
        self.H, self.W = m.height, m.width
        n_total = m.n_pixels
        self.band = _band_range(n_total, self.rank, self.world)
        n = self.band[1] - self.band[0]
        self.n = n
        target = torch.as_tensor(target, dtype=torch.float32, device=dev).reshape(-1, 3)
        if target.shape[0] != n_total:
            raise ValueError('target must be [H, W, 3] = [%d, %d, 3]' % (self.H, self.W))
        self.target = target[self.band[0]:self.band[1]].contiguous()
        # leaves: scene tensors that require grad; they must be the caller's own tensors (no marshalling copy)
        self.leaf_slots = [i for i, t in enumerate(m.floats) if t.requires_grad]
        if not self.leaf_slots:
            raise ValueError('MSEStep: no scene tensor requires grad')
        self.leaves = [m.floats[i] for i in self.leaf_slots]
        originals = self._scene_tensors(scene)
        for t in self.leaves:
            if not any(t is o for o in originals):
                raise ValueError('MSEStep: a leaf had to be copied while marshalling (it must be a contiguous float32 CUDA '
                                 'tensor on the render device) - its gradient would not reach the optimizer')
        # ONE packed buffer: [grad of leaf 0 | grad of leaf 1 | ... | loss]; the leaves' .grad are views into it
        sizes = [t.numel() for t in self.leaves]
        self.packed = torch.zeros(sum(sizes) + 1, dtype=torch.float32, device=dev)
        grads = [None] * len(m.floats)
        off = 0
        for i, t, sz in zip(self.leaf_slots, self.leaves, sizes):
            view = self.packed[off:off + sz].view_as(t)
            grads[i] = view
            t.grad = view
            off += sz
        self.loss = self.packed[off:off + 1].view(())
        self._grads = grads
        # buffers of the call
        opt = make_options(self.params, self.band)
        self._opt = opt
        n_lights = int(m.floats[m.i_light_pos].shape[0])
        self.ws_bytes = lib().surf_workspace_bytes_ex(m.total_prims, n, n_lights, int(opt.shadow), m.proj, 1)
        self.workspace = torch.empty(self.ws_bytes, dtype=torch.uint8, device=dev)
        self.depth = torch.empty(n, dtype=torch.float32, device=dev)
        self.nearest = torch.empty(n, dtype=torch.int64, device=dev)
        keep = set(keep_outputs or ())
        self.image = torch.empty(n, 3, dtype=torch.float32, device=dev) if 'image' in keep else None
        self.normal = torch.empty(n, 3, dtype=torch.float32, device=dev) if 'normal' in keep else None
        self.pos = torch.empty(n, 3, dtype=torch.float32, device=dev) if 'pos' in keep else None
        ptr = lambda t: None if t is None else t.data_ptr()      # noqa: E731
        self._out = _abi.SurfOutputs(ptr(self.image), self.depth.data_ptr(), ptr(self.normal), ptr(self.pos),
                                     self.nearest.data_ptr(), None)
        self._step = _abi.SurfStepMSE(self.target.data_ptr(), 1.0 / (3.0 * n_total), self.loss.data_ptr(), None)
        self._sc, self._cam = m.c_scene(), m.c_camera()
        self._sg = m.c_grads(grads)
        self.launches = 0

    @staticmethod
    def _scene_tensors(scene):
        out = []

        def walk(v):
            if isinstance(v, torch.Tensor):
                out.append(v)
            elif isinstance(v, dict):
                for x in v.values():
                    walk(x)
        walk(scene)
        return out

    def __call__(self):
        """Enqueue one step on the current stream; returns the loss (0-d tensor, a view into the packed buffer)."""
        self.packed.zero_()
        with torch.cuda.device(self.device):
            check(lib().surf_step_mse(C.byref(self._sc), C.byref(self._cam), C.byref(self._opt), self.workspace.data_ptr(),
                                      self.ws_bytes, C.byref(self._out), C.byref(self._step), C.byref(self._sg), _stream_ptr()))
        self.launches = lib().surf_last_launch_count()
        if self.world > 1:
            dist.all_reduce(self.packed, op=dist.ReduceOp.SUM, group=self.group)
        for t, i in zip(self.leaves, self.leaf_slots):      # an optimizer with set_to_none=True may have dropped them
            if t.grad is not self._grads[i]:
                t.grad = self._grads[i]
        return self.loss

    def gather_image(self):
        """The full frame [H, W, 3] of the last step (all_gather of the bands; not part of the step)."""
        if self.image is None:
            raise RuntimeError("MSEStep was built without keep_outputs=('image',)")
        if self.world == 1:
            return self.image.view(self.H, self.W, 3)
        n_total = self.H * self.W
        sizes = [_band_range(n_total, r, self.world) for r in range(self.world)]
        mx = max(b - a for a, b in sizes)
        pad = self.image if self.n == mx else torch.cat((self.image, self.image.new_zeros(mx - self.n, 3)))
        parts = torch.empty(self.world, mx, 3, dtype=torch.float32, device=self.device)
        dist.all_gather_into_tensor(parts, pad.contiguous(), group=self.group)
        return torch.cat([parts[r, :b - a] for r, (a, b) in enumerate(sizes)], dim=0).view(self.H, self.W, 3)


class PackedAdam:
    """optimizer.step() of the reference loop (test_optimization.py:122: torch.optim.Adam, no amsgrad / weight decay) for
    the leaves of an MSEStep, whose gradients are views of the plan's packed buffer: ONE kernel over the packed layout
    (surf_adam_step) instead of torch's multi-tensor launch.  Step count and bias corrections live on the device, so
    plan() + opt.step() captures into a CUDA graph.

        plan = MSEStep(scene, target);  opt = PackedAdam(plan, lr=1e-3)
        for it in range(300):  loss = plan();  opt.step()
    """

    def __init__(self, plan, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        if len(plan.leaves) > _abi.SURF_ADAM_MAX_TENSORS:
            raise ValueError('PackedAdam handles at most %d leaves' % _abi.SURF_ADAM_MAX_TENSORS)
        self.plan = plan
        self.lr, self.betas, self.eps = float(lr), (float(betas[0]), float(betas[1])), float(eps)
        n = plan.packed.numel() - 1                              # (the last float of the packed buffer is the loss)
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=plan.device)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=plan.device)
        self.state = torch.zeros(4, dtype=torch.float32, device=plan.device)
        self._tn = _abi.SurfAdamTensors()
        self._tn.count = len(plan.leaves)
        for j, t in enumerate(plan.leaves):
            self._tn.param[j] = t.data_ptr()
            self._tn.size[j] = t.numel()

    def step(self):
        with torch.cuda.device(self.plan.device):
            check(lib().surf_adam_step(C.byref(self._tn), self.plan.packed.data_ptr(), self.exp_avg.data_ptr(),
                                       self.exp_avg_sq.data_ptr(), self.state.data_ptr(), self.lr, self.betas[0],
                                       self.betas[1], self.eps, _stream_ptr()))

    def zero_grad(self, set_to_none=False):
        self.plan.packed.zero_()


def render_mse_step(scene, target, group=None, **params):
    """One-off form of MSEStep: builds the plan, runs one step, returns (loss, plan).  Prefer keeping the plan."""
    plan = MSEStep(scene, target, group=group, **params)
    return plan(), plan
