"""Multi-GPU sharding of the render path: one process per GPU (torch.distributed, NCCL over NVLink).

Pixels are independent (reference: README.md:81-84, renderer.py:170-198), so a frame shards into contiguous
row-major pixel bands with replicated scene inputs and no halo:

  * tiles over GPUs  - rank r renders flat pixels [r*N/G, (r+1)*N/G) against all primitives; the output bands
    are all-gathered, and after backward the partial scene gradients are summed with ONE all-reduce over a
    single packed fp32 buffer.  No collective sits inside the intersection kernel's data path.
  * scenes over GPUs - a batch of independent scenes is split over ranks (render_batch_sharded: contiguous blocks of
    a stacked batch; shard_scenes / gather_scene_outputs: round-robin helpers for lists); outputs are all-gathered,
    gradients of parameters shared by all scenes are all-reduced.

The collectives are plain torch.distributed calls (payloads are a few MB; latency-bound).  The render function is
injectable so the host-side logic is testable on CPU with gloo (tests use the oracle as the stand-in).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def band_range(n_pixels, rank, world):
    """Contiguous, balanced split of the flat pixel range; the first n_pixels % world ranks get one more."""
    base, rem = divmod(n_pixels, world)
    p0 = rank * base + min(rank, rem)
    return p0, p0 + base + (1 if rank < rem else 0)


class _GatherBands(torch.autograd.Function):
    """all_gather of per-rank bands along dim 0; backward hands each rank the slice of its own band.
    Every rank evaluates the same loss on the gathered frame, so no reduction is needed here - the partial
    parameter gradients are summed later by allreduce_gradients()."""

    @staticmethod
    def forward(ctx, band, sizes, rank, group):
        ctx.sizes, ctx.rank = sizes, rank
        world = len(sizes)
        if world == 1:
            return band.clone()
        if min(sizes) == max(sizes):          # equal bands: gather straight into the full-frame tensor
            out = band.new_empty((sum(sizes),) + tuple(band.shape[1:]))
            dist.all_gather_into_tensor(out, band.contiguous(), group=group)
            return out
        mx = max(sizes)
        pad = band
        if band.shape[0] < mx:
            pad = torch.cat((band, band.new_zeros((mx - band.shape[0],) + tuple(band.shape[1:]))), dim=0)
        parts = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad.contiguous(), group=group)
        return torch.cat([p[:s] for p, s in zip(parts, sizes)], dim=0)

    @staticmethod
    def backward(ctx, g):
        lo = sum(ctx.sizes[:ctx.rank])
        return g[lo:lo + ctx.sizes[ctx.rank]].contiguous(), None, None, None


def render_bands(scene, render_flat_fn=None, group=None, gather=('image', 'depth', 'nearest'), **params):
    """Tiles-over-GPUs render of one frame.  Returns the reference's output dict with the gathered full-frame
    tensors for the keys in `gather` (others hold this rank's band, flat)."""
    if render_flat_fn is None:
        from .renderer import render_flat as render_flat_fn
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    vp = scene['camera']['viewport']
    W, H = int(vp[2] - vp[0]), int(vp[3] - vp[1])
    n = W * H
    p0, p1 = band_range(n, rank, world)
    (image, depth, normal, pos, nearest, ray_dir), _ = render_flat_fn(scene, (p0, p1), **params)
    local = {'image': image, 'depth': depth, 'normal': normal, 'pos': pos, 'nearest': nearest}
    sizes = [band_range(n, r, world)[1] - band_range(n, r, world)[0] for r in range(world)]
    out = {'band': (p0, p1), 'ray_dir': ray_dir, 'ray_dist': None}
    shapes = {'image': (H, W, 3), 'depth': (H, W), 'normal': (H, W, 3), 'pos': (H, W, 3), 'nearest': (H, W)}
    for k, v in local.items():
        if k in gather:
            full = _GatherBands.apply(v, sizes, rank, group)
            out[k] = full.view(*shapes[k])
        else:
            out[k] = v
    return out


def render_batch_sharded(scene, render_batch_fn=None, group=None, gather=('image', 'depth', 'nearest'), **params):
    """Scenes-over-GPUs render of a batched scene dict (render_batch's stacked form; BASELINE config D: 64 scenes
    over 8 GPUs).  Rank r renders the contiguous block band_range(B, r, world) of scenes; the outputs named in
    `gather` come back as full [B, H, W, ...] tensors on every rank (one all_gather each), the others hold this
    rank's block.  Differentiable like render_bands: each rank back-propagates through its own block; gradients of
    tensors shared by all scenes are partial per rank - sum them with allreduce_gradients()."""
    from .marshal import batched_scene_size, select_scenes
    if render_batch_fn is None:
        from .renderer import render_batch as render_batch_fn
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    B = batched_scene_size(scene)
    if B is None:
        raise ValueError('render_batch_sharded needs a batched scene dict')
    if B < world:
        raise ValueError('fewer scenes (%d) than ranks (%d)' % (B, world))
    b0, b1 = band_range(B, rank, world)
    res = render_batch_fn(select_scenes(scene, slice(b0, b1)), **params)
    sizes = [band_range(B, r, world)[1] - band_range(B, r, world)[0] for r in range(world)]
    out = {'block': (b0, b1), 'ray_dist': None}
    for k, v in res.items():
        out[k] = _GatherBands.apply(v, sizes, rank, group) if (k in gather and v is not None) else v
    return out


def allreduce_gradients(tensors, group=None):
    """Sum the .grad of `tensors` over ranks with one all-reduce of a packed fp32 buffer."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    grads = [t.grad for t in tensors if t.grad is not None]
    if world == 1 or not grads:
        return 0
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    off = 0
    for g in grads:
        g.copy_(flat[off:off + g.numel()].view_as(g))
        off += g.numel()
    return flat.numel() * 4


class GradBucket:
    """The .grad of `tensors` as views of ONE flat fp32 buffer, so that the gradient all-reduce is a single in-place
    collective with no packing (torch.cat) before it and no copy-back after it.  autograd accumulates into an existing
    .grad in place, so the views stay attached as long as zero_grad(set_to_none=False) / bucket.zero_() is used."""

    def __init__(self, tensors):
        self.tensors = [t for t in tensors]
        dev = self.tensors[0].device
        self.flat = torch.zeros(sum(t.numel() for t in self.tensors), dtype=torch.float32, device=dev)
        self.views, off = [], 0
        for t in self.tensors:
            v = self.flat[off:off + t.numel()].view_as(t)
            t.grad = v
            self.views.append(v)
            off += t.numel()

    def zero_(self):
        self.flat.zero_()
        for t, v in zip(self.tensors, self.views):
            if t.grad is not v:
                t.grad = v

    def all_reduce(self, group=None):
        if dist.is_initialized() and dist.get_world_size(group) > 1:
            for t, v in zip(self.tensors, self.views):      # a .grad that was re-created by autograd: fold it back in
                if t.grad is not v:
                    if t.grad is not None:
                        v.copy_(t.grad)
                    t.grad = v
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
        return self.flat.numel() * 4


class ShardedBatchStep:
    """Scenes-over-GPUs step on a stacked batch (BASELINE configs[3]; the reference renders such a batch in
    a Python loop of render() calls, GAN/gan.py:326-377, and feeds the frames to a loss network).

    Built ONCE from a batched scene dict (render_batch's stacked form; its batched tensors may live in pinned host
    memory).  Rank r owns the contiguous block band_range(B, r, G) of scenes: the block's batched tensors are copied to
    the device and become the plan's leaves (`plan.leaves`: name -> tensor), shared tensors named in `shared_leaves`
    become leaves whose gradients are all-reduced.  One step is

        forward of the block (surf_forward_strided, five launches)
        ONE all_gather of the block's images into the full [B, H, W, 3] batch
        loss = loss_fn(images)                         # the caller's loss on the whole batch (every rank)
        backward of the block (surf_backward_strided) from d(loss)/d(images)[block]
        ONE all_reduce over the packed gradients of the shared leaves

    with the marshalling, the workspace, the output and gradient buffers reused by every step.  Per-scene leaves (splat
    positions, normals) get their gradients locally - no collective touches them."""

    def __init__(self, scene, device=None, group=None, block_leaves=('objects/disk/pos', 'objects/disk/normal'),
                 shared_leaves=('lights/pos',), **params):
        from .marshal import MarshalledBatch, batched_scene_size, make_options, select_scenes
        from .renderer import _resolve_device
        self.group = None if group in (None, True) else group
        self.world = dist.get_world_size(self.group) if (group is not None and dist.is_initialized()) else 1
        self.rank = dist.get_rank(self.group) if self.world > 1 else 0
        B = batched_scene_size(scene)
        if B is None:
            raise ValueError('ShardedBatchStep needs a batched scene dict')
        if B % self.world:
            raise ValueError('%d scenes do not split evenly over %d ranks' % (B, self.world))
        self.B = B
        self.block = band_range(B, self.rank, self.world)
        b0, b1 = self.block
        dev = torch.device(device) if device is not None else _resolve_device(scene)
        self.device = dev
        self.params = dict(params)
        # this rank's block on the device; remember which tensors came from host memory for upload()
        self._uploads = []
        local = select_scenes(scene, slice(b0, b1))

        def to_dev(path, v):
            if not isinstance(v, torch.Tensor):
                return v
            src = v
            t = v.detach().to(dev, non_blocking=True)
            if t is v.detach() or t.data_ptr() == v.data_ptr():
                t = t.clone()
            if not src.is_cuda and src.is_floating_point():
                self._uploads.append((t, src))
            return t
        self.scene = {}
        for key, val in local.items():
            if key == 'objects':
                self.scene[key] = {kind: {f: to_dev('objects/%s/%s' % (kind, f), v) for f, v in prim.items()} for kind, prim in val.items()}
            elif isinstance(val, dict):
                self.scene[key] = {f: to_dev('%s/%s' % (key, f), v) for f, v in val.items()}
            else:
                self.scene[key] = to_dev(key, val)

        def at(path):
            node = self.scene
            for part in path.split('/'):
                node = node[part]
            return node
        self.leaves = {}
        for path in tuple(block_leaves) + tuple(shared_leaves):
            t = at(path)
            t.requires_grad_(True)
            self.leaves[path] = t
        self.mb = MarshalledBatch(self.scene, dev)
        m = self.mb.m
        self.H, self.W = m.height, m.width
        n = m.n_pixels
        nb = b1 - b0
        self.nb = nb
        self._opt = make_options(self.params)
        from ._lib import lib
        self.ws_bytes = (lib().surf_workspace_bytes_ex(m.total_prims, n, int(m.floats[m.i_light_pos].shape[0]), int(self._opt.shadow), m.proj, 0) + 255) // 256 * 256
        self.workspace = torch.empty(nb, self.ws_bytes, dtype=torch.uint8, device=dev)
        self.image_block = torch.empty(nb, n, 3, dtype=torch.float32, device=dev)
        self.image_full = torch.empty(B, n, 3, dtype=torch.float32, device=dev) if self.world > 1 else self.image_block
        self.depth = torch.empty(nb, n, dtype=torch.float32, device=dev)
        self.nearest = torch.empty(nb, n, dtype=torch.int64, device=dev)
        # gradient buffers: block leaves local, shared leaves in one bucket for the all-reduce
        fulls = self.mb.fulls
        self._grads = [None] * len(fulls)
        shared = [self.leaves[p] for p in shared_leaves]
        self.bucket = GradBucket(shared) if shared else None
        self._block_grads = []
        for i, t in enumerate(fulls):
            for path, leaf in self.leaves.items():
                if t is leaf:
                    if path in shared_leaves:
                        self._grads[i] = leaf.grad
                    else:
                        g = torch.zeros_like(leaf)
                        leaf.grad = g
                        self._grads[i] = g
                        self._block_grads.append(g)
        missing = [p for p, leaf in self.leaves.items() if not any(t is leaf for t in fulls)]
        if missing:
            raise ValueError('ShardedBatchStep: leaves were copied while marshalling: %s' % missing)
        self.launches = 0
        self.upload_bytes = sum(t.numel() * t.element_size() for t, _ in self._uploads)

    def upload(self):
        """copy the block's host-resident inputs to the device again (the per-step H2D of an end-to-end measurement)"""
        with torch.no_grad():
            for t, src in self._uploads:
                t.copy_(src, non_blocking=True)

    def _structs(self):
        if getattr(self, '_cached', None) is None:          # the leaves are updated in place: pointers never change
            m, mb = self.mb.m, self.mb
            self._views = mb.views0(mb.fulls)
            self._cached = (m.c_scene(self._views), m.c_camera(), mb.c_layout())
        return self._cached

    def forward(self):
        import ctypes as C
        from . import _abi
        from ._lib import check, lib
        from .renderer import _stream_ptr
        sc, cam, lay = self._structs()
        out = _abi.SurfOutputs(self.image_block.data_ptr(), self.depth.data_ptr(), None, None, self.nearest.data_ptr(), None)
        with torch.cuda.device(self.device):
            check(lib().surf_forward_strided(self.nb, C.byref(sc), C.byref(cam), C.byref(lay), C.byref(self._opt),
                                             self.workspace.data_ptr(), self.ws_bytes, C.byref(out), _stream_ptr()))
        self.launches = lib().surf_last_launch_count()
        if self.world > 1:
            dist.all_gather_into_tensor(self.image_full, self.image_block, group=self.group)
        return self.image_full.view(self.B, self.H, self.W, 3)

    def backward(self, g_image_full):
        import ctypes as C
        from . import _abi
        from ._lib import check, lib
        from .renderer import _stream_ptr
        b0, b1 = self.block
        g = g_image_full.reshape(self.B, -1, 3)[b0:b1]
        if not g.is_contiguous():
            g = g.contiguous()
        for gb in self._block_grads:
            gb.zero_()
        if self.bucket is not None:
            self.bucket.zero_()
        sc, cam, lay = self._structs()
        og = _abi.SurfOutGrads(g.data_ptr(), None, None, None)
        sg = self.mb.m.c_grads(self._grads)
        opt = self._opt
        opt.forced_nearest = 2
        with torch.cuda.device(self.device):
            check(lib().surf_backward_strided(self.nb, C.byref(sc), C.byref(cam), C.byref(lay), C.byref(opt),
                                              self.workspace.data_ptr(), self.ws_bytes, self.nearest.data_ptr(),
                                              self.depth.data_ptr(), C.byref(og), C.byref(sg), _stream_ptr()))
        opt.forced_nearest = 0
        self.launches += lib().surf_last_launch_count()
        if self.bucket is not None:
            self.bucket.all_reduce(self.group)

    def step(self, loss_fn):
        """forward -> all_gather -> loss_fn(images [B,H,W,3]) -> backward -> all_reduce; returns the loss tensor"""
        images = self.forward().detach().requires_grad_(True)
        loss = loss_fn(images)
        (g,) = torch.autograd.grad(loss, images)
        self.backward(g)
        return loss.detach()


def shard_scenes(n_scenes, rank, world):
    """Scenes-over-GPUs (GAN batch, gan.py:326-377): scene b -> rank b mod world."""
    return list(range(rank, n_scenes, world))


def gather_scene_outputs(local, n_scenes, group=None):
    """all_gather a [n_local, ...] stack of per-scene outputs back into batch order [n_scenes, ...]."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if world == 1:
        return local
    counts = [len(shard_scenes(n_scenes, r, world)) for r in range(world)]
    mx = max(counts)
    pad = local
    if local.shape[0] < mx:
        pad = torch.cat((local, local.new_zeros((mx - local.shape[0],) + tuple(local.shape[1:]))), dim=0)
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad.contiguous(), group=group)
    out = local.new_empty((n_scenes,) + tuple(local.shape[1:]))
    for r in range(world):
        idx = shard_scenes(n_scenes, r, world)
        if idx:
            out[idx] = parts[r][:len(idx)]
    return out
