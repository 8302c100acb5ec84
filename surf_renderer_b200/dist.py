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
