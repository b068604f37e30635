"""Multi-GPU driver: one process per GPU (torchrun), torch.distributed for the plumbing (NCCL over
NVLink on the B200 box, gloo in the CPU tests).

The reference is a single process; the render path shards naturally
 because every (pixel, sample) and every photon path owns its random stream:

  * photons   rank r traces paths [r*n/G, (r+1)*n/G) of every light; the stored particles are
              all-gathered (variable counts) and concatenated in (light, path) order, so every rank
              ends up with the list -- and therefore the kd-tree -- the single-GPU run builds;
  * pixels    interleaved 16x16 tiles round-robin over ranks ("tile": bit-identical to 1 GPU, the reduce
              adds zeros) or sample-index ranges ("sample": each rank renders all pixels for its samples);
  * frame     the fp32 per-pixel sums and the int32 hit counters are sum-reduced to rank 0, which
              composites over the background on its GPU (Renderer.cpp:262-265) and writes the PPM.

Nothing here computes on the CPU: the functions move torch tensors and call the C ABI.
"""
from __future__ import annotations

import numpy as np


def path_range(per_light: int, rank: int, world: int):
    """Photon paths [first, first+count) of every light traced by `rank`."""
    first = per_light * rank // world
    return first, per_light * (rank + 1) // world - first


def sample_range(num_rays: int, rank: int, world: int):
    first = num_rays * rank // world
    return first, num_rays * (rank + 1) // world - first


def sample_shard_kwargs(num_rays: int, rank: int, world: int) -> dict:
    """Renderer keyword arguments for `rank`'s share of the sample indices.  rt_params treats
    sample_count <= 0 as "all samples", so a rank whose share is empty (more ranks than samples) is
    given the empty range that starts at num_rays."""
    first, count = sample_range(num_rays, rank, world)
    if count <= 0:
        return dict(sample_first=num_rays, sample_count=0)
    return dict(sample_first=first, sample_count=count)


def gather_photons(local_photons: np.ndarray, per_light_counts: np.ndarray, device=None, group=None) -> np.ndarray:
    """All-gather the shards of the photon list and splice them into (light, path) order.

    local_photons: [n,7] float32 in (light, path) order for this rank's path range; per_light_counts: [L].
    Every rank returns the same [N,7] array: for each light, rank 0's particles, then rank 1's, ...
    """
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    L = len(per_light_counts)
    counts = torch.as_tensor(np.asarray(per_light_counts, np.int64), device=device)
    all_counts = [torch.zeros_like(counts) for _ in range(world)]
    dist.all_gather(all_counts, counts, group=group)
    all_counts = torch.stack(all_counts).cpu().numpy()  # [world, L]
    cap = int(all_counts.sum(axis=1).max())
    buf = torch.zeros((max(cap, 1), 7), dtype=torch.float32, device=device)
    if len(local_photons):
        buf[:len(local_photons)] = torch.as_tensor(np.ascontiguousarray(local_photons, np.float32), device=device)
    gathered = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(gathered, buf, group=group)
    gathered = [g.cpu().numpy() for g in gathered]
    parts = []
    for light in range(L):
        for r in range(world):
            start = int(all_counts[r, :light].sum())
            parts.append(gathered[r][start:start + int(all_counts[r, light])])
    return np.concatenate(parts) if parts else np.zeros((0, 7), np.float32)


def reduce_frame(sum_rgb, counter, dst=0, group=None):
    """Sum-reduce the fp32 sums and int32 counters (torch tensors, any device) to rank `dst`."""
    import torch.distributed as dist

    dist.reduce(sum_rgb, dst, op=dist.ReduceOp.SUM, group=group)
    dist.reduce(counter, dst, op=dist.ReduceOp.SUM, group=group)
    return sum_rgb, counter


def build_photon_map_distributed(renderer, device=None, group=None):
    """Sharded emission + all-gather + kd-tree on every rank (Renderer.cpp:209-213 across G GPUs)."""
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    first, count = path_range(renderer.photons_per_light(), rank, world)
    local, counts, _ = renderer.emit_photons(first, count)
    full = gather_photons(local, counts, device=device, group=group)
    renderer.set_photons(full)
    return full


def render_distributed(scene, num_rays, mode, num_photons=0, k=5, *, background, seed=1, shard="tile", device=None,
                       local_device=0, group=None):
    """The whole path on G ranks.  Returns the composited image [H,W,3] on rank 0 (None elsewhere)."""
    import torch
    import torch.distributed as dist
    from . import Renderer

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    kw = dict(seed=seed, device=local_device)
    if shard == "tile":
        kw.update(shard_rank=rank, shard_count=world)
    else:
        kw.update(sample_shard_kwargs(num_rays, rank, world))
    r = Renderer(scene, num_rays, mode, None, num_photons, k, **kw)
    if num_photons > 0:
        build_photon_map_distributed(r, device=device, group=group)
    H, W = r.height, r.width
    sum_t = torch.zeros((H, W, 3), dtype=torch.float32, device=device)
    cnt_t = torch.zeros((H, W), dtype=torch.int32, device=device)
    if sum_t.is_cuda:
        torch.cuda.synchronize(sum_t.device)  # torch's fill kernels vs the context's own stream
    r.render_accumulate_device(sum_t.data_ptr(), cnt_t.data_ptr())
    reduce_frame(sum_t, cnt_t, 0, group)
    out = None
    if rank == 0:
        if sum_t.is_cuda:  # composite on the device, one D2H of the frame
            # the reduce runs on NCCL's stream and the composite on the context's own: wait for the collective first
            torch.cuda.synchronize(sum_t.device)
            out = r.composite_device(num_rays, sum_t.data_ptr(), cnt_t.data_ptr(), background)
        else:              # gloo tests: host tensors
            out = Renderer.composite(num_rays, sum_t.numpy(), cnt_t.numpy(), background)
    r.close()
    return out
