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

Stream discipline: the library runs on its own CUDA stream and every entry point returns only when its work is
complete, so data it produced is safe to hand to NCCL.  The other direction needs care: a torch tensor that a pending
collective still reads must not be freed (and recycled by torch's allocator for a buffer this library then writes
through a raw pointer) before the collective has finished -- the functions below synchronise the device on every rank
before such a tensor goes out of scope.
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


def gather_photons_device(renderer, device, group=None):
    """The same on the GPU, where the particles are: every rank emits and compacts its share on the device, the
    per-light counts (world x L int64) and then the padded shards themselves are all-gathered with NCCL straight
    from device memory, a splice kernel rearranges them into the single-process (light, path) order, and the list is
    installed with ONE device->host copy (the kd-tree build is libstdc++'s nth_element on the host).
    Returns (list as a [N,7] CUDA tensor, N)."""
    import torch
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    per, L = renderer.photons_per_light(), renderer.scene.L
    first, count = path_range(per, rank, world)
    cap = max(1, max(path_range(per, r, world)[1] for r in range(world)) * max(L, 1))
    local = torch.empty((cap, 7), dtype=torch.float32, device=device)
    counts, _ = renderer.emit_photons_device(first, count, local.data_ptr(), cap)  # synchronous: `local` is complete
    counts_t = torch.as_tensor(np.asarray(counts, np.int64), device=device)
    all_counts = torch.empty((world, max(L, 1)), dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(all_counts, counts_t.reshape(1, -1) if L else torch.zeros((1, 1), dtype=torch.int64, device=device), group=group)
    gathered = torch.empty((world * cap, 7), dtype=torch.float32, device=device)
    dist.all_gather_into_tensor(gathered, local, group=group)
    all_counts = all_counts.cpu().numpy()[:, :L]  # also waits for the collectives (NCCL stream) before the splice
    torch.cuda.synchronize(device)
    total = int(all_counts.sum())
    out = torch.empty((max(total, 1), 7), dtype=torch.float32, device=device)
    got = renderer.splice_photons_device(gathered.data_ptr(), world, cap, all_counts, out.data_ptr(), max(total, 1))
    assert got == total
    return out, total


def reduce_packed(packed, dst=0, group=None):
    """Sum-reduce the packed fp32 frame {sum r, g, b, counter} to rank `dst`: one collective per frame."""
    import torch.distributed as dist

    dist.reduce(packed, dst, op=dist.ReduceOp.SUM, group=group)
    return packed


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
    if device is not None and str(device).startswith("cuda"):  # the particles never leave the GPUs until the kd build
        out, total = gather_photons_device(renderer, device, group)
        renderer.set_photons_device(out.data_ptr(), total)
        return total
    first, count = path_range(renderer.photons_per_light(), rank, world)
    local, counts, _ = renderer.emit_photons(first, count)
    full = gather_photons(local, counts, device=device, group=group)
    renderer.set_photons(full)
    return len(full)


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
    on_gpu = device is not None and str(device).startswith("cuda")
    out = None
    if on_gpu:  # sums and counter as one packed frame: a single reduce, composite on rank 0's GPU, one D2H
        packed = torch.empty((H, W, 4), dtype=torch.float32, device=device)
        torch.cuda.synchronize(device)
        r.render_accumulate_packed_device(packed.data_ptr())
        reduce_packed(packed, 0, group)
        # EVERY rank waits for the collective here.  Rank 0 because the reduce runs on NCCL's stream and the composite
        # on the context's own.  The others because `packed` is freed when this function returns: torch's allocator may
        # hand the block to the next tensor at once (it only orders reuse against torch's streams), and this library
        # writes through raw pointers on its own stream -- a rank that raced ahead would overwrite its contribution
        # while its NCCL kernel is still waiting for the slower ranks (seen on 8 GPUs: one 64 KiB chunk of the frame
        # wrong in 2 of 6 runs of scripts/dist_check.py).
        torch.cuda.synchronize(device)
        if rank == 0:
            out = r.composite_packed_device(num_rays, packed.data_ptr(), background)
    else:       # host tensors (gloo): separate sums and counters
        sum_np, cnt_np = r.render_accumulate()
        sum_t, cnt_t = torch.from_numpy(sum_np), torch.from_numpy(cnt_np)
        reduce_frame(sum_t, cnt_t, 0, group)
        if rank == 0:
            out = Renderer.composite(num_rays, sum_t.numpy(), cnt_t.numpy(), background)
    r.close()
    return out
