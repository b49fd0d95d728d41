"""Multi-GPU use of the hot path: one process and one native context per GPU.

* ``shard_bounds`` / ``search_sharded``: bsdiff match search sharded by new-data range over a REPLICATED suffix
  array.  Rank ``src`` sorts ``old`` (or supplies I), the index is broadcast, every rank answers its contiguous
  slice of scan positions, the (pos, len) slices are all-gathered.  No collective sits inside the search itself.
  With the NCCL backend the index and the results stay on the device (device-pointer C ABI); with gloo (the CPU
  tests) they travel as host tensors.
* Independent (old, new) pairs need no code here: each rank calls ``bsdiff.create_streams`` on its own objects
  (that is what ``bench.py --gpus N`` measures).
"""
import numpy as np

from .suffix_sort import as_bytes_array


def shard_bounds(m, world, rank):
    """Contiguous, balanced [begin, end) of `m` scan positions for `rank` of `world`."""
    base, extra = divmod(m, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def search_sharded(old, new, suffix_sort, group=None, src=0):
    """(pos, len) of Diff.Search at every scan position of `new`, computed by all ranks of `group`.

    Every rank passes the same `old` and `new` (host buffers) and its own CudaSuffixSort; every rank returns
    the full arrays."""
    import torch
    import torch.distributed as dist

    o = as_bytes_array(old, "oldData")
    w = as_bytes_array(new, "newData")
    n, m = int(o.size), int(w.size)
    ctx = suffix_sort.context
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        I = np.zeros(n + 1, dtype=np.int32)
        suffix_sort.sort(o, I[:n])
        pos = np.empty(m, dtype=np.int32)
        ln = np.empty(m, dtype=np.int32)
        ctx.bsdiff_search(o, None, w, 0, m, pos, ln)
        return pos, ln

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    begin, end = shard_bounds(m, world, rank)
    count = end - begin
    maxc = (m + world - 1) // world
    on_device = dist.get_backend(group) == "nccl"

    if on_device:
        dev = torch.device("cuda", torch.cuda.current_device())
        d_old = torch.from_numpy(o).to(dev)
        d_new = torch.from_numpy(w).to(dev)
        d_I = torch.zeros(max(n, 1), dtype=torch.int32, device=dev)
        torch.cuda.synchronize()
        if rank == src and n:
            ctx.suffix_sort_device(d_old.data_ptr(), n, d_I.data_ptr())
        dist.broadcast(d_I, src=src, group=group)           # replicate the suffix array (4n bytes per rank)
        out = torch.zeros(2, maxc, dtype=torch.int32, device=dev)
        torch.cuda.synchronize()
        if count:
            ctx.bsdiff_search_device(d_old.data_ptr(), n, d_I.data_ptr(), d_new.data_ptr(), m, begin, count,
                                     out[0].data_ptr(), out[1].data_ptr())
        parts = [torch.empty_like(out) for _ in range(world)]
        dist.all_gather(parts, out, group=group)
        parts = [p.cpu().numpy() for p in parts]
    else:
        I = np.zeros(n + 1, dtype=np.int32)
        if rank == src:
            suffix_sort.sort(o, I[:n])
        t_I = torch.from_numpy(I)
        dist.broadcast(t_I, src=src, group=group)
        out = np.zeros((2, maxc), dtype=np.int32)
        if count:
            pos = np.empty(count, dtype=np.int32)
            ln = np.empty(count, dtype=np.int32)
            ctx.bsdiff_search(o, I, w, begin, count, pos, ln)
            out[0, :count] = pos
            out[1, :count] = ln
        t_out = torch.from_numpy(out)
        parts = [torch.empty_like(t_out) for _ in range(world)]
        dist.all_gather(parts, t_out, group=group)
        parts = [p.numpy() for p in parts]

    pos_all = np.empty(m, dtype=np.int32)
    len_all = np.empty(m, dtype=np.int32)
    for r in range(world):
        b, e = shard_bounds(m, world, r)
        pos_all[b:e] = parts[r][0, :e - b]
        len_all[b:e] = parts[r][1, :e - b]
    return pos_all, len_all
