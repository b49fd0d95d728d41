"""Multi-GPU use of the hot path: one process and one native context per GPU.

* ``shard_bounds`` / ``search_sharded``: bsdiff match search sharded by new-data range over a REPLICATED suffix
  array.  Rank ``src`` sorts ``old`` (or supplies I), the index is broadcast, every rank answers its contiguous
  slice of scan positions, the (pos, len) slices are all-gathered.  No collective sits inside the search itself.
  With the NCCL backend the index and the results stay on the device (device-pointer C ABI); with gloo (the CPU
  tests) they travel as host tensors.
* Independent (old, new) pairs need no code here: each rank calls ``bsdiff.create_streams`` on its own objects
  (that is what ``bench.py --gpus N`` measures).
"""
import os

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


# ---------------------------------------------------------------------------------------------------------
# Suffix sort of ONE text sharded across the GPUs of a process group (BASELINE configs
# #4/#5): distributed prefix doubling.
#
#   * key buckets:  round 0 packs 8-byte keys per position slice, an all-reduced 16-bit-prefix histogram gives
#                   every rank the same splitters, one all-to-all-v moves each (key, suffix) tuple to the GPU
#                   that owns its bucket.  Groups never cross buckets, so every later sort is GPU-local
#                   (dq_cuda_dist_round0 / dq_cuda_dist_round: the same onesweep + rank kernels as on one GPU).
#   * ISA by position: rank r owns ISA[r << kb, (r+1) << kb).  Each round the unresolved suffixes ask the
#                   owners for ISA[sa+h] (all-to-all-v request, local gather, all-to-all-v reply) and send the
#                   new ranks back as (position, rank) updates (all-to-all-v, local scatter).  Every random access
#                   is local to one 4n/G-byte slice.
# torch.distributed (NCCL on GPUs, gloo in the CPU tests) is the plumbing; partitions use the library's own
# radix pass (dq_cuda_radix_sort_pairs_device); gathers/scatters around the exchanges are torch indexing ops.

def _bit_length(v):
    return max(1, int(v).bit_length())


class _Exchanger:
    """Collectives of the sharded sort.  NCCL moves device tensors directly; with gloo and device tensors (several
    ranks sharing one GPU, as in the single-GPU test of the multi-rank path) the payload is staged through the host."""

    def __init__(self, group, device):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.device = device
        self.world = dist.get_world_size(group)
        self.stage = device.type == "cuda" and dist.get_backend(group) != "nccl"

    def _in(self, t):
        return t.cpu() if self.stage else t

    def _out(self, t):
        return t.to(self.device) if self.stage else t

    def counts(self, send_counts):
        import torch
        sc = torch.tensor(send_counts, dtype=torch.int64, device="cpu" if self.stage else self.device)
        rc = torch.empty_like(sc)
        self.dist.all_to_all_single(rc, sc, group=self.group)
        return [int(x) for x in rc.tolist()]

    def data(self, t, send_counts, recv_counts):
        import torch
        out = torch.empty(sum(recv_counts), dtype=t.dtype, device="cpu" if self.stage else self.device)
        self.dist.all_to_all_single(out, self._in(t.contiguous()), recv_counts, send_counts, group=self.group)
        return self._out(out)

    def all_reduce(self, t):
        if not self.stage:
            self.dist.all_reduce(t, group=self.group)
            return t
        h = t.cpu()
        self.dist.all_reduce(h, group=self.group)
        t.copy_(h)
        return t

    def all_gather(self, t):
        import torch
        src = self._in(t)
        parts = [torch.empty_like(src) for _ in range(self.world)]
        self.dist.all_gather(parts, src, group=self.group)
        return [self._out(p) for p in parts]


def suffix_sort_sharded(text, suffix_sort, group=None, gather=True, profile=None, out=None):
    """Suffix array of `text`, sorted cooperatively by all ranks of `group` (every rank passes the same text).

    gather=True : returns the full suffix array (numpy int32) on every rank.
    gather=False: returns (slot_base, bucket) -- this rank's bucket covers SA slots [slot_base, slot_base+len);
                  with `out` (an int32 numpy array, ideally pinned: Context.pinned) the bucket is copied into
                  out[:len] and that view is returned.
    profile: optional dict receiving wall seconds per phase (adds synchronisation)."""
    import time as _time

    import torch
    import torch.distributed as dist

    o = as_bytes_array(text)
    n = int(o.size)
    ctx = suffix_sort.context
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        sa = np.empty(n, dtype=np.int32)
        suffix_sort.sort(o, sa)
        return sa if gather else (0, sa)

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if world > 256:
        raise ValueError("suffix_sort_sharded supports at most 256 ranks")
    # device tensors whenever the context is the CUDA library (NCCL, or gloo with host staging); host tensors only for
    # the CPU logic emulator of the tests
    on_gpu = dist.get_backend(group) == "nccl" or os.path.basename(ctx.lib.path) == "libdeltaq_cuda.so"
    nccl = on_gpu
    dev = torch.device("cuda", torch.cuda.current_device()) if on_gpu else torch.device("cpu")
    ex = _Exchanger(group, dev)
    i32, i64 = torch.int32, torch.int64

    def sync():
        if nccl:
            torch.cuda.synchronize()

    _t = [_time.perf_counter()]

    def mark(name):
        if profile is not None:
            sync()
            now = _time.perf_counter()
            profile[name] = profile.get(name, 0.0) + (now - _t[0])
            _t[0] = now

    def empty(count, dtype):
        return torch.empty(max(int(count), 1), dtype=dtype, device=dev)[:int(count)]

    # ownership of text positions (ISA slices): power-of-two slices so that owner(pos) = pos >> kb
    per = (n + world - 1) // world
    kb = 0
    while (1 << kb) < max(per, 1):
        kb += 1
    own_begin = min(n, rank << kb)
    own_end = min(n, (rank + 1) << kb)
    own_cnt = own_end - own_begin

    # ---- round 0: pack my position slice (+ 16-bit prefix histogram), agree on splitters, exchange tuples
    halo = min(n, own_end + 8) - own_begin                       # keys read up to 7 bytes past the slice
    T = torch.zeros(own_cnt + 64, dtype=torch.uint8, device=dev)
    if halo > 0:
        T[:halo] = torch.from_numpy(o[own_begin:own_begin + halo]).to(dev)
    keys = empty(own_cnt, i64)
    vals = empty(own_cnt, i32)
    hist = torch.zeros(65536, dtype=i64, device=dev)
    mark("h2d_text_slice")
    sync()
    ctx.dist_pack(T.data_ptr(), own_begin, own_cnt, keys.data_ptr(), vals.data_ptr(), hist.data_ptr())
    ex.all_reduce(hist)
    cum = torch.cumsum(hist, 0) - hist
    lut64 = torch.clamp((cum * world) // max(n, 1), max=world - 1)
    bucket_cnt = torch.zeros(world, dtype=i64, device=dev).index_add_(0, lut64, hist)
    cnts = [int(x) for x in bucket_cnt.tolist()]
    slot_base = sum(cnts[:rank])
    my_cnt = cnts[rank]
    lut = lut64.to(torch.uint8)
    mark("r0_pack_splitters")

    keys_s = empty(own_cnt, i64)
    vals_s = empty(own_cnt, i32)
    sync()
    counts = ctx.dist_partition(keys.data_ptr(), vals.data_ptr(), own_cnt, lut.data_ptr(), keys_s.data_ptr(),
                                vals_s.data_ptr())
    send_counts = [int(c) for c in counts[:world]]
    recv_counts = ex.counts(send_counts)
    keys_r = ex.data(keys_s, send_counts, recv_counts)
    vals_r = ex.data(vals_s, send_counts, recv_counts)
    assert keys_r.numel() == my_cnt
    del keys, vals, keys_s, vals_s
    # equal keys must arrive in descending suffix order (dq_suffix.cuh, end-of-text rule): sources hold ascending
    # position slices, each chunk is descending inside, so lay the chunks out from the last source to the first
    if my_cnt:
        offs = np.concatenate([[0], np.cumsum(recv_counts)])
        order = [slice(int(offs[g]), int(offs[g + 1])) for g in reversed(range(world))]
        keys_r = torch.cat([keys_r[s] for s in order])
        vals_r = torch.cat([vals_r[s] for s in order])
    mark("r0_partition_exchange")

    sa_local = empty(my_cnt, i32)
    upd_pos = empty(my_cnt, i64)
    upd_rank = empty(my_cnt, i32)
    isa_local = torch.zeros(max(own_cnt, 1), dtype=i32, device=dev)
    sync()
    a = ctx.dist_round0(keys_r.data_ptr(), vals_r.data_ptr(), my_cnt, n, slot_base, sa_local.data_ptr(),
                        upd_pos.data_ptr(), upd_rank.data_ptr())
    del keys_r, vals_r
    mark("r0_local_sort_rank")
    owner_bits = _bit_length(world - 1)

    def route_updates(count):
        """(position, rank) updates -> the GPUs that own the positions; applied to their ISA slices"""
        sync()
        hist_o = ctx.radix_sort_pairs_device(upd_pos.data_ptr(), upd_rank.data_ptr(), count, kb, owner_bits, want_hist=True)
        sc = [int(c) for c in hist_o[:world]]
        rc = ex.counts(sc)
        pos_r = ex.data(upd_pos[:count], sc, rc)
        rank_r = ex.data(upd_rank[:count], sc, rc)
        if pos_r.numel():
            isa_local.index_copy_(0, pos_r - own_begin, rank_r)

    route_updates(my_cnt)
    mark("r0_route_updates")

    # ---- doubling rounds: fetch ISA[sa+h] from the position owners, sort locally, send the new ranks back
    h = 8
    tot = ex.all_reduce(torch.tensor([a], dtype=i64, device=dev))
    rounds = 1
    while int(tot) > 0:
        q = empty(a, i64)
        origin = empty(a, i32)
        sync()
        ctx.dist_requests(h, q.data_ptr(), origin.data_ptr())
        # partition the requests by owner = (q >> kb); a request past the end of the text may land on any rank
        # (digit wraps): whoever gets it answers 0
        hist_o = ctx.radix_sort_pairs_device(q.data_ptr(), origin.data_ptr(), a, kb, 8, want_hist=True)
        sc = [0] * world
        for d_, c_ in enumerate(hist_o):
            if c_:
                sc[min(d_, world - 1)] += int(c_)
        if any(hist_o[world:]):
            # digits >= world only occur for q >= n; keep them contiguous at the end of the last rank's chunk
            pass
        rc = ex.counts(sc)
        q_r = ex.data(q, sc, rc)
        inside = (q_r >= own_begin) & (q_r < own_end)
        idx = torch.clamp(q_r - own_begin, min=0, max=max(own_cnt - 1, 0))
        resp = torch.where(inside, isa_local.index_select(0, idx) + 1, torch.zeros((), dtype=i32, device=dev))
        back = ex.data(resp, rc, sc)
        r2 = empty(a, i32)
        if a:
            r2.index_copy_(0, origin.long(), back)
        prev = a
        mark("rounds_fetch_isa")
        sync()
        a = ctx.dist_round(r2.data_ptr(), n, slot_base, sa_local.data_ptr(), upd_pos.data_ptr(), upd_rank.data_ptr())
        mark("rounds_local_sort_rank")
        route_updates(prev)
        mark("rounds_route_updates")
        h *= 2
        rounds += 1
        tot = ex.all_reduce(torch.tensor([a], dtype=i64, device=dev))
    if profile is not None:
        profile["rounds"] = rounds

    mine = sa_local[:my_cnt]
    if not gather:
        if out is not None:
            t_out = torch.from_numpy(out)[:my_cnt]
            t_out.copy_(mine, non_blocking=True)
            sync()
            mark("d2h_bucket")
            return slot_base, out[:my_cnt]
        res = mine.cpu().numpy()
        mark("d2h_bucket")
        return slot_base, res
    # all-gather-v of the buckets (bucket g owns slots [base_g, base_g + cnt_g))
    maxc = max(max(cnts), 1)
    pad = torch.zeros(maxc, dtype=i32, device=dev)
    pad[:my_cnt] = mine
    parts = ex.all_gather(pad)
    return torch.cat([parts[g][:cnts[g]] for g in range(world)]).cpu().numpy()
