"""Multi-GPU use of the hot path.

The sharding lives in the native library: a context created over several devices (``CudaSuffixSort(device=[0, 1, ...])``
-> ``dq_cuda_create(ctx, devices, ndev)``) is a device GROUP driven by one process.  Inputs of at least ``DQ_SHARD_MIN``
bytes (default 128 MiB) are worked on by all of its GPUs through the ordinary entry points:

* ``sort``           one text sorted by all GPUs: distributed prefix doubling, every exchange a partition kernel that
                     scatters straight into peer memory (csrc/dq_dist.cuh, csrc/dq_group.inl);
* ``bsdiff_search``  scan positions sharded by new-data range over the index replicated by peer copies.

Independent (old, new) pairs need no code here: one process and one single-device context per GPU, each calling
``bsdiff.create_streams`` on its own objects (what ``bench.py --gpus N`` reports as ``value`` / ``e2e``).

The helpers below are conveniences over a group context for callers that hold host buffers.
"""
import numpy as np

from .suffix_sort import as_bytes_array


def shard_bounds(m, world, rank):
    """Contiguous, balanced [begin, end) of `m` scan positions for shard `rank` of `world` (the split the library
    uses for the sharded search: csrc/dq_group.inl, group_search)."""
    base, extra = divmod(m, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def suffix_sort_sharded(text, suffix_sort, out=None):
    """Suffix array of `text` by all GPUs of `suffix_sort`'s device group (int32 numpy array; `out`, ideally pinned
    host memory, receives it when given)."""
    t = as_bytes_array(text)
    sa = out[:t.size] if out is not None else np.empty(t.size, dtype=np.int32)
    suffix_sort.sort(t, sa)
    return sa


def search_sharded(old, new, suffix_sort, I=None):
    """(pos, len) of Diff.Search at every scan position of `new`, by all GPUs of `suffix_sort`'s device group.
    With I=None the group sorts `old` first and searches over the index it left on the devices."""
    o = as_bytes_array(old, "oldData")
    w = as_bytes_array(new, "newData")
    ctx = suffix_sort.context
    if I is None:
        sa = np.empty(o.size, dtype=np.int32)
        suffix_sort.sort(o, sa)
    pos = np.empty(w.size, dtype=np.int32)
    ln = np.empty(w.size, dtype=np.int32)
    ctx.bsdiff_search(o, I, w, 0, w.size, pos, ln)
    return pos, ln
