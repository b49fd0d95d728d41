"""N > 1 paths: one text sorted by all shards of a device group, and the match search
sharded by new-data range over the replicated index -- both inside the library (csrc/dq_group.inl, dq_dist.cuh),
reached through the ordinary C-ABI entry points of a context created over several devices.

CPU (`-m "not gpu"`): the logic emulator (tests/emu) with 2..8 logical shards.
GPU (`-m gpu`): the real library; on a one-GPU box the shards are logical (the same ordinal listed several times),
so every kernel and every exchange of the group path runs there too; with >= 2 GPUs (gpurun --gpus 2) the exchanges
cross NVLink.  DQ_SHARD_MIN=1 makes small inputs take the sharded path."""
import os
import sys

import numpy as np
import pytest

import oracle
from conftest import adversarial_texts, load_asset, random_bytes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _sort_texts():
    t = adversarial_texts()
    out = {k: t[k] for k in ("zeros_tail", "all_zero_1000", "period8_zero_end", "fibonacci", "binary_random",
                             "repeated_paragraph", "period3_tail0", "runs_mixed", "runs_long_zero_islands")}
    out["fuzz3"] = load_asset("fuzz3")
    out["crash-gosais"] = load_asset("crash-gosais-force-alloc")
    out["random_20000"] = random_bytes(20000)
    out["tiny_5"] = np.array([3, 1, 2, 1, 0], np.uint8)
    out["two_equal"] = np.array([1, 1], np.uint8)
    out["single"] = np.array([9], np.uint8)
    out["empty"] = np.zeros(0, np.uint8)
    return out


def _pairs():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from search_cases import structured_pairs
    p = structured_pairs()
    return {k: p[k] for k in ("zero_runs_with_islands", "point_edits", "periodic_records", "new_is_suffix_of_old",
                              "empty_old")}


@pytest.fixture()
def shard_everything(monkeypatch):
    monkeypatch.setenv("DQ_SHARD_MIN", "1")
    monkeypatch.setenv("DQ_SUB_MIN_LOG", "3")   # sub-ranges of the position exchanges from 8 positions up
    monkeypatch.setenv("DQ_EARLY_COPY_MIN", "1")  # early copy of the suffix array whatever the text length


def _check_sort(sorter):
    for name, t in _sort_texts().items():
        sa = np.full(t.size + 1, -7, dtype=np.int32)
        sorter.sort(t, sa[:t.size])
        assert sa[t.size] == -7, name
        assert np.array_equal(sa[:t.size], oracle.sais(t)), name


def _check_search(sorter):
    from deltaq_b200.parallel import search_sharded
    for name, (old, new) in _pairs().items():
        I = oracle.make_I(oracle.sais(old))
        rp, rl = oracle.search_all(I, old, new)
        pos, ln = search_sharded(old, new, sorter)             # group sort, then search over the replicated index
        assert np.array_equal(pos, rp) and np.array_equal(ln, rl), name
        pos, ln = search_sharded(old, new, sorter, I=I)        # caller-supplied suffix array
        assert np.array_equal(pos, rp) and np.array_equal(ln, rl), name
        if old.size:                                           # the LCP array of the group-sorted text (dq_cuda_lcp)
            sa = np.empty(old.size, np.int32)
            sorter.sort(old, sa)
            assert np.array_equal(sorter.lcp_array(old), oracle.lcp_array(old, sa)), name
        b, c = new.size // 3, new.size // 2                    # a sub-range of scan positions
        p3 = np.empty(c, np.int32)
        l3 = np.empty(c, np.int32)
        sorter.context.bsdiff_search(old, I, new, b, c, p3, l3)
        assert np.array_equal(p3, rp[b:b + c]) and np.array_equal(l3, rl[b:b + c]), name


def test_shard_bounds_cover_everything():
    from deltaq_b200.parallel import shard_bounds
    for m in (0, 1, 7, 64, 1000):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(m, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == m
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(e - b for b, e in spans) - min(e - b for b, e in spans) <= 1


# DQ_DIRECT_MAX: rounds with at most that many unresolved suffixes read and write ISA through peer pointers, larger
# ones use the request / reply / update exchanges -- 0: exchanges only, 3000: both kinds in one sort, default: all of
# these small texts go the direct way
@pytest.mark.parametrize("shards,direct_max", [(2, "0"), (3, "3000"), (8, "0"), (8, None), (5, "3000")])
def test_group_sort_emulator(shard_everything, monkeypatch, shards, direct_max):
    import emu
    from deltaq_b200 import CudaSuffixSort
    if direct_max is not None:
        monkeypatch.setenv("DQ_DIRECT_MAX", direct_max)
    with CudaSuffixSort(device=[0] * shards, _lib=emu.library()) as sorter:
        _check_sort(sorter)


def test_group_sort_small_alphabets_emulator(shard_everything, monkeypatch):
    """Recoded keys (dq_suffix.cuh, "small alphabets") through the sharded path: one code for all shards, 64-character
    halos."""
    import emu
    from conftest import small_alphabet_texts
    from deltaq_b200 import CudaSuffixSort
    monkeypatch.setenv("DQ_COMPACT_MIN", "1")
    monkeypatch.setenv("DQ_DIRECT_MAX", "2000")
    with CudaSuffixSort(device=[0] * 3, _lib=emu.library()) as sorter:
        for name, t in small_alphabet_texts().items():
            sa = np.empty(t.size, np.int32)
            sorter.sort(t, sa)
            assert np.array_equal(sa, oracle.sais(t)), name


def test_group_search_emulator(shard_everything):
    import emu
    from deltaq_b200 import CudaSuffixSort
    with CudaSuffixSort(device=[0, 0, 0], _lib=emu.library()) as sorter:
        _check_search(sorter)


def _check_streams(sorter):
    """dq_cuda_bsdiff_streams on a device group: sort by all shards, search by all shards into shard 0's table, then the
    coding and the host loop -- streams identical to the oracle's."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from search_cases import structured_pairs
    for name, (old, new) in structured_pairs().items():
        got = sorter.context.bsdiff_streams(old, new)
        ref = oracle.bsdiff_streams(old, new)
        for k in ("ctrl", "diff", "extra"):
            assert got[k] == ref[k], (name, k)
        assert got["search_visits"] == ref["search_calls"], name


def test_group_streams_emulator(shard_everything):
    import emu
    from deltaq_b200 import CudaSuffixSort
    with CudaSuffixSort(device=[0, 0, 0], _lib=emu.library()) as sorter:
        _check_streams(sorter)


def test_small_inputs_stay_on_one_device():
    """Below DQ_SHARD_MIN a group context behaves like a single-device one (BASELINE: single-GPU-sized inputs stay
    on one GPU)."""
    import emu
    from deltaq_b200 import CudaSuffixSort
    t = random_bytes(5000)
    with CudaSuffixSort(device=[0, 0], _lib=emu.library()) as sorter:
        sa = np.empty(t.size, np.int32)
        sorter.sort(t, sa)
        assert np.array_equal(sa, oracle.sais(t))
        # the plain path keeps the index on shard 0: a search with I=None finds it
        pos = np.empty(t.size, np.int32)
        ln = np.empty(t.size, np.int32)
        sorter.context.bsdiff_search(t, None, t, 0, t.size, pos, ln)
        assert np.array_equal(ln, np.arange(t.size, 0, -1, dtype=np.int32))


def test_too_many_devices_is_an_error():
    import emu
    from deltaq_b200 import CudaSuffixSort, _native
    with pytest.raises(_native.NativeError) as ei:
        CudaSuffixSort(device=[0] * 17, _lib=emu.library())
    assert ei.value.status == _native.DQ_ERR_INVALID_ARGUMENT


# ---- the real library ----------------------------------------------------------------------------------------

def _devices(shards):
    import torch
    k = torch.cuda.device_count()
    return [i % k for i in range(shards)]


@pytest.mark.gpu
@pytest.mark.parametrize("shards,direct_max", [(2, "0"), (3, "3000"), (8, "0"), (8, None)])
def test_group_sort_gpu(shard_everything, monkeypatch, shards, direct_max):
    """Logical shards on a one-GPU box; real peers (NVLink) when the box has several GPUs."""
    from deltaq_b200 import CudaSuffixSort
    if direct_max is not None:
        monkeypatch.setenv("DQ_DIRECT_MAX", direct_max)
    with CudaSuffixSort(device=_devices(shards)) as sorter:
        _check_sort(sorter)


@pytest.mark.gpu
def test_group_sort_small_alphabets_gpu(shard_everything, monkeypatch):
    from conftest import small_alphabet_texts
    from deltaq_b200 import CudaSuffixSort
    monkeypatch.setenv("DQ_COMPACT_MIN", "1")
    monkeypatch.setenv("DQ_DIRECT_MAX", "2000")
    with CudaSuffixSort(device=_devices(3)) as sorter:
        for name, t in small_alphabet_texts().items():
            sa = np.empty(t.size, np.int32)
            sorter.sort(t, sa)
            assert np.array_equal(sa, oracle.sais(t)), name


@pytest.mark.gpu
def test_group_search_gpu(shard_everything):
    from deltaq_b200 import CudaSuffixSort
    with CudaSuffixSort(device=_devices(3)) as sorter:
        _check_search(sorter)


@pytest.mark.gpu
def test_group_streams_gpu(shard_everything):
    from deltaq_b200 import CudaSuffixSort
    with CudaSuffixSort(device=_devices(4)) as sorter:
        _check_streams(sorter)


@pytest.mark.gpu
def test_group_sort_gpu_midsize(shard_everything, monkeypatch):
    """A few MiB per shard: many tiles per pass, several doubling rounds, exchanges of millions of entries."""
    from deltaq_b200 import CudaSuffixSort
    rng = np.random.default_rng(11)
    texts = {
        "acgt_6M": np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, 6_000_000)],
        "uniform_5M": rng.integers(0, 256, 5_000_001, dtype=np.uint8),
        "repeats_4M": np.tile(rng.integers(0, 256, 70_001, dtype=np.uint8), 60)[:4_000_000],
    }
    monkeypatch.setenv("DQ_DIRECT_MAX", str(1 << 20))   # big rounds by exchange, the small ones after them direct
    with CudaSuffixSort(device=_devices(4)) as sorter:
        for name, t in texts.items():
            sa = np.empty(t.size, np.int32)
            sorter.sort(t, sa)
            assert oracle.sufcheck(t, sa) == 0, name
            assert sorter.stats()["n"] == t.size
