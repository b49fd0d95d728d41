"""BASELINE's large configurations at FULL size on the GPU: C4 (512 MiB genome-like text) and C5
(2,040,109,466-byte executable-like pair, near the int32 suffix-array limit), through the C ABI.

Full-size oracle comparisons are out of reach in minutes, so these use the reference's own O(n) checker (sufcheck =
LDSSChecker.cs:23-119 restated) and Verify (LibDivSufSortTests.cs:43-59) on the suffix arrays, the bspatch round trip
on the delta streams, and a byte-for-byte comparison with the oracle's streams at the largest size the oracle finishes
in about a minute."""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu

MIB = 1 << 20


@pytest.fixture(scope="module")
def sorter():
    from deltaq_b200 import CudaSuffixSort
    s = CudaSuffixSort()
    yield s
    s.dispose()


def test_c4_full_size_sufcheck_and_verify(sorter):
    from deltaq_b200 import workloads as w
    t = w.c4_genome()
    assert t.size == 512 * MIB
    sa = np.empty(t.size, np.int32)
    sorter.sort(t, sa)
    oracle.verify(t, sa)                     # sufcheck == 0 and every adjacent pair strictly increasing


def test_c4_full_size_by_a_device_group(monkeypatch):
    """The same text by a device group (real peers when the box has several GPUs, logical shards otherwise)."""
    import torch
    from deltaq_b200 import CudaSuffixSort, workloads as w
    monkeypatch.setenv("DQ_SHARD_MIN", str(32 * MIB))
    k = torch.cuda.device_count()
    t = w.c4_genome()
    sa = np.empty(t.size, np.int32)
    with CudaSuffixSort(device=[i % k for i in range(max(2, min(k, 8)))]) as grp:
        grp.sort(t, sa)
        assert grp.stats()["n"] == t.size
    assert oracle.sufcheck(t, sa) == 0


def test_c5_full_size_streams_round_trip_and_sufcheck(sorter):
    """dq_cuda_bsdiff_streams on BASELINE's C5 recipe at full size: the bspatch round trip reproduces `new`; the suffix
    array of `old` passes the reference's checker."""
    from deltaq_b200 import bsdiff, workloads as w
    old, new = w.c5_pair(workers=8)
    assert old.size == 2_040_109_466
    if new.size > 2_100_000_000:
        new = np.ascontiguousarray(new[:2_100_000_000])
    st = sorter.context.bsdiff_streams(old, new, copy=False)
    rebuilt = bsdiff.apply_streams(old, st["ctrl"], st["diff"], st["extra"], new.size)
    assert rebuilt == new.tobytes()
    del rebuilt, st
    sa = np.empty(old.size, np.int32)
    sorter.sort(old, sa)
    assert oracle.sufcheck(old, sa) == 0


def test_c5_recipe_128mib_streams_identical_to_oracle(sorter):
    """The largest C5-recipe pair the oracle finishes in about a minute: ctrl/diff/extra byte-identical."""
    from deltaq_b200 import bsdiff, workloads as w
    old, new = w.c5_pair(128 * MIB, workers=4)
    got = bsdiff.create_streams(old, new, sorter)
    ref = oracle.bsdiff_streams(old, new)
    for k in ("ctrl", "diff", "extra"):
        assert got[k] == ref[k], k
    assert got["search_visits"] == ref["search_calls"]
