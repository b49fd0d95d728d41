"""N > 1 path: match search sharded by new-range over a replicated suffix array.
World size 2 over gloo on the CPU (contexts = the logic emulator, tests/emu); the same code runs over NCCL with
device pointers on GPUs (test_search_sharded_nccl, needs >= 2 GPUs: `gpurun --gpus 2 -- pytest -m gpu ...`)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _pair():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from search_cases import structured_pairs
    old, new = structured_pairs()["zero_runs_with_islands"]
    return old, new


def _worker(rank, world, port, backend, use_emu, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from deltaq_b200 import CudaSuffixSort
    from deltaq_b200.parallel import search_sharded
    if backend == "nccl":
        torch.cuda.set_device(rank)
    dist.init_process_group(backend, init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    if use_emu:
        import emu
        sorter = CudaSuffixSort(_lib=emu.library())
    elif backend == "nccl":
        sorter = CudaSuffixSort(device=rank)
    else:
        torch.cuda.set_device(0)            # both ranks share GPU 0
        sorter = CudaSuffixSort(device=0)
    old, new = _pair()
    pos, ln = search_sharded(old, new, sorter)
    q.put((rank, pos, ln))
    dist.barrier()
    dist.destroy_process_group()
    sorter.dispose()


def _run(world, backend, use_emu):
    import torch.multiprocessing as mp
    import oracle
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, backend, use_emu, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    old, new = _pair()
    I = oracle.make_I(oracle.sais(old))
    rp, rl = oracle.search_all(I, old, new)
    for rank, pos, ln in results:
        assert np.array_equal(pos, rp) and np.array_equal(ln, rl), rank


def test_shard_bounds_cover_everything():
    from deltaq_b200.parallel import shard_bounds
    for m in (0, 1, 7, 64, 1000):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(m, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == m
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(e - b for b, e in spans) - min(e - b for b, e in spans) <= 1


def test_search_sharded_gloo_world2():
    import emu
    emu.build()
    _run(2, "gloo", True)


@pytest.mark.gpu
def test_search_sharded_nccl():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    _run(2, "nccl", False)


# ---- one text sorted by several ranks (distributed prefix doubling) ---------------------------------------

def _sort_texts():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import adversarial_texts, load_asset, random_bytes
    t = adversarial_texts()
    out = {k: t[k] for k in ("zeros_tail", "all_zero_1000", "period8_zero_end", "fibonacci", "binary_random",
                             "repeated_paragraph", "period3_tail0")}
    out["fuzz3"] = load_asset("fuzz3")
    out["crash-gosais"] = load_asset("crash-gosais-force-alloc")
    out["random_20000"] = random_bytes(20000)
    out["tiny_5"] = np.array([3, 1, 2, 1, 0], np.uint8)
    out["single"] = np.array([9], np.uint8)
    out["empty"] = np.zeros(0, np.uint8)
    return out


def _sort_worker(rank, world, port, backend, use_emu, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from deltaq_b200 import CudaSuffixSort
    from deltaq_b200.parallel import suffix_sort_sharded
    if backend == "nccl":
        torch.cuda.set_device(rank)
    dist.init_process_group(backend, init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    if use_emu:
        import emu
        sorter = CudaSuffixSort(_lib=emu.library())
    elif backend == "nccl":
        sorter = CudaSuffixSort(device=rank)
    else:
        torch.cuda.set_device(0)            # several ranks share GPU 0, collectives staged through gloo
        sorter = CudaSuffixSort(device=0)
    res = {}
    for name, t in _sort_texts().items():
        res[name] = suffix_sort_sharded(t, sorter)
    q.put((rank, res))
    dist.barrier()
    dist.destroy_process_group()
    sorter.dispose()


def _run_sort(world, backend, use_emu):
    import torch.multiprocessing as mp
    import oracle
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sort_worker, args=(r, world, port, backend, use_emu, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=600) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    texts = _sort_texts()
    for rank, res in results:
        for name, t in texts.items():
            assert np.array_equal(res[name], oracle.sais(t)), (rank, name)


@pytest.mark.parametrize("world", [2, 3])
def test_suffix_sort_sharded_gloo(world):
    import emu
    emu.build()
    _run_sort(world, "gloo", True)


@pytest.mark.gpu
def test_suffix_sort_sharded_nccl():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    _run_sort(2, "nccl", False)


@pytest.mark.gpu
def test_suffix_sort_sharded_two_ranks_one_gpu():
    """The multi-rank sort on a single GPU: two processes share cuda:0 (gloo moves the exchanged tuples through the
    host), so the distributed kernels and the exchange logic run in every single-GPU `-m gpu` pass."""
    _run_sort(2, "gloo", False)


@pytest.mark.gpu
def test_search_sharded_two_ranks_one_gpu():
    _run(2, "gloo", False)
