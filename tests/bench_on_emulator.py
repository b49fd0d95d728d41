"""Runs bench.py's main() -- the GPU arm, N = 1 -- on the CPU logic emulator (tests/emu; test infrastructure) with small
inputs, so that every line of the driver-facing script is executed before it meets a GPU: torch's CUDA calls are
replaced by host stand-ins (the emulator's "device" pointers are host pointers), the library is the emulator build,
the workloads are scaled down.  Prints bench.py's JSON line.  Used by tests/test_bench_contract.py; the numbers mean
nothing.

    python tests/bench_on_emulator.py [bench.py arguments]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port P \
        tests/bench_on_emulator.py --gpus 2 ...      # the N > 1 flow: gloo stands in for NCCL, logical shards for GPUs
"""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import numpy as np
    import torch
    import emu
    from deltaq_b200 import _native, workloads as w

    # the emulator build instead of libdeltaq_cuda.so
    lib = emu.library()
    _native.default_library = lambda: lib

    # host stand-ins for the few torch.cuda calls bench.py makes
    torch.cuda.is_available = lambda: True
    torch.cuda.set_device = lambda *_a, **_k: None
    torch.cuda.synchronize = lambda *_a, **_k: None
    torch.cuda.empty_cache = lambda: None
    torch.Tensor.cuda = lambda self, *_a, **_k: self
    real_empty = torch.empty

    def empty(*a, **k):
        k.pop("device", None)
        return real_empty(*a, **k)
    torch.empty = empty
    real_tensor = torch.tensor

    def tensor(*a, **k):
        k.pop("device", None)
        return real_tensor(*a, **k)
    torch.tensor = tensor
    import torch.distributed as dist
    real_init = dist.init_process_group

    def init_process_group(backend=None, **k):   # N > 1: gloo stands in for NCCL
        k.pop("device_id", None)
        return real_init("gloo", **k)
    dist.init_process_group = init_process_group

    # the workloads, scaled down (same generators)
    c2, c3, c4, c1 = w.c2_exe_pair, w.c3_repetitive, w.c4_genome, w.c1_uniform
    w.c2_exe_pair = lambda seed_old=1, seed_new=2, **_k: c2(120_000, 128_000, seed_old=seed_old, seed_new=seed_new)
    w.c3_repetitive = lambda *_a, **_k: c3(40_000)
    w.c4_genome = lambda *_a, **_k: c4(60_000)
    w.c1_uniform = lambda *_a, **_k: c1(30_000)

    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    # N > 1: the group record on logical shards of the one emulated device, tiny inputs
    real_sharded = bench.sharded_record
    bench.sharded_record = lambda devices, workers, **_k: real_sharded([0] * len(devices), workers=1, lib=lib, scale=2e-5)
    os.environ.setdefault("DQ_SHARD_MIN", "1")
    bench.main()


if __name__ == "__main__":
    main()
