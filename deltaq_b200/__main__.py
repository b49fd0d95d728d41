"""`dq bsdiff` / `dq bspatch` with the CUDA provider -- the `-ss cuda` arm of the reference's command line
(/root/reference/src/DeltaQ.CommandLine/Commands.BsDiff.cs:16-70, Commands.BsPatch.cs:12-57, Program.cs:28-29; SURVEY.md
section 8(f) rank 4).  Same arguments, same messages:

    python -m deltaq_b200 bsdiff  <oldfile> <newfile> <deltafile> [-ss cuda] [--devices 0,1,...]
    python -m deltaq_b200 bspatch <oldfile> <deltafile> <newfile>
    python -m deltaq_b200 bench   [--sizes 0,1,...] [--reps K] [--devices 0,1,...]
    python -m deltaq_b200 fuzz    [file ...]          (no file: the input is read from stdin)

`bench` is the `cuda` column of the reference's suffix-sorting benchmark (bench/DeltaQ.Benchmarks/SuffixSortingBenchmarks.cs:
`[Benchmark] public void cuda(string name, byte[] asset) => CUDA.Sort(asset).Dispose();` over its `Randoms`, :27-57).

`fuzz` is the reference's fuzz target (Fuzzing/Commands.Fuzz.cs:22-38, built there under `#if FUZZ`): sort the input, then
SuffixSortingVerifier.Verify (Fuzzing/SuffixSortingVerifier.cs:7-22) -- every suffix strictly below the next -- with the CUDA
provider in the place of GoSAIS.

`-ss` accepts only `cuda` here (the reference's `sais` / `divsufsort` providers live in the reference; this package has
no CPU sorter).  bspatch is host code (dq_cuda_bspatch) and needs no GPU.
"""
import argparse
import os
import sys
import time

import numpy as np


def _elapsed(seconds):
    whole = int(seconds)
    return f"{whole // 3600:02d}:{whole % 3600 // 60:02d}:{whole % 60:02d}.{int((seconds - whole) * 1e7):07d}"


def _size(n):
    for unit in ("B", "KB", "MB", "GB"):
        if n < 1000 or unit == "GB":
            return f"{n:.4g} {unit}"
        n /= 1000.0


def bsdiff_command(args):
    from . import CudaSuffixSort
    from .bsdiff import Diff
    if args.suffix_sort not in (None, "cuda"):
        print(f"Unknown suffix sort library '{args.suffix_sort}': this build offers [cuda]", file=sys.stderr)
        return -1
    print("Generating BsDiff delta between")
    print(f'Old file: "{args.oldfile}"')
    print(f'New file: "{args.newfile}"')
    if args.suffix_sort:
        print("with suffix sort CudaSuffixSort")
    print()
    try:
        device = [int(x) for x in args.devices.split(",")] if args.devices else None
        if device is not None and len(device) == 1:
            device = device[0]
        old_bytes = np.fromfile(args.oldfile, dtype=np.uint8)
        new_bytes = np.fromfile(args.newfile, dtype=np.uint8)
        with CudaSuffixSort(device=device) as sort, open(args.deltafile, "w+b") as delta:
            t0 = time.perf_counter()
            Diff.create(old_bytes, new_bytes, delta, sort)
            dt = time.perf_counter() - t0
    except Exception:
        print("Failed to create delta", file=sys.stderr)
        raise
    print(f"Finished in {dt * 1e3:.0f} milliseconds [{_elapsed(dt)}]")
    print(f'Delta file: "{args.deltafile}"')
    delta_size = os.path.getsize(args.deltafile)
    ratio = delta_size / max(1, os.path.getsize(args.oldfile) + os.path.getsize(args.newfile))
    print(f"Delta size: {_size(delta_size)} ({ratio:.2%})")
    return 0


def bspatch_command(args):
    from . import _native
    print("Applying BsDiff delta between")
    print(f'Old file:   "{args.oldfile}"')
    print(f'Delta file: "{args.deltafile}"')
    print()
    try:
        old_bytes = np.fromfile(args.oldfile, dtype=np.uint8)
        delta = np.fromfile(args.deltafile, dtype=np.uint8)
        t0 = time.perf_counter()
        new_bytes = _native.bspatch(old_bytes, delta)
        dt = time.perf_counter() - t0
        new_bytes.tofile(args.newfile)
    except Exception:
        print("Failed to apply delta", file=sys.stderr)
        raise
    print(f"Finished in {dt * 1e3:.0f} milliseconds [{_elapsed(dt)}]")
    print(f'New file: "{args.newfile}"')
    return 0


def benchmark_sizes():
    """SuffixSortingBenchmarks.Sizes (SuffixSortingBenchmarks.cs:27-53)."""
    return [0, 1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384, 32768] + \
        [i * 1024 for i in range(64, 1025, 64)]


def bench_command(args, _lib=None):
    """One line per size: the call the reference's harness times -- Sort(asset) -> owner, Dispose() -- mean and best of
    `reps` after one warm-up call.  Random bytes seeded like the reference's buffers (63*13*63*13; .NET's generator is
    not reproducible here, any seeded uniform stream is the same benchmark)."""
    from . import CudaSuffixSort
    sizes = [int(x) for x in args.sizes.split(",")] if args.sizes else benchmark_sizes()
    device = [int(x) for x in args.devices.split(",")] if args.devices else None
    if device is not None and len(device) == 1:
        device = device[0]
    print("| Method | name | Mean | Best | MB/s (best) |")
    print("|------- |----- |-----:|-----:|------------:|")
    with CudaSuffixSort(device=device, _lib=_lib) as sort:
        for size in sizes:
            asset = np.random.default_rng(63 * 13 * 63 * 13).integers(0, 256, size, dtype=np.uint8)
            sort.sort(asset).dispose()
            times = []
            for _ in range(max(1, args.reps)):
                t0 = time.perf_counter()
                sort.sort(asset).dispose()
                times.append(time.perf_counter() - t0)
            mean, best = sum(times) / len(times), min(times)
            rate = f"{size / best / 1e6:.1f}" if size else "-"
            print(f"| cuda | {size} | {mean * 1e6:,.1f} us | {best * 1e6:,.1f} us | {rate} |")
    return 0


def suffix_less(t, a, b):
    """t[a:] < t[b:] as Span.SequenceCompareTo orders them (first difference, else the shorter one first), without copying
    the suffixes whole: windows that double in size."""
    step = 64
    while True:
        x, y = t[a:a + step], t[b:b + step]
        if x != y:
            return x < y
        if len(x) < step:       # both ran out together: only when a == b
            return False
        a += step
        b += step
        step = min(step * 2, 1 << 20)


def verify_suffix_array(t, sa):
    """SuffixSortingVerifier.Verify: InvalidOperationException("Input was unsorted") -> RuntimeError with (i, j)."""
    for i in range(len(t) - 1):
        if not suffix_less(t, int(sa[i]), int(sa[i + 1])):
            raise RuntimeError(f"Input was unsorted (i = {i}, j = {i + 1})")


def fuzz_command(args, _lib=None):
    from . import CudaSuffixSort
    inputs = [open(p, "rb").read() for p in args.files] if args.files else [sys.stdin.buffer.read()]
    with CudaSuffixSort(_lib=_lib) as sort:
        for data in inputs:
            with sort.sort(np.frombuffer(data, dtype=np.uint8)) as owner:
                verify_suffix_array(data, owner.memory)
    return 0


def main(argv=None):
    ap = argparse.ArgumentParser(prog="dq", description="DeltaQ binary diff and patch tool (CUDA provider)")
    sub = ap.add_subparsers(dest="command")
    d = sub.add_parser("bsdiff", help="Generate a BSDIFF-compatible delta (difference) between two files")
    d.add_argument("oldfile", help="Original file (input)")
    d.add_argument("newfile", help="New file (input)")
    d.add_argument("deltafile", help="Delta file (output)")
    d.add_argument("-ss", "--suffix-sort", metavar="LIB", help="Suffix sort library: [cuda]")
    d.add_argument("--devices", help="CUDA ordinals, comma separated; several = one device group (large inputs are sharded)")
    p = sub.add_parser("bspatch", help="Apply a BSDIFF-compatible delta (patch) to an original file and generate an "
                                       "output file")
    p.add_argument("oldfile", help="Original file (input)")
    p.add_argument("deltafile", help="Delta file (input)")
    p.add_argument("newfile", help="New file (output)")
    b = sub.add_parser("bench", help="Suffix-sorting benchmark: the cuda column of SuffixSortingBenchmarks")
    b.add_argument("--sizes", help="comma separated sizes (default: the reference's list, 0 .. 1 MiB)")
    b.add_argument("--reps", type=int, default=10)
    b.add_argument("--devices", help="CUDA ordinals, comma separated")
    f = sub.add_parser("fuzz", help="Fuzz target: sort the input with the CUDA provider and verify the suffix array")
    f.add_argument("files", nargs="*", help="inputs (default: stdin)")
    args = ap.parse_args(argv)
    if args.command == "fuzz":
        return fuzz_command(args)
    if args.command == "bench":
        return bench_command(args)
    if args.command == "bsdiff":
        return bsdiff_command(args)
    if args.command == "bspatch":
        return bspatch_command(args)
    ap.print_help()
    return 0


if __name__ == "__main__":
    sys.exit(main())
