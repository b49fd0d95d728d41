"""`dq bsdiff` / `dq bspatch` with the CUDA provider -- the `-ss cuda` arm of the reference's command line
(/root/reference/src/DeltaQ.CommandLine/Commands.BsDiff.cs:16-70, Commands.BsPatch.cs:12-57, Program.cs:28-29; SURVEY.md
section 8(f) rank 4).  Same arguments, same messages:

    python -m deltaq_b200 bsdiff  <oldfile> <newfile> <deltafile> [-ss cuda] [--devices 0,1,...]
    python -m deltaq_b200 bspatch <oldfile> <deltafile> <newfile>

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
    args = ap.parse_args(argv)
    if args.command == "bsdiff":
        return bsdiff_command(args)
    if args.command == "bspatch":
        return bspatch_command(args)
    ap.print_help()
    return 0


if __name__ == "__main__":
    sys.exit(main())
