"""`python -m deltaq_b200 bsdiff|bspatch`: the reference's command line (Commands.BsDiff.cs, Commands.BsPatch.cs) with the
CUDA provider.  bspatch is host code and runs anywhere; bsdiff needs the GPU."""
import bz2
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle
from conftest import random_bytes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*argv):
    return subprocess.run([sys.executable, "-m", "deltaq_b200", *argv], cwd=ROOT, capture_output=True, text=True,
                          timeout=600)


def _pair():
    old = random_bytes(50_000)
    new = np.concatenate([old[:20_000], random_bytes(300, seed=4), old[19_000:]])
    return old, new


def test_bspatch_cli(tmp_path):
    old, new = _pair()
    st = oracle.bsdiff_streams(old, new)
    secs = [bz2.compress(st[k]) for k in ("ctrl", "diff", "extra")]
    patch = b"BSDIFF40" + len(secs[0]).to_bytes(8, "little") + len(secs[1]).to_bytes(8, "little") + \
        int(new.size).to_bytes(8, "little") + b"".join(secs)
    (tmp_path / "old").write_bytes(old.tobytes())
    (tmp_path / "delta").write_bytes(patch)
    r = _run("bspatch", str(tmp_path / "old"), str(tmp_path / "delta"), str(tmp_path / "new"))
    assert r.returncode == 0, r.stderr
    assert "Applying BsDiff delta between" in r.stdout and "Finished in" in r.stdout
    assert (tmp_path / "new").read_bytes() == new.tobytes()
    (tmp_path / "delta").write_bytes(b"BSDIFF41" + patch[8:])
    r = _run("bspatch", str(tmp_path / "old"), str(tmp_path / "delta"), str(tmp_path / "new2"))
    assert r.returncode != 0 and "Failed to apply delta" in r.stderr and "Corrupt patch" in r.stderr


@pytest.mark.gpu
def test_bsdiff_cli_roundtrip(tmp_path):
    old, new = _pair()
    (tmp_path / "old").write_bytes(old.tobytes())
    (tmp_path / "new").write_bytes(new.tobytes())
    r = _run("bsdiff", str(tmp_path / "old"), str(tmp_path / "new"), str(tmp_path / "delta"), "-ss", "cuda")
    assert r.returncode == 0, r.stderr
    assert "with suffix sort CudaSuffixSort" in r.stdout and "Delta size:" in r.stdout
    r = _run("bspatch", str(tmp_path / "old"), str(tmp_path / "delta"), str(tmp_path / "rebuilt"))
    assert r.returncode == 0, r.stderr
    assert (tmp_path / "rebuilt").read_bytes() == new.tobytes()
    r = _run("bsdiff", str(tmp_path / "old"), str(tmp_path / "new"), str(tmp_path / "delta"), "-ss", "sais")
    assert r.returncode != 0


def test_bench_sizes_are_the_references():
    """SuffixSortingBenchmarks.cs:27-53: 0, the powers of two up to 32768, then 64 KiB .. 1 MiB in steps of 64 KiB."""
    from deltaq_b200.__main__ import benchmark_sizes
    s = benchmark_sizes()
    assert s[:4] == [0, 1, 2, 4] and s[16] == 32768 and s[17] == 65536 and s[-1] == 1048576 and len(s) == 17 + 16


def test_bench_command_on_the_emulator(capsys):
    """The `cuda` column of the reference's benchmark, on the emulator with a few small sizes (the numbers mean nothing)."""
    import argparse
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import emu
    from deltaq_b200.__main__ import bench_command
    rc = bench_command(argparse.Namespace(sizes="0,1,51,5000", reps=2, devices=None), _lib=emu.library())
    out = capsys.readouterr().out
    assert rc == 0 and out.count("| cuda |") == 4 and "| cuda | 5000 |" in out


@pytest.mark.gpu
def test_bench_cli_gpu():
    r = _run("bench", "--sizes", "0,1,4096,65536,1048576", "--reps", "3")
    assert r.returncode == 0, r.stderr
    assert r.stdout.count("| cuda |") == 5 and "| cuda | 1048576 |" in r.stdout


def test_fuzz_target_on_the_emulator():
    """The reference's fuzz target (Commands.Fuzz.cs) with the CUDA provider, over the reference's own fuzz corpus
    (test/assets, copied to tests/golden/assets) on the emulator; its verifier must reject a wrong array."""
    import argparse
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import emu
    from conftest import GOLDEN, asset_names
    from deltaq_b200.__main__ import fuzz_command, suffix_less, verify_suffix_array
    files = [os.path.join(GOLDEN, "assets", n) for n in asset_names()]
    assert len(files) == 13
    assert fuzz_command(argparse.Namespace(files=files), _lib=emu.library()) == 0
    t = b"banana"
    verify_suffix_array(t, [5, 3, 1, 0, 4, 2])
    with pytest.raises(RuntimeError, match="Input was unsorted"):
        verify_suffix_array(t, [5, 1, 3, 0, 4, 2])
    with pytest.raises(RuntimeError, match="Input was unsorted"):
        verify_suffix_array(t, [5, 5, 3, 1, 0, 4])              # a repeated entry is not "strictly below"
    long = bytes(1000) + b"\x01" + bytes(1000)
    assert suffix_less(long, 0, 1) and not suffix_less(long, 1, 0)   # the longer run of zeros in front of the 1 sorts first
    assert suffix_less(long, 1500, 1400) and not suffix_less(long, 1400, 1500)   # all zeros: the shorter suffix first


@pytest.mark.gpu
def test_fuzz_cli_gpu():
    from conftest import GOLDEN, asset_names
    r = _run("fuzz", *[os.path.join(GOLDEN, "assets", n) for n in asset_names()])
    assert r.returncode == 0, r.stderr
    r = subprocess.run([sys.executable, "-m", "deltaq_b200", "fuzz"], cwd=ROOT, input=b"mississippi", capture_output=True,
                       timeout=600)
    assert r.returncode == 0, r.stderr
