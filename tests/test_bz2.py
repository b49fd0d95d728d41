"""Block-parallel bzip2 producer (deltaq_b200/csrc/dq_bz2_host.h) against serial libbz2 (Python's bz2).

The reference writes the ctrl / diff / extra sections through BZip2OutputStream (Diff.cs:14-19, :85-87); Patch.cs:52-93
reads them back with BZip2InputStream.  The producer here must emit ONE ordinary stream per section; stronger, its bytes
must equal what serial libbz2 writes at the same level.  Pure host code: runs without a GPU.
"""
import bz2
import ctypes
import ctypes.util

import numpy as np
import pytest

from deltaq_b200 import _native


_libbz2 = ctypes.CDLL(ctypes.util.find_library("bz2") or "libbz2.so.1.0")


def serial(data, level):
    """Serial libbz2, one BZ2_bzBuffToBuffCompress call -- the entry the producer uses per piece.  (Python's bz2.compress
    feeds BZ_RUN then BZ_FINISH; it writes the same bytes except when the LAST input byte is the one that fills a
    block, where it ends with an extra block holding that byte alone: see test_block_boundaries_every_offset.)"""
    data = bytes(data)
    cap = len(data) + len(data) // 100 + 600
    dst = ctypes.create_string_buffer(cap)
    n = ctypes.c_uint(cap)
    rc = _libbz2.BZ2_bzBuffToBuffCompress(dst, ctypes.byref(n), data, len(data), level, 0, 0)
    assert rc == 0
    return dst.raw[:n.value]


def cases():
    rng = np.random.default_rng(11)
    out = {}
    out["empty"] = np.zeros(0, np.uint8)
    out["one"] = np.array([7], np.uint8)
    out["four_equal"] = np.full(4, 9, np.uint8)
    out["random_250k"] = rng.integers(0, 256, 250_000, dtype=np.uint8)
    out["random_1M"] = rng.integers(0, 256, 1_000_003, dtype=np.uint8)
    out["zeros_3M"] = np.zeros(3_000_000, np.uint8)
    # diff-stream like: mostly zero with islands (C2's diff section is 96 % zeros)
    d = np.zeros(2_500_000, np.uint8)
    idx = rng.integers(0, d.size, 120_000)
    d[idx] = rng.integers(1, 256, idx.size, dtype=np.uint8)
    out["diff_like"] = d
    # runs of every length around the run-length stage's thresholds (3, 4, 5, 255, 256, 259, 510)
    parts = []
    for i in range(9000):
        L = [1, 2, 3, 4, 5, 254, 255, 256, 259, 510, 511][i % 11]
        parts.append(np.full(L, rng.integers(0, 256), np.uint8))
    out["runs_mixed"] = np.concatenate(parts)
    # text-like, several level-1 blocks
    words = [bytes(rng.integers(97, 123, int(rng.integers(2, 9)), dtype=np.uint8)) for _ in range(500)]
    out["words"] = np.frombuffer(b" ".join(words[int(i)] for i in rng.integers(0, 500, 120_000)), dtype=np.uint8).copy()
    return out


CASES = cases()


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("level", [1, 2, 5, 9])
def test_bytes_equal_serial_libbz2(name, level):
    data = CASES[name]
    info = []
    (z,) = _native.bz2_compress([data], level=level, threads=4, info=info)
    assert info[0][0] == level and info[0][2] == 0, info     # no serial fallback: the split rule holds
    assert bz2.decompress(z) == data.tobytes()
    assert z == serial(data, level)


def test_block_boundaries_every_offset():
    """Inputs whose length walks across a block boundary byte by byte (level 1 closes a block at 99981 coded bytes),
    random and run-heavy: the split must agree with libbz2 at every offset, including a run piece that overshoots the
    limit and the length at which the last byte is the one that fills the block."""
    rng = np.random.default_rng(3)
    base = rng.integers(0, 256, 100_200, dtype=np.uint8)
    runs = np.repeat(rng.integers(0, 256, 40_000, dtype=np.uint8), rng.integers(1, 9, 40_000))
    # length of the runs text at which its coded size reaches the limit
    coded, cut = 0, None
    edges = np.flatnonzero(np.diff(runs.astype(np.int16)) != 0) + 1
    for a, b in zip(np.concatenate([[0], edges]), np.concatenate([edges, [runs.size]])):
        coded += (b - a) if b - a < 4 else 5
        if coded >= 99_981:
            cut = int(b)
            break
    assert cut is not None
    differs_from_python = 0
    for src, lens in ((base, range(99_960, 100_020)), (runs, range(cut - 30, cut + 30))):
        sections = [src[:n] for n in lens]
        info = []
        got = _native.bz2_compress(sections, level=1, threads=4, info=info)
        for n, z, inf in zip(lens, got, info):
            assert inf[2] == 0, n
            assert z == serial(src[:n], 1), n
            assert bz2.decompress(z) == src[:n].tobytes()
            differs_from_python += z != bz2.compress(src[:n].tobytes(), 1)
    assert differs_from_python <= 2      # only the fills-the-block lengths


def test_many_sections_one_crew_and_auto_level():
    rng = np.random.default_rng(5)
    ctrl = rng.integers(0, 256, 2592, dtype=np.uint8)
    diff = CASES["diff_like"]
    extra = CASES["random_1M"]
    info = []
    z = _native.bz2_compress([ctrl, diff, extra], level=0, threads=8, info=info)
    for data, comp, inf in zip((ctrl, diff, extra), z, info):
        assert inf[2] == 0 and 1 <= inf[0] <= 9
        assert comp == serial(data, inf[0])
    assert info[2][1] >= 10          # 1 MB of random bytes at the level chosen for 8 threads: about 10 blocks
    one = _native.bz2_compress([extra], level=0, threads=1, info=info)
    assert info[0][0] == 9 and one[0] == serial(extra, 9)   # one thread: nothing to gain from small blocks


def test_thread_counts_agree():
    data = CASES["words"]
    ref = _native.bz2_compress([data], level=1, threads=1)[0]
    for t in (2, 3, 16):
        assert _native.bz2_compress([data], level=1, threads=t)[0] == ref


def test_bad_arguments():
    with pytest.raises(_native.NativeError):
        _native.bz2_compress([b"abc"], level=10)
    with pytest.raises(_native.NativeError):
        _native.bz2_compress([b"abc"], level=-1)


# ---- the reader's side: block-parallel decode, Patch.Apply on the file -------------------------------------------------

@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("level", [1, 9])
def test_parallel_decode_of_serial_streams(name, level):
    """Streams written by serial libbz2 (what a foreign bsdiff writes), decoded block by block."""
    data = CASES[name].tobytes()
    z = serial(data, level)
    info = []
    assert _native.bz2_decompress(z, threads=4, info=info) == data
    assert info[1] == 0                                   # no serial fallback
    assert (info[0] == 0) == (not data)                   # an empty stream has no block
    if name == "random_1M" and level == 1:
        assert info[0] >= 10


def test_parallel_decode_of_the_producers_streams():
    for name in ("diff_like", "runs_mixed", "words"):
        data = CASES[name].tobytes()
        (z,) = _native.bz2_compress([data], level=0, threads=8)
        assert _native.bz2_decompress(z, threads=3) == data


def test_decode_rejects_damage_and_handles_oddities():
    data = CASES["words"].tobytes()
    z = bytearray(serial(data, 1))
    # a flipped bit inside a block: the block's CRC (or its Huffman tables) no longer fit
    for at in (len(z) // 2, 200, len(z) - 20):
        bad = bytearray(z)
        bad[at] ^= 0x10
        with pytest.raises(RuntimeError, match="Corrupt patch"):
            _native.bz2_decompress(bytes(bad))
    # truncated
    with pytest.raises(RuntimeError, match="Corrupt patch"):
        _native.bz2_decompress(bytes(z[:len(z) // 2]))
    with pytest.raises(RuntimeError, match="Corrupt patch"):
        _native.bz2_decompress(b"not a bzip2 stream at all")
    with pytest.raises(RuntimeError, match="Corrupt patch"):
        _native.bz2_decompress(b"")
    # trailing bytes after the stream (the reference's extra section is "the rest of the file"): the serial decoder takes
    # it and stops at the end of the first stream, as libbz2 does
    info = []
    assert _native.bz2_decompress(bytes(z) + b"\0\0\0", info=info) == data and info[1] == 1
    # the payload contains the block magic itself, byte aligned and bit shifted: candidates that are not blocks
    magic = bytes.fromhex("314159265359")
    rng = np.random.default_rng(8)
    tricky = b"".join(magic + bytes(rng.integers(0, 256, 50, dtype=np.uint8)) for _ in range(3000))
    for level in (1, 9):
        assert _native.bz2_decompress(serial(tricky, level)) == tricky
    # stored (incompressible) data made of the magic at every bit offset cannot appear verbatim in a bzip2 stream, but a
    # forged "block" can: a stream followed by a second stream is not what a patch section is -- serial path, first stream
    two = serial(b"abc" * 1000, 9) + serial(b"xyz", 9)
    assert _native.bz2_decompress(two) == b"abc" * 1000


def _patch_file(old, new, level=9):
    """A BSDIFF40 file written the reference's way: oracle streams, serial bzip2 sections (Diff.cs:54-70, :226-241)."""
    import oracle
    st = oracle.bsdiff_streams(old, new)
    secs = [bz2.compress(st[k], level) for k in ("ctrl", "diff", "extra")]
    head = b"BSDIFF40" + len(secs[0]).to_bytes(8, "little") + len(secs[1]).to_bytes(8, "little") + \
        int(new.size).to_bytes(8, "little")
    return head + b"".join(secs)


def test_bspatch_on_patch_files():
    """dq_cuda_bspatch = Patch.Apply (Patch.cs:25-168) on the file: serial-bzip2 files and the producer's own, empty and
    tiny inputs (BsPatchTests.cs:18-38), and the reference's corrupt-patch cases."""
    from conftest import random_bytes
    from deltaq_b200 import workloads as w
    pairs = [(random_bytes(0), random_bytes(0)), (random_bytes(1), random_bytes(1, seed=2)),
             (random_bytes(0), random_bytes(100)), (random_bytes(100), random_bytes(0)),
             (random_bytes(4096), random_bytes(4099, seed=3)), w.c2_exe_pair(300_000, 330_000)]
    for old, new in pairs:
        for level in (1, 9):
            patch = _patch_file(old, new, level)
            assert _native.bspatch(old, patch).tobytes() == new.tobytes()
    old, new = pairs[-1]
    patch = bytearray(_patch_file(old, new))
    for field in (0, 8, 16, 24):                       # signature; negative lengths (sign bit of a packed long)
        bad = bytearray(patch)
        bad[field + 7] ^= 0x80
        with pytest.raises(RuntimeError, match="Corrupt patch"):
            _native.bspatch(old, bytes(bad))
    with pytest.raises(RuntimeError, match="Corrupt patch"):
        _native.bspatch(old, bytes(patch[:31]))
    big = bytearray(patch)
    big[8:16] = (len(patch)).to_bytes(8, "little")     # ctrl section longer than the file
    with pytest.raises(RuntimeError, match="Corrupt patch"):
        _native.bspatch(old, bytes(big))
    with pytest.raises(RuntimeError, match="Corrupt patch"):
        _native.bspatch(old[:1000], bytes(patch))       # wrong old file: reads past its end
    short = bytearray(patch)
    short[24:32] = (int(new.size) + 5).to_bytes(8, "little")   # header promises more than the streams hold
    with pytest.raises(RuntimeError, match="Corrupt patch"):
        _native.bspatch(old, bytes(short))


def test_randomised_sweep_against_serial():
    """Random mixtures of literal stretches and runs, random lengths and levels: bytes equal serial libbz2's, decode
    (parallel and Python's) gives the input back."""
    rng = np.random.default_rng(2024)
    sections, levels = [], []
    for _ in range(40):
        parts = []
        total = int(rng.integers(1, 400_000))
        sigma = int(rng.choice([2, 4, 16, 256]))
        while sum(p.size for p in parts) < total:
            if rng.random() < 0.5:
                parts.append(rng.integers(0, sigma, int(rng.integers(1, 5000)), dtype=np.uint8))
            else:
                parts.append(np.full(int(rng.choice([1, 2, 3, 4, 5, 100, 255, 256, 1000, 70_000])),
                                     int(rng.integers(0, sigma)), np.uint8))
        sections.append(np.concatenate(parts)[:total])
        levels.append(int(rng.integers(1, 4)))
    for level in (1, 2, 3):
        pick = [s for s, lv in zip(sections, levels) if lv == level]
        info = []
        got = _native.bz2_compress(pick, level=level, threads=4, info=info)
        for data, z, inf in zip(pick, got, info):
            assert inf[2] == 0
            assert z == serial(data, level)
            assert _native.bz2_decompress(z, threads=4) == data.tobytes()


def test_bspatch_does_not_unpack_a_bomb():
    """A section that decodes to far more than the patch can use (here 64 MiB of zeros in 60 bytes... of diff for a
    100-byte file) is rejected instead of being unpacked."""
    old = np.zeros(100, np.uint8)
    ctrl = bz2.compress((100).to_bytes(8, "little") + (0).to_bytes(8, "little") + (0).to_bytes(8, "little"))
    bomb = bz2.compress(bytes(64 << 20))
    extra = bz2.compress(b"")
    head = b"BSDIFF40" + len(ctrl).to_bytes(8, "little") + len(bomb).to_bytes(8, "little") + (100).to_bytes(8, "little")
    with pytest.raises(RuntimeError, match="Corrupt patch"):
        _native.bspatch(old, head + ctrl + bomb + extra)
    ok = bz2.compress(bytes(100))
    head = b"BSDIFF40" + len(ctrl).to_bytes(8, "little") + len(ok).to_bytes(8, "little") + (100).to_bytes(8, "little")
    assert _native.bspatch(old, head + ctrl + ok + extra).tobytes() == bytes(100)
