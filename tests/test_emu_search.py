"""bsdiff match search on the CPU logic emulator (tests/emu; NOT the product) against the oracle's literal
replay of Diff.Search (oracle/bsdiff.c) and its greedy loop."""
import io
import os

import numpy as np
import pytest

import oracle
from conftest import GOLDEN, random_bytes
from search_cases import small_random_pairs, structured_pairs


@pytest.fixture(scope="module")
def sorter():
    import emu
    from deltaq_b200 import CudaSuffixSort
    s = CudaSuffixSort(_lib=emu.library())
    yield s
    s.dispose()


def check_pair(sorter, old, new, with_streams=True):
    from deltaq_b200 import bsdiff
    I = oracle.make_I(oracle.sais(old))
    rp, rl = oracle.search_all(I, old, new)
    pos, ln = bsdiff.search_all(old, new, sorter, I=I)              # caller-supplied I
    assert np.array_equal(ln, rl) and np.array_equal(pos, rp)
    pos, ln = bsdiff.search_all(old, new, sorter)                   # I resident from the sort
    assert np.array_equal(ln, rl) and np.array_equal(pos, rp)
    if new.size > 10:                                               # sub-range
        b, c = new.size // 3, new.size // 2
        pos, ln = bsdiff.search_all(old, new, sorter, I=I, scan_begin=b, count=c)
        assert np.array_equal(ln, rl[b:b + c]) and np.array_equal(pos, rp[b:b + c])
    if with_streams:
        got = bsdiff.create_streams(old, new, sorter)
        ref = oracle.bsdiff_streams(old, new, I)
        for k in ("ctrl", "diff", "extra"):
            assert got[k] == ref[k], k
        assert got["search_visits"] == ref["search_calls"]


def test_small_random_pairs(sorter):
    for old, new in small_random_pairs():
        check_pair(sorter, old, new)


@pytest.mark.parametrize("name", sorted(structured_pairs()))
def test_structured(sorter, name):
    old, new = structured_pairs()[name]
    check_pair(sorter, old, new)


def test_golden_bsdiff_cases(sorter):
    from deltaq_b200 import bsdiff
    g = np.load(os.path.join(GOLDEN, "bsdiff_cases.npz"))
    for k in range(int(g["count"])):
        old, new = g[f"c{k}_old"], g[f"c{k}_new"]
        got = bsdiff.create_streams(old, new, sorter)
        for s in ("ctrl", "diff", "extra"):
            assert got[s] == g[f"c{k}_{s}"].tobytes(), (k, s)
        pos, ln = bsdiff.search_all(old, new, sorter)
        visited = g[f"c{k}_trace_len"] >= 0
        assert np.array_equal(pos[visited], g[f"c{k}_trace_pos"][visited])
        assert np.array_equal(ln[visited], g[f"c{k}_trace_len"][visited])


@pytest.mark.parametrize("size", [0, 1, 512, 999, 1024, 4096])
def test_diff_create_roundtrip(sorter, size):
    # BsDiffTests.cs:30-78: Create then Apply reproduces new
    import io
    from conftest import random_bytes
    from deltaq_b200.bsdiff import Diff, Patch
    old = random_bytes(size)
    for new in (old.copy(), random_bytes(size + 3, seed=9)):
        out = io.BytesIO()
        Diff.create(old, new, out, sorter)
        patch = out.getvalue()
        assert patch[:8] == b"BSDIFF40"
        rebuilt = io.BytesIO()
        Patch.apply(old, patch, rebuilt)
        assert rebuilt.getvalue() == new.tobytes()


def test_diff_create_argument_validation(sorter):
    # BsDiffTests.cs:80-100
    import io
    from deltaq_b200.bsdiff import Diff
    with pytest.raises(TypeError):
        Diff.create(b"", b"", None, sorter)
    with pytest.raises(TypeError):
        Diff.create(b"", b"", io.BytesIO(), None)

    class NotSeekable(io.BytesIO):
        def seekable(self):
            return False

    class NotWritable(io.BytesIO):
        def writable(self):
            return False

    with pytest.raises(ValueError):
        Diff.create(b"", b"", NotSeekable(), sorter)
    with pytest.raises(ValueError):
        Diff.create(b"", b"", NotWritable(), sorter)


def test_coded_table_overflow_falls_back_to_full_table(monkeypatch):
    # dq_cuda_bsdiff_streams ships the table as code bytes + match heads; a head list that does not fit must give
    # the same streams through the full-table path (DQ_HEADS_CAP shrinks the list for this test)
    import emu
    from deltaq_b200 import CudaSuffixSort, bsdiff
    rng = np.random.default_rng(77)
    old = rng.integers(0, 4, 6000, dtype=np.uint8)
    new = np.concatenate([old[3000:4000], rng.integers(0, 4, 500, dtype=np.uint8), old[:2500], old[5000:]])
    ref = oracle.bsdiff_streams(old, new, oracle.make_I(oracle.sais(old)))
    for cap, fallbacks in (("1", 1), ("100000", 0)):
        monkeypatch.setenv("DQ_HEADS_CAP", cap)
        s = CudaSuffixSort(_lib=emu.library())
        try:
            got = bsdiff.create_streams(old, new, s)
            assert s._ctx.stats()["table_fallbacks"] == fallbacks
        finally:
            s.dispose()
        for k in ("ctrl", "diff", "extra"):
            assert got[k] == ref[k], (cap, k)


@pytest.mark.parametrize("per", ["1", "3", "64"])
def test_seed_level_of_the_head_kernels(sorter, per, monkeypatch):
    # lcp_heads_kernel / search_heads_kernel run a sparse seed level first when the text has long repeats;
    # DQ_SEEDS_PER forces it on (with that many supers per warp) for inputs of any size
    monkeypatch.setenv("DQ_SEEDS_PER", per)
    for old, new in list(small_random_pairs(count=12, seed=5)) + [structured_pairs()[k] for k in sorted(structured_pairs())[:6]]:
        check_pair(sorter, old, new)
    rng = np.random.default_rng(9)
    blk = rng.integers(0, 256, 3000, dtype=np.uint8)
    old = np.concatenate([blk, rng.integers(0, 256, 500, dtype=np.uint8), blk, blk[:1500], blk])   # long repeats, > 1 super
    new = np.concatenate([blk[100:], blk, rng.integers(0, 256, 300, dtype=np.uint8), old[2000:9000]])
    check_pair(sorter, old, new)


def test_prefix3_table_for_scratch_searches(sorter, monkeypatch):
    # from a megabyte of text the from-scratch searches start from a 3-byte prefix table instead of the 2-byte
    # buckets; DQ_PREFIX3=1 turns it on for small inputs (the two short suffixes at the end of old are the edge)
    monkeypatch.setenv("DQ_PREFIX3", "1")
    pairs = list(small_random_pairs(count=3, seed=77))
    pairs += [structured_pairs()[k] for k in sorted(structured_pairs())[:2]]
    rng = np.random.default_rng(3)
    for tail in (b"", b"\x00", b"\x00\x00", b"ab", b"\xff\xff\xff"):                   # old ends in short suffixes that
        old = np.frombuffer(rng.integers(0, 3, 300, dtype=np.uint8).tobytes() + tail, dtype=np.uint8)  # pad to real prefixes
        new = np.concatenate([rng.integers(0, 3, 200, dtype=np.uint8), np.frombuffer(tail + b"\x00\x00\x01", dtype=np.uint8), old[50:150]])
        pairs.append((old, new))
    for old, new in pairs:
        check_pair(sorter, old, new, with_streams=False)


def test_prefix3_table_from_the_sorts_round0_keys(sorter, monkeypatch):
    # once a context has searched, its sorts build the 3-byte prefix table from their round-0 keys (1 MiB and up;
    # DQ_PREFIX3_SORTED_MIN lowers that for this test).  check_pair searches with a supplied I first, then sorts.
    monkeypatch.setenv("DQ_PREFIX3_SORTED_MIN", "1")
    pairs = list(small_random_pairs(count=6, seed=78)) + [structured_pairs()[k] for k in sorted(structured_pairs())[2:4]]
    rng = np.random.default_rng(4)
    for tail in (b"", b"\x00", b"\x00\x00", b"ab", b"\xff\xff\xff"):
        old = np.frombuffer(rng.integers(0, 3, 300, dtype=np.uint8).tobytes() + tail, dtype=np.uint8)
        new = np.concatenate([rng.integers(0, 3, 200, dtype=np.uint8), np.frombuffer(tail + b"\x00\x00\x01", dtype=np.uint8), old[50:150]])
        pairs.append((old, new))
    for old, new in pairs:
        check_pair(sorter, old, new)


def _broken_suffix_arrays(n):
    """Suffix arrays a faulty ISuffixSort provider might hand to Diff.Create: in range but not a
    permutation, an entry == n, a negative entry."""
    good = np.arange(n, dtype=np.int32)
    zeros = np.zeros(n + 1, np.int32)
    too_big = np.concatenate([good, [0]]).astype(np.int32)
    too_big[n // 2] = n
    negative = np.concatenate([good, [0]]).astype(np.int32)
    negative[3] = -5
    dup = np.concatenate([good, [0]]).astype(np.int32)
    dup[7] = dup[8]
    return {"all_zero": zeros, "entry_eq_n": too_big, "negative": negative, "duplicate": dup}


@pytest.mark.parametrize("kind", ["all_zero", "entry_eq_n", "negative", "duplicate"])
def test_broken_provider_is_rejected(sorter, kind):
    """The search validates a caller-supplied suffix array on the device before following its entries."""
    from deltaq_b200 import _native
    n = 5000
    old, new = random_bytes(n), random_bytes(300, seed=3)
    I = _broken_suffix_arrays(n)[kind]
    pos = np.zeros(new.size, np.int32)
    with pytest.raises(_native.NativeError) as ei:
        sorter.context.bsdiff_search(old, I, new, 0, new.size, pos, pos.copy())
    assert ei.value.status == _native.DQ_ERR_INVALID_ARGUMENT
    assert "permutation" in str(ei.value)
    # the context is still usable
    sa = np.empty(n, np.int32)
    sorter.sort(old, sa)
    assert np.array_equal(sa, oracle.sais(old))


def test_native_patch_file(sorter):
    """dq_cuda_bsdiff_patch: header (Diff.cs:54-70) + three sections, each ONE ordinary bzip2 stream holding exactly the
    bytes of dq_cuda_bsdiff_streams; Patch.apply rebuilds `new`; every level gives the same streams."""
    import bz2
    from deltaq_b200 import bsdiff
    from search_cases import structured_pairs
    pairs = list(structured_pairs().values())[:4] + [(random_bytes(70000), random_bytes(70003, seed=9))]
    for old, new in pairs:
        ref = bsdiff.create_streams(old, new, sorter)
        for level in (0, 1, 9):
            patch = sorter.context.bsdiff_patch(old, new, level=level)
            assert patch[:8] == b"BSDIFF40"
            cl, dl, size = (bsdiff.read_packed_long(patch[8 + 8 * i:16 + 8 * i]) for i in range(3))
            assert size == new.size
            assert bz2.decompress(patch[32:32 + cl]) == ref["ctrl"]
            assert bz2.decompress(patch[32 + cl:32 + cl + dl]) == ref["diff"]
            assert bz2.decompress(patch[32 + cl + dl:]) == ref["extra"]
            rebuilt = io.BytesIO()
            bsdiff.Patch.apply(old, patch, rebuilt)
            assert rebuilt.getvalue() == new.tobytes()
        out = io.BytesIO()
        out.write(b"xx")                      # Diff.Create writes at the stream's current position (Diff.cs:56)
        bsdiff.Diff.create(old, new, out, sorter)
        assert out.getvalue()[2:] == sorter.context.bsdiff_patch(old, new) and out.tell() == len(out.getvalue())


def _lcp_texts():
    from conftest import adversarial_texts, small_alphabet_texts
    out = dict(adversarial_texts())
    out.update({k: v for k, v in small_alphabet_texts().items() if k in ("acgt_tandem", "bin_zero_tail", "one_value_5000")})
    out["random_300k"] = random_bytes(300_000)
    out["fuzz3"] = np.fromfile(os.path.join(GOLDEN, "assets", "fuzz3"), dtype=np.uint8)
    out["single"] = np.array([9], np.uint8)
    out["empty"] = np.zeros(0, np.uint8)
    return out


def test_lcp_array_export(sorter):
    """dq_cuda_lcp: the LCP array under the resident suffix array and under a caller-supplied one,
    against the definition (oracle.lcp_array); a following search reuses the index; a broken array is rejected."""
    from deltaq_b200 import _native, bsdiff
    for name, t in _lcp_texts().items():
        sa = np.empty(t.size, np.int32)
        sorter.sort(t, sa)
        ref = oracle.lcp_array(t, sa)
        assert np.array_equal(sorter.lcp_array(t), ref), name            # resident
        assert np.array_equal(sorter.lcp_array(t, sa), ref), name        # adopted
        if t.size > 100:
            new = np.concatenate([t[50:], t[:60]])
            pos, ln = bsdiff.search_all(t, new, sorter, I=oracle.make_I(sa))
            lcp_again = sorter.lcp_array(t)                                # index still resident after the search
            assert np.array_equal(lcp_again, ref), name
            rp, rl = oracle.search_all(oracle.make_I(sa), t, new)
            assert np.array_equal(pos, rp) and np.array_equal(ln, rl), name
    t = random_bytes(5000)
    with pytest.raises(_native.NativeError):
        sorter.lcp_array(t, np.zeros(t.size, np.int32))                   # not a permutation
    with pytest.raises(_native.NativeError):
        sorter.sort(random_bytes(100), np.empty(100, np.int32))
        sorter.lcp_array(t)                                               # resident array has another length


def test_tiny_old_against_a_large_new(sorter):
    """`old` empty or a few bytes while the coded table is active at scale; the all-zero pair overflows the
    head list (full-table path at 200 KB, not only at the 6 KB of the DQ_HEADS_CAP test)."""
    from deltaq_b200 import bsdiff
    from search_cases import tiny_old_pairs
    fallbacks = 0
    for name, old, new in tiny_old_pairs(200_000):
        got = bsdiff.create_streams(old, new, sorter)
        ref = oracle.bsdiff_streams(old, new)
        for k in ("ctrl", "diff", "extra"):
            assert got[k] == ref[k], (name, k)
        assert got["search_visits"] == ref["search_calls"], name
        fallbacks += sorter.stats()["table_fallbacks"]
    assert fallbacks >= 1


def test_patch_and_output_are_flushed_through_buffering_streams(sorter):
    """BsPatchTests.cs:18-38 (BsPatchFlushesOutput): Create and Apply write through buffering wrappers, and the test reads
    the wrapped streams directly afterwards -- so both must have pushed their bytes down."""
    import io
    from conftest import random_bytes
    from deltaq_b200.bsdiff import Diff, Patch
    old, new = random_bytes(0x123, seed=21), random_bytes(0x4567, seed=22)
    patch_ms = io.BytesIO()
    wrapped_patch = io.BufferedRandom(patch_ms, buffer_size=1 << 20)      # (kept alive: closing it closes patch_ms)
    Diff.create(old, new, wrapped_patch, sorter)
    patch = patch_ms.getvalue()
    assert patch[:8] == b"BSDIFF40"
    rebuilt_ms = io.BytesIO()
    wrapped_rebuilt = io.BufferedRandom(rebuilt_ms, buffer_size=1 << 20)
    Patch.apply(old, patch, wrapped_rebuilt)
    assert rebuilt_ms.getvalue() == new.tobytes()


@pytest.mark.parametrize("size", [0, 1, 512, 999, 1024, 4096])
def test_roundtrip_from_streams(sorter, size):
    """BsDiffTests.cs:57-78 (BsDiffRoundtripFromStreams): Create writes into a fixed 0x2000-byte memory stream; Apply takes
    the old data as a stream and the patch through an open-stream callback (offset, length; length 0 = the rest -- so
    the last section is followed by the unused tail of the buffer, which the reader must leave alone)."""
    import io
    from conftest import random_bytes
    from deltaq_b200.bsdiff import Diff, Patch
    old, new = random_bytes(size), random_bytes(size, seed=77)
    memory = io.BytesIO(bytes(0x2000))
    Diff.create(old, new, memory, sorter)
    buf = memory.getvalue()
    assert len(buf) == 0x2000

    def open_patch_stream(start, length):
        return io.BytesIO(buf[start:start + length] if length > 0 else buf[start:])

    out = io.BytesIO()
    Patch.apply(io.BytesIO(old.tobytes()), open_patch_stream, out)
    assert out.getvalue() == new.tobytes()


def test_apply_stream_argument_checks(sorter):
    """Patch.cs:60-63, :97-102: unreadable / unseekable patch or input streams and an unwritable output are argument
    errors; a bad signature or negative header fields are a corrupt patch (:68-78)."""
    import io
    from conftest import random_bytes
    from deltaq_b200.bsdiff import Diff, Patch
    old, new = random_bytes(600), random_bytes(700, seed=5)
    ms = io.BytesIO()
    Diff.create(old, new, ms, sorter)
    buf = ms.getvalue()

    class Pipe(io.BytesIO):
        def seekable(self):
            return False

    class WriteOnly(io.BytesIO):
        def readable(self):
            return False

    def opener(kind=io.BytesIO, data=buf):
        return lambda start, length: kind(data[start:start + length] if length > 0 else data[start:])

    with pytest.raises(ValueError, match="Patch stream must be seekable"):
        Patch.apply(io.BytesIO(old.tobytes()), opener(Pipe), io.BytesIO())
    with pytest.raises(ValueError, match="Patch stream must be readable"):
        Patch.apply(io.BytesIO(old.tobytes()), opener(WriteOnly), io.BytesIO())
    with pytest.raises(ValueError, match="Input stream must be seekable"):
        Patch.apply(Pipe(old.tobytes()), opener(), io.BytesIO())
    with pytest.raises(ValueError, match="Input stream must be readable"):
        Patch.apply(WriteOnly(old.tobytes()), opener(), io.BytesIO())

    class ReadOnly(io.BytesIO):
        def writable(self):
            return False

    with pytest.raises(ValueError, match="Output stream must be writable"):
        Patch.apply(io.BytesIO(old.tobytes()), opener(), ReadOnly())
    with pytest.raises(RuntimeError, match="Corrupt patch"):
        Patch.apply(io.BytesIO(old.tobytes()), opener(data=b"BSDIFF41" + buf[8:]), io.BytesIO())
    negative = buf[:15] + bytes([buf[15] | 0x80]) + buf[16:]          # ctrl length with the sign bit set
    with pytest.raises(RuntimeError, match="Corrupt patch"):
        Patch.apply(io.BytesIO(old.tobytes()), opener(data=negative), io.BytesIO())
    out = io.BytesIO()
    Patch.apply(io.BytesIO(old.tobytes()), opener(), out)
    assert out.getvalue() == new.tobytes()
