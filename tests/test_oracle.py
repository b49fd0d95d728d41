"""Pins the CPU oracle against the reference's own fixtures and checkers.

CPU only.  The reference ships no golden SA: its tests check the property (Verify + sufcheck),
and so do these, on the same fixture files, string and sizes."""
import numpy as np
import pytest

import oracle
from conftest import (LDSS_FIXTURES, REF_RANDOM_SIZES, SHRUGGY, adversarial_texts, asset_names,
                      load_asset, load_golden_sa, random_bytes)


@pytest.mark.parametrize("name", asset_names())
def test_fixture_files(name):
    # SAISTester.cs:53-70 (every file) and LibDivSufSortTests.cs:108-124 (CheckFile)
    t = load_asset(name)
    sa = oracle.sais(t)
    oracle.verify(t, sa)
    assert np.array_equal(sa, oracle.sa_naive(t))
    assert np.array_equal(sa, load_golden_sa(name))


def test_ldss_fixture_list_is_present():
    assert set(LDSS_FIXTURES) <= set(asset_names())
    assert len(asset_names()) == 13


def test_shruggy():
    # LibDivSufSortTests.cs:66-77
    t = np.frombuffer(SHRUGGY, dtype=np.uint8)
    sa = oracle.sais(t)
    oracle.verify(t, sa)
    assert np.array_equal(sa, oracle.sa_naive(t))


@pytest.mark.parametrize("size", REF_RANDOM_SIZES)
def test_random_buffer(size):
    # LibDivSufSortTests.cs:126-148, SAISTester.cs:35-50
    t = random_bytes(size)
    sa = oracle.sais(t)
    assert sa.size == size
    oracle.verify(t, sa)
    if size <= 0x1000:
        assert np.array_equal(sa, oracle.sa_naive(t))


@pytest.mark.parametrize("name", sorted(adversarial_texts()))
def test_adversarial(name):
    t = adversarial_texts()[name]
    sa = oracle.sais(t)
    oracle.verify(t, sa)
    assert np.array_equal(sa, oracle.sa_naive(t))


def test_sufcheck_rejects_wrong_arrays():
    t = random_bytes(200)
    sa = oracle.sais(t)
    assert oracle.sufcheck(t, sa[:-1]) == -1          # BadArguments
    bad = sa.copy(); bad[3] = 200
    assert oracle.sufcheck(t, bad) == -2              # OutOfRange
    bad = sa.copy(); bad[[0, -1]] = bad[[-1, 0]]
    assert oracle.sufcheck(t, bad) in (-3, -4)
    z = np.zeros(10, np.uint8)
    assert oracle.sufcheck(z, np.arange(10, dtype=np.int32)) == -4   # needs 9,8,...,0
    assert oracle.sufcheck(z, np.arange(9, -1, -1, dtype=np.int32)) == 0


def test_packed_long():
    # SpanExtensions.cs:7-30 sign-magnitude little-endian
    assert oracle.packed_long(0) == bytes(8)
    assert oracle.packed_long(1) == b"\x01" + bytes(7)
    assert oracle.packed_long(-1) == b"\x01" + bytes(6) + b"\x80"
    assert oracle.packed_long(0x3034464649445342) == b"BSDIFF40"   # Constants.cs:12
    assert oracle.packed_long(-0x0102030405060708) == bytes([8, 7, 6, 5, 4, 3, 2, 0x81])


def _search_by_definition(sa, old, q):
    """Search is a function of L = #{old suffixes < q}."""
    n = len(old)
    ob, qb = bytes(old), bytes(q)
    L = sum(1 for p in sa if ob[p:] < qb)
    I = list(sa) + [0]
    if n == 0:
        s = e = 0
    else:
        e = max(L, 1); s = e - 1

    def ml(p):
        k = 0
        while p + k < n and k < len(qb) and ob[p + k] == qb[k]:
            k += 1
        return k
    x, y = ml(I[s]), ml(I[e])
    return (I[s], x) if x > y else (I[e], y)


@pytest.mark.parametrize("seed", range(6))
def test_search_matches_definition(seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(0, 60))
    sigma = int(rng.integers(1, 4))
    old = rng.integers(0, sigma, n, dtype=np.uint8)
    new = rng.integers(0, sigma, int(rng.integers(0, 60)), dtype=np.uint8)
    sa = oracle.sais(old)
    I = oracle.make_I(sa)
    pos, ln = oracle.search_all(I, old, new)
    for j in range(new.size):
        assert (pos[j], ln[j]) == _search_by_definition(sa, old, new[j:]), (seed, j)


def test_bsdiff_golden_cases_reproduce():
    g = np.load("tests/golden/bsdiff_cases.npz") if False else np.load(
        __import__("os").path.join(__import__("conftest").GOLDEN, "bsdiff_cases.npz"))
    for k in range(int(g["count"])):
        r = oracle.bsdiff_streams(g[f"c{k}_old"], g[f"c{k}_new"], trace=True)
        for s in ("ctrl", "diff", "extra"):
            assert r[s] == g[f"c{k}_{s}"].tobytes(), (k, s)
        assert np.array_equal(r["trace_pos"], g[f"c{k}_trace_pos"])


def _apply_streams(old, ctrl, diff, extra, newsize):
    """Patch.ApplyInternal (Patch.cs:95-168) on uncompressed streams."""
    def rd(b):
        y = int.from_bytes(b[:7], "little") | ((b[7] & 0x7F) << 56)
        return -y if b[7] & 0x80 else y
    out = bytearray()
    op = cp = dp = ep = 0
    while len(out) < newsize:
        add, copy, seek = rd(ctrl[cp:cp + 8]), rd(ctrl[cp + 8:cp + 16]), rd(ctrl[cp + 16:cp + 24])
        cp += 24
        assert len(out) + add <= newsize
        seg = np.frombuffer(diff[dp:dp + add], np.uint8) + np.frombuffer(bytes(old[op:op + add]), np.uint8)
        out += seg.astype(np.uint8).tobytes(); dp += add; op += add
        assert len(out) + copy <= newsize
        out += extra[ep:ep + copy]; ep += copy
        op += seek
    return bytes(out)


@pytest.mark.parametrize("size", [0, 1, 512, 999, 1024, 4096])
def test_bsdiff_roundtrip(size):
    # BsDiffTests.cs:30-78 (old == new there, both from one seed) + an unrelated pair (BsPatchTests.cs:18-38)
    old = random_bytes(size)
    for new in (old.copy(), random_bytes(size + 7, seed=5)):
        r = oracle.bsdiff_streams(old, new)
        assert _apply_streams(old.tobytes(), r["ctrl"], r["diff"], r["extra"], new.size) == new.tobytes()


# ---- LibDivSufSort restatement (oracle/divsufsort.c): the reference's DEFAULT sorter ---------------------------

def _dss_texts():
    from deltaq_b200 import workloads as w
    out = dict(adversarial_texts())
    for name in asset_names():
        out[name] = load_asset(name)
    for size in REF_RANDOM_SIZES:
        out[f"random_{size}"] = random_bytes(size)
    out["shruggy"] = np.frombuffer(SHRUGGY, dtype=np.uint8)
    # sizes that reach the parts small fixtures do not: sssort's block merges (> 1024 B* suffixes per bucket),
    # trsort's budget and tandem-repeat copies, the in-place merge with a small buffer
    out["random_300k"] = random_bytes(300_000, seed=5)
    out["repetitive_1m"] = w.c3_repetitive(1 << 20)
    out["genome_2m"] = w.c4_genome(2 << 20)
    out["exe_like_2m"] = w.c2_exe_pair(2 << 20, (2 << 20) + 100)[0]
    out["fibonacci_300k"] = w.c3_fibonacci(300_000)
    out["two_letters_200k"] = np.random.default_rng(3).integers(0, 2, 200_000, dtype=np.uint8)
    return out


def test_divsufsort_matches_the_reference_checkers_and_sais():
    """LibDivSufSortTests.cs: every fixture, the shruggy string and the random sizes pass Verify (sufcheck + strictly
    increasing adjacent suffixes); the result equals the SA-IS restatement's (the suffix array is unique)."""
    for name, t in _dss_texts().items():
        sa = oracle.divsufsort(t)
        assert sa.size == t.size, name
        oracle.verify(t, sa)
        assert np.array_equal(sa, oracle.sais(t)), name


def test_divsufsort_does_not_need_a_zeroed_buffer_and_writes_n_entries():
    # LibDivSufSort.cs:14 allocates without clearing; Diff.cs:90 relies on I[n] staying untouched
    import ctypes
    t = random_bytes(5000)
    buf = np.full(t.size + 1, -7, dtype=np.int32)
    assert oracle.lib().oracle_divsufsort(ctypes.c_void_p(t.ctypes.data), t.size, ctypes.c_void_p(buf.ctypes.data)) == 0
    assert buf[t.size] == -7 and np.array_equal(buf[:t.size], oracle.sais(t))


# ---- second pin of the delta streams: literal Python transliteration of Diff.cs (tests/diff_transliteration.py) -----

def test_python_transliteration_of_diff_cs_agrees_with_the_oracle_on_the_golden_cases():
    import os
    from conftest import GOLDEN
    import diff_transliteration as dt
    g = np.load(os.path.join(GOLDEN, "bsdiff_cases.npz"))
    checked = 0
    for k in range(int(g["count"])):
        old, new = g[f"c{k}_old"], g[f"c{k}_new"]
        if old.size > 6000:
            continue                      # the literal per-byte Python loops are for small cases
        ctrl, diff, extra, trace = dt.diff_streams(old.tobytes(), new.tobytes())
        assert ctrl == g[f"c{k}_ctrl"].tobytes(), k
        assert diff == g[f"c{k}_diff"].tobytes(), k
        assert extra == g[f"c{k}_extra"].tobytes(), k
        tp = np.full(new.size, -1, np.int32)          # the golden trace: (pos, len) at the visited scan positions
        tl = np.full(new.size, -1, np.int32)
        for scan, p, ln in trace:
            tp[scan], tl[scan] = p, ln
        assert np.array_equal(tp, g[f"c{k}_trace_pos"]) and np.array_equal(tl, g[f"c{k}_trace_len"]), k
        r = oracle.bsdiff_streams(old, new)
        assert (r["ctrl"], r["diff"], r["extra"]) == (ctrl, diff, extra), k
        checked += 1
    assert checked >= 8


def test_python_transliteration_on_structured_pairs():
    import diff_transliteration as dt
    rng = np.random.default_rng(21)
    base = rng.integers(0, 256, 1500, dtype=np.uint8)
    pairs = {
        "zero_runs": (np.concatenate([np.zeros(700, np.uint8), base[:300], np.zeros(500, np.uint8)]),
                      np.concatenate([np.zeros(650, np.uint8), base[:310], np.zeros(540, np.uint8)])),
        "shifted": (base, np.concatenate([rng.integers(0, 256, 37, dtype=np.uint8), base])),
        "periodic": (np.tile(np.array([1, 2, 3, 4, 5], np.uint8), 300), np.tile(np.array([1, 2, 3, 4, 5], np.uint8), 280)[3:]),
        "old_one_byte": (np.array([7], np.uint8), base[:200]),
        "old_empty": (np.zeros(0, np.uint8), base[:100]),
        "new_empty": (base[:100], np.zeros(0, np.uint8)),
    }
    for name, (old, new) in pairs.items():
        ctrl, diff, extra, _ = dt.diff_streams(old.tobytes(), new.tobytes())
        r = oracle.bsdiff_streams(old, new)
        assert (r["ctrl"], r["diff"], r["extra"]) == (ctrl, diff, extra), name
