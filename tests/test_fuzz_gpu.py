"""Short randomised parity sweep on the GPU (scripts/gpu_fuzz.py): sorts and searches vs the oracle."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [11, 12])
def test_fuzz(seed):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "gpu_fuzz.py"), str(seed), "12"],
                       capture_output=True, text=True, cwd=ROOT, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "0 mismatches" in r.stdout


def _mid_size_pair(rng):
    """A 0.3-3 MB pair with the ingredients the coded table and the host loop care about: long unchanged stretches,
    point damage (matches the current alignment almost fits), inserted unrelated data (short matches), moved blocks,
    equal-byte runs and periodic records."""
    import numpy as np
    n = int(rng.integers(300_000, 3_000_000))
    sigma = int(rng.choice([4, 16, 256]))
    old = rng.integers(0, sigma, n, dtype=np.uint8)
    a = int(rng.integers(0, n // 2))
    old[a:a + int(rng.integers(1000, 200_000))] = 0
    b = int(rng.integers(0, n // 2))
    rec = rng.integers(0, 256, int(rng.integers(8, 64)), dtype=np.uint8)
    k = int(rng.integers(1000, 150_000))
    old[b:b + k] = np.resize(rec, k)[:old[b:b + k].size]
    parts, cur = [], 0
    for c in np.sort(rng.integers(0, n, int(rng.integers(3, 40)))):
        c = int(max(c, cur))
        parts.append(old[cur:c])
        op = int(rng.integers(0, 4))
        k = int(rng.integers(1, 60_000))
        if op == 0:
            parts.append(rng.integers(0, sigma, k, dtype=np.uint8)); cur = min(n, c + k)
        elif op == 1:
            parts.append(rng.integers(0, 256, k, dtype=np.uint8)); cur = c
        elif op == 2:
            cur = min(n, c + k)
        else:
            s = int(rng.integers(0, n - k)); parts.append(old[s:s + k]); cur = c
    parts.append(old[cur:])
    new = np.concatenate(parts).copy()
    hits = rng.integers(0, new.size, int(rng.integers(0, 200)))
    new[hits] ^= 1
    return old, new


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_streams_mid_size_pairs_identical_to_oracle(seed):
    # dq_cuda_bsdiff_streams end to end (coded table in slices on 8 streams, scan / extender + crew / writers)
    import numpy as np
    import oracle
    from deltaq_b200 import CudaSuffixSort, bsdiff
    rng = np.random.default_rng(seed)
    with CudaSuffixSort() as s:
        for _ in range(3):
            old, new = _mid_size_pair(rng)
            ref = oracle.bsdiff_streams(old, new)
            got = bsdiff.create_streams(old, new, s)
            for k in ("ctrl", "diff", "extra"):
                assert got[k] == ref[k], (seed, k, old.size, new.size)
            assert got["search_visits"] == ref["search_calls"]
            assert s._ctx.stats()["table_fallbacks"] == 0
