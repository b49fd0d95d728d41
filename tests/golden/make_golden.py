"""Regenerates tests/golden/ (run in the build container, where /root/reference exists).

* assets/<name>      byte-for-byte copies of the reference's 13 suffix-sort fixture inputs
                     (/root/reference/test/assets/*: LibDivSufSortTests.cs:87-106, SAISTester.cs:53-55).
                     They are test DATA (minimised AFL crashers), not source.
* sa/<name>.npy      the suffix array of each fixture, from oracle.sais after it passed the restated
                     reference checkers (sufcheck + Verify) and matched the definition-level sorter.
* bsdiff_cases.npz   small seeded (old, new) pairs with the oracle's uncompressed ctrl/diff/extra
                     streams and the (pos, len) Search trace at the positions Diff.Create visits.

The reference itself (managed C#) cannot run in this image (no .NET), so known answers come from the
oracle restatement pinned by the reference's own property checks; see oracle/README.md.
"""
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle  # noqa: E402

REF_ASSETS = "/root/reference/test/assets"


def bsdiff_case(seed, n, kind):
    rng = np.random.default_rng(seed)
    if kind == "identical":
        old = rng.integers(0, 256, n, dtype=np.uint8)
        new = old.copy()
    elif kind == "unrelated":
        old = rng.integers(0, 256, n, dtype=np.uint8)
        new = rng.integers(0, 256, n + n // 7, dtype=np.uint8)
    elif kind == "mutated":
        old = rng.integers(0, 4, n, dtype=np.uint8) if seed % 2 else rng.integers(0, 256, n, dtype=np.uint8)
        new = bytearray(old.tobytes())
        for _ in range(max(1, n // 400)):
            p = int(rng.integers(0, max(1, len(new))))
            op = int(rng.integers(0, 3))
            k = int(rng.integers(1, 40))
            if op == 0:
                new[p:p + k] = rng.integers(0, 256, k, dtype=np.uint8).tobytes()
            elif op == 1:
                new[p:p] = rng.integers(0, 256, k, dtype=np.uint8).tobytes()
            else:
                del new[p:p + k]
        new = np.frombuffer(bytes(new), dtype=np.uint8)
    elif kind == "zeros":
        old = np.zeros(n, dtype=np.uint8)
        old[rng.integers(0, n, max(1, n // 50))] = 1
        new = np.zeros(n + 17, dtype=np.uint8)
        new[rng.integers(0, n + 17, max(1, n // 60))] = 1
    else:
        raise ValueError(kind)
    return old, new


def main():
    os.makedirs(os.path.join(HERE, "assets"), exist_ok=True)
    os.makedirs(os.path.join(HERE, "sa"), exist_ok=True)
    for name in sorted(os.listdir(REF_ASSETS)):
        src = os.path.join(REF_ASSETS, name)
        dst = os.path.join(HERE, "assets", name)
        shutil.copyfile(src, dst)
        os.chmod(dst, 0o644)
        t = np.fromfile(dst, dtype=np.uint8)
        sa = oracle.sais(t)
        oracle.verify(t, sa)
        assert np.array_equal(sa, oracle.sa_naive(t)), name
        np.save(os.path.join(HERE, "sa", name + ".npy"), sa)
        print(f"{name}: n={t.size} distinct={np.unique(t).size}")

    cases = {}
    specs = [(1, 0, "identical"), (2, 1, "identical"), (3, 512, "identical"), (4, 999, "unrelated"),
             (5, 1024, "mutated"), (6, 4096, "mutated"), (7, 3000, "zeros"), (8, 0x123, "unrelated"),
             (9, 20000, "mutated"), (10, 1, "unrelated"), (11, 2, "mutated")]
    for k, (seed, n, kind) in enumerate(specs):
        old, new = bsdiff_case(seed, n, kind)
        r = oracle.bsdiff_streams(old, new, trace=True)
        cases[f"c{k}_old"] = old
        cases[f"c{k}_new"] = new
        for s in ("ctrl", "diff", "extra"):
            cases[f"c{k}_{s}"] = np.frombuffer(r[s], dtype=np.uint8)
        cases[f"c{k}_trace_pos"] = r["trace_pos"]
        cases[f"c{k}_trace_len"] = r["trace_len"]
        print(f"case {k} {kind} n={old.size} m={new.size} ctrl={len(r['ctrl'])//24} triples "
              f"diff={len(r['diff'])} extra={len(r['extra'])} searches={r['search_calls']}")
    cases["count"] = np.array(len(specs))
    np.savez_compressed(os.path.join(HERE, "bsdiff_cases.npz"), **cases)


if __name__ == "__main__":
    main()
