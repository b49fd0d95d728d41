"""Seeded synthetic inputs for the five BASELINE.json configurations (recipes: SURVEY.md section 8(d)).

numpy only; used by bench.py and the tests so that both sides see the same bytes."""
import numpy as np

MIB = 1 << 20


def c1_uniform(n=MIB, seed=670761):
    """C1: uniform random bytes; the reference's benchmark seed constant (SuffixSortingBenchmarks.cs:15)."""
    return np.random.default_rng(seed).integers(0, 256, n, dtype=np.uint8)


def _exe_like(n, rng):
    """Concatenation of sections: 40 % Zipf 'code', 20 % repeated records, 15 % zero runs, 25 % uniform."""
    parts = []
    remaining = n
    zipf_p = 1.0 / np.arange(1, 257) ** 1.2
    zipf_p /= zipf_p.sum()
    perm = rng.permutation(256).astype(np.uint8)
    while remaining > 0:
        sec = int(min(remaining, rng.integers(64 * 1024, 1024 * 1024)))
        kind = rng.choice(4, p=[0.40, 0.20, 0.15, 0.25])
        if kind == 0:
            part = perm[rng.choice(256, size=sec, p=zipf_p)]
        elif kind == 1:
            rec_len = int(rng.integers(16, 65))
            nrec = sec // rec_len + 1
            base = rng.integers(0, 256, rec_len, dtype=np.uint8)
            recs = np.tile(base, (nrec, 1))
            # a few fields vary per record (table / relocation entries)
            cols = rng.choice(rec_len, size=max(1, rec_len // 8), replace=False)
            recs[:, cols] = rng.integers(0, 256, (nrec, cols.size), dtype=np.uint8)
            part = recs.reshape(-1)[:sec]
        elif kind == 2:
            part = np.zeros(sec, dtype=np.uint8)
            # sprinkle a few non-zero islands so the runs have different lengths
            for _ in range(int(rng.integers(1, 6))):
                p = int(rng.integers(0, sec))
                k = int(min(sec - p, rng.integers(1, 64)))
                part[p:p + k] = rng.integers(1, 256, k, dtype=np.uint8)
        else:
            part = rng.integers(0, 256, sec, dtype=np.uint8)
        parts.append(part)
        remaining -= sec
    return np.concatenate(parts)[:n]


def _mutate(old, rng, target_len, regions=200, frac=0.13):
    """~frac of the bytes inside `regions` mutated regions (overwrite / insert / delete), lengths
    log-uniform in [64 B, 256 KiB] scaled to the input; net growth to target_len by insertions."""
    n = old.size
    scale = n / (16 * MIB)
    lens = np.exp(rng.uniform(np.log(64), np.log(256 * 1024 * max(scale, 1e-3)), regions)).astype(np.int64)
    lens = np.maximum(1, (lens * (frac * n / max(1, lens.sum()))).astype(np.int64))
    starts = np.sort(rng.integers(0, max(1, n - 1), regions))
    out = []
    cur = 0
    grow = target_len - n
    for s, ln in zip(starts, lens):
        s = int(max(s, cur))
        if s >= n:
            break
        out.append(old[cur:s])
        op = int(rng.integers(0, 3))
        if op == 0:      # overwrite
            e = min(n, s + int(ln))
            out.append(rng.integers(0, 256, e - s, dtype=np.uint8))
            cur = e
        elif op == 1:    # insert
            out.append(rng.integers(0, 256, int(ln), dtype=np.uint8))
            cur = s
        else:            # delete
            cur = min(n, s + int(ln))
    out.append(old[cur:])
    new = np.concatenate(out)
    if new.size < target_len:
        extra = rng.integers(0, 256, target_len - new.size, dtype=np.uint8)
        p = int(rng.integers(0, new.size))
        new = np.concatenate([new[:p], extra, new[p:]])
    return np.ascontiguousarray(new[:target_len])


def c2_exe_pair(n_old=16 * MIB, n_new=17 * MIB, seed_old=1, seed_new=2):
    """C2: 16 MiB -> 17 MiB executable-like pair, ~13 % of bytes in ~200 mutated regions."""
    old = _exe_like(n_old, np.random.default_rng(seed_old))
    new = _mutate(old, np.random.default_rng(seed_new), n_new)
    return old, new


def c3_repetitive(n=64 * MIB, seed=3):
    """C3: a 4 KiB paragraph repeated, 0.1 % random point edits (long LCPs, many doubling rounds)."""
    rng = np.random.default_rng(seed)
    vocab = [rng.integers(97, 123, int(rng.integers(2, 10)), dtype=np.uint8) for _ in range(200)]
    words = []
    size = 0
    while size < 4096:
        w = vocab[int(rng.integers(0, len(vocab)))]
        words.append(w)
        words.append(np.array([32], dtype=np.uint8))
        size += w.size + 1
    para = np.concatenate(words)[:4096]
    text = np.tile(para, n // 4096 + 1)[:n].copy()
    k = max(1, n // 1000)
    text[rng.integers(0, n, k)] = rng.integers(0, 256, k, dtype=np.uint8)
    return text


def c3_fibonacci(n):
    a, b = b"a", b"ab"
    while len(b) < n:
        a, b = b, b + a
    return np.frombuffer(b[:n], dtype=np.uint8).copy()


def c4_genome(n=512 * MIB, seed=4):
    """C4: iid {A,C,G,T} with ~1 % of the bytes inside 1-10 KiB tandem repeats."""
    rng = np.random.default_rng(seed)
    text = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, n, dtype=np.uint8)]
    budget = n // 100
    while budget > 0 and n > 4096:
        unit = int(rng.integers(2, 64))
        total = int(min(budget, rng.integers(1024, 10 * 1024)))
        p = int(rng.integers(0, n - total - unit))
        reps = total // unit + 1
        text[p:p + total] = np.tile(text[p:p + unit], reps)[:total]
        budget -= total
    return text


C5_PIECES = 8


def _c5_piece(args):
    n, seed, i = args
    return _exe_like(n, np.random.default_rng([seed, i]))


def c5_old(n_old=2_040_109_466, seed=5, workers=1):
    """C5's old file: C2's generator scaled to ~1.9 GiB (near the int32 suffix-array limit), built as C5_PIECES
    independently seeded stretches of sections so that `workers` processes can generate it side by side (the bytes do
    not depend on `workers`)."""
    sizes = [n_old // C5_PIECES + (1 if i < n_old % C5_PIECES else 0) for i in range(C5_PIECES)]
    jobs = [(sz, seed, i) for i, sz in enumerate(sizes)]
    if workers > 1:
        import multiprocessing as mp
        from concurrent.futures import ProcessPoolExecutor
        with ProcessPoolExecutor(max_workers=workers, mp_context=mp.get_context("spawn")) as ex:
            parts = list(ex.map(_c5_piece, jobs))
    else:
        parts = [_c5_piece(j) for j in jobs]
    return np.concatenate(parts)


def c5_pair(n_old=2_040_109_466, seed=5, workers=1):
    """C5: (old, new) with new = old mutated like C2's pair, ~6 % longer."""
    old = c5_old(n_old, seed, workers)
    new = _mutate(old, np.random.default_rng(seed + 100), n_old + n_old // 16, regions=2000)
    return old, new
