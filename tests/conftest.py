import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")

# The host loop skips stretches of `new` the scan has proved equal to `old` (Cert, dq_diff_host.h); under the tests every
# such stretch is compared byte for byte as well, and a false one fails the call.
os.environ.setdefault("DQ_CHECK_CERTS", "1")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def asset_names():
    return sorted(os.listdir(os.path.join(GOLDEN, "assets")))


def load_asset(name):
    return np.fromfile(os.path.join(GOLDEN, "assets", name), dtype=np.uint8)


def load_golden_sa(name):
    return np.load(os.path.join(GOLDEN, "sa", name + ".npy"))


def random_bytes(size, seed=63 * 13 * 63 * 13):
    """Seeded uniform bytes.  The reference seeds System.Random with 63*13*63*13
    (LibDivSufSortTests.cs:29); .NET's generator is not reproducible here, and the checks are
    property-based, so any seeded uniform stream is the same test."""
    return np.random.default_rng(seed).integers(0, 256, size, dtype=np.uint8)


# Reference test sizes: LibDivSufSortTests.cs:126-137, SAISTester.cs:35
REF_RANDOM_SIZES = [0, 1, 2, 4, 8, 16, 32, 51, 0x1000, 0x8000 - 1, 0x8000]

# LDSS-listed fixtures (LibDivSufSortTests.cs:87-106); SAISTester.cs:53-55 runs every file
LDSS_FIXTURES = [
    "fuzz1", "fuzz2", "fuzz3",
    "crash-cf8673530fdca659e0ddf070b4718b9c0bb504ec",
    "crash-ce407adf7cf638d3fa89b5637a94355d7d658872",
    "crash-c792e788de61771b6cd65c1aa5670c62e57a33c4",
    "crash-90b42d1c55ee90a8b004fb9db1853429ceb4c4ba",
    "crash-8765ef2258178ca027876eab83e01d6d58db9ca0",
    "crash-4f8c31dec8c3678a07e0fbacc6bd69e7cc9037fb",
    "crash-16356e91966a827f79e49167170194fc3088a7ab",
    "crash-aoob-ss_mintrosort",
]

SHRUGGY = "¯\\_(ツ)_/¯".encode("utf-8")  # LibDivSufSortTests.cs:69-71


def adversarial_texts():
    """Inputs aimed at the hard parts listed in SURVEY.md §7 (H1 end-of-text, long LCPs)."""
    rng = np.random.default_rng(7)
    out = {}
    out["zeros_tail"] = np.concatenate([rng.integers(0, 256, 100, dtype=np.uint8), np.zeros(9, np.uint8)])
    out["all_zero_7"] = np.zeros(7, np.uint8)
    out["all_zero_8"] = np.zeros(8, np.uint8)
    out["all_zero_9"] = np.zeros(9, np.uint8)
    out["all_zero_1000"] = np.zeros(1000, np.uint8)
    out["all_ff_777"] = np.full(777, 255, np.uint8)
    out["zero_then_one"] = np.concatenate([np.zeros(300, np.uint8), np.ones(1, np.uint8), np.zeros(300, np.uint8)])
    out["period2"] = np.tile(np.array([0xFF, 0xF3], np.uint8), 1500)
    out["period3_tail0"] = np.concatenate([np.tile(np.array([1, 0, 0], np.uint8), 700), np.zeros(5, np.uint8)])
    out["period8"] = np.tile(np.arange(8, dtype=np.uint8), 400)
    out["period8_zero_end"] = np.concatenate([np.tile(np.array([0, 0, 0, 0, 0, 0, 0, 1], np.uint8), 300),
                                              np.zeros(8, np.uint8)])
    fib_a, fib_b = b"a", b"ab"
    while len(fib_b) < 5000:
        fib_a, fib_b = fib_b, fib_b + fib_a
    out["fibonacci"] = np.frombuffer(fib_b, dtype=np.uint8).copy()
    out["binary_random"] = rng.integers(0, 2, 6000, dtype=np.uint8)
    out["acgt"] = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, 9000)]
    para = rng.integers(97, 123, 512, dtype=np.uint8)
    rep = np.tile(para, 24)
    rep[rng.integers(0, rep.size, 12)] = 32
    out["repeated_paragraph"] = rep
    # equal-byte runs (run-length refinement of round 1): equal lengths, both classes (next byte below / above the
    # run byte), runs touching each other and the text end
    runs = []
    for b, L, c in [(5, 40, 3), (5, 40, 9), (5, 41, 3), (5, 39, 9), (5, 40, 3), (0, 100, 1), (255, 100, 0), (5, 8, 3),
                    (5, 9, 9), (7, 300, 7), (5, 40, 5)]:
        runs.append(np.full(L, b, np.uint8))
        runs.append(np.array([c, int(rng.integers(0, 256))], np.uint8))
    out["runs_mixed"] = np.concatenate(runs)
    out["runs_to_end"] = np.concatenate(runs + [np.full(64, 5, np.uint8)])
    out["runs_two_bytes"] = np.concatenate([np.full(int(L), int(b), np.uint8)
                                            for b, L in zip(rng.integers(0, 2, 60), rng.integers(1, 90, 60))])
    out["runs_long_zero_islands"] = np.zeros(30000, np.uint8)
    out["runs_long_zero_islands"][[5000, 5001, 12000, 20000, 20001, 20002, 29990]] = [9, 9, 1, 200, 0, 3, 7]
    return out


def small_alphabet_texts():
    """Texts over at most 16 byte values (the sorter then packs 16 / 32 / 64 characters per key: dq_suffix.cuh, "small
    alphabets"), plus neighbours of the thresholds; ends in the smallest character exercise the end-of-text rule with
    keys longer than 8 characters."""
    rng = np.random.default_rng(5)
    out = {}
    out["bin_70000"] = rng.integers(0, 2, 70000, dtype=np.uint8)
    out["acgt_50000"] = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, 50000)]
    out["acgtn_30000"] = np.frombuffer(b"ACGNT", dtype=np.uint8)[rng.integers(0, 5, 30000)]
    out["hex16_30000"] = rng.integers(100, 116, 30000, dtype=np.uint8)
    out["sigma17_9000"] = rng.integers(0, 17, 9000, dtype=np.uint8)
    out["one_value_5000"] = np.full(5000, 7, np.uint8)
    out["bin_zero_tail"] = np.concatenate([rng.integers(0, 2, 6000, dtype=np.uint8), np.zeros(70, np.uint8)])
    out["acgt_a_tail"] = np.concatenate([np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, 7000)],
                                         np.full(40, ord("A"), np.uint8)])
    out["acgt_tandem"] = np.tile(np.frombuffer(b"ACGTTGCAAC", dtype=np.uint8), 900)
    out["bin_4163"] = rng.integers(0, 2, 4163, dtype=np.uint8)
    return out
