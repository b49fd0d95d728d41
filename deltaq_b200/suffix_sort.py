"""CudaSuffixSort -- host-side mirror of the reference's suffix-sort provider contract.

Mirrors ``DeltaQ.SuffixSorting.ISuffixSort`` (/root/reference/src/DeltaQ.SuffixSorting.Abstractions/
ISuffixSort.cs:9-28) as the reference's providers implement it (LibDivSufSort.cs:10-32, SAIS.cs:11-45):

    Sort(ReadOnlySpan<byte> text) -> IMemoryOwner<int>          ->  sort(text) -> SuffixArrayOwner
    Sort(ReadOnlySpan<byte> text, Span<int> suffixes)           ->  sort(text, suffixes)

Same argument meaning and error behaviour: the two-buffer overload raises ``ValueError("Text and suffix
buffers should have the same length")`` (ArgumentException, LibDivSufSort.cs:23-31); empty input yields
an empty owner; only ``suffixes[0:n]`` is written.  The C# provider that binds the same C ABI is in
csharp/DeltaQ.SuffixSorting.Cuda/ (see INTEGRATION.md).
"""
import numpy as np

from . import _native

LENGTH_MISMATCH = "Text and suffix buffers should have the same length"  # LibDivSufSort.cs:31, SAIS.cs:44


def as_bytes_array(buf, name="text"):
    if buf is None:
        raise TypeError(f"{name} must not be None")  # ArgumentNullException
    if isinstance(buf, np.ndarray):
        if buf.dtype != np.uint8:
            raise TypeError(f"{name} must be a uint8 buffer")
        return np.ascontiguousarray(buf).reshape(-1)
    return np.frombuffer(bytes(buf) if not isinstance(buf, (bytes, bytearray, memoryview)) else buf, dtype=np.uint8)


class SuffixArrayOwner:
    """IMemoryOwner<int> analogue: ``.memory`` is an int32 array of length n over pinned host memory;
    ``dispose()`` (or the context manager) releases it (README.md:108-110 of the reference)."""

    def __init__(self, pinned, n, release=None):
        self._pinned = pinned
        self._release = release        # hands the buffer back to the provider's pool; None: free it
        self.memory = pinned.array[:n]

    def dispose(self):
        if self._pinned is not None:
            self.memory = None
            pinned, self._pinned = self._pinned, None
            if self._release is not None:
                self._release(pinned)
            else:
                pinned.free()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.dispose()


class CudaSuffixSort:
    """ISuffixSort provider backed by libdeltaq_cuda.  Holds a native context (device, stream, scratch);
    disposable; one call in flight per instance.  Raises if the CUDA library or device is missing."""

    # The owners Sort(text) hands out sit on pinned host memory, and pinning is dear (cudaHostAlloc of a few MiB costs as
    # much as sorting them).  The reference rents its owners from a pool (MemoryOwner<int>.Allocate -> ArrayPool,
    # LibDivSufSort.cs:14), so Sort(asset).Dispose() in a loop -- its benchmark, SuffixSortingBenchmarks.cs:63-73 -- reuses
    # one array; so does this provider: a disposed owner's buffer is kept (a few, of moderate size) for the next Sort.
    _POOL_KEEP = 4
    _POOL_MAX_BYTES = 256 << 20

    def __init__(self, device=None, _lib=None):
        import threading
        self._ctx = _native.Context(device=device, lib=_lib)
        self._pool = []
        self._pool_lock = threading.Lock()
        self._closed = False

    def _rent(self, n):
        """A pinned int32 buffer of at least n entries: the smallest kept one that fits without being more than twice too
        big, else a new one (capacities below 16 Mi entries are rounded up to a power of two, like ArrayPool's buckets)."""
        with self._pool_lock:
            fit = [b for b in self._pool if n <= b.array.size <= max(2 * n, 1024)]
            if fit:
                best = min(fit, key=lambda b: b.array.size)
                self._pool.remove(best)
                return best
        cap = max(1, n)
        if cap < (16 << 20):
            cap = 1 << (cap - 1).bit_length()
        return self._ctx.pinned(cap, np.int32)

    def _give_back(self, pinned):
        with self._pool_lock:
            if not self._closed and pinned.array.nbytes <= self._POOL_MAX_BYTES and len(self._pool) < self._POOL_KEEP:
                self._pool.append(pinned)
                return
        pinned.free()

    # -- IDisposable
    def dispose(self):
        with self._pool_lock:
            self._closed = True
            pool, self._pool = self._pool, []
        for b in pool:
            b.free()
        self._ctx.close()

    close = dispose

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.dispose()

    @property
    def context(self):
        return self._ctx

    def stats(self):
        return self._ctx.stats()

    def sort(self, text, suffixes=None):
        t = as_bytes_array(text)
        if suffixes is None:
            # overload 1: ISuffixSort.cs:18 -- allocate, sort, hand the owner to the caller
            pinned = self._rent(t.size)
            owner = SuffixArrayOwner(pinned, t.size, release=self._give_back)
            try:
                self._ctx.suffix_sort(t, pinned.array)
            except BaseException:
                owner.dispose()
                raise
            return owner
        # overload 2: ISuffixSort.cs:27
        if suffixes is None or not isinstance(suffixes, np.ndarray) or suffixes.dtype != np.int32:
            raise TypeError("suffixes must be an int32 numpy array")
        if suffixes.size != t.size:
            raise ValueError(LENGTH_MISMATCH)
        if not suffixes.flags.c_contiguous:
            raise ValueError("suffixes must be contiguous")
        self._ctx.suffix_sort(t, suffixes)
        return None

    def lcp_array(self, text, suffixes=None):
        """LCP array of `text` (int32, lcp[0] = 0, lcp[r] = common prefix of suffixes r-1 and r in sorted order): the
        extension SURVEY.md 8(f) rank 4 proposes for ISuffixSort providers.  suffixes=None uses the suffix array this
        provider's last sort(text, ...) left on the device; otherwise any suffix array of `text` (validated)."""
        t = as_bytes_array(text)
        I = None
        if suffixes is not None:
            I = np.ascontiguousarray(suffixes, dtype=np.int32)
            if I.size != t.size:
                raise ValueError(LENGTH_MISMATCH)
        return self._ctx.lcp(t, I)
