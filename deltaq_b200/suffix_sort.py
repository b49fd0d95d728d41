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

    def __init__(self, pinned, n):
        self._pinned = pinned
        self.memory = pinned.array[:n]

    def dispose(self):
        if self._pinned is not None:
            self.memory = None
            self._pinned.free()
            self._pinned = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.dispose()


class CudaSuffixSort:
    """ISuffixSort provider backed by libdeltaq_cuda.  Holds a native context (device, stream, scratch);
    disposable; one call in flight per instance.  Raises if the CUDA library or device is missing."""

    def __init__(self, device=None, _lib=None):
        self._ctx = _native.Context(device=device, lib=_lib)

    # -- IDisposable
    def dispose(self):
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
            pinned = self._ctx.pinned(max(1, t.size), np.int32)
            owner = SuffixArrayOwner(pinned, t.size)
            self._ctx.suffix_sort(t, pinned.array)
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
