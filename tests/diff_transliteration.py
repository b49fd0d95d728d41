"""A SECOND, independent restatement of the reference's Diff.Create loop: a literal Python transliteration of
/root/reference/src/DeltaQ.BsDiff/Diff.cs:92-298 and SpanExtensions.cs:7-30, statement by statement, sharing no code
with oracle/bsdiff.c.  TEST INFRASTRUCTURE ONLY.

Why: the reference holds no golden delta and cannot run here (no .NET), so the delta streams are pinned only by
restatements written in this repository.  Two restatements in two languages that agree byte for byte on the golden
cases are a stronger pin than one -- it is STILL not an execution of the reference (oracle/README.md).

The suffix array comes from a definition-level sort (sorted() over the suffixes themselves), so nothing here depends on
oracle/ or on the CUDA library either."""


def write_packed_long(y):          # SpanExtensions.cs:7-30
    buf = bytearray(8)
    if y < 0:
        y = -y
        buf[7] = ((y >> 56) | 0x80) & 0xFF
    else:
        buf[7] = (y >> 56) & 0xFF
    buf[6] = (y >> 48) & 0xFF
    buf[5] = (y >> 40) & 0xFF
    buf[4] = (y >> 32) & 0xFF
    buf[3] = (y >> 24) & 0xFF
    buf[2] = (y >> 16) & 0xFF
    buf[1] = (y >> 8) & 0xFF
    buf[0] = y & 0xFF
    return bytes(buf)


def suffix_array(old):             # what any ISuffixSort returns: suffix starts in Span.SequenceCompareTo order
    return sorted(range(len(old)), key=lambda i: old[i:])   # bytes compare: first difference, else shorter first


def compare_bytes(left, right):    # Diff.cs:245-246, Span.SequenceCompareTo
    return -1 if left < right else (1 if left > right else 0)


def match_length(old_data, new_data):   # Diff.cs:249-265
    i = 0
    while i < len(old_data) and i < len(new_data):
        if old_data[i] != new_data[i]:
            break
        i += 1
    return i


def search(I, old_data, new_data, start, end):    # Diff.cs:267-298 -> (len, pos)
    while True:
        if end - start < 2:
            x = match_length(old_data[I[start]:], new_data)
            y = match_length(old_data[I[end]:], new_data)
            if x > y:
                return x, I[start]
            return y, I[end]
        mid_point = start + (end - start) // 2
        if compare_bytes(old_data[I[mid_point]:], new_data) < 0:
            start = mid_point
        else:
            end = mid_point


def diff_streams(old_data, new_data):
    """Diff.cs:78-223 -> (ctrl, diff, extra) uncompressed, and (scan, pos, len) of every Search call in order."""
    old_data = bytes(old_data)
    new_data = bytes(new_data)
    I = suffix_array(old_data) + [0]       # Diff.cs:78, :90 -- n+1 entries, the last one stays 0
    ctrl = bytearray()
    diff = bytearray()
    extra = bytearray()
    trace = []
    scan = 0
    pos = 0
    length = 0
    lastscan = 0
    lastpos = 0
    lastoffset = 0
    while scan < len(new_data):
        oldscore = 0
        scan += length
        scsc = scan
        while scan < len(new_data):
            length, pos = search(I, old_data, new_data[scan:], 0, len(old_data))
            trace.append((scan, pos, length))
            while scsc < scan + length:
                if scsc + lastoffset < len(old_data) and old_data[scsc + lastoffset] == new_data[scsc]:
                    oldscore += 1
                scsc += 1
            if (length == oldscore and length != 0) or (length > oldscore + 8):
                break
            if scan + lastoffset < len(old_data) and old_data[scan + lastoffset] == new_data[scan]:
                oldscore -= 1
            scan += 1
        if length != oldscore or scan == len(new_data):
            s = 0
            sf = 0
            lenf = 0
            i = 0
            while lastscan + i < scan and lastpos + i < len(old_data):
                if old_data[lastpos + i] == new_data[lastscan + i]:
                    s += 1
                i += 1
                if s * 2 - i > sf * 2 - lenf:
                    sf = s
                    lenf = i
            lenb = 0
            if scan < len(new_data):
                s = 0
                sb = 0
                i = 1
                while scan >= lastscan + i and pos >= i:
                    if old_data[pos - i] == new_data[scan - i]:
                        s += 1
                    if s * 2 - i > sb * 2 - lenb:
                        sb = s
                        lenb = i
                    i += 1
            if lastscan + lenf > scan - lenb:
                overlap = (lastscan + lenf) - (scan - lenb)
                s = 0
                ss = 0
                lens = 0
                for i in range(overlap):
                    if new_data[lastscan + lenf - overlap + i] == old_data[lastpos + lenf - overlap + i]:
                        s += 1
                    if new_data[scan - lenb + i] == old_data[pos - lenb + i]:
                        s -= 1
                    if s > ss:
                        ss = s
                        lens = i + 1
                lenf += lens - overlap
                lenb -= lens
            for i in range(lenf):
                diff.append((new_data[lastscan + i] - old_data[lastpos + i]) & 0xFF)
            extra_length = (scan - lenb) - (lastscan + lenf)
            if extra_length > 0:
                extra += new_data[lastscan + lenf:lastscan + lenf + extra_length]
            ctrl += write_packed_long(lenf)
            ctrl += write_packed_long(extra_length)
            ctrl += write_packed_long((pos - lenb) - (lastpos + lenf))
            lastscan = scan - lenb
            lastpos = pos - lenb
            lastoffset = pos - scan
    return bytes(ctrl), bytes(diff), bytes(extra), trace
