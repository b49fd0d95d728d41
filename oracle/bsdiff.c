/*
 * oracle/bsdiff.c -- CPU restatement of the reference's bsdiff match search and
 * of the greedy scan/emit loop that consumes it.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/README.md).
 *
 * Follows (paths relative to /root/reference):
 *   src/DeltaQ.BsDiff/Diff.cs:245-246   CompareBytes = Span.SequenceCompareTo
 *   src/DeltaQ.BsDiff/Diff.cs:249-265   MatchLength  = common prefix length
 *   src/DeltaQ.BsDiff/Diff.cs:267-298   Search (binary search over I, I has n+1
 *                                       entries and I[n] == 0: Diff.cs:78,90)
 *   src/DeltaQ.BsDiff/Diff.cs:92-223    greedy scan / extend / overlap / emit loop
 *   src/DeltaQ.BsDiff/SpanExtensions.cs:7-30  WritePackedLong (sign-magnitude LE)
 *
 * The three streams are returned UNCOMPRESSED: the reference pipes them through
 * SharpZipLib 1.4.2 BZip2OutputStream (DeltaQ.BsDiff.csproj:41-43), a
 * third-party dependency that is not vendored in /root/reference, and the
 * reference's tests pin the compressed bytes only by round trip
 * (BsDiffTests.cs:30-78) -- compressed-byte parity is unpinned; stream parity
 * one layer down is what the tests compare.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ---- Diff.cs:245-265 ---------------------------------------------------- */

static int compare_bytes(const uint8_t *l, int64_t ll, const uint8_t *r, int64_t rl)
{
    int64_t k = ll < rl ? ll : rl;
    int c = memcmp(l, r, (size_t)k);
    if (c != 0)
        return c;
    return (ll > rl) - (ll < rl);
}

/* MatchLength: Span.CommonPrefixLength on net7+ (Diff.cs:250-251, vectorised by the BCL) or the byte loop
 * of Diff.cs:253-263 -- same value; stepped 8 bytes at a time here so the timed CPU arm is not handicapped. */
static int32_t match_length(const uint8_t *a, int64_t al, const uint8_t *b, int64_t bl)
{
    int64_t i = 0, k = al < bl ? al : bl;
    while (i + 8 <= k) {
        uint64_t x, y;
        memcpy(&x, a + i, 8);
        memcpy(&y, b + i, 8);
        if (x != y)
            return (int32_t)(i + (__builtin_ctzll(x ^ y) >> 3));
        i += 8;
    }
    for (; i < k; ++i)
        if (a[i] != b[i])
            break;
    return (int32_t)i;
}

/* ---- Diff.cs:267-298 ---------------------------------------------------- */

int32_t oracle_search(const int32_t *I, const uint8_t *old, int32_t n,
                      const uint8_t *q, int32_t qlen,
                      int32_t start, int32_t end, int32_t *pos)
{
    for (;;) {
        if (end - start < 2) {
            int32_t x = match_length(old + I[start], n - I[start], q, qlen);
            int32_t y = match_length(old + I[end], n - I[end], q, qlen);
            if (x > y) {
                *pos = I[start];
                return x;
            }
            *pos = I[end];
            return y;
        }
        int32_t mid = start + (end - start) / 2;
        if (compare_bytes(old + I[mid], n - I[mid], q, qlen) < 0)
            start = mid;
        else
            end = mid;
    }
}

/* Search() at every scan position in [scan_begin, scan_begin+count): the
 * literal, quadratic replay.  Small inputs only. */
void oracle_search_all(const int32_t *I, const uint8_t *old, int32_t n,
                       const uint8_t *nw, int32_t m,
                       int32_t scan_begin, int32_t count,
                       int32_t *pos_out, int32_t *len_out)
{
    for (int32_t k = 0; k < count; ++k) {
        int32_t scan = scan_begin + k;
        int32_t pos = 0;
        len_out[k] = oracle_search(I, old, n, nw + scan, m - scan, 0, n, &pos);
        pos_out[k] = pos;
    }
}

/* ---- SpanExtensions.cs:7-30 --------------------------------------------- */

void oracle_write_packed_long(uint8_t *span, int64_t y)
{
    uint64_t u;
    uint8_t top = 0;
    if (y < 0) {
        u = (uint64_t)0 - (uint64_t)y;
        top = 0x80;
    } else {
        u = (uint64_t)y;
    }
    span[7] = (uint8_t)((u >> 56) | top);
    span[6] = (uint8_t)(u >> 48);
    span[5] = (uint8_t)(u >> 40);
    span[4] = (uint8_t)(u >> 32);
    span[3] = (uint8_t)(u >> 24);
    span[2] = (uint8_t)(u >> 16);
    span[1] = (uint8_t)(u >> 8);
    span[0] = (uint8_t)u;
}

/* ---- growable byte sink -------------------------------------------------- */

typedef struct {
    uint8_t *p;
    int64_t len, cap;
} sink_t;

static int sink_put(sink_t *s, const void *src, int64_t k)
{
    if (s->len + k > s->cap) {
        int64_t nc = s->cap ? s->cap * 2 : 4096;
        while (nc < s->len + k)
            nc *= 2;
        uint8_t *np = (uint8_t *)realloc(s->p, (size_t)nc);
        if (!np)
            return -1;
        s->p = np;
        s->cap = nc;
    }
    memcpy(s->p + s->len, src, (size_t)k);
    s->len += k;
    return 0;
}

typedef struct oracle_bsdiff_result {
    uint8_t *ctrl, *diff, *extra;
    int64_t ctrl_len, diff_len, extra_len;
    int64_t search_calls;
} oracle_bsdiff_result;

void oracle_bsdiff_free(oracle_bsdiff_result *r)
{
    if (!r)
        return;
    free(r->ctrl);
    free(r->diff);
    free(r->extra);
    free(r);
}

/*
 * Diff.cs:92-223 with Search inlined as the reference calls it (Diff.cs:106).
 * I: n+1 entries, I[n] == 0.  trace_pos/trace_len (optional, m entries each,
 * caller-initialised to -1) receive the Search result at every scan position
 * the loop visits, so a bulk (pos,len) table can be compared at exactly the
 * positions the reference evaluates.
 */
oracle_bsdiff_result *oracle_bsdiff_run(const uint8_t *old, int32_t n,
                                        const uint8_t *nw, int32_t m,
                                        const int32_t *I,
                                        int32_t *trace_pos, int32_t *trace_len)
{
    oracle_bsdiff_result *res = (oracle_bsdiff_result *)calloc(1, sizeof *res);
    if (!res)
        return NULL;
    sink_t ctrl = {0}, diff = {0}, extra = {0};
    uint8_t buf[8];
    int fail = 0;

    int32_t scan = 0, pos = 0, len = 0;
    int32_t lastscan = 0, lastpos = 0, lastoffset = 0;

    while (scan < m) {
        int32_t oldscore = 0;
        int32_t scsc;

        for (scsc = scan += len; scan < m; scan++) {
            len = oracle_search(I, old, n, nw + scan, m - scan, 0, n, &pos);
            res->search_calls++;
            if (trace_pos) {
                trace_pos[scan] = pos;
                trace_len[scan] = len;
            }

            for (; scsc < scan + len; scsc++)
                if (scsc + lastoffset < n && old[scsc + lastoffset] == nw[scsc])
                    oldscore++;

            if ((len == oldscore && len != 0) || len > oldscore + 8)
                break;

            if (scan + lastoffset < n && old[scan + lastoffset] == nw[scan])
                oldscore--;
        }

        if (len != oldscore || scan == m) {
            int32_t s = 0, sf = 0, lenf = 0, i;
            for (i = 0; lastscan + i < scan && lastpos + i < n;) {
                if (old[lastpos + i] == nw[lastscan + i])
                    s++;
                i++;
                if (s * 2 - i > sf * 2 - lenf) {
                    sf = s;
                    lenf = i;
                }
            }

            int32_t lenb = 0;
            if (scan < m) {
                int32_t sb = 0;
                s = 0;
                for (i = 1; scan >= lastscan + i && pos >= i; i++) {
                    if (old[pos - i] == nw[scan - i])
                        s++;
                    if (s * 2 - i > sb * 2 - lenb) {
                        sb = s;
                        lenb = i;
                    }
                }
            }

            if (lastscan + lenf > scan - lenb) {
                int32_t overlap = (lastscan + lenf) - (scan - lenb);
                int32_t ss = 0, lens = 0;
                s = 0;
                for (i = 0; i < overlap; i++) {
                    if (nw[lastscan + lenf - overlap + i] == old[lastpos + lenf - overlap + i])
                        s++;
                    if (nw[scan - lenb + i] == old[pos - lenb + i])
                        s--;
                    if (s > ss) {
                        ss = s;
                        lens = i + 1;
                    }
                }
                lenf += lens - overlap;
                lenb -= lens;
            }

            for (i = 0; i < lenf; i++) {
                uint8_t d = (uint8_t)(nw[lastscan + i] - old[lastpos + i]);
                fail |= sink_put(&diff, &d, 1);
            }

            int32_t extra_len = (scan - lenb) - (lastscan + lenf);
            if (extra_len > 0)
                fail |= sink_put(&extra, nw + lastscan + lenf, extra_len);

            oracle_write_packed_long(buf, lenf);
            fail |= sink_put(&ctrl, buf, 8);
            oracle_write_packed_long(buf, extra_len);
            fail |= sink_put(&ctrl, buf, 8);
            oracle_write_packed_long(buf, (int64_t)(pos - lenb) - (int64_t)(lastpos + lenf));
            fail |= sink_put(&ctrl, buf, 8);

            lastscan = scan - lenb;
            lastpos = pos - lenb;
            lastoffset = pos - scan;
        }
        if (fail)
            break;
    }

    res->ctrl = ctrl.p;
    res->ctrl_len = ctrl.len;
    res->diff = diff.p;
    res->diff_len = diff.len;
    res->extra = extra.p;
    res->extra_len = extra.len;
    if (fail) {
        oracle_bsdiff_free(res);
        return NULL;
    }
    return res;
}
