/*
 * oracle/sufcheck.c -- CPU restatement of the reference's suffix-array checkers.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/README.md).
 *
 * Follows (paths relative to /root/reference):
 *   test/DeltaQ.SuffixSorting.LibDivSufSort.Tests/LDSSChecker.cs:23-119
 *       `Check` (libdivsufsort's sufcheck): range, first-character order and the
 *       inverse-induction position check.  Return codes 0/-1/-2/-3/-4.
 *       (test/DeltaQ.SuffixSorting.SAIS.Tests/SAISChecker.cs:7-100 is the same
 *       check with int return codes.)
 *   test/DeltaQ.SuffixSorting.LibDivSufSort.Tests/LibDivSufSortTests.cs:43-59
 *       `Verify`: every adjacent pair strictly increasing under
 *       Span.SequenceCompareTo (first differing byte, else the shorter first).
 *   src/DeltaQ.CommandLine/Fuzzing/SuffixSortingVerifier.cs:7-22 -- same property.
 *
 * These two checks ARE the reference's parity contract for the suffix array:
 * the reference ships no golden SA, and the SA of a byte string is unique.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ALPHABET_SIZE 256

/* LDSSChecker.cs:23-119 */
int oracle_sufcheck(const uint8_t *T, int32_t n, const int32_t *SA, int32_t sa_len)
{
    int32_t C[ALPHABET_SIZE];
    int32_t i, p, q, t;
    int c;

    if (n != sa_len)
        return -1; /* BadArguments */
    if (n == 0)
        return 0;

    for (i = 0; i < n; ++i)
        if (SA[i] < 0 || n <= SA[i])
            return -2; /* OutOfRange */

    for (i = 1; i < n; ++i)
        if (T[SA[i - 1]] > T[SA[i]])
            return -3; /* WrongOrder */

    memset(C, 0, sizeof C);
    for (i = 0; i < n; ++i)
        ++C[T[i]];
    for (i = 0, p = 0; i < ALPHABET_SIZE; ++i) {
        t = C[i];
        C[i] = p;
        p += t;
    }

    q = C[T[n - 1]];
    C[T[n - 1]] += 1;
    for (i = 0; i < n; ++i) {
        p = SA[i];
        if (0 < p) {
            c = T[--p];
            t = C[c];
        } else {
            c = T[p = n - 1];
            t = q;
        }
        if (t < 0 || p != SA[t])
            return -4; /* WrongPosition */
        if (t != q) {
            ++C[c];
            if (n <= C[c] || T[SA[C[c]]] != c)
                C[c] = -1;
        }
    }
    return 0;
}

/* Span.SequenceCompareTo on two suffixes of T: <0, 0, >0 */
static int suffix_compare(const uint8_t *T, int32_t n, int32_t a, int32_t b)
{
    int32_t la = n - a, lb = n - b;
    int32_t l = la < lb ? la : lb;
    int r = memcmp(T + a, T + b, (size_t)l);
    if (r != 0)
        return r;
    return (la > lb) - (la < lb);
}

/*
 * LibDivSufSortTests.cs:43-59.  Returns -1 when sorted, else the index i of the
 * first adjacent pair (i, i+1) that is not strictly increasing.
 * Entries must already be in range (run oracle_sufcheck first).
 */
int32_t oracle_verify_sorted(const uint8_t *T, int32_t n, const int32_t *SA)
{
    for (int32_t i = 0; i + 1 < n; ++i)
        if (!(suffix_compare(T, n, SA[i], SA[i + 1]) < 0))
            return i;
    return -1;
}

/*
 * The same comparison (LibDivSufSortTests.cs:46-59) at the k sampled indices idx[] only: number of pairs
 * (SA[i], SA[i+1]) that are not strictly increasing or not in range.  For inputs where the full O(n) checks above
 * take minutes (2 G suffixes), next to an exact comparison with an independently computed array.
 */
int64_t oracle_verify_pairs(const uint8_t *T, int32_t n, const int32_t *SA, const int64_t *idx, int64_t k)
{
    int64_t bad = 0;
    for (int64_t j = 0; j < k; ++j) {
        int64_t i = idx[j];
        if (i < 0 || i + 1 >= n)
            continue;
        int32_t a = SA[i], b = SA[i + 1];
        if (a < 0 || a >= n || b < 0 || b >= n || !(suffix_compare(T, n, a, b) < 0))
            ++bad;
    }
    return bad;
}

/* definition-level sorter for tiny inputs: qsort with the comparison above */
static const uint8_t *g_T;
static int32_t g_n;
static int cmp_suffix(const void *x, const void *y)
{
    return suffix_compare(g_T, g_n, *(const int32_t *)x, *(const int32_t *)y);
}
void oracle_sa_naive(const uint8_t *T, int32_t n, int32_t *SA)
{
    for (int32_t i = 0; i < n; ++i)
        SA[i] = i;
    g_T = T;
    g_n = n;
    qsort(SA, (size_t)n, sizeof(int32_t), cmp_suffix);
}

/* LCP array by definition, for the on-device LCP export (the reference has none -- this is the
 * textbook definition, LCP[r] = longest common prefix of suffixes SA[r-1] and SA[r], LCP[0] = 0), computed with Kasai's
 * order of evaluation (along the text, each comparison resuming one below the previous length) so that large inputs
 * finish quickly.  rank is scratch of n entries. */
void oracle_lcp_kasai(const uint8_t *T, int32_t n, const int32_t *SA, int32_t *rank, int32_t *LCP)
{
    for (int32_t r = 0; r < n; ++r)
        rank[SA[r]] = r;
    int32_t h = 0;
    for (int32_t i = 0; i < n; ++i) {
        int32_t r = rank[i];
        if (r == 0) {
            LCP[0] = 0;
            h = 0;
            continue;
        }
        int32_t j = SA[r - 1];
        while (i + h < n && j + h < n && T[i + h] == T[j + h])
            ++h;
        LCP[r] = h;
        if (h > 0)
            --h;
    }
}
