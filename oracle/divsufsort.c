/*
 * oracle/divsufsort.c -- CPU restatement of the reference's DEFAULT suffix sorter, LibDivSufSort.
 * TEST INFRASTRUCTURE ONLY (see oracle/README.md): used by tests/ as a second sorter to pin the oracle, and by
 * bench.py's CPU arms as the timed baseline.  The product (deltaq_b200/) never links or loads it.
 *
 * Follows, function by function, /root/reference/src/DeltaQ.SuffixSorting.LibDivSufSort/:
 *   DivSufSort.cs   divsufsort :18-42, construct_SA :44-153, sort_typeBstar :186-511, bucket accessors :162-184
 *   SsSort.cs       sssort :23-149, ss_compare :155-194, ss_inplacemerge :196-280, ss_rotate :282-365,
 *                   ss_blockswap :367-374, ss_swapmerge :376-561, ss_mergebackward :563-771, ss_mergeforward :773-897,
 *                   ss_mintrosort :934-1273, ss_pivot/median :1275-1372, ss_partition :1374-1432,
 *                   ss_insertionsort :1434-1490, ss_heapsort/fixdown :1492-1567, ss_isqrt :1573-1642
 *   TrSort.cs       trsort :19-102, tr_introsort :148-769, tr_pivot/median :771-863, tr_heapsort/fixdown :865-952,
 *                   tr_insertionsort :954-1006, tr_partialcopy :1008-1090, tr_copy :1092-1146, tr_partition :1148-1322
 *   Budget.cs :2-34, TdPAStarAccessor.cs :7-22, Utils.cs (ss_ilg / tr_ilg :56-98)
 * in the configuration a net6+/net8 build executes: BitOperations.Log2 for ss_ilg/tr_ilg (Utils.cs:29-31), and the
 * non-lookup ss_isqrt, (int)MathF.Sqrt(x) capped at SS_BLOCKSIZE (SsSort.cs:1634-1642).  As in the C#, SA positions
 * ("SAPtr") are plain indices into the one array SA; PA, ISA, ISAd and buf are offsets into it.  The crosscheck /
 * SA_dump tracing of the C# is compile-time-off there ([Conditional("CROSSCHECK")]) and is not restated.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef int32_t idx_t;

#define ALPHABET_SIZE 256
#define BUCKET_A_SIZE ALPHABET_SIZE
#define BUCKET_B_SIZE (ALPHABET_SIZE * ALPHABET_SIZE)
#define SS_BLOCKSIZE 1024
#define SS_INSERTIONSORT_THRESHOLD 8
#define SS_STACK_SIZE 16
#define MERGE_STACK_SIZE 32
#define TR_INSERTIONSORT_THRESHOLD 8
#define TR_STACK_SIZE 64

#define SWAP_AT(A, i, j)        \
    do {                        \
        idx_t t_ = (A)[i];      \
        (A)[i] = (A)[j];        \
        (A)[j] = t_;            \
    } while (0)
#define SWAP_VAR(x, y)  \
    do {                \
        idx_t t_ = (x); \
        (x) = (y);      \
        (y) = t_;       \
    } while (0)

/* Utils.FastLog2 = BitOperations.Log2((uint)n): floor(log2), 0 for 0 */
static inline idx_t ilg(idx_t n)
{
    uint32_t v = (uint32_t)n;
    return v ? 31 - __builtin_clz(v) : 0;
}
#define ss_ilg ilg
#define tr_ilg ilg

/* SsSort.cs:1634-1642 */
static inline idx_t ss_isqrt(idx_t x)
{
    if (x >= SS_BLOCKSIZE * SS_BLOCKSIZE)
        return SS_BLOCKSIZE;
    return (idx_t)sqrtf((float)x);
}

/* ================================================ SsSort.cs ================================================ */

/* TdPAStarAccessor.cs: T[PA[SA[index]] + depth] */
#define TDPA(i) ((idx_t)T[SA[PA + SA[(i)]] + depth])

/* SsSort.cs:155-194; SAp1/SAp2 are arrays, p1/p2 indices into them */
static inline int ss_compare(const uint8_t *T, const idx_t *SAp1, idx_t p1, const idx_t *SAp2, idx_t p2, idx_t depth)
{
    idx_t U1 = depth + SAp1[p1];
    idx_t U2 = depth + SAp2[p2];
    idx_t U1n = SAp1[p1 + 1] + 2;
    idx_t U2n = SAp2[p2 + 1] + 2;
    while (U1 < U1n && U2 < U2n && T[U1] == T[U2]) {
        U1 += 1;
        U2 += 1;
    }
    if (U1 < U1n)
        return U2 < U2n ? (int)T[U1] - (int)T[U2] : 1;
    return U2 < U2n ? -1 : 0;
}

/* SsSort.cs:367-374 */
static inline void ss_blockswap(idx_t *SA, idx_t a, idx_t b, idx_t n)
{
    for (idx_t i = 0; i < n; i++)
        SWAP_AT(SA, a + i, b + i);
}

/* SsSort.cs:282-365 */
static void ss_rotate(idx_t *SA, idx_t first, idx_t middle, idx_t last)
{
    idx_t a, b, t, l, r;
    l = middle - first;
    r = last - middle;
    while (0 < l && 0 < r) {
        if (l == r) {
            ss_blockswap(SA, first, middle, l);
            break;
        }
        if (l < r) {
            a = last - 1;
            b = middle - 1;
            t = SA[a];
            for (;;) {
                SA[a] = SA[b];
                a -= 1;
                SA[b] = SA[a];
                b -= 1;
                if (b < first) {
                    SA[a] = t;
                    last = a;
                    r -= l + 1;
                    if (r <= l)
                        break;
                    a -= 1;
                    b = middle - 1;
                    t = SA[a];
                }
            }
        } else {
            a = first;
            b = middle;
            t = SA[a];
            for (;;) {
                SA[a] = SA[b];
                a += 1;
                SA[b] = SA[a];
                b += 1;
                if (last <= b) {
                    SA[a] = t;
                    first = a + 1;
                    l -= r + 1;
                    if (l <= r)
                        break;
                    a += 1;
                    b = middle;
                    t = SA[a];
                }
            }
        }
    }
}

/* SsSort.cs:196-280 */
static void ss_inplacemerge(const uint8_t *T, idx_t *SA, idx_t PA, idx_t first, idx_t middle, idx_t last, idx_t depth)
{
    idx_t p, a, b, len, half, q, r, x;
    for (;;) {
        if (SA[last - 1] < 0) {
            x = 1;
            p = PA + ~SA[last - 1];
        } else {
            x = 0;
            p = PA + SA[last - 1];
        }
        a = first;
        len = middle - first;
        half = len >> 1;
        r = -1;
        while (0 < len) {
            b = a + half;
            q = ss_compare(T, SA, PA + (0 <= SA[b] ? SA[b] : ~SA[b]), SA, p, depth);
            if (q < 0) {
                a = b + 1;
                half -= (len & 1) ^ 1;
            } else {
                r = q;
            }
            len = half;
            half >>= 1;
        }
        if (a < middle) {
            if (r == 0)
                SA[a] = ~SA[a];
            ss_rotate(SA, a, middle, last);
            last -= middle - a;
            middle = a;
            if (first == middle)
                break;
        }
        last -= 1;
        if (x != 0) {
            last -= 1;
            while (SA[last] < 0)
                last -= 1;
        }
        if (middle == last)
            break;
    }
}

/* SsSort.cs:773-897 */
static void ss_mergeforward(const uint8_t *T, idx_t *SA, idx_t PA, idx_t first, idx_t middle, idx_t last, idx_t buf,
                            idx_t depth)
{
    idx_t a, b, c, bufend, t, r;
    bufend = buf + (middle - first) - 1;
    ss_blockswap(SA, buf, first, middle - first);
    a = first;
    t = SA[a];
    b = buf;
    c = middle;
    for (;;) {
        r = ss_compare(T, SA, PA + SA[b], SA, PA + SA[c], depth);
        if (r < 0) {
            for (;;) {
                SA[a] = SA[b];
                a += 1;
                if (bufend <= b) {
                    SA[bufend] = t;
                    return;
                }
                SA[b] = SA[a];
                b += 1;
                if (!(SA[b] < 0))
                    break;
            }
        } else if (r > 0) {
            for (;;) {
                SA[a] = SA[c];
                a += 1;
                SA[c] = SA[a];
                c += 1;
                if (last <= c) {
                    while (b < bufend) {
                        SA[a] = SA[b];
                        a += 1;
                        SA[b] = SA[a];
                        b += 1;
                    }
                    SA[a] = SA[b];
                    SA[b] = t;
                    return;
                }
                if (!(SA[c] < 0))
                    break;
            }
        } else {
            SA[c] = ~SA[c];
            for (;;) {
                SA[a] = SA[b];
                a += 1;
                if (bufend <= b) {
                    SA[bufend] = t;
                    return;
                }
                SA[b] = SA[a];
                b += 1;
                if (!(SA[b] < 0))
                    break;
            }
            for (;;) {
                SA[a] = SA[c];
                a += 1;
                SA[c] = SA[a];
                c += 1;
                if (last <= c) {
                    while (b < bufend) {
                        SA[a] = SA[b];
                        a += 1;
                        SA[b] = SA[a];
                        b += 1;
                    }
                    SA[a] = SA[b];
                    SA[b] = t;
                    return;
                }
                if (!(SA[c] < 0))
                    break;
            }
        }
    }
}

/* SsSort.cs:563-771 */
static void ss_mergebackward(const uint8_t *T, idx_t *SA, idx_t PA, idx_t first, idx_t middle, idx_t last, idx_t buf,
                             idx_t depth)
{
    idx_t p1, p2, a, b, c, bufend, t, r, x;
    bufend = buf + (last - middle) - 1;
    ss_blockswap(SA, buf, middle, last - middle);
    x = 0;
    if (SA[bufend] < 0) {
        p1 = PA + ~SA[bufend];
        x |= 1;
    } else {
        p1 = PA + SA[bufend];
    }
    if (SA[middle - 1] < 0) {
        p2 = PA + ~SA[middle - 1];
        x |= 2;
    } else {
        p2 = PA + SA[middle - 1];
    }
    a = last - 1;
    t = SA[a];
    b = bufend;
    c = middle - 1;
    for (;;) {
        r = ss_compare(T, SA, p1, SA, p2, depth);
        if (0 < r) {
            if ((x & 1) > 0) {
                for (;;) {
                    SA[a] = SA[b];
                    a -= 1;
                    SA[b] = SA[a];
                    b -= 1;
                    if (!(SA[b] < 0))
                        break;
                }
                x ^= 1;
            }
            SA[a] = SA[b];
            a -= 1;
            if (b <= buf) {
                SA[buf] = t;
                break;
            }
            SA[b] = SA[a];
            b -= 1;
            if (SA[b] < 0) {
                p1 = PA + ~SA[b];
                x |= 1;
            } else {
                p1 = PA + SA[b];
            }
        } else if (r < 0) {
            if ((x & 2) > 0) {
                for (;;) {
                    SA[a] = SA[c];
                    a -= 1;
                    SA[c] = SA[a];
                    c -= 1;
                    if (~SA[c] < 0)
                        break;
                }
                x ^= 2;
            }
            SA[a] = SA[c];
            a -= 1;
            SA[c] = SA[a];
            c -= 1;
            if (c < first) {
                while (buf < b) {
                    SA[a] = SA[b];
                    a -= 1;
                    SA[b] = SA[a];
                    b -= 1;
                }
                SA[a] = SA[b];
                SA[b] = t;
                break;
            }
            if (SA[c] < 0) {
                p2 = PA + ~SA[c];
                x |= 2;
            } else {
                p2 = PA + SA[c];
            }
        } else {
            if ((x & 1) > 0) {
                for (;;) {
                    SA[a] = SA[b];
                    a -= 1;
                    SA[b] = SA[a];
                    b -= 1;
                    if (!(SA[b] < 0))
                        break;
                }
                x ^= 1;
            }
            SA[a] = ~SA[b];
            a -= 1;
            if (b <= buf) {
                SA[buf] = t;
                break;
            }
            SA[b] = SA[a];
            b -= 1;
            if ((x & 2) > 0) {
                for (;;) {
                    SA[a] = SA[c];
                    a -= 1;
                    SA[c] = SA[a];
                    c -= 1;
                    if (!(SA[c] < 0))
                        break;
                }
                x ^= 2;
            }
            SA[a] = SA[c];
            a -= 1;
            SA[c] = SA[a];
            c -= 1;
            if (c < first) {
                while (buf < b) {
                    SA[a] = SA[b];
                    a -= 1;
                    SA[b] = SA[a];
                    b -= 1;
                }
                SA[a] = SA[b];
                SA[b] = t;
                break;
            }
            if (SA[b] < 0) {
                p1 = PA + ~SA[b];
                x |= 1;
            } else {
                p1 = PA + SA[b];
            }
            if (SA[c] < 0) {
                p2 = PA + ~SA[c];
                x |= 2;
            } else {
                p2 = PA + SA[c];
            }
        }
    }
}

/* SsSort.cs:899-931 (SsStack; also used as the merge stack) */
typedef struct {
    idx_t a, b, c, d;
} ss_item;
typedef struct {
    ss_item *items;
    int size;
} ss_stack;
static inline void ss_push(ss_stack *s, idx_t a, idx_t b, idx_t c, idx_t d)
{
    ss_item *it = &s->items[s->size++];
    it->a = a;
    it->b = b;
    it->c = c;
    it->d = d;
}
static inline int ss_pop(ss_stack *s, idx_t *a, idx_t *b, idx_t *c, idx_t *d)
{
    if (s->size == 0)
        return 0;
    ss_item *it = &s->items[--s->size];
    *a = it->a;
    *b = it->b;
    *c = it->c;
    *d = it->d;
    return 1;
}

static inline idx_t get_idx(idx_t a) { return 0 <= a ? a : ~a; }

/* SsSort.cs:380-393 (local function merge_check) */
static inline void merge_check(const uint8_t *T, idx_t *SA, idx_t PA, idx_t depth, idx_t a, idx_t b, idx_t c)
{
    if (((c & 1) > 0) ||
        (((c & 2) > 0) && (ss_compare(T, SA, PA + get_idx(SA[a - 1]), SA, PA + SA[a], depth) == 0)))
        SA[a] = ~SA[a];
    if (((c & 4) > 0) && (ss_compare(T, SA, PA + get_idx(SA[b - 1]), SA, PA + SA[b], depth) == 0))
        SA[b] = ~SA[b];
}

/* SsSort.cs:376-561 */
static void ss_swapmerge(const uint8_t *T, idx_t *SA, idx_t PA, idx_t first, idx_t middle, idx_t last, idx_t buf,
                         idx_t bufsize, idx_t depth)
{
    ss_item items[MERGE_STACK_SIZE];
    ss_stack stack = {items, 0};
    idx_t l, r, lm, rm, m, len, half, check, next;
    memset(items, 0, sizeof items);
    check = 0;
    for (;;) {
        if ((last - middle) <= bufsize) {
            if (first < middle && middle < last)
                ss_mergebackward(T, SA, PA, first, middle, last, buf, depth);
            merge_check(T, SA, PA, depth, first, last, check);
            if (!ss_pop(&stack, &first, &middle, &last, &check))
                return;
            continue;
        }
        if ((middle - first) <= bufsize) {
            if (first < middle)
                ss_mergeforward(T, SA, PA, first, middle, last, buf, depth);
            merge_check(T, SA, PA, depth, first, last, check);
            if (!ss_pop(&stack, &first, &middle, &last, &check))
                return;
            continue;
        }
        m = 0;
        len = (middle - first) < (last - middle) ? (middle - first) : (last - middle);
        half = len >> 1;
        while (0 < len) {
            if (ss_compare(T, SA, PA + get_idx(SA[middle + m + half]), SA, PA + get_idx(SA[middle - m - half - 1]),
                           depth) < 0) {
                m += half + 1;
                half -= (len & 1) ^ 1;
            }
            len = half;
            half >>= 1;
        }
        if (0 < m) {
            lm = middle - m;
            rm = middle + m;
            ss_blockswap(SA, lm, middle, m);
            r = middle;
            l = middle;
            next = 0;
            if (rm < last) {
                if (SA[rm] < 0) {
                    SA[rm] = ~SA[rm];
                    if (first < lm) {
                        l -= 1;
                        while (SA[l] < 0)
                            l -= 1;
                        next |= 4;
                    }
                    next |= 1;
                } else if (first < lm) {
                    while (SA[r] < 0)
                        r += 1;
                    next |= 2;
                }
            }
            if ((l - first) <= (last - r)) {
                ss_push(&stack, r, rm, last, (next & 3) | (check & 4));
                middle = lm;
                last = l;
                check = (check & 3) | (next & 4);
            } else {
                if (((next & 2) > 0) && (r == middle))
                    next ^= 6;
                ss_push(&stack, first, lm, l, (check & 3) | (next & 4));
                first = r;
                middle = rm;
                check = (next & 3) | (check & 4);
            }
        } else {
            if (ss_compare(T, SA, PA + get_idx(SA[middle - 1]), SA, PA + SA[middle], depth) == 0)
                SA[middle] = ~SA[middle];
            merge_check(T, SA, PA, depth, first, last, check);
            if (!ss_pop(&stack, &first, &middle, &last, &check))
                return;
        }
    }
}

/* SsSort.cs:1327-1372; get[v] = T[PA[SA[v]] + Td] */
#define GETV(v) ((idx_t)T[SA[PA + SA[(v)]] + Td])
static inline idx_t ss_median3(const uint8_t *T, idx_t Td, const idx_t *SA, idx_t PA, idx_t v1, idx_t v2, idx_t v3)
{
    if (GETV(v1) > GETV(v2))
        SWAP_VAR(v1, v2);
    if (GETV(v2) > GETV(v3)) {
        if (GETV(v1) > GETV(v3))
            return v1;
        return v3;
    }
    return v2;
}
static inline idx_t ss_median5(const uint8_t *T, idx_t Td, const idx_t *SA, idx_t PA, idx_t v1, idx_t v2, idx_t v3,
                               idx_t v4, idx_t v5)
{
    if (GETV(v2) > GETV(v3))
        SWAP_VAR(v2, v3);
    if (GETV(v4) > GETV(v5))
        SWAP_VAR(v4, v5);
    if (GETV(v2) > GETV(v4)) {
        SWAP_VAR(v2, v4);
        SWAP_VAR(v3, v5);
    }
    if (GETV(v1) > GETV(v3))
        SWAP_VAR(v1, v3);
    if (GETV(v1) > GETV(v4)) {
        SWAP_VAR(v1, v4);
        SWAP_VAR(v3, v5);
    }
    if (GETV(v3) > GETV(v4))
        return v4;
    return v3;
}
/* SsSort.cs:1275-1306 */
static inline idx_t ss_pivot(const uint8_t *T, idx_t Td, const idx_t *SA, idx_t PA, idx_t first, idx_t last)
{
    idx_t t = last - first;
    idx_t middle = first + (t / 2);
    if (t <= 512) {
        if (t <= 32)
            return ss_median3(T, Td, SA, PA, first, middle, last - 1);
        t >>= 2;
        return ss_median5(T, Td, SA, PA, first, first + t, middle, last - 1 - t, last - 1);
    }
    t >>= 3;
    first = ss_median3(T, Td, SA, PA, first, first + t, first + (t << 1));
    middle = ss_median3(T, Td, SA, PA, middle - t, middle, middle + t);
    last = ss_median3(T, Td, SA, PA, last - 1 - (t << 1), last - 1 - t, last - 1);
    return ss_median3(T, Td, SA, PA, first, middle, last);
}
#undef GETV

/* SsSort.cs:1374-1432 */
static inline idx_t ss_partition(idx_t *SA, idx_t PA, idx_t first, idx_t last, idx_t depth)
{
    idx_t a = first - 1;
    idx_t b = last;
    for (;;) {
        for (;;) {
            a += 1;
            if (!(a < b))
                break;
            if (!((SA[PA + SA[a]] + depth) >= (SA[PA + SA[a] + 1] + 1)))
                break;
            SA[a] = ~SA[a];
        }
        for (;;) {
            b -= 1;
            if (!(a < b))
                break;
            if (!((SA[PA + SA[b]] + depth) < (SA[PA + SA[b] + 1] + 1)))
                break;
        }
        if (b <= a)
            break;
        idx_t t = ~SA[b];
        SA[b] = SA[a];
        SA[a] = t;
    }
    if (first < a)
        SA[first] = ~SA[first];
    return a;
}

/* SsSort.cs:1434-1490 */
static void ss_insertionsort(const uint8_t *T, idx_t *SA, idx_t PA, idx_t first, idx_t last, idx_t depth)
{
    idx_t i, j, t, r;
    i = last - 2;
    while (first <= i) {
        t = SA[i];
        j = i + 1;
        for (;;) {
            r = ss_compare(T, SA, PA + t, SA, PA + SA[j], depth);
            if (!(0 < r))
                break;
            for (;;) {
                SA[j - 1] = SA[j];
                j += 1;
                if (!((j < last) && SA[j] < 0))
                    break;
            }
            if (last <= j)
                break;
        }
        if (r == 0)
            SA[j] = ~SA[j];
        SA[j - 1] = t;
        i -= 1;
    }
}

/* SsSort.cs:1530-1567; T already offset by depth, PA and SA are array views */
static inline void ss_fixdown(const uint8_t *T, const idx_t *PA, idx_t *SA, idx_t i, idx_t size)
{
    idx_t j, v, c, d, e, k;
    v = SA[i];
    c = T[PA[v]];
    for (;;) {
        j = 2 * i + 1;
        if (!(j < size))
            break;
        k = j;
        j += 1;
        d = T[PA[SA[k]]];
        e = T[PA[SA[j]]];
        if (d < e) {
            k = j;
            d = e;
        }
        if (d <= c)
            break;
        SA[i] = SA[k];
        i = k;
    }
    SA[i] = v;
}
/* SsSort.cs:1492-1528 */
static void ss_heapsort(const uint8_t *T, const idx_t *PA, idx_t *SA, idx_t size)
{
    idx_t i, m = size, t;
    if ((size % 2) == 0) {
        m -= 1;
        if (T[PA[SA[m / 2]]] < T[PA[SA[m]]])
            SWAP_AT(SA, m, m / 2);
    }
    for (i = (m / 2) - 1; i >= 0; i--)
        ss_fixdown(T, PA, SA, i, m);
    if ((size % 2) == 0) {
        SWAP_AT(SA, 0, m);
        ss_fixdown(T, PA, SA, 0, m);
    }
    for (i = m - 1; i > 0; i--) {
        t = SA[0];
        SA[0] = SA[i];
        ss_fixdown(T, PA, SA, 0, i);
        SA[i] = t;
    }
}

/* SsSort.cs:934-1273; TDOFF(x) = TdPAStar.AsOffset(x) = T[x + depth] */
#define TDOFF(x) ((idx_t)T[(x) + depth])
static void ss_mintrosort(const uint8_t *T, idx_t *SA, idx_t PA, idx_t first, idx_t last, idx_t depth)
{
    ss_item items[SS_STACK_SIZE];
    ss_stack stack = {items, 0};
    idx_t a, b, c, d, e, f, s, t, limit, v, x = 0;
    limit = ss_ilg(last - first);
    for (;;) {
        if ((last - first) <= SS_INSERTIONSORT_THRESHOLD) {
            if (1 < (last - first))
                ss_insertionsort(T, SA, PA, first, last, depth);
            if (!ss_pop(&stack, &first, &last, &depth, &limit))
                return;
            continue;
        }
        idx_t old_limit = limit;
        limit -= 1;
        if (old_limit == 0)
            ss_heapsort(T + depth, SA + PA, SA + first, last - first);
        if (limit < 0) {
            a = first + 1;
            v = TDPA(first);
            while (a < last) {
                x = TDPA(a);
                if (x != v) {
                    if (1 < (a - first))
                        break;
                    v = x;
                    first = a;
                }
                a += 1;
            }
            if (TDOFF(SA[PA + SA[first]] - 1) < v)
                first = ss_partition(SA, PA, first, a, depth);
            if ((a - first) <= (last - a)) {
                if (1 < (a - first)) {
                    ss_push(&stack, a, last, depth, -1);
                    last = a;
                    depth += 1;
                    limit = ss_ilg(a - first);
                } else {
                    first = a;
                    limit = -1;
                }
            } else {
                if (1 < (last - a)) {
                    ss_push(&stack, first, a, depth + 1, ss_ilg(a - first));
                    first = a;
                    limit = -1;
                } else {
                    last = a;
                    depth += 1;
                    limit = ss_ilg(a - first);
                }
            }
            continue;
        }
        /* choose pivot */
        a = ss_pivot(T, depth, SA, PA, first, last);
        v = TDPA(a);
        SWAP_AT(SA, first, a);
        /* partition */
        b = first;
        for (;;) {
            b += 1;
            if (!(b < last))
                break;
            x = TDPA(b);
            if (!(x == v))
                break;
        }
        a = b;
        if ((a < last) && (x < v)) {
            for (;;) {
                b += 1;
                if (!(b < last))
                    break;
                x = TDPA(b);
                if (!(x <= v))
                    break;
                if (x == v) {
                    SWAP_AT(SA, b, a);
                    a += 1;
                }
            }
        }
        c = last;
        for (;;) {
            c -= 1;
            if (!(b < c))
                break;
            x = TDPA(c);
            if (!(x == v))
                break;
        }
        d = c;
        if ((b < d) && (x > v)) {
            for (;;) {
                c -= 1;
                if (!(b < c))
                    break;
                x = TDPA(c);
                if (!(x >= v))
                    break;
                if (x == v) {
                    SWAP_AT(SA, c, d);
                    d -= 1;
                }
            }
        }
        while (b < c) {
            SWAP_AT(SA, b, c);
            for (;;) {
                b += 1;
                if (!(b < c))
                    break;
                x = TDPA(b);
                if (!(x <= v))
                    break;
                if (x == v) {
                    SWAP_AT(SA, b, a);
                    a += 1;
                }
            }
            for (;;) {
                c -= 1;
                if (!(b < c))
                    break;
                x = TDPA(c);
                if (!(x >= v))
                    break;
                if (x == v) {
                    SWAP_AT(SA, c, d);
                    d -= 1;
                }
            }
        }
        if (a <= d) {
            c = b - 1;
            s = a - first;
            t = b - a;
            if (s > t)
                s = t;
            e = first;
            f = b - s;
            while (0 < s) {
                SWAP_AT(SA, e, f);
                s -= 1;
                e += 1;
                f += 1;
            }
            s = d - c;
            t = last - d - 1;
            if (s > t)
                s = t;
            e = b;
            f = last - s;
            while (0 < s) {
                SWAP_AT(SA, e, f);
                s -= 1;
                e += 1;
                f += 1;
            }
            a = first + (b - a);
            c = last - (d - c);
            b = v <= TDOFF(SA[PA + SA[a]] - 1) ? a : ss_partition(SA, PA, a, c, depth);
            if ((a - first) <= (last - c)) {
                if ((last - c) <= (c - b)) {
                    ss_push(&stack, b, c, depth + 1, ss_ilg(c - b));
                    ss_push(&stack, c, last, depth, limit);
                    last = a;
                } else if ((a - first) <= (c - b)) {
                    ss_push(&stack, c, last, depth, limit);
                    ss_push(&stack, b, c, depth + 1, ss_ilg(c - b));
                    last = a;
                } else {
                    ss_push(&stack, c, last, depth, limit);
                    ss_push(&stack, first, a, depth, limit);
                    first = b;
                    last = c;
                    depth += 1;
                    limit = ss_ilg(c - b);
                }
            } else {
                if ((a - first) <= (c - b)) {
                    ss_push(&stack, b, c, depth + 1, ss_ilg(c - b));
                    ss_push(&stack, first, a, depth, limit);
                    first = c;
                } else if ((last - c) <= (c - b)) {
                    ss_push(&stack, first, a, depth, limit);
                    ss_push(&stack, b, c, depth + 1, ss_ilg(c - b));
                    first = c;
                } else {
                    ss_push(&stack, first, a, depth, limit);
                    ss_push(&stack, c, last, depth, limit);
                    first = b;
                    last = c;
                    depth += 1;
                    limit = ss_ilg(c - b);
                }
            }
        } else {
            limit += 1;
            if (TDOFF(SA[PA + SA[first]] - 1) < v) {
                first = ss_partition(SA, PA, first, last, depth);
                limit = ss_ilg(last - first);
            }
            depth += 1;
        }
    }
}
#undef TDOFF

/* SsSort.cs:23-149 */
static void sssort(const uint8_t *T, idx_t *SA, idx_t PA, idx_t first, idx_t last, idx_t buf, idx_t bufsize, idx_t depth,
                   idx_t n, int lastsuffix)
{
    idx_t a, b, middle, curbuf, j, k, curbufsize, limit, i;
    if (lastsuffix)
        first += 1;
    limit = ss_isqrt(last - first);
    if ((bufsize < SS_BLOCKSIZE) && (bufsize < (last - first)) && (bufsize < limit)) {
        if (SS_BLOCKSIZE < limit)
            limit = SS_BLOCKSIZE;
        middle = last - limit;
        buf = middle;
        bufsize = limit;
    } else {
        middle = last;
        limit = 0;
    }
    a = first;
    i = 0;
    while (SS_BLOCKSIZE < (middle - a)) {
        ss_mintrosort(T, SA, PA, a, a + SS_BLOCKSIZE, depth);
        curbufsize = last - (a + SS_BLOCKSIZE);
        curbuf = a + SS_BLOCKSIZE;
        if (curbufsize <= bufsize) {
            curbufsize = bufsize;
            curbuf = buf;
        }
        b = a;
        k = SS_BLOCKSIZE;
        j = i;
        while ((j & 1) > 0) {
            ss_swapmerge(T, SA, PA, b - k, b, b + k, curbuf, curbufsize, depth);
            b -= k;
            k <<= 1;
            j >>= 1;
        }
        a += SS_BLOCKSIZE;
        i += 1;
    }
    ss_mintrosort(T, SA, PA, a, middle, depth);
    k = SS_BLOCKSIZE;
    while (i != 0) {
        if ((i & 1) > 0) {
            ss_swapmerge(T, SA, PA, a - k, a, middle, buf, bufsize, depth);
            a -= k;
        }
        k <<= 1;
        i >>= 1;
    }
    if (limit != 0) {
        ss_mintrosort(T, SA, PA, middle, last, depth);
        ss_inplacemerge(T, SA, PA, first, middle, last, depth);
    }
    if (lastsuffix) {
        /* insert the last type B* suffix */
        idx_t PAi[2] = {SA[PA + SA[first - 1]], n - 2};
        a = first;
        i = SA[first - 1];
        while ((a < last) && ((SA[a] < 0) || (0 < ss_compare(T, PAi, 0, SA, PA + SA[a], depth)))) {
            SA[a - 1] = SA[a];
            a += 1;
        }
        SA[a - 1] = i;
    }
}

/* ================================================ TrSort.cs ================================================ */

/* Budget.cs:2-34 */
typedef struct {
    idx_t chance, remain, incval, count;
} budget_t;
static inline int budget_check(budget_t *b, idx_t size)
{
    if (size <= b->remain) {
        b->remain -= size;
        return 1;
    }
    if (b->chance == 0) {
        b->count += size;
        return 0;
    }
    b->remain += b->incval - size;
    b->chance -= 1;
    return 1;
}

/* TrSort.cs:104-146 */
typedef struct {
    idx_t a, b, c, d, e;
} tr_item;
typedef struct {
    tr_item *items;
    int size;
} tr_stack;
static inline void tr_push(tr_stack *s, idx_t a, idx_t b, idx_t c, idx_t d, idx_t e)
{
    tr_item *it = &s->items[s->size++];
    it->a = a;
    it->b = b;
    it->c = c;
    it->d = d;
    it->e = e;
}
static inline int tr_pop(tr_stack *s, idx_t *a, idx_t *b, idx_t *c, idx_t *d, idx_t *e)
{
    if (s->size == 0)
        return 0;
    tr_item *it = &s->items[--s->size];
    *a = it->a;
    *b = it->b;
    *c = it->c;
    *d = it->d;
    *e = it->e;
    return 1;
}

#define ISAD(i) SA[ISAd + (i)]

/* TrSort.cs:1148-1322 */
static inline void tr_partition(idx_t *SA, idx_t ISAd, idx_t first, idx_t middle, idx_t last, idx_t *pa, idx_t *pb,
                                idx_t v)
{
    idx_t a, b, c, d, e, f, t, s, x = 0;
    b = middle - 1;
    for (;;) {
        b += 1;
        if (!(b < last))
            break;
        x = ISAD(SA[b]);
        if (!(x == v))
            break;
    }
    a = b;
    if ((a < last) && (x < v)) {
        for (;;) {
            b += 1;
            if (!(b < last))
                break;
            x = ISAD(SA[b]);
            if (!(x <= v))
                break;
            if (x == v) {
                SWAP_AT(SA, b, a);
                a += 1;
            }
        }
    }
    c = last;
    for (;;) {
        c -= 1;
        if (!(b < c))
            break;
        x = ISAD(SA[c]);
        if (!(x == v))
            break;
    }
    d = c;
    if ((b < d) && (x > v)) {
        for (;;) {
            c -= 1;
            if (!(b < c))
                break;
            x = ISAD(SA[c]);
            if (!(x >= v))
                break;
            if (x == v) {
                SWAP_AT(SA, c, d);
                d -= 1;
            }
        }
    }
    while (b < c) {
        SWAP_AT(SA, b, c);
        for (;;) {
            b += 1;
            if (!(b < c))
                break;
            x = ISAD(SA[b]);
            if (!(x <= v))
                break;
            if (x == v) {
                SWAP_AT(SA, b, a);
                a += 1;
            }
        }
        for (;;) {
            c -= 1;
            if (!(b < c))
                break;
            x = ISAD(SA[c]);
            if (!(x >= v))
                break;
            if (x == v) {
                SWAP_AT(SA, c, d);
                d -= 1;
            }
        }
    }
    if (a <= d) {
        c = b - 1;
        s = a - first;
        t = b - a;
        if (s > t)
            s = t;
        e = first;
        f = b - s;
        while (0 < s) {
            SWAP_AT(SA, e, f);
            s -= 1;
            e += 1;
            f += 1;
        }
        s = d - c;
        t = last - d - 1;
        if (s > t)
            s = t;
        e = b;
        f = last - s;
        while (0 < s) {
            SWAP_AT(SA, e, f);
            s -= 1;
            e += 1;
            f += 1;
        }
        first += (b - a);
        last -= (d - c);
    }
    *pa = first;
    *pb = last;
}

/* TrSort.cs:1092-1146 */
static void tr_copy(idx_t ISA, idx_t *SA, idx_t first, idx_t a, idx_t b, idx_t last, idx_t depth)
{
    idx_t c, d, e, s, v;
    v = b - 1;
    c = first;
    d = a - 1;
    while (c <= d) {
        s = SA[c] - depth;
        if ((0 <= s) && (SA[ISA + s] == v)) {
            d += 1;
            SA[d] = s;
            SA[ISA + s] = d;
        }
        c += 1;
    }
    c = last - 1;
    e = d + 1;
    d = b;
    while (e < d) {
        s = SA[c] - depth;
        if ((0 <= s) && (SA[ISA + s] == v)) {
            d -= 1;
            SA[d] = s;
            SA[ISA + s] = d;
        }
        c -= 1;
    }
}

/* TrSort.cs:1008-1090 */
static void tr_partialcopy(idx_t ISA, idx_t *SA, idx_t first, idx_t a, idx_t b, idx_t last, idx_t depth)
{
    idx_t c, d, e, s, v, rank, lastrank, newrank = -1;
    v = b - 1;
    lastrank = -1;
    c = first;
    d = a - 1;
    while (c <= d) {
        s = SA[c] - depth;
        if ((0 <= s) && (SA[ISA + s] == v)) {
            d += 1;
            SA[d] = s;
            rank = SA[ISA + s + depth];
            if (lastrank != rank) {
                lastrank = rank;
                newrank = d;
            }
            SA[ISA + s] = newrank;
        }
        c += 1;
    }
    lastrank = -1;
    e = d;
    while (first <= e) {
        rank = SA[ISA + SA[e]];
        if (lastrank != rank) {
            lastrank = rank;
            newrank = e;
        }
        if (newrank != rank)
            SA[ISA + SA[e]] = newrank;
        e -= 1;
    }
    lastrank = -1;
    c = last - 1;
    e = d + 1;
    d = b;
    while (e < d) {
        s = SA[c] - depth;
        if ((0 <= s) && (SA[ISA + s] == v)) {
            d -= 1;
            SA[d] = s;
            rank = SA[ISA + s + depth];
            if (lastrank != rank) {
                lastrank = rank;
                newrank = d;
            }
            SA[ISA + s] = newrank;
        }
        c -= 1;
    }
}

/* TrSort.cs:954-1006 */
static void tr_insertionsort(idx_t *SA, idx_t ISAd, idx_t first, idx_t last)
{
    idx_t a, b, t, r;
    a = first + 1;
    while (a < last) {
        t = SA[a];
        b = a - 1;
        for (;;) {
            r = ISAD(t) - ISAD(SA[b]);
            if (!(0 > r))
                break;
            for (;;) {
                SA[b + 1] = SA[b];
                b -= 1;
                if (!((first <= b) && (SA[b] < 0)))
                    break;
            }
            if (b < first)
                break;
        }
        if (r == 0)
            SA[b] = ~SA[b];
        SA[b + 1] = t;
        a += 1;
    }
}

/* TrSort.cs:914-952; ISAd and SA are array views */
static inline void tr_fixdown(const idx_t *ISAd_, idx_t *SA, idx_t i, idx_t size)
{
    idx_t j, k, d, e;
    idx_t v = SA[i];
    idx_t c = ISAd_[v];
    for (;;) {
        j = 2 * i + 1;
        if (!(j < size))
            break;
        k = j;
        d = ISAd_[SA[k]];
        j += 1;
        e = ISAd_[SA[j]];
        if (d < e) {
            k = j;
            d = e;
        }
        if (d <= c)
            break;
        SA[i] = SA[k];
        i = k;
    }
    SA[i] = v;
}
/* TrSort.cs:865-912 */
static void tr_heapsort(idx_t ISAd, idx_t *SA_top, idx_t first, idx_t size)
{
    idx_t i, m, t;
    const idx_t *ISAd_ = SA_top + ISAd;
    idx_t *SA = SA_top + first;
    m = size;
    if ((size % 2) == 0) {
        m -= 1;
        if (ISAd_[SA[m / 2]] < ISAd_[SA[m]])
            SWAP_AT(SA_top, first + m, first + (m / 2));
    }
    for (i = (m / 2) - 1; i >= 0; i--)
        tr_fixdown(ISAd_, SA, i, m);
    if ((size % 2) == 0) {
        SWAP_AT(SA_top, first + 0, first + m);
        tr_fixdown(ISAd_, SA, 0, m);
    }
    for (i = m - 1; i > 0; i--) {
        t = SA[0];
        SA[0] = SA[i];
        tr_fixdown(ISAd_, SA, 0, i);
        SA[i] = t;
    }
}

/* TrSort.cs:797-863 */
static inline idx_t tr_median3(const idx_t *SA, idx_t ISAd, idx_t v1, idx_t v2, idx_t v3)
{
    if (ISAD(SA[v1]) > ISAD(SA[v2]))
        SWAP_VAR(v1, v2);
    if (ISAD(SA[v2]) > ISAD(SA[v3])) {
        if (ISAD(SA[v1]) > ISAD(SA[v3]))
            return v1;
        return v3;
    }
    return v2;
}
static inline idx_t tr_median5(const idx_t *SA, idx_t ISAd, idx_t v1, idx_t v2, idx_t v3, idx_t v4, idx_t v5)
{
    if (ISAD(SA[v2]) > ISAD(SA[v3]))
        SWAP_VAR(v2, v3);
    if (ISAD(SA[v4]) > ISAD(SA[v5]))
        SWAP_VAR(v4, v5);
    if (ISAD(SA[v2]) > ISAD(SA[v4])) {
        SWAP_VAR(v2, v4);
        SWAP_VAR(v3, v5);
    }
    if (ISAD(SA[v1]) > ISAD(SA[v3]))
        SWAP_VAR(v1, v3);
    if (ISAD(SA[v1]) > ISAD(SA[v4])) {
        SWAP_VAR(v1, v4);
        SWAP_VAR(v3, v5);
    }
    if (ISAD(SA[v3]) > ISAD(SA[v4]))
        return v4;
    return v3;
}
/* TrSort.cs:771-795 */
static inline idx_t tr_pivot(const idx_t *SA, idx_t ISAd, idx_t first, idx_t last)
{
    idx_t t = last - first;
    idx_t middle = first + t / 2;
    if (t <= 512) {
        if (t <= 32)
            return tr_median3(SA, ISAd, first, middle, last - 1);
        t >>= 2;
        return tr_median5(SA, ISAd, first, first + t, middle, last - 1 - t, last - 1);
    }
    t >>= 3;
    first = tr_median3(SA, ISAd, first, first + t, first + (t << 1));
    middle = tr_median3(SA, ISAd, middle - t, middle, middle + t);
    last = tr_median3(SA, ISAd, last - 1 - (t << 1), last - 1 - t, last - 1);
    return tr_median3(SA, ISAd, first, middle, last);
}

#define POP_OR_RETURN()                                               \
    do {                                                              \
        if (!tr_pop(&stack, &ISAd, &first, &last, &limit, &trlink))   \
            return;                                                   \
    } while (0)

/* TrSort.cs:148-769.  ISA / ISAd are offsets into SA ("isaOffset" / "isadOffset" in the C#). */
static void tr_introsort(idx_t ISA, idx_t ISAd, idx_t *SA, idx_t first, idx_t last, budget_t *budget)
{
    idx_t a = 0, b = 0, c, v, x, next, trlink = -1;
    idx_t incr = ISAd - ISA;
    tr_item items[TR_STACK_SIZE];
    tr_stack stack = {items, 0};
    memset(items, 0, sizeof items);
    idx_t limit = tr_ilg(last - first);
    for (;;) {
        if (limit < 0) {
            if (limit == -1) {
                /* tandem repeat partition */
                tr_partition(SA, ISAd - incr, first, first, last, &a, &b, last - 1);
                /* update ranks */
                if (a < last) {
                    c = first;
                    v = a - 1;
                    while (c < a) {
                        SA[ISA + SA[c]] = v;
                        c += 1;
                    }
                }
                if (b < last) {
                    c = a;
                    v = b - 1;
                    while (c < b) {
                        SA[ISA + SA[c]] = v;
                        c += 1;
                    }
                }
                /* push */
                if (1 < (b - a)) {
                    tr_push(&stack, 0, a, b, 0, 0);
                    tr_push(&stack, ISAd - incr, first, last, -2, trlink);
                    trlink = stack.size - 2;
                }
                if ((a - first) <= (last - b)) {
                    if (1 < (a - first)) {
                        tr_push(&stack, ISAd, b, last, tr_ilg(last - b), trlink);
                        last = a;
                        limit = tr_ilg(a - first);
                    } else if (1 < (last - b)) {
                        first = b;
                        limit = tr_ilg(last - b);
                    } else {
                        POP_OR_RETURN();
                    }
                } else {
                    if (1 < (last - b)) {
                        tr_push(&stack, ISAd, first, a, tr_ilg(a - first), trlink);
                        first = b;
                        limit = tr_ilg(last - b);
                    } else if (1 < (a - first)) {
                        last = a;
                        limit = tr_ilg(a - first);
                    } else {
                        POP_OR_RETURN();
                    }
                }
            } else if (limit == -2) {
                /* tandem repeat copy */
                tr_item *item = &stack.items[--stack.size];
                a = item->b;
                b = item->c;
                if (item->d == 0) {
                    tr_copy(ISA, SA, first, a, b, last, ISAd - ISA);
                } else {
                    if (0 <= trlink)
                        stack.items[trlink].d = -1;
                    tr_partialcopy(ISA, SA, first, a, b, last, ISAd - ISA);
                }
                POP_OR_RETURN();
            } else {
                /* sorted partition */
                if (0 <= SA[first]) {
                    a = first;
                    for (;;) {
                        SA[ISA + SA[a]] = a;
                        a += 1;
                        if (!((a < last) && (0 <= SA[a])))
                            break;
                    }
                    first = a;
                }
                if (first < last) {
                    a = first;
                    for (;;) {
                        SA[a] = ~SA[a];
                        a += 1;
                        if (!(SA[a] < 0))
                            break;
                    }
                    next = SA[ISA + SA[a]] != ISAD(SA[a]) ? tr_ilg(a - first + 1) : -1;
                    a += 1;
                    if (a < last) {
                        b = first;
                        v = a - 1;
                        while (b < a) {
                            SA[ISA + SA[b]] = v;
                            b += 1;
                        }
                    }
                    /* push */
                    if (budget_check(budget, a - first)) {
                        if ((a - first) <= (last - a)) {
                            tr_push(&stack, ISAd, a, last, -3, trlink);
                            ISAd += incr;
                            last = a;
                            limit = next;
                        } else {
                            if (1 < (last - a)) {
                                tr_push(&stack, ISAd + incr, first, a, next, trlink);
                                first = a;
                                limit = -3;
                            } else {
                                ISAd += incr;
                                last = a;
                                limit = next;
                            }
                        }
                    } else {
                        if (0 <= trlink)
                            stack.items[trlink].d = -1;
                        if (1 < (last - a)) {
                            first = a;
                            limit = -3;
                        } else {
                            POP_OR_RETURN();
                        }
                    }
                } else {
                    POP_OR_RETURN();
                }
            }
            continue;
        }
        if ((last - first) <= TR_INSERTIONSORT_THRESHOLD) {
            tr_insertionsort(SA, ISAd, first, last);
            limit = -3;
            continue;
        }
        idx_t old_limit = limit;
        limit -= 1;
        if (old_limit == 0) {
            tr_heapsort(ISAd, SA, first, last - first);
            a = last - 1;
            while (first < a) {
                x = ISAD(SA[a]);
                b = a - 1;
                while ((first <= b) && (ISAD(SA[b]) == x)) {
                    SA[b] = ~SA[b];
                    b -= 1;
                }
                a = b;
            }
            limit = -3;
            continue;
        }
        /* choose pivot */
        a = tr_pivot(SA, ISAd, first, last);
        SWAP_AT(SA, first, a);
        v = ISAD(SA[first]);
        /* partition */
        tr_partition(SA, ISAd, first, first + 1, last, &a, &b, v);
        if ((last - first) != (b - a)) {
            next = SA[ISA + SA[a]] != v ? tr_ilg(b - a) : -1;
            /* update ranks */
            c = first;
            v = a - 1;
            while (c < a) {
                SA[ISA + SA[c]] = v;
                c += 1;
            }
            if (b < last) {
                c = a;
                v = b - 1;
                while (c < b) {
                    SA[ISA + SA[c]] = v;
                    c += 1;
                }
            }
            /* push */
            if ((1 < (b - a)) && budget_check(budget, b - a)) {
                if ((a - first) <= (last - b)) {
                    if ((last - b) <= (b - a)) {
                        if (1 < (a - first)) {
                            tr_push(&stack, ISAd + incr, a, b, next, trlink);
                            tr_push(&stack, ISAd, b, last, limit, trlink);
                            last = a;
                        } else if (1 < (last - b)) {
                            tr_push(&stack, ISAd + incr, a, b, next, trlink);
                            first = b;
                        } else {
                            ISAd += incr;
                            first = a;
                            last = b;
                            limit = next;
                        }
                    } else if ((a - first) <= (b - a)) {
                        if (1 < (a - first)) {
                            tr_push(&stack, ISAd, b, last, limit, trlink);
                            tr_push(&stack, ISAd + incr, a, b, next, trlink);
                            last = a;
                        } else {
                            tr_push(&stack, ISAd, b, last, limit, trlink);
                            ISAd += incr;
                            first = a;
                            last = b;
                            limit = next;
                        }
                    } else {
                        tr_push(&stack, ISAd, b, last, limit, trlink);
                        tr_push(&stack, ISAd, first, a, limit, trlink);
                        ISAd += incr;
                        first = a;
                        last = b;
                        limit = next;
                    }
                } else {
                    if ((a - first) <= (b - a)) {
                        if (1 < (last - b)) {
                            tr_push(&stack, ISAd + incr, a, b, next, trlink);
                            tr_push(&stack, ISAd, first, a, limit, trlink);
                            first = b;
                        } else if (1 < (a - first)) {
                            tr_push(&stack, ISAd + incr, a, b, next, trlink);
                            last = a;
                        } else {
                            ISAd += incr;
                            first = a;
                            last = b;
                            limit = next;
                        }
                    } else if ((last - b) <= (b - a)) {
                        if (1 < (last - b)) {
                            tr_push(&stack, ISAd, first, a, limit, trlink);
                            tr_push(&stack, ISAd + incr, a, b, next, trlink);
                            first = b;
                        } else {
                            tr_push(&stack, ISAd, first, a, limit, trlink);
                            ISAd += incr;
                            first = a;
                            last = b;
                            limit = next;
                        }
                    } else {
                        tr_push(&stack, ISAd, first, a, limit, trlink);
                        tr_push(&stack, ISAd, b, last, limit, trlink);
                        ISAd += incr;
                        first = a;
                        last = b;
                        limit = next;
                    }
                }
            } else {
                if ((1 < (b - a)) && (0 <= trlink))
                    stack.items[trlink].d = -1;
                if ((a - first) <= (last - b)) {
                    if (1 < (a - first)) {
                        tr_push(&stack, ISAd, b, last, limit, trlink);
                        last = a;
                    } else if (1 < (last - b)) {
                        first = b;
                    } else {
                        POP_OR_RETURN();
                    }
                } else {
                    if (1 < (last - b)) {
                        tr_push(&stack, ISAd, first, a, limit, trlink);
                        first = b;
                    } else if (1 < (a - first)) {
                        last = a;
                    } else {
                        POP_OR_RETURN();
                    }
                }
            }
        } else {
            if (budget_check(budget, last - first)) {
                limit = tr_ilg(last - first);
                ISAd += incr;
            } else {
                if (0 <= trlink)
                    stack.items[trlink].d = -1;
                POP_OR_RETURN();
            }
        }
    }
}
#undef POP_OR_RETURN

/* TrSort.cs:19-102 */
static void trsort(idx_t ISA, idx_t *SA, idx_t n, idx_t depth)
{
    idx_t ISAd, first, last, t, skip, unsorted;
    budget_t budget = {tr_ilg(n) * 2 / 3, n, n, 0};
    ISAd = ISA + depth;
    while (-n < SA[0]) {
        first = 0;
        skip = 0;
        unsorted = 0;
        for (;;) {
            t = SA[first];
            if (t < 0) {
                first -= t;
                skip += t;
            } else {
                if (skip != 0) {
                    SA[first + skip] = skip;
                    skip = 0;
                }
                last = SA[ISA + t] + 1;
                if (1 < (last - first)) {
                    budget.count = 0;
                    tr_introsort(ISA, ISAd, SA, first, last, &budget);
                    if (budget.count != 0)
                        unsorted += budget.count;
                    else
                        skip = first - last;
                } else if ((last - first) == 1) {
                    skip = -1;
                }
                first = last;
            }
            if (!(first < n))
                break;
        }
        if (skip != 0)
            SA[first + skip] = skip;
        if (unsorted == 0)
            break;
        ISAd += ISAd - ISA;
    }
}
#undef ISAD

/* =============================================== DivSufSort.cs =============================================== */

/* DivSufSort.cs:162-184: BStarBucket[(c0,c1)] = B[(c0<<8)|c1], BBucket[(c0,c1)] = B[(c1<<8)|c0] */
#define BSTAR(c0, c1) B[((c0) << 8) | (c1)]
#define BB(c0, c1) B[((c1) << 8) | (c0)]

/* DivSufSort.cs:186-511; returns m */
static idx_t sort_typeBstar(const uint8_t *T, idx_t *SA, idx_t *A, idx_t *B, idx_t n)
{
    idx_t c0, c1, i, j, k, t, m;
    /* count the first one or two characters of each type A, B and B* suffix; store the B* positions in SA */
    i = n - 1;
    m = n;
    c0 = T[n - 1];
    while (0 <= i) {
        /* type A suffix */
        for (;;) {
            c1 = c0;
            A[c1] += 1;
            i -= 1;
            if (0 > i)
                break;
            c0 = T[i];
            if (c0 < c1)
                break;
        }
        if (0 <= i) {
            /* type B* suffix */
            BSTAR(c0, c1) += 1;
            m -= 1;
            SA[m] = i;
            /* type B suffix */
            i -= 1;
            c1 = c0;
            for (;;) {
                if (0 > i)
                    break;
                c0 = T[i];
                if (c0 > c1)
                    break;
                BB(c0, c1) += 1;
                i -= 1;
                c1 = c0;
            }
        }
    }
    m = n - m;

    /* start/end point of each bucket */
    i = 0;
    j = 0;
    for (c0 = 0; c0 < ALPHABET_SIZE; c0++) {
        t = i + A[c0];
        A[c0] = i + j; /* start point */
        i = t + BB(c0, c0);
        for (c1 = c0 + 1; c1 < ALPHABET_SIZE; c1++) {
            j += BSTAR(c0, c1);
            BSTAR(c0, c1) = j; /* end point */
            i += BB(c0, c1);
        }
    }

    if (0 < m) {
        /* sort the type B* suffixes by their first two characters */
        idx_t PAb = n - m;
        idx_t ISAb = m;
        for (i = m - 2; i >= 0; i--) {
            t = SA[PAb + i];
            c0 = T[t];
            c1 = T[t + 1];
            BSTAR(c0, c1) -= 1;
            SA[BSTAR(c0, c1)] = i;
        }
        t = SA[PAb + m - 1];
        c0 = T[t];
        c1 = T[t + 1];
        BSTAR(c0, c1) -= 1;
        SA[BSTAR(c0, c1)] = m - 1;

        /* sort the type B* substrings using sssort */
        idx_t buf = m;
        idx_t bufsize = n - (2 * m);
        c0 = ALPHABET_SIZE - 2;
        j = m;
        while (0 < j) {
            c1 = ALPHABET_SIZE - 1;
            while (c0 < c1) {
                i = BSTAR(c0, c1);
                if (1 < (j - i))
                    sssort(T, SA, PAb, i, j, buf, bufsize, 2, n, SA[i] == (m - 1));
                j = i;
                c1 -= 1;
            }
            c0 -= 1;
        }

        /* compute ranks of type B* substrings */
        i = m - 1;
        while (0 <= i) {
            if (0 <= SA[i]) {
                j = i;
                for (;;) {
                    SA[ISAb + SA[i]] = i;
                    i -= 1;
                    if (!((0 <= i) && (0 <= SA[i])))
                        break;
                }
                SA[i + 1] = i - j;
                if (i <= 0)
                    break;
            }
            j = i;
            for (;;) {
                SA[i] = ~SA[i];
                SA[ISAb + SA[i]] = j;
                i -= 1;
                if (!(SA[i] < 0))
                    break;
            }
            SA[ISAb + SA[i]] = j;
            i -= 1;
        }

        /* inverse suffix array of the type B* suffixes using trsort */
        trsort(ISAb, SA, m, 1);

        /* set the sorted order of type B* suffixes */
        i = n - 1;
        j = m;
        c0 = T[n - 1];
        while (0 <= i) {
            i -= 1;
            c1 = c0;
            for (;;) {
                if (!(0 <= i))
                    break;
                c0 = T[i];
                if (!(c0 >= c1))
                    break;
                i -= 1;
                c1 = c0;
            }
            if (0 <= i) {
                t = i;
                i -= 1;
                c1 = c0;
                for (;;) {
                    if (!(0 <= i))
                        break;
                    c0 = T[i];
                    if (!(c0 <= c1))
                        break;
                    i -= 1;
                    c1 = c0;
                }
                j -= 1;
                {
                    idx_t pos = SA[ISAb + j];
                    SA[pos] = (t == 0 || (1 < (t - i))) ? t : ~t;
                }
            }
        }

        /* start/end point of each bucket; move the type B* suffixes to their final places */
        BB(ALPHABET_SIZE - 1, ALPHABET_SIZE - 1) = n; /* end point */
        c0 = ALPHABET_SIZE - 2;
        k = m - 1;
        while (0 <= c0) {
            i = A[c0 + 1] - 1;
            c1 = ALPHABET_SIZE - 1;
            while (c0 < c1) {
                t = i - BB(c0, c1);
                BB(c0, c1) = i; /* end point */
                i = t;
                j = BSTAR(c0, c1);
                while (j <= k) {
                    SA[i] = SA[k];
                    i -= 1;
                    k -= 1;
                }
                c1 -= 1;
            }
            BSTAR(c0, c0 + 1) = i - BB(c0, c0) + 1;
            BB(c0, c0) = i; /* end point */
            c0 -= 1;
        }
    }
    return m;
}

/* DivSufSort.cs:44-153 */
static void construct_SA(const uint8_t *T, idx_t *SA, idx_t *A, idx_t *B, idx_t n, idx_t m)
{
    idx_t i, j, k, s, c0, c1, c2;
    if (0 < m) {
        /* sorted order of type B suffixes from the sorted order of type B* suffixes */
        c1 = ALPHABET_SIZE - 2;
        while (0 <= c1) {
            /* scan the suffix array from right to left */
            i = BSTAR(c1, c1 + 1);
            j = A[c1 + 1] - 1;
            k = 0;
            c2 = -1;
            while (i <= j) {
                s = SA[j];
                if (0 < s) {
                    SA[j] = ~s;
                    s -= 1;
                    c0 = T[s];
                    if ((0 < s) && (T[s - 1] > c0))
                        s = ~s;
                    if (c0 != c2) {
                        if (0 <= c2)
                            BB(c2, c1) = k;
                        c2 = c0;
                        k = BB(c2, c1);
                    }
                    SA[k] = s;
                    k -= 1;
                } else {
                    SA[j] = ~s;
                }
                j -= 1;
            }
            c1 -= 1;
        }
    }
    /* the suffix array from the sorted order of type B suffixes */
    c2 = T[n - 1];
    k = A[c2];
    SA[k] = T[n - 2] < c2 ? ~(n - 1) : n - 1;
    k += 1;
    /* scan the suffix array from left to right */
    i = 0;
    j = n;
    while (i < j) {
        s = SA[i];
        if (0 < s) {
            s -= 1;
            c0 = T[s];
            if ((s == 0) || (T[s - 1] < c0))
                s = ~s;
            if (c0 != c2) {
                A[c2] = k;
                c2 = c0;
                k = A[c2];
            }
            SA[k] = s;
            k += 1;
        } else {
            SA[i] = ~s;
        }
        i += 1;
    }
}

/*
 * LibDivSufSort.Sort(text, suffixes) -> DivSufSort.divsufsort (DivSufSort.cs:18-42).
 * Returns 0, or -1 when the bucket tables cannot be allocated.  SA need not be zeroed (LibDivSufSort.cs:14).
 */
int oracle_divsufsort(const uint8_t *T, int32_t n, int32_t *SA)
{
    if (n == 0)
        return 0;
    if (n == 1) {
        SA[0] = 0;
        return 0;
    }
    if (n == 2) {
        if (T[0] < T[1]) {
            SA[0] = 0;
            SA[1] = 1;
        } else {
            SA[0] = 1;
            SA[1] = 0;
        }
        return 0;
    }
    /* "These MUST be zeroed first" (DivSufSort.cs:190-192) */
    idx_t *A = (idx_t *)calloc(BUCKET_A_SIZE, sizeof(idx_t));
    idx_t *B = (idx_t *)calloc(BUCKET_B_SIZE, sizeof(idx_t));
    if (!A || !B) {
        free(A);
        free(B);
        return -1;
    }
    idx_t m = sort_typeBstar(T, SA, A, B, n);
    construct_SA(T, SA, A, B, n, m);
    free(A);
    free(B);
    return 0;
}
