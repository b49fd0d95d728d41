/* A plain C (C99) caller of include/deltaq_cuda.h -- what a P/Invoke or cgo stub does, without Python in between.
 * Built and run by tests/test_abi.py:
 *   gcc -std=c99 -Wall -Werror -I include tests/c_abi_caller.c -L deltaq_b200 -ldeltaq_cuda -o ...
 * With a device: ISuffixSort.Sort of a known string (LibDivSufSortTests.cs:43-64 checks the same property: the suffixes
 * named by the array are in strictly increasing order), Diff.Create's streams of a small pair, Patch.ApplyInternal on
 * them (BsDiffTests.cs:30-78: the round trip reproduces `new`), the whole patch file and Patch.Apply on it.
 * Without a device: dq_cuda_create must say DQ_ERR_NO_DEVICE (there is no CPU fallback); the host-only entry points
 * (bzip2 sections, Patch.Apply) still run.  Exit code 0 = every check held. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "deltaq_cuda.h"

#define CHECK(cond, what)                                                \
    do {                                                                 \
        if (!(cond)) {                                                   \
            fprintf(stderr, "FAILED: %s (line %d)\n", what, __LINE__);   \
            return 1;                                                    \
        }                                                                \
    } while (0)

static int suffix_less(const uint8_t *t, int32_t n, int32_t a, int32_t b)
{
    while (a < n && b < n) {
        if (t[a] != t[b]) return t[a] < t[b];
        ++a;
        ++b;
    }
    return a == n && b != n; /* the shorter suffix is the smaller one (Span.SequenceCompareTo) */
}

static int host_only_checks(void)
{
    /* one section through the block-parallel producer and back */
    enum { N = 300000 };
    uint8_t *src = (uint8_t *)malloc(N), *back = (uint8_t *)malloc(N);
    uint32_t x = 12345u;
    int64_t i, cap, out_len = 0, got = 0;
    int32_t info[3] = {0, 0, 0}, dinfo[2] = {0, 0};
    uint8_t *packed;
    const uint8_t *srcs[1];
    uint8_t *outs[1];
    int64_t lens[1], caps[1];
    CHECK(src && back, "malloc");
    for (i = 0; i < N; ++i) {
        x = x * 1664525u + 1013904223u;
        src[i] = (uint8_t)((i / 1000) % 7 == 0 ? 0 : (x >> 24) & 15);
    }
    cap = dq_cuda_bz2_bound(N);
    CHECK(cap > 0, "dq_cuda_bz2_bound");
    packed = (uint8_t *)malloc((size_t)cap);
    CHECK(packed, "malloc");
    srcs[0] = src;
    lens[0] = N;
    outs[0] = packed;
    caps[0] = cap;
    CHECK(dq_cuda_bz2_compress(srcs, lens, 1, 1, 2, outs, caps, &out_len, info) == DQ_OK, "dq_cuda_bz2_compress");
    CHECK(out_len > 4 && packed[0] == 'B' && packed[1] == 'Z' && packed[2] == 'h' && packed[3] == '1', "bzip2 header");
    CHECK(dq_cuda_bz2_decompress(packed, out_len, 2, back, N, &got, dinfo) == DQ_OK, "dq_cuda_bz2_decompress");
    CHECK(got == N && memcmp(src, back, N) == 0, "bzip2 round trip");
    /* a capacity that is too small reports the size and writes nothing past it */
    CHECK(dq_cuda_bz2_decompress(packed, out_len, 2, back, 10, &got, dinfo) == DQ_ERR_INVALID_ARGUMENT && got == N,
          "dq_cuda_bz2_decompress with a short buffer");
    free(packed);
    free(back);
    free(src);
    {
        /* Patch.ApplyInternal on hand-made streams: add 3 bytes of diff onto old[0..3), copy 2 extra bytes, seek +1,
         * then add 2 more (packed longs: 8-byte little-endian magnitude, sign in bit 63; SpanExtensions.cs:7-30) */
        const uint8_t old_[6] = {10, 20, 30, 40, 50, 60};
        uint8_t ctrl[48], out[7];
        const uint8_t diff[5] = {1, 2, 3, 0, 255}, extra[2] = {7, 8};
        const uint8_t want[7] = {11, 22, 33, 7, 8, 50, 59};
        memset(ctrl, 0, sizeof ctrl);
        ctrl[0] = 3;
        ctrl[8] = 2;
        ctrl[16] = 1;
        ctrl[24] = 2;
        CHECK(dq_cuda_patch_apply(old_, 6, ctrl, 48, diff, 5, extra, 2, out, 7) == DQ_OK, "dq_cuda_patch_apply");
        CHECK(memcmp(out, want, 7) == 0, "Patch.ApplyInternal result");
        ctrl[7] = 0x80; /* add = -3 */
        CHECK(dq_cuda_patch_apply(old_, 6, ctrl, 48, diff, 5, extra, 2, out, 7) == DQ_ERR_CORRUPT_PATCH,
              "a negative add length is a corrupt patch (Patch.cs:128)");
    }
    return 0;
}

static int device_checks(dq_ctx *ctx)
{
    /* ISuffixSort.Sort: exactly n entries are written (Diff.Create relies on I[n] staying 0: Diff.cs:78,90) */
    static const char text_s[] = "mississippi$mississippi#the quick brown fox jumps over the lazy dog; mississippi";
    const int32_t n = (int32_t)(sizeof text_s - 1);
    const uint8_t *text = (const uint8_t *)text_s;
    int32_t sa[sizeof text_s + 1];
    uint8_t seen[sizeof text_s];
    int32_t i;
    dq_stats st;
    sa[n] = -7;
    CHECK(dq_cuda_suffix_sort(ctx, text, n, sa) == DQ_OK, dq_cuda_last_error(ctx));
    CHECK(sa[n] == -7, "dq_cuda_suffix_sort wrote past n entries");
    memset(seen, 0, sizeof seen);
    for (i = 0; i < n; ++i) {
        CHECK(sa[i] >= 0 && sa[i] < n && !seen[sa[i]], "suffix array is not a permutation");
        seen[sa[i]] = 1;
        if (i) CHECK(suffix_less(text, n, sa[i - 1], sa[i]), "suffixes out of order");
    }
    CHECK(dq_cuda_get_stats(ctx, &st) == DQ_OK && st.n == n && st.kernel_launches > 0, "dq_cuda_get_stats");
    CHECK(dq_cuda_suffix_sort(ctx, text, 0, sa) == DQ_OK, "n == 0 is a no-op");
    CHECK(dq_cuda_suffix_sort(ctx, NULL, 5, sa) == DQ_ERR_INVALID_ARGUMENT, "null text is an argument error");
    {
        /* Diff.Create: streams, Patch.ApplyInternal on them, the patch file, Patch.Apply on it */
        enum { N = 70000, M = 70300 };
        uint8_t *old_ = (uint8_t *)malloc(N), *new_ = (uint8_t *)malloc(M), *out = (uint8_t *)malloc(M);
        uint32_t x = 99u;
        dq_diff_streams s;
        const uint8_t *patch = NULL;
        int64_t patch_len = 0, new_size = -1;
        int32_t *pos = (int32_t *)malloc(sizeof(int32_t) * M), *len = (int32_t *)malloc(sizeof(int32_t) * M);
        CHECK(old_ && new_ && out && pos && len, "malloc");
        for (i = 0; i < N; ++i) {
            x = x * 1664525u + 1013904223u;
            old_[i] = (uint8_t)((x >> 24) & 3);
        }
        memcpy(new_, old_, 30000);
        for (i = 0; i < 300; ++i) {
            x = x * 1664525u + 1013904223u;
            new_[30000 + i] = (uint8_t)(x >> 24);
        }
        memcpy(new_ + 30300, old_ + 30000, N - 30000);
        new_[100] ^= 1;
        new_[60000] ^= 2;
        CHECK(dq_cuda_bsdiff_streams(ctx, old_, N, new_, M, &s) == DQ_OK, dq_cuda_last_error(ctx));
        CHECK(s.ctrl_len > 0 && s.ctrl_len % 24 == 0 && s.diff_len + s.extra_len == M, "stream sizes");
        CHECK(dq_cuda_patch_apply(old_, N, s.ctrl, s.ctrl_len, s.diff, s.diff_len, s.extra, s.extra_len, out, M) == DQ_OK,
              "dq_cuda_patch_apply on the streams");
        CHECK(memcmp(out, new_, M) == 0, "the streams do not reproduce new");
        /* the same in three calls -- Sort, Search at every scan position, the host consumer over the table -- must give
         * the same streams (the context's stream buffers are valid until the next call on it: keep a copy) */
        {
            const int64_t c0 = s.ctrl_len, d0 = s.diff_len, e0 = s.extra_len;
            uint8_t *ctrl_copy = (uint8_t *)malloc((size_t)c0);
            int32_t *sa = (int32_t *)malloc(sizeof(int32_t) * (N + 1));
            CHECK(ctrl_copy && sa, "malloc");
            memcpy(ctrl_copy, s.ctrl, (size_t)c0);
            CHECK(dq_cuda_suffix_sort(ctx, old_, N, sa) == DQ_OK, dq_cuda_last_error(ctx));
            CHECK(dq_cuda_bsdiff_search(ctx, old_, N, NULL, new_, M, 0, M, pos, len) == DQ_OK, dq_cuda_last_error(ctx));
            CHECK(dq_cuda_greedy_emit(ctx, old_, N, new_, M, pos, len, &s) == DQ_OK, dq_cuda_last_error(ctx));
            CHECK(s.ctrl_len == c0 && s.diff_len == d0 && s.extra_len == e0 && memcmp(ctrl_copy, s.ctrl, (size_t)c0) == 0,
                  "Sort + Search + dq_cuda_greedy_emit give other streams than dq_cuda_bsdiff_streams");
            /* a caller-supplied suffix array (any ISuffixSort, Diff.cs:90) gives the same table */
            sa[N] = 0;
            {
                int32_t *pos2 = (int32_t *)malloc(sizeof(int32_t) * M), *len2 = (int32_t *)malloc(sizeof(int32_t) * M);
                CHECK(pos2 && len2, "malloc");
                CHECK(dq_cuda_bsdiff_search(ctx, old_, N, sa, new_, M, 0, M, pos2, len2) == DQ_OK, dq_cuda_last_error(ctx));
                CHECK(memcmp(pos, pos2, sizeof(int32_t) * M) == 0 && memcmp(len, len2, sizeof(int32_t) * M) == 0,
                      "the table under a caller-supplied suffix array differs");
                free(pos2);
                free(len2);
            }
            free(sa);
            free(ctrl_copy);
        }
        CHECK(dq_cuda_bsdiff_patch(ctx, old_, N, new_, M, 9, &patch, &patch_len) == DQ_OK, dq_cuda_last_error(ctx));
        CHECK(patch_len > 32 && memcmp(patch, "BSDIFF40", 8) == 0, "BSDIFF40 header (Constants.cs:5-12)");
        CHECK(dq_cuda_bspatch(old_, N, patch, patch_len, 0, NULL, 0, &new_size) == DQ_ERR_INVALID_ARGUMENT && new_size == M,
              "dq_cuda_bspatch size query");
        memset(out, 0, M);
        CHECK(dq_cuda_bspatch(old_, N, patch, patch_len, 0, out, M, &new_size) == DQ_OK, "dq_cuda_bspatch");
        CHECK(memcmp(out, new_, M) == 0, "Patch.Apply does not reproduce new");
        free(pos);
        free(len);
        free(out);
        free(new_);
        free(old_);
    }
    return 0;
}

int main(void)
{
    dq_ctx *ctx = NULL;
    int rc = dq_cuda_create(&ctx, NULL, 0);
    if (host_only_checks()) return 1;
    if (rc == DQ_ERR_NO_DEVICE) {
        CHECK(ctx == NULL, "a failed dq_cuda_create must not hand out a context");
        printf("no device: %s\nhost-only entry points ok\n", dq_cuda_last_error(NULL));
        return 0;
    }
    CHECK(rc == DQ_OK && ctx != NULL, dq_cuda_last_error(NULL));
    if (device_checks(ctx)) return 1;
    CHECK(dq_cuda_destroy(ctx) == DQ_OK, "dq_cuda_destroy");
    printf("device path ok\n");
    return 0;
}
