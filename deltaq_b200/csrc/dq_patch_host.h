// dq_patch_host.h -- Patch.ApplyInternal on uncompressed streams (host code).
//
// Mirrors /root/reference/src/DeltaQ.BsDiff/Patch.cs:95-168: per control triple (add, copy, seek) -- add `add` bytes
// of the old file to `add` bytes of the diff stream (:143-144, here 16 bytes per step), copy `copy` bytes of the
// extra stream, seek the old file by `seek`.  The reference's sanity checks (:128, :151) raise "Corrupt patch"; this
// restatement also rejects what makes the reference misbehave instead of failing -- negative sizes, a short control
// stream, reads past the end of old/diff/extra (the reference's chunk loop would spin on a zero-byte read).
#pragma once
#include <cstdint>
#include <cstring>
#include <emmintrin.h>

namespace dq {
namespace patchhost {

// SpanExtensions.ReadPackedLong (SpanExtensions.cs:20-30): sign-magnitude little-endian
inline int64_t read_packed_long(const uint8_t *b)
{
    uint64_t y = 0;
    for (int i = 7; i >= 0; --i) y = (y << 8) | (uint64_t)(i == 7 ? (b[i] & 0x7f) : b[i]);
    return (b[7] & 0x80) ? -(int64_t)y : (int64_t)y;
}

// returns 0, or -1 for a corrupt patch
inline int apply_streams(const uint8_t *old_, int64_t n, const uint8_t *ctrl, int64_t ctrl_len, const uint8_t *diff,
                         int64_t diff_len, const uint8_t *extra, int64_t extra_len, uint8_t *out, int64_t new_size)
{
    int64_t out_pos = 0, old_pos = 0, cp = 0, dp = 0, ep = 0;
    while (out_pos < new_size) {
        if (cp + 24 > ctrl_len) return -1;
        const int64_t add = read_packed_long(ctrl + cp);
        const int64_t copy = read_packed_long(ctrl + cp + 8);
        const int64_t seek = read_packed_long(ctrl + cp + 16);
        cp += 24;
        if (add < 0 || copy < 0) return -1;
        if (out_pos + add > new_size) return -1;                        // Patch.cs:128
        if (old_pos < 0 || old_pos + add > n || dp + add > diff_len) return -1;
        const uint8_t *d = diff + dp, *o = old_ + old_pos;
        uint8_t *w = out + out_pos;
        int64_t i = 0;
        for (; i + 16 <= add; i += 16)                                    // Patch.cs:143-144
            _mm_storeu_si128(reinterpret_cast<__m128i *>(w + i),
                             _mm_add_epi8(_mm_loadu_si128(reinterpret_cast<const __m128i *>(d + i)),
                                          _mm_loadu_si128(reinterpret_cast<const __m128i *>(o + i))));
        for (; i < add; ++i) w[i] = (uint8_t)(d[i] + o[i]);
        dp += add;
        old_pos += add;
        out_pos += add;
        if (out_pos + copy > new_size) return -1;                       // Patch.cs:151
        if (ep + copy > extra_len) return -1;
        if (copy) memcpy(out + out_pos, extra + ep, (size_t)copy);
        ep += copy;
        out_pos += copy;
        old_pos += seek;                                                // Patch.cs:163
    }
    return 0;
}

}  // namespace patchhost
}  // namespace dq
