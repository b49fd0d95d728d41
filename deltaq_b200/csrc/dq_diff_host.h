// dq_diff_host.h -- host consumer of the bulk (pos, len) table: the reference's greedy scan / extend / emit
// loop, /root/reference/src/DeltaQ.BsDiff/Diff.cs:92-223, kept statement for statement except that the call
//     len = Search(I, oldData, newData[scan..], 0, oldData.Length, out pos);          (Diff.cs:106)
// is the array read  len = len_tab[scan]; pos = pos_tab[scan];  -- the hook SURVEY.md section 8(b) describes.
// Emits the three UNCOMPRESSED streams (ctrl triples as packed longs, SpanExtensions.cs:7-30).
#pragma once
#include <cstdint>
#include <vector>

namespace dq {
namespace diffhost {

struct Streams {
    std::vector<uint8_t> ctrl, diff, extra;
    int64_t visits = 0;
};

inline void put_packed_long(std::vector<uint8_t> &out, int64_t y)
{
    uint64_t u = y < 0 ? (uint64_t)0 - (uint64_t)y : (uint64_t)y;
    uint8_t b[8];
    for (int i = 0; i < 8; ++i) b[i] = (uint8_t)(u >> (8 * i));
    if (y < 0) b[7] |= 0x80;
    out.insert(out.end(), b, b + 8);
}

inline void greedy_emit(const uint8_t *oldData, int32_t oldLen, const uint8_t *newData, int32_t newLen,
                        const int32_t *pos_tab, const int32_t *len_tab, Streams &out)
{
    out.ctrl.clear();
    out.diff.clear();
    out.extra.clear();
    out.visits = 0;
    out.diff.reserve((size_t)newLen);

    int32_t scan = 0, pos = 0, len = 0;
    int32_t lastscan = 0, lastpos = 0, lastoffset = 0;

    while (scan < newLen) {
        int32_t oldscore = 0;

        for (int32_t scsc = scan += len; scan < newLen; scan++) {
            len = len_tab[scan];
            pos = pos_tab[scan];
            out.visits++;

            for (; scsc < scan + len; scsc++)
                if ((scsc + lastoffset < oldLen) && (oldData[scsc + lastoffset] == newData[scsc])) oldscore++;

            if ((len == oldscore && len != 0) || (len > oldscore + 8)) break;

            if ((scan + lastoffset < oldLen) && (oldData[scan + lastoffset] == newData[scan])) oldscore--;
        }

        if (len != oldscore || scan == newLen) {
            int32_t s = 0, sf = 0, lenf = 0;
            for (int32_t i = 0; (lastscan + i < scan) && (lastpos + i < oldLen);) {
                if (oldData[lastpos + i] == newData[lastscan + i]) s++;
                i++;
                if (s * 2 - i > sf * 2 - lenf) {
                    sf = s;
                    lenf = i;
                }
            }

            int32_t lenb = 0;
            if (scan < newLen) {
                s = 0;
                int32_t sb = 0;
                for (int32_t i = 1; (scan >= lastscan + i) && (pos >= i); i++) {
                    if (oldData[pos - i] == newData[scan - i]) s++;
                    if (s * 2 - i > sb * 2 - lenb) {
                        sb = s;
                        lenb = i;
                    }
                }
            }

            if (lastscan + lenf > scan - lenb) {
                const int32_t overlap = (lastscan + lenf) - (scan - lenb);
                s = 0;
                int32_t ss = 0, lens = 0;
                for (int32_t i = 0; i < overlap; i++) {
                    if (newData[lastscan + lenf - overlap + i] == oldData[lastpos + lenf - overlap + i]) s++;
                    if (newData[scan - lenb + i] == oldData[pos - lenb + i]) s--;
                    if (s > ss) {
                        ss = s;
                        lens = i + 1;
                    }
                }
                lenf += lens - overlap;
                lenb -= lens;
            }

            for (int32_t i = 0; i < lenf; i++) out.diff.push_back((uint8_t)(newData[lastscan + i] - oldData[lastpos + i]));

            const int32_t extraLength = (scan - lenb) - (lastscan + lenf);
            if (extraLength > 0)
                out.extra.insert(out.extra.end(), newData + lastscan + lenf, newData + lastscan + lenf + extraLength);

            put_packed_long(out.ctrl, lenf);
            put_packed_long(out.ctrl, extraLength);
            put_packed_long(out.ctrl, (int64_t)((pos - lenb) - (lastpos + lenf)));

            lastscan = scan - lenb;
            lastpos = pos - lenb;
            lastoffset = pos - scan;
        }
    }
}

}  // namespace diffhost
}  // namespace dq
