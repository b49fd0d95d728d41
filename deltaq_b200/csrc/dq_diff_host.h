// dq_diff_host.h -- host consumer of the bulk (pos, len) table: the reference's greedy scan / extend / emit
// loop, /root/reference/src/DeltaQ.BsDiff/Diff.cs:92-223, kept statement for statement except that the call
//     len = Search(I, oldData, newData[scan..], 0, oldData.Length, out pos);          (Diff.cs:106)
// is the array read  len = len_tab[scan]; pos = pos_tab[scan];  -- the hook SURVEY.md section 8(b) describes.
// Emits the three UNCOMPRESSED streams (ctrl triples as packed longs, SpanExtensions.cs:7-30).
//
// Same statements, faster stepping: the byte-at-a-time loops of the reference are advanced 32 bytes at a time
// wherever a whole block behaves uniformly (all bytes equal, or the bound check cannot fire), which leaves
// every variable exactly as the reference's loop would (proofs at each site); anything else falls back to the
// reference's own single-byte step.
#pragma once
#include <cstdint>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

namespace dq {
namespace diffhost {

struct Streams {
    std::vector<uint8_t> ctrl, diff, extra;
    int64_t visits = 0;
};

// bit k of the result is set iff a[k] == b[k], k in [0, 32)
inline uint32_t eq_mask32(const uint8_t *a, const uint8_t *b)
{
#if defined(__SSE2__)
    const __m128i a0 = _mm_loadu_si128(reinterpret_cast<const __m128i *>(a));
    const __m128i b0 = _mm_loadu_si128(reinterpret_cast<const __m128i *>(b));
    const __m128i a1 = _mm_loadu_si128(reinterpret_cast<const __m128i *>(a + 16));
    const __m128i b1 = _mm_loadu_si128(reinterpret_cast<const __m128i *>(b + 16));
    return (uint32_t)_mm_movemask_epi8(_mm_cmpeq_epi8(a0, b0)) |
           ((uint32_t)_mm_movemask_epi8(_mm_cmpeq_epi8(a1, b1)) << 16);
#else
    uint32_t m = 0;
    for (int k = 0; k < 32; ++k) m |= (uint32_t)(a[k] == b[k]) << k;
    return m;
#endif
}

// any of v[0..16) > 8 ?
inline bool any_greater_than_8(const int32_t *v)
{
#if defined(__SSE2__)
    const __m128i eight = _mm_set1_epi32(8);
    __m128i g = _mm_cmpgt_epi32(_mm_loadu_si128(reinterpret_cast<const __m128i *>(v)), eight);
    g = _mm_or_si128(g, _mm_cmpgt_epi32(_mm_loadu_si128(reinterpret_cast<const __m128i *>(v + 4)), eight));
    g = _mm_or_si128(g, _mm_cmpgt_epi32(_mm_loadu_si128(reinterpret_cast<const __m128i *>(v + 8)), eight));
    g = _mm_or_si128(g, _mm_cmpgt_epi32(_mm_loadu_si128(reinterpret_cast<const __m128i *>(v + 12)), eight));
    return _mm_movemask_epi8(g) != 0;
#else
    for (int k = 0; k < 16; ++k)
        if (v[k] > 8) return true;
    return false;
#endif
}

inline int32_t count_equal(const uint8_t *a, const uint8_t *b, int32_t len)
{
    int32_t cnt = 0, k = 0;
    for (; k + 32 <= len; k += 32) cnt += __builtin_popcount(eq_mask32(a + k, b + k));
    for (; k < len; ++k) cnt += (a[k] == b[k]);
    return cnt;
}

inline void put_packed_long(std::vector<uint8_t> &out, int64_t y)
{
    uint64_t u = y < 0 ? (uint64_t)0 - (uint64_t)y : (uint64_t)y;
    uint8_t b[8];
    for (int i = 0; i < 8; ++i) b[i] = (uint8_t)(u >> (8 * i));
    if (y < 0) b[7] |= 0x80;
    out.insert(out.end(), b, b + 8);
}

// One stop of the scan: forward/backward extension, overlap split and emission (Diff.cs:127-222).  lastscan/lastpos
// chain from one stop to the next.
struct EmitState {
    int32_t lastscan = 0, lastpos = 0;
};

inline void emit_stop(const uint8_t *oldData, int32_t oldLen, const uint8_t *newData, int32_t newLen, int32_t scan,
                      int32_t pos, EmitState &st, Streams &out)
{
    int32_t &lastscan = st.lastscan, &lastpos = st.lastpos;
                int32_t s = 0, sf = 0, lenf = 0;
                {
                    // Diff.cs:132-145.  Over a block whose bytes all match, s*2-i rises by one per byte, so if it
                    // ends above the running best the last byte of the block is the (strictly improving) final
                    // update, and if not there is no update at all: one step does the whole block.
                    const int32_t span = (scan - lastscan) < (oldLen - lastpos) ? (scan - lastscan) : (oldLen - lastpos);
                    const uint8_t *po = oldData + lastpos, *pn = newData + lastscan;
                    int32_t i = 0;
                    while (i < span) {
                        if (i + 32 <= span) {
                            const uint32_t eq = eq_mask32(po + i, pn + i);
                            if (eq == 0xffffffffu) {
                                s += 32;
                                i += 32;
                                if (s * 2 - i > sf * 2 - lenf) {
                                    sf = s;
                                    lenf = i;
                                }
                                continue;
                            }
                            {   // p matching bytes can lift the score by at most p inside the block: if even that
                                // does not beat the running best, no statement of the loop body fires here
                                const int32_t p = __builtin_popcount(eq);
                                if (s * 2 - i + p <= sf * 2 - lenf) {
                                    s += p;
                                    i += 32;
                                    continue;
                                }
                            }
                            for (int k = 0; k < 32; ++k) {
                                s += (int32_t)((eq >> k) & 1u);
                                i++;
                                if (s * 2 - i > sf * 2 - lenf) {
                                    sf = s;
                                    lenf = i;
                                }
                            }
                            continue;
                        }
                        if (po[i] == pn[i]) s++;
                        i++;
                        if (s * 2 - i > sf * 2 - lenf) {
                            sf = s;
                            lenf = i;
                        }
                    }
                }

                int32_t lenb = 0;
                if (scan < newLen) {
                    s = 0;
                    int32_t sb = 0;
                    // Diff.cs:152-164, same block argument as the forward loop (walking backwards)
                    const int32_t span = (scan - lastscan) < pos ? (scan - lastscan) : pos;
                    int32_t i = 1;
                    while (i <= span) {
                        if (i + 31 <= span) {
                            // bit k <-> step i + 31 - k (the block is read forwards, the loop walks backwards)
                            const uint32_t eq = eq_mask32(oldData + pos - i - 31, newData + scan - i - 31);
                            if (eq == 0xffffffffu) {
                                s += 32;
                                i += 31;
                                if (s * 2 - i > sb * 2 - lenb) {
                                    sb = s;
                                    lenb = i;
                                }
                                i++;
                                continue;
                            }
                            {
                                const int32_t p = __builtin_popcount(eq);
                                if (s * 2 - (i - 1) + p <= sb * 2 - lenb) {
                                    s += p;
                                    i += 32;
                                    continue;
                                }
                            }
                            for (int k = 31; k >= 0; --k) {
                                s += (int32_t)((eq >> k) & 1u);
                                if (s * 2 - i > sb * 2 - lenb) {
                                    sb = s;
                                    lenb = i;
                                }
                                i++;
                            }
                            continue;
                        }
                        if (oldData[pos - i] == newData[scan - i]) s++;
                        if (s * 2 - i > sb * 2 - lenb) {
                            sb = s;
                            lenb = i;
                        }
                        i++;
                    }
                }

                if (lastscan + lenf > scan - lenb) {
                    const int32_t overlap = (lastscan + lenf) - (scan - lenb);
                    s = 0;
                    int32_t ss = 0, lens = 0;
                    // Diff.cs:172-188.  s <= ss holds after every step; over a 32-byte block s can rise by at most
                    // the number of positions where only the first comparison matches, so a block that cannot lift
                    // s above ss is applied in one step.
                    const uint8_t *n1 = newData + lastscan + lenf - overlap, *o1 = oldData + lastpos + lenf - overlap;
                    const uint8_t *n2 = newData + scan - lenb, *o2 = oldData + pos - lenb;
                    int32_t i = 0;
                    while (i < overlap) {
                        if (i + 32 <= overlap) {
                            const uint32_t ma = eq_mask32(n1 + i, o1 + i), mb = eq_mask32(n2 + i, o2 + i);
                            if (s + __builtin_popcount(ma & ~mb) <= ss) {
                                s += __builtin_popcount(ma) - __builtin_popcount(mb);
                                i += 32;
                                continue;
                            }
                            for (int k = 0; k < 32; ++k, ++i) {
                                s += (int32_t)((ma >> k) & 1u) - (int32_t)((mb >> k) & 1u);
                                if (s > ss) {
                                    ss = s;
                                    lens = i + 1;
                                }
                            }
                            continue;
                        }
                        if (n1[i] == o1[i]) s++;
                        if (n2[i] == o2[i]) s--;
                        if (s > ss) {
                            ss = s;
                            lens = i + 1;
                        }
                        i++;
                    }
                    lenf += lens - overlap;
                    lenb -= lens;
                }

                {
                    // Diff.cs:197-200
                    const size_t at = out.diff.size();
                    out.diff.resize(at + (size_t)(lenf > 0 ? lenf : 0));
                    uint8_t *dst = out.diff.data() + at;
                    const uint8_t *pn = newData + lastscan, *po = oldData + lastpos;
                    int32_t i = 0;
    #if defined(__SSE2__)
                    for (; i + 16 <= lenf; i += 16)
                        _mm_storeu_si128(reinterpret_cast<__m128i *>(dst + i),
                                         _mm_sub_epi8(_mm_loadu_si128(reinterpret_cast<const __m128i *>(pn + i)),
                                                      _mm_loadu_si128(reinterpret_cast<const __m128i *>(po + i))));
    #endif
                    for (; i < lenf; i++) dst[i] = (uint8_t)(pn[i] - po[i]);
                }

                const int32_t extraLength = (scan - lenb) - (lastscan + lenf);
                if (extraLength > 0)
                    out.extra.insert(out.extra.end(), newData + lastscan + lenf, newData + lastscan + lenf + extraLength);

                put_packed_long(out.ctrl, lenf);
                put_packed_long(out.ctrl, extraLength);
                put_packed_long(out.ctrl, (int64_t)((pos - lenb) - (lastpos + lenf)));


    lastscan = scan - lenb;
    lastpos = pos - lenb;
}

// ready(upto): returns once table entries [0, min(upto, newLen)) are valid (the table may still be arriving from
// the device in slices while the loop runs)
template <typename Ready, typename Sink>
inline void greedy_scan(const uint8_t *oldData, int32_t oldLen, const uint8_t *newData, int32_t newLen,
                        const int32_t *pos_tab, const int32_t *len_tab, Streams &out, Ready &&ready, Sink &&sink)
{
    int32_t scan = 0, pos = 0, len = 0;
    int32_t lastoffset = 0;

    while (scan < newLen) {
        int32_t oldscore = 0;

        for (int32_t scsc = scan += len; scan < newLen; scan++) {
            ready(scan + 64);
            // Block step over positions that cannot end the scan.  While oldscore == 0 and no byte of
            // new[scan .. scan+kBlk+8) equals its partner at the current offset, nothing can be added to or
            // taken from oldscore by the next kBlk positions (their matches are <= 8 long, so scsc stays inside
            // that window): each of them evaluates "(len == 0 && len != 0) || len > 8" = false.  The reference
            // would visit them one by one with the same outcome; scsc ends at the same running maximum.
            {
                constexpr int32_t kBlk = 16;
                while (oldscore == 0 && scan + 32 + kBlk <= newLen && scsc <= scan + kBlk + 8 &&
                       (int64_t)scan + lastoffset + 32 <= (int64_t)oldLen) {
                    if (eq_mask32(oldData + scan + lastoffset, newData + scan) & 0x00ffffffu) break;
                    if (any_greater_than_8(len_tab + scan)) break;
                    int32_t e = scsc;
                    for (int32_t i = 0; i < kBlk; ++i) {
                        const int32_t t = scan + i + len_tab[scan + i];
                        e = t > e ? t : e;
                    }
                    scsc = e;
                    scan += kBlk;
                    out.visits += kBlk;
                    ready(scan + 64);
                }
            }
            len = len_tab[scan];
            pos = pos_tab[scan];
            out.visits++;

            {
                // Diff.cs:108-114.  scsc + lastoffset >= lastpos >= 0, so only the upper bound can fire: count
                // equal bytes over the part of [scsc, scan+len) that stays inside oldData.
                const int64_t lim = (int64_t)oldLen - lastoffset;  // scsc < lim  <=>  scsc + lastoffset < oldLen
                const int32_t stop = scan + len;
                const int32_t inb = (int32_t)(lim < stop ? (lim > scsc ? lim : scsc) : stop);
                if (inb > scsc) oldscore += count_equal(oldData + lastoffset + scsc, newData + scsc, inb - scsc);
                if (scsc < stop) scsc = stop;
            }

            if ((len == oldscore && len != 0) || (len > oldscore + 8)) break;

            if ((scan + lastoffset < oldLen) && (oldData[scan + lastoffset] == newData[scan])) oldscore--;
        }

        if (len != oldscore || scan == newLen) {
            // Diff.cs:127-222 for this stop of the scan.  lastoffset (all the scan needs) does not depend on the
            // extension results, so the emission may run on the consumer side (see greedy_emit_pipelined).
            sink(scan, pos);
            lastoffset = pos - scan;
        }
    }
}

inline void reset_streams(Streams &out, int32_t newLen)
{
    out.ctrl.clear();
    out.diff.clear();
    out.extra.clear();
    out.visits = 0;
    out.diff.reserve((size_t)newLen);
}

// the whole loop on the calling thread
template <typename Ready>
inline void greedy_emit(const uint8_t *oldData, int32_t oldLen, const uint8_t *newData, int32_t newLen,
                        const int32_t *pos_tab, const int32_t *len_tab, Streams &out, Ready &&ready)
{
    reset_streams(out, newLen);
    EmitState st;
    greedy_scan(oldData, oldLen, newData, newLen, pos_tab, len_tab, out, ready,
                [&](int32_t scan, int32_t pos) { emit_stop(oldData, oldLen, newData, newLen, scan, pos, st, out); });
}

inline void greedy_emit(const uint8_t *oldData, int32_t oldLen, const uint8_t *newData, int32_t newLen,
                        const int32_t *pos_tab, const int32_t *len_tab, Streams &out)
{
    greedy_emit(oldData, oldLen, newData, newLen, pos_tab, len_tab, out, [](int32_t) {});
}

// Two host threads: the caller runs the scan (it may block on table slices still in flight from the device) and
// hands every stop to a consumer thread that does the extensions and the emission.  The scan needs only
// lastoffset = pos - scan from a stop, never the extension results, so both sides compute exactly what the
// single-threaded loop computes, in the same order.
template <typename Ready>
inline void greedy_emit_pipelined(const uint8_t *oldData, int32_t oldLen, const uint8_t *newData, int32_t newLen,
                                  const int32_t *pos_tab, const int32_t *len_tab, Streams &out, Ready &&ready)
{
    reset_streams(out, newLen);
    struct Stop {
        int32_t scan, pos;
    };
    std::mutex mu;
    std::condition_variable cv;
    std::deque<Stop> q;
    bool done = false;
    std::thread consumer([&]() {
        EmitState st;
        for (;;) {
            Stop sp;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return done || !q.empty(); });
                if (q.empty()) return;
                sp = q.front();
                q.pop_front();
            }
            emit_stop(oldData, oldLen, newData, newLen, sp.scan, sp.pos, st, out);
        }
    });
    int64_t visits = 0;
    Streams scan_side;  // the scan only counts visits; keep its counter off the consumer's Streams
    greedy_scan(oldData, oldLen, newData, newLen, pos_tab, len_tab, scan_side, ready, [&](int32_t scan, int32_t pos) {
        {
            std::lock_guard<std::mutex> lk(mu);
            q.push_back(Stop{scan, pos});
        }
        cv.notify_one();
    });
    visits = scan_side.visits;
    {
        std::lock_guard<std::mutex> lk(mu);
        done = true;
    }
    cv.notify_one();
    consumer.join();
    out.visits = visits;
}

}  // namespace diffhost
}  // namespace dq
