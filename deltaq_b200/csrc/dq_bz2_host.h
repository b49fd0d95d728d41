// dq_bz2_host.h -- block-parallel bzip2 producer for the ctrl / diff / extra sections of a BSDIFF40 patch (host code;
// SURVEY.md section 8(f) rank 2).
//
// The reference wraps each section in a SharpZipLib BZip2OutputStream and feeds it serially (/root/reference/src/
// DeltaQ.BsDiff/Diff.cs:14-19, :85-87, :197-207, :226-241).  Once sort + search + greedy loop take milliseconds, that
// serial bzip2 is two orders of magnitude above everything else in Diff.Create (C2: 7.5 ms against ~1 s).  bzip2 blocks
// are independent (own BWT, own Huffman tables, own CRC), so a stream can be produced block by block on many threads and
// stitched together afterwards.  This file does that on top of the system's libbz2 (dlopen'ed; the compressor itself
// is the library's):
//
//   1. split_blocks() finds the input positions where a SERIAL libbz2 of the same level would start each block.  The
//      first stage of bzip2 turns runs of 4..255 equal bytes into 4 bytes + a count, and a block is closed as soon as
//      100000*level - 19 bytes of that run-length coded text have been collected -- a rule that depends on the input
//      alone, restated here from the published format (bzip2 1.0.x, compress stage "RLE1").
//   2. every piece is compressed as its own one-block stream by BZ2_bzBuffToBuffCompress on a crew of threads;
//   3. stitch() cuts the block out of every piece (bit range after the 32-bit stream header, before the 48-bit
//      end-of-stream magic), concatenates the blocks bit by bit behind one header, and closes the stream with the
//      combined CRC (rotate-left-1, xor block CRC).
//
// The output is therefore BIT-IDENTICAL to what the serial library produces at that level (tests/test_bz2.py compares
// with Python's bz2 for every level), and any bzip2 reader -- SharpZipLib's BZip2InputStream in Patch.cs:52-93
// included -- decodes it as one ordinary stream.  With level 0 ("auto") the level is lowered until there are about two
// pieces per thread: more, smaller blocks (C2's diff section: +0.5 % bytes at level 1 against level 9, 10x the
// parallelism).  If a piece does not come back as exactly one block (it cannot, unless the split rule is wrong for this
// libbz2), the stream is compressed serially instead and the fallback is counted.
#pragma once
#include <dlfcn.h>
#if defined(__linux__)
#include <sched.h>
#endif

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <climits>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

namespace dq {
namespace bz2host {

constexpr uint64_t kBlockMagic = 0x314159265359ull;  // 48 bits in front of every block (pi)
constexpr uint64_t kEndMagic = 0x177245385090ull;    // 48 bits in front of the stream CRC (sqrt(pi))

// int BZ2_bzBuffToBuffCompress(char *dest, unsigned *destLen, char *source, unsigned sourceLen, int blockSize100k,
//                              int verbosity, int workFactor) -- libbz2's one-shot call; 0 = BZ_OK
using BuffToBuffFn = int (*)(char *, unsigned *, char *, unsigned, int, int, int);

inline BuffToBuffFn libbz2()
{
    static BuffToBuffFn fn = [] {
        for (const char *name : {"libbz2.so.1.0", "libbz2.so.1", "libbz2.so"}) {
            if (void *h = dlopen(name, RTLD_NOW | RTLD_LOCAL)) {
                if (void *s = dlsym(h, "BZ2_bzBuffToBuffCompress")) return reinterpret_cast<BuffToBuffFn>(s);
            }
        }
        return static_cast<BuffToBuffFn>(nullptr);
    }();
    return fn;
}

// what a caller must provide for n input bytes (libbz2's own bound for one stream: 1 % + 600 bytes; the stitched
// stream is never longer than the serial one)
inline int64_t bound(int64_t n) { return n + n / 100 + 600; }

// bytes the run-length stage emits for a run piece of L equal bytes, 1 <= L <= 255
inline int64_t coded_size(int64_t L) { return L < 4 ? L : 5; }

// end of the run of equal bytes that starts at i (eight bytes per step once the run has two)
inline int64_t run_end(const uint8_t *p, int64_t i, int64_t n)
{
    const uint8_t c = p[i];
    int64_t j = i + 1;
    if (j < n && p[j] == c) {
        const uint64_t pattern = 0x0101010101010101ull * c;
        ++j;
        while (j + 8 <= n) {
            uint64_t w;
            memcpy(&w, p + j, 8);
            w ^= pattern;
            if (w) return j + (__builtin_ctzll(w) >> 3);
            j += 8;
        }
        while (j < n && p[j] == c) ++j;
    }
    return j;
}

// Run-length coded size of the whole input (what the blocks are filled with); used to choose the level.
inline int64_t coded_total(const uint8_t *p, int64_t n)
{
    int64_t total = 0, i = 0;
    while (i < n) {
        const int64_t j = run_end(p, i, n);
        const int64_t L = j - i;
        total += (L / 255) * 5 + coded_size(L % 255) * (L % 255 ? 1 : 0);
        i = j;
    }
    return total;
}

// Input offsets at which serial libbz2 starts a block (first entry 0; empty input -> no blocks).
// The compressor consumes byte after byte; a run piece (at most 255 equal bytes) is added to the block when the byte
// AFTER it has been consumed, and the block is closed as soon as it holds >= cap coded bytes -- unless that byte was the
// last of the input, in which case it still joins the same block.  The byte consumed last opens the next block.
template <class Emit>
inline void split_blocks(const uint8_t *p, int64_t n, int level, Emit &&emit)
{
    if (n <= 0) return;
    emit((int64_t)0);
    const int64_t cap = 100000ll * level - 19;
    int64_t held = 0, i = 0;
    while (i < n) {
        const int64_t j = run_end(p, i, n);
        int64_t L = j - i;
        if (held + (L / 255 + 1) * 5 < cap) {  // the whole run cannot fill the block: no need to walk its pieces
            held += (L / 255) * 5 + (L % 255 ? coded_size(L % 255) : 0);
            i = j;
            continue;
        }
        while (L > 0) {
            const int64_t s = std::min<int64_t>(L, 255);
            held += coded_size(s);
            i += s;
            L -= s;
            if (held >= cap && i + 1 < n) {  // byte i consumed, more input behind it: the block closes, i opens the next
                emit(i);
                held = 0;
            }
        }
    }
}

inline uint64_t read_bits(const uint8_t *b, int64_t bit, int count)  // count <= 56, big-endian bit order
{
    uint64_t v = 0;
    int64_t byte = bit >> 3;
    const int skip = (int)(bit & 7);
    const int need = (skip + count + 7) >> 3;
    for (int k = 0; k < need; ++k) v = (v << 8) | b[byte + k];
    v >>= (need * 8 - skip - count);
    return count == 64 ? v : (v & ((1ull << count) - 1));
}

class BitWriter {
public:
    BitWriter(uint8_t *out, int64_t cap) : out_(out), cap_(cap) {}
    bool put(uint64_t v, int count)  // count <= 32
    {
        acc_ = (acc_ << count) | (v & ((count == 64 ? 0 : (1ull << count)) - 1));
        have_ += count;
        while (have_ >= 8) {
            if (len_ >= cap_) return false;
            out_[len_++] = (uint8_t)(acc_ >> (have_ - 8));
            have_ -= 8;
        }
        return true;
    }
    // bits [begin, end) of src
    bool append(const uint8_t *src, int64_t begin, int64_t end)
    {
        int64_t bit = begin;
        while (bit < end && ((bit & 7) != 0)) {  // up to a byte boundary of the source
            if (!put(read_bits(src, bit, 1), 1)) return false;
            ++bit;
        }
        if (have_ == 0) {  // both sides byte aligned: plain copy
            const int64_t bytes = (end - bit) >> 3;
            if (len_ + bytes > cap_) return false;
            memcpy(out_ + len_, src + (bit >> 3), (size_t)bytes);
            len_ += bytes;
            bit += bytes * 8;
        } else {
            const int sh = have_;  // 1..7 bits pending
            uint64_t pending = acc_ & ((1ull << sh) - 1);
            const int64_t bytes = (end - bit) >> 3;
            if (len_ + bytes > cap_) return false;
            const uint8_t *s = src + (bit >> 3);
            uint8_t *d = out_ + len_;
            for (int64_t k = 0; k < bytes; ++k) {
                const uint64_t x = (pending << 8) | s[k];
                d[k] = (uint8_t)(x >> sh);
                pending = x & ((1ull << sh) - 1);
            }
            len_ += bytes;
            bit += bytes * 8;
            acc_ = pending;
        }
        while (bit < end) {
            if (!put(read_bits(src, bit, 1), 1)) return false;
            ++bit;
        }
        return true;
    }
    bool finish()  // zero padding to a byte boundary
    {
        return have_ ? put(0, 8 - have_) : true;
    }
    int64_t size() const { return len_; }

private:
    uint8_t *out_;
    int64_t cap_, len_ = 0;
    uint64_t acc_ = 0;
    int have_ = 0;
};

// growable byte buffer without the zero fill of std::vector (realloc moves large blocks by remapping, not copying)
struct RawBuf {
    uint8_t *p = nullptr;
    size_t len = 0, cap = 0;
    RawBuf() = default;
    RawBuf(const RawBuf &) = delete;
    RawBuf &operator=(const RawBuf &) = delete;
    RawBuf(RawBuf &&o) noexcept : p(o.p), len(o.len), cap(o.cap) { o.p = nullptr; o.len = o.cap = 0; }
    ~RawBuf() { free(p); }
    void reserve(size_t want)
    {
        if (want <= cap) return;
        void *q = realloc(p, want);
        if (!q) throw std::bad_alloc();
        p = static_cast<uint8_t *>(q);
        cap = want;
    }
    void shrink()
    {
        if (len && len < cap) {
            if (void *q = realloc(p, len)) p = static_cast<uint8_t *>(q), cap = len;
        }
    }
    void swap(RawBuf &o)
    {
        std::swap(p, o.p);
        std::swap(len, o.len);
        std::swap(cap, o.cap);
    }
    const uint8_t *data() const { return p; }
    size_t size() const { return len; }
    bool empty() const { return len == 0; }
};

struct Piece {
    int stream = 0;
    const uint8_t *src = nullptr;
    int64_t len = 0;
    RawBuf z;                     // the piece as its own bzip2 stream
    int64_t block_end = 0;        // bit offset of the end-of-stream magic in z
    uint32_t crc = 0;             // CRC of the one block in z
    bool ok = false;
};

// z is a stream of exactly one block?  Then find where the block ends.
inline bool locate_block(Piece &pc, int level)
{
    const uint8_t *z = pc.z.data();
    const int64_t bits = (int64_t)pc.z.size() * 8;
    if (pc.z.size() < 4 + 10 + 10 || z[0] != 'B' || z[1] != 'Z' || z[2] != 'h' || z[3] != '0' + level) return false;
    if (read_bits(z, 32, 48) != kBlockMagic) return false;
    pc.crc = (uint32_t)read_bits(z, 80, 32);
    int found = 0;
    for (int pad = 0; pad < 8; ++pad) {
        const int64_t at = bits - pad - 80;
        if (at < 112) break;
        if (read_bits(z, at, 48) != kEndMagic) continue;
        if ((uint32_t)read_bits(z, at + 48, 32) != pc.crc) continue;  // one block: stream CRC == block CRC
        if (pad && read_bits(z, bits - pad, pad) != 0) continue;
        pc.block_end = at;
        ++found;
    }
    return found == 1;
}

struct StreamJob {
    const uint8_t *src;
    int64_t len;
    uint8_t *out;
    int64_t cap;
    int64_t out_len = 0;
    int level = 9;
    int pieces = 0;
    bool fell_back = false;
    int status = 0;  // 0 ok, -1 output buffer too small, -2 libbz2 failed
};

inline unsigned usable_cpus()
{
    unsigned hw = std::max(1u, std::thread::hardware_concurrency());
#if defined(__linux__)
    cpu_set_t set;
    if (sched_getaffinity(0, sizeof set, &set) == 0) {
        const unsigned allowed = (unsigned)CPU_COUNT(&set);
        if (allowed && allowed < hw) hw = allowed;
    }
#endif
    return hw;
}

inline int serial_compress(StreamJob &job, int level)
{
    BuffToBuffFn fn = libbz2();
    if (job.len > (int64_t)0xfffffff0u) return -2;
    unsigned dl = (unsigned)std::min<int64_t>(job.cap, 0xfffffff0u);
    const int rc = fn(reinterpret_cast<char *>(job.out), &dl,
                      const_cast<char *>(reinterpret_cast<const char *>(job.src)), (unsigned)job.len, level, 0, 0);
    if (rc == -8) return -1;  // BZ_OUTBUFF_FULL
    if (rc != 0) return -2;
    job.out_len = dl;
    job.level = level;
    return 0;
}

// Estimate of coded_total from 64 evenly spaced windows of 16 KiB (exact for inputs up to 1 MiB): only the choice of the
// level hangs on it.
inline int64_t coded_estimate(const uint8_t *p, int64_t n)
{
    constexpr int64_t kWindow = 16 << 10, kWindows = 64;
    if (n <= kWindow * kWindows) return coded_total(p, n);
    int64_t coded = 0;
    for (int64_t w = 0; w < kWindows; ++w) coded += coded_total(p + (n - kWindow) / (kWindows - 1) * w, kWindow);
    return (int64_t)((double)coded / (double)(kWindow * kWindows) * (double)n);
}

// run(): every job cut and compressed; level 1..9, or 0 = choose per stream; threads 0 = the CPUs this process may run
// on.  Returns 0, -2 (libbz2 failed), -3 (libbz2 not found).  The jobs' out / cap are not used by the class: size(s) is the
// exact length of section s and write(s, out, cap) stitches it wherever the caller wants it.
//
// The calling thread cuts the sections (largest first) and hands every piece to the crew the moment its end is known, so
// the cut (1-2 GB/s, sequential by nature: where a block ends depends on where the one before it ended) runs
// beside the compression instead of in front of it; then it joins the crew.
class SectionCompressor {
    StreamJob *jobs_ = nullptr;
    int count_ = 0;
    std::deque<Piece> pieces;  // in cut order; references stay valid while the cutter appends
    std::vector<size_t> first_piece;
    std::vector<RawBuf> serial_;  // sections that had to take the serial path

public:
    // cut and compress; afterwards size(s) is exact and write(s, ...) stitches section s
    int run(StreamJob *jobs, int count, int level, int threads);
    int64_t size(int s) const;
    bool write(int s, uint8_t *out, int64_t cap);
};

inline int SectionCompressor::run(StreamJob *jobs, int count, int level, int threads)
{
    if (!libbz2()) return -3;
    jobs_ = jobs;
    count_ = count;
    pieces.clear();
    const bool trace = std::getenv("DQ_TRACE") != nullptr;
    const auto t0 = std::chrono::steady_clock::now();
    auto since = [&] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); };
    const int crew = threads > 0 ? std::min(threads, 256) : (int)usable_cpus();

    int64_t expect = 0;  // pieces, roughly: how many threads are worth starting
    for (int s = 0; s < count; ++s) {
        StreamJob &job = jobs[s];
        job.out_len = 0;
        job.pieces = 0;
        job.fell_back = false;
        job.status = 0;
        const int64_t coded = (level <= 0 || crew > 1) ? coded_estimate(job.src, job.len) : 0;
        int lv = level;
        if (lv <= 0)  // about two pieces per thread, blocks no smaller than level 1's
            lv = crew == 1 ? 9 : (int)std::max<int64_t>(1, std::min<int64_t>(9, coded / (2ll * crew * 100000ll)));
        job.level = std::min(lv, 9);
        expect += coded / (100000ll * job.level - 19) + 1;
    }

    std::mutex mu;
    std::condition_variable cv;
    size_t next = 0;
    bool closed = false;
    std::atomic<int> failed{0};
    auto work = [&] {
        BuffToBuffFn fn = libbz2();
        for (;;) {
            Piece *pc;
            {
                std::unique_lock<std::mutex> lock(mu);
                cv.wait(lock, [&] { return next < pieces.size() || closed; });
                if (next >= pieces.size()) return;
                pc = &pieces[next++];
            }
            try {
                pc->z.reserve((size_t)bound(pc->len));
            } catch (...) {
                failed.store(1);
                continue;
            }
            unsigned dl = (unsigned)pc->z.cap;
            const int rc = fn(reinterpret_cast<char *>(pc->z.p), &dl,
                              const_cast<char *>(reinterpret_cast<const char *>(pc->src)), (unsigned)pc->len,
                              jobs[pc->stream].level, 0, 0);
            if (rc != 0) {
                failed.store(1);
                continue;
            }
            pc->z.len = dl;
            pc->z.shrink();  // the bound is the size of the INPUT piece: give the rest back
            pc->ok = locate_block(*pc, jobs[pc->stream].level);
        }
    };
    std::vector<std::thread> pool;
    try {
        const int64_t helpers = std::min<int64_t>(crew - 1, expect - 1);
        for (int64_t t = 0; t < helpers; ++t) pool.emplace_back(work);
    } catch (...) {  // fewer threads than asked for: the ones that started (and this one) do the work
    }

    // cut: the sections in descending size, each piece published as soon as the next block start is known
    std::vector<int> by_size(count);
    for (int s = 0; s < count; ++s) by_size[s] = s;
    std::stable_sort(by_size.begin(), by_size.end(), [&](int a, int b) { return jobs[a].len > jobs[b].len; });
    first_piece.assign((size_t)count, 0);
    bool cut_failed = false;
    for (int s : by_size) {
        StreamJob &job = jobs[s];
        {
            std::lock_guard<std::mutex> lock(mu);
            first_piece[s] = pieces.size();
        }
        int64_t begin = 0;
        auto publish = [&](int64_t end) {
            Piece pc;
            pc.stream = s;
            pc.src = job.src + begin;
            pc.len = end - begin;
            begin = end;
            ++job.pieces;
            {
                std::lock_guard<std::mutex> lock(mu);
                pieces.push_back(std::move(pc));
            }
            cv.notify_one();
        };
        try {
            split_blocks(job.src, job.len, job.level, [&](int64_t start) { if (start > 0) publish(start); });
            if (job.len > 0) publish(job.len);
        } catch (...) {
            cut_failed = true;
            break;
        }
    }
    {
        std::lock_guard<std::mutex> lock(mu);
        closed = true;
    }
    cv.notify_all();
    if (trace) fprintf(stderr, "[dq trace] bz2: %zu pieces of %d sections cut at %.3f ms\n", pieces.size(), count, since());
    work();
    for (auto &t : pool) t.join();
    if (failed.load() || cut_failed) return -2;
    if (trace) fprintf(stderr, "[dq trace] bz2: pieces compressed by %zu threads at %.3f ms\n", pool.size() + 1, since());

    // a section with a piece that did not come back as one block is compressed serially, here, so that its size is known
    serial_.clear();
    serial_.resize((size_t)count);
    int worst = 0;
    for (int s = 0; s < count; ++s) {
        StreamJob &job = jobs[s];
        bool all_ok = true;
        for (int k = 0; k < job.pieces; ++k) all_ok = all_ok && pieces[first_piece[s] + (size_t)k].ok;
        if (all_ok) continue;
        job.fell_back = true;
        RawBuf &z = serial_[(size_t)s];
        z.reserve((size_t)bound(job.len));
        StreamJob tmp = job;
        tmp.out = z.p;
        tmp.cap = (int64_t)z.cap;
        job.status = serial_compress(tmp, job.level);
        z.len = (size_t)tmp.out_len;
        worst = std::min(worst, job.status);
    }
    if (trace) fprintf(stderr, "[dq trace] bz2: pieces ready at %.3f ms\n", since());
    return worst;
}

inline int64_t SectionCompressor::size(int s) const
{
    const StreamJob &job = jobs_[s];
    if (job.fell_back) return (int64_t)serial_[(size_t)s].len;
    int64_t bits = 32 + 80;  // stream header; end magic + combined CRC
    for (int k = 0; k < job.pieces; ++k) bits += pieces[first_piece[(size_t)s] + (size_t)k].block_end - 32;
    return (bits + 7) / 8;
}

inline bool SectionCompressor::write(int s, uint8_t *out, int64_t cap)
{
    StreamJob &job = jobs_[s];
    job.out_len = 0;
    if (cap < size(s)) {
        job.status = -1;
        return false;
    }
    if (job.fell_back) {
        const RawBuf &z = serial_[(size_t)s];
        if (z.len) memcpy(out, z.p, z.len);
        job.out_len = (int64_t)z.len;
        return true;
    }
    BitWriter w(out, cap);
    bool fits = w.put('B', 8) && w.put('Z', 8) && w.put('h', 8) && w.put((uint64_t)('0' + job.level), 8);
    uint32_t combined = 0;
    for (int k = 0; k < job.pieces && fits; ++k) {
        const Piece &pc = pieces[first_piece[(size_t)s] + (size_t)k];
        fits = w.append(pc.z.data(), 32, pc.block_end);
        combined = ((combined << 1) | (combined >> 31)) ^ pc.crc;
    }
    fits = fits && w.put(kEndMagic >> 24, 24) && w.put(kEndMagic & 0xffffffu, 24) && w.put(combined, 32) && w.finish();
    job.status = fits ? 0 : -1;
    job.out_len = fits ? w.size() : 0;
    return fits;
}

// Compress every job into its out buffer; level 1..9, or 0 = choose per stream; threads 0 = the CPUs this process may
// run on.  Returns 0, -1 (some output buffer too small), -2 (libbz2 failed), -3 (libbz2 not found).
inline int compress_streams(StreamJob *jobs, int count, int level, int threads)
{
    SectionCompressor c;
    const int rc = c.run(jobs, count, level, threads);
    if (rc != 0) return rc;
    int worst = 0;
    for (int s = 0; s < count; ++s)
        if (!c.write(s, jobs[s].out, jobs[s].cap)) worst = -1;
    return worst;
}

// ---- the reader's side: Patch.CreatePatchStreams (Patch.cs:52-93) un-bzip2's three sections serially ---------------------
// Blocks can be decoded independently as well, once their bit offsets are known: every block starts with the 48-bit
// magic, so the section is scanned for it at all eight bit alignments (a table on the byte that follows the first,
// partial, byte keeps that to about one look-up per input byte), every block is wrapped into a one-block stream of its
// own (header + the block's bits + end magic + the block's CRC) and decoded by libbz2 on the crew, and the pieces are laid
// end to end.  The combined CRC of the block CRCs must equal the stream's.  Anything unexpected -- a magic that turns
// out to be compressed data (2^-48 per bit), trailing bytes, several streams in a row -- sends the section to the serial
// decoder instead.

// bz_stream of libbz2 1.0.x (public, stable layout)
struct BzStream {
    char *next_in;
    unsigned avail_in, total_in_lo32, total_in_hi32;
    char *next_out;
    unsigned avail_out, total_out_lo32, total_out_hi32;
    void *state;
    void *(*bzalloc)(void *, int, int);
    void (*bzfree)(void *, void *);
    void *opaque;
};
struct DecodeLib {
    int (*init)(BzStream *, int, int) = nullptr;                                    // BZ2_bzDecompressInit
    int (*step)(BzStream *) = nullptr;                                              // BZ2_bzDecompress
    int (*end)(BzStream *) = nullptr;                                               // BZ2_bzDecompressEnd
    bool ok() const { return init && step && end; }
};
inline const DecodeLib &libbz2_decode()
{
    static DecodeLib lib = [] {
        DecodeLib d;
        for (const char *name : {"libbz2.so.1.0", "libbz2.so.1", "libbz2.so"}) {
            if (void *h = dlopen(name, RTLD_NOW | RTLD_LOCAL)) {
                d.init = reinterpret_cast<decltype(d.init)>(dlsym(h, "BZ2_bzDecompressInit"));
                d.step = reinterpret_cast<decltype(d.step)>(dlsym(h, "BZ2_bzDecompress"));
                d.end = reinterpret_cast<decltype(d.end)>(dlsym(h, "BZ2_bzDecompressEnd"));
                if (d.ok()) break;
            }
        }
        return d;
    }();
    return lib;
}

// serial decode of the first stream in [z, z+len): 0, or -2 (not a bzip2 stream / damaged / more than `limit` bytes)
inline int serial_decompress(const uint8_t *z, int64_t len, RawBuf &out, int64_t limit = INT64_MAX)
{
    const DecodeLib &lib = libbz2_decode();
    if (!lib.ok()) return -2;
    BzStream st;
    memset(&st, 0, sizeof st);
    if (lib.init(&st, 0, 0) != 0) return -2;
    out.len = 0;
    out.reserve((size_t)std::max<int64_t>(1 << 16, len * 8));
    int64_t fed = 0;
    int rc = 0;
    for (;;) {
        if (st.avail_in == 0 && fed < len) {
            const int64_t take = std::min<int64_t>(len - fed, 1 << 30);
            st.next_in = const_cast<char *>(reinterpret_cast<const char *>(z + fed));
            st.avail_in = (unsigned)take;
            fed += take;
        }
        if (out.len == out.cap) out.reserve(out.cap * 2);
        const size_t room = std::min<size_t>(out.cap - out.len, 1u << 30);
        st.next_out = reinterpret_cast<char *>(out.p + out.len);
        st.avail_out = (unsigned)room;
        rc = lib.step(&st);
        out.len += room - st.avail_out;
        if ((int64_t)out.len > limit) {  // a section that decodes to more than its patch can use: not worth the memory
            rc = -4;
            break;
        }
        if (rc != 0) break;  // 4 = BZ_STREAM_END, negative = damaged
        if (st.avail_in == 0 && fed >= len && st.avail_out != 0) {  // input exhausted in mid-stream
            rc = -7;                                                   // BZ_UNEXPECTED_EOF
            break;
        }
    }
    lib.end(&st);
    return rc == 4 ? 0 : -2;
}

struct Magics {
    uint16_t by_second_byte[256];  // bit s: block magic may start s bits into the byte before; bit 8+s: end magic
    Magics()
    {
        memset(by_second_byte, 0, sizeof by_second_byte);
        for (int s = 0; s < 8; ++s) {
            by_second_byte[(kBlockMagic >> (32 + s)) & 0xff] |= (uint16_t)(1u << s);
            by_second_byte[(kEndMagic >> (32 + s)) & 0xff] |= (uint16_t)(1u << (8 + s));
        }
    }
};

struct DecodeJob {
    const uint8_t *src;
    int64_t len;
    RawBuf out;                // the decoded section
    int64_t limit = INT64_MAX; // a section that decodes to more than this counts as damaged
    int blocks = 0;
    bool fell_back = false;
    int status = 0;            // 0, -2 damaged
};

struct DecodePiece {
    int stream;
    int64_t begin, end;        // bit range of the block in the section
    RawBuf data;               // decoded
    uint32_t crc = 0;
    bool ok = false;
};

// bit offsets of the blocks of a single, complete stream that fills [z, z+len) exactly; false = leave it to the serial path
inline bool find_blocks(const uint8_t *z, int64_t len, std::vector<int64_t> &blocks, int64_t &end_at)
{
    static const Magics magics;
    blocks.clear();
    if (len < 14 || z[0] != 'B' || z[1] != 'Z' || z[2] != 'h' || z[3] < '1' || z[3] > '9') return false;
    const int64_t bits = len * 8;
    end_at = -1;
    for (int64_t q = 5; q < len; ++q) {  // q: the byte after the one the magic starts in
        const uint16_t m = magics.by_second_byte[z[q]];
        if (!m) continue;
        for (int s = 0; s < 8; ++s) {
            const int64_t at = (q - 1) * 8 + s;
            if (at < 32 || at + 48 > bits) continue;
            if ((m >> s) & 1) {
                if (read_bits(z, at, 48) == kBlockMagic && at + 80 <= bits) blocks.push_back(at);
            }
            if ((m >> (8 + s)) & 1) {
                // the end magic, its CRC and the padding must close the section exactly
                if (read_bits(z, at, 48) == kEndMagic && (at + 80 + 7) / 8 == len) end_at = at;
            }
        }
    }
    if (end_at < 0) return false;
    while (!blocks.empty() && blocks.back() >= end_at) blocks.pop_back();
    if (blocks.empty()) return end_at == 32;   // an empty stream has no block
    return blocks.front() == 32;
}

// Decode every job; threads 0 = the CPUs this process may run on.  Returns 0, -2 (some section damaged), -3 (no libbz2).
inline int decompress_streams(DecodeJob *jobs, int count, int threads)
{
    const DecodeLib &lib = libbz2_decode();
    if (!lib.ok()) return -3;
    const int crew = threads > 0 ? std::min(threads, 256) : (int)usable_cpus();
    std::vector<DecodePiece> pieces;
    std::vector<size_t> first_piece(count, 0);
    std::vector<int64_t> end_at(count, -1);
    std::vector<int64_t> blocks;
    for (int s = 0; s < count; ++s) {
        DecodeJob &job = jobs[s];
        job.out.len = 0;
        job.blocks = 0;
        job.status = 0;
        job.fell_back = !find_blocks(job.src, job.len, blocks, end_at[s]);
        first_piece[s] = pieces.size();
        if (job.fell_back) {  // the whole section is one serial piece
            DecodePiece pc;
            pc.stream = s;
            pc.begin = pc.end = -1;
            pieces.push_back(std::move(pc));
            continue;
        }
        job.blocks = (int)blocks.size();
        for (size_t k = 0; k < blocks.size(); ++k) {
            DecodePiece pc;
            pc.stream = s;
            pc.begin = blocks[k];
            pc.end = k + 1 < blocks.size() ? blocks[k + 1] : end_at[s];
            pieces.push_back(std::move(pc));
        }
    }
    std::atomic<size_t> next{0};
    auto work = [&] {
        std::vector<uint8_t> wrapped;
        for (;;) {
            const size_t k = next.fetch_add(1);
            if (k >= pieces.size()) return;
            DecodePiece &pc = pieces[k];
            const DecodeJob &job = jobs[pc.stream];
            try {
                if (pc.begin < 0) {
                    pc.ok = serial_decompress(job.src, job.len, pc.data, job.limit) == 0;
                    continue;
                }
                pc.crc = (uint32_t)read_bits(job.src, pc.begin + 48, 32);
                wrapped.resize((size_t)((pc.end - pc.begin + 7) / 8 + 4 + 10 + 1));
                BitWriter w(wrapped.data(), (int64_t)wrapped.size());
                bool fits = w.put('B', 8) && w.put('Z', 8) && w.put('h', 8) && w.put((uint64_t)job.src[3], 8) &&
                            w.append(job.src, pc.begin, pc.end) && w.put(kEndMagic >> 24, 24) &&
                            w.put(kEndMagic & 0xffffffu, 24) && w.put(pc.crc, 32) && w.finish();
                if (!fits) continue;
                // the decoded size is not known in advance (a block of zeros expands 51x): the streaming decoder grows
                // its output as it goes
                pc.ok = serial_decompress(wrapped.data(), w.size(), pc.data, job.limit) == 0;
            } catch (...) {
                pc.ok = false;
            }
        }
    };
    {
        std::vector<std::thread> pool;
        try {
            const size_t helpers = std::min<size_t>((size_t)crew, pieces.size());
            for (size_t t = 1; t < helpers; ++t) pool.emplace_back(work);
        } catch (...) {
        }
        work();
        for (auto &t : pool) t.join();
    }
    int worst = 0;
    for (int s = 0; s < count; ++s) {
        DecodeJob &job = jobs[s];
        const size_t first = first_piece[s];
        if (job.fell_back) {
            job.status = pieces[first].ok ? 0 : -2;
            if (pieces[first].ok) job.out.swap(pieces[first].data);
        } else {
            bool all_ok = true;
            uint32_t combined = 0;
            size_t total = 0;
            for (int k = 0; k < job.blocks; ++k) {
                const DecodePiece &pc = pieces[first + k];
                all_ok = all_ok && pc.ok;
                combined = ((combined << 1) | (combined >> 31)) ^ pc.crc;
                total += pc.data.size();
            }
            if (all_ok && (int64_t)total > job.limit) {
                job.status = -2;
            } else if (all_ok && combined == (uint32_t)read_bits(job.src, end_at[s] + 48, 32)) {
                if (job.blocks == 1) {
                    job.out.swap(pieces[first].data);
                } else {
                    job.out.reserve(std::max<size_t>(total, 1));
                    job.out.len = total;
                    size_t at = 0;
                    for (int k = 0; k < job.blocks; ++k) {
                        const DecodePiece &pc = pieces[first + k];
                        if (!pc.data.empty()) memcpy(job.out.p + at, pc.data.data(), pc.data.size());
                        at += pc.data.size();
                    }
                }
            } else {
                // a candidate block was not a block, or the stream is damaged: the serial decoder decides
                job.fell_back = true;
                job.status = serial_decompress(job.src, job.len, job.out, job.limit) == 0 ? 0 : -2;
            }
        }
        worst = std::min(worst, job.status);
    }
    return worst;
}

}  // namespace bz2host
}  // namespace dq
