// dq_search.cuh -- the bsdiff match search: Diff.Search for EVERY scan position of `new`, bit-exact.
//
// Reference: /root/reference/src/DeltaQ.BsDiff/Diff.cs:267-298 (Search), :245-246 (CompareBytes =
// Span.SequenceCompareTo), :249-265 (MatchLength), call site :106 with start = 0, end = n, I[n] == 0.
//
// The reference's result is a pure function of L = #{old suffixes < query}: the leaf it
// reaches is (start, end) = (max(L,1)-1, max(L,1)), and it returns the longer of the two matches there
// (ties -> end).  A literal per-position replay costs O(match length) per probe, i.e. quadratic work inside
// unchanged regions that the reference's greedy loop hops over (scan += len, Diff.cs:104).  These kernels
// compute the same L -- and the same two match lengths -- for all positions with amortised O(1) probes:
//
//   * if the query at j matched old suffix p for l bytes, the query at j+s matches suffix p+s for exactly
//     l-s bytes and lies on the same side of it, so rank ISA[p+s] is an exact anchor for j+s;
//   * from an anchor, L is found by walking the suffix array with the LCP array of `old`
//     (LCP[r] = lcp(SA[r-1], SA[r])): a neighbour that shares less than the anchor does with the query is
//     decided without touching the text, one that shares more is skipped (64- and 4096-rank block minima
//     skip whole blocks), and only an exact tie extends the byte comparison from where it stopped;
//   * chains of `kChunk` consecutive positions are walked by one thread; their heads are produced by a
//     stride-kChunk chain, and only every kChunk^2-th position is searched from scratch (binary search with
//     Manber-Myers lcp skipping).  The LCP array itself is built by the same two-level chaining (Kasai's
//     PLCP[i+1] >= PLCP[i]-1 along the text).
#pragma once
#include "dq_common.cuh"

namespace dq {
namespace search {

constexpr int kChunk = 64;              // positions per chain
constexpr int kSuper = kChunk * kChunk; // positions per from-scratch search
constexpr int kThreads = 128;
constexpr uint32_t kNone = 0xffffffffu;
constexpr int kBlk1 = 64, kBlk2 = 4096; // LCP block-minimum levels

// 8 bytes at an arbitrary address; the buffers are padded so that reading up to 15 bytes past p is safe
__device__ __forceinline__ uint64_t load64u(const uint8_t *p)
{
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint64_t *w = reinterpret_cast<const uint64_t *>(a & ~(uintptr_t)7);
    const unsigned sh = (unsigned)(a & 7u) * 8u;
    const uint64_t lo = __ldg(w);
    if (sh == 0) return lo;
    const uint64_t hi = __ldg(w + 1);
    return (lo >> sh) | (hi << (64u - sh));
}

// number of equal leading bytes of a[0..la) and b[0..lb)
__device__ __forceinline__ uint32_t common_prefix(const uint8_t *a, uint32_t la, const uint8_t *b, uint32_t lb)
{
    const uint32_t lim = min(la, lb);
    uint32_t k = 0;
    while (k < lim) {
        const uint64_t x = load64u(a + k) ^ load64u(b + k);
        if (x) {
            k += (uint32_t)(__ffsll((long long)x) - 1) >> 3;
            return min(k, lim);
        }
        k += 8;
    }
    return lim;
}

struct Texts {
    const uint8_t *old_;
    const uint8_t *new_;
    uint32_t n, m;
};

// lcp of old suffix p with the query at j, given that the first `known` bytes agree; *less = (suffix < query)
// under Span.SequenceCompareTo: first differing byte, else the shorter one is smaller.
__device__ __forceinline__ uint32_t match_from(const Texts &t, uint32_t p, uint32_t j, uint32_t known, bool *less)
{
    const uint32_t la = t.n - p, lq = t.m - j;
    const uint32_t c = known + common_prefix(t.old_ + p + known, la - known, t.new_ + j + known, lq - known);
    if (c == la)
        *less = c < lq;
    else if (c == lq)
        *less = false;
    else
        *less = t.old_[p + c] < t.new_[j + c];
    return c;
}

struct Index {
    const int32_t *SA;    // n
    const uint32_t *ISA;  // n
    const uint32_t *LCP;  // n, LCP[0] = 0
    const uint32_t *min1; // ceil(n/64)   block minima of LCP
    const uint32_t *min2; // ceil(n/4096)
};

struct Bracket {
    uint32_t L;  // #{old suffixes < query}
    uint32_t x;  // lcp(query, suffix SA[L-1])   (valid when L > 0)
    uint32_t y;  // lcp(query, suffix SA[L])     (valid when L < n)
};

// binary search from scratch, with Manber-Myers skipping of the bytes both ends are known to share
__device__ __forceinline__ Bracket locate_scratch(const Texts &t, const Index &ix, uint32_t j)
{
    int64_t lo = -1, hi = t.n;
    uint32_t llo = 0, lhi = 0;
    while (hi - lo > 1) {
        const uint32_t mid = (uint32_t)((lo + hi) >> 1);
        bool less;
        const uint32_t c = match_from(t, (uint32_t)ix.SA[mid], j, min(llo, lhi), &less);
        if (less) {
            lo = mid;
            llo = c;
        } else {
            hi = mid;
            lhi = c;
        }
    }
    return Bracket{(uint32_t)hi, llo, lhi};
}

// anchor: rank r whose suffix shares exactly c bytes with the query and is (less ? < : >=) the query
__device__ __forceinline__ Bracket locate_anchor(const Texts &t, const Index &ix, uint32_t j, uint32_t r, uint32_t c,
                                                 bool less)
{
    const uint32_t n = t.n;
    if (less) {
        uint32_t lo = r, llo = c, u = r + 1;
        for (;;) {
            if (u >= n) return Bracket{n, llo, 0};
            if ((u & (kBlk2 - 1)) == 0 && u + kBlk2 <= n && ix.min2[u / kBlk2] > llo) {
                lo = u + kBlk2 - 1;
                u += kBlk2;
                continue;
            }
            if ((u & (kBlk1 - 1)) == 0 && u + kBlk1 <= n && ix.min1[u / kBlk1] > llo) {
                lo = u + kBlk1 - 1;
                u += kBlk1;
                continue;
            }
            const uint32_t g = ix.LCP[u];
            if (g > llo) {
                lo = u++;
                continue;
            }
            if (g < llo) return Bracket{u, llo, g};
            bool ls;
            const uint32_t cc = match_from(t, (uint32_t)ix.SA[u], j, llo, &ls);
            if (!ls) return Bracket{u, llo, cc};
            lo = u++;
            llo = cc;
        }
    } else {
        uint32_t hi = r, lhi = c;
        for (;;) {
            if (hi == 0) return Bracket{0, 0, lhi};
            // LCP[hi-k+1 .. hi] all > lhi  =>  suffixes hi-k .. hi-1 are >= query with the same lcp
            if ((hi & (kBlk2 - 1)) == kBlk2 - 1 && ix.min2[hi / kBlk2] > lhi) {
                hi -= kBlk2;
                continue;
            }
            if ((hi & (kBlk1 - 1)) == kBlk1 - 1 && ix.min1[hi / kBlk1] > lhi) {
                hi -= kBlk1;
                continue;
            }
            const uint32_t g = ix.LCP[hi];
            if (g > lhi) {
                --hi;
                continue;
            }
            if (g < lhi) return Bracket{hi, g, lhi};
            bool ls;
            const uint32_t cc = match_from(t, (uint32_t)ix.SA[hi - 1], j, lhi, &ls);
            if (ls) return Bracket{hi, cc, lhi};
            --hi;
            lhi = cc;
        }
    }
}

// chain state carried from one position to the next: matched old suffix, match length, side
struct Carry {
    uint32_t p, l;
    bool less;
};

__device__ __forceinline__ Bracket locate_step(const Texts &t, const Index &ix, uint32_t j, const Carry &cy,
                                               uint32_t stride, bool have)
{
    if (have && cy.l > stride) return locate_anchor(t, ix, j, ix.ISA[cy.p + stride], cy.l - stride, cy.less);
    return locate_scratch(t, ix, j);
}

__device__ __forceinline__ Carry carry_of(const Texts &t, const Index &ix, const Bracket &b)
{
    const bool hx = b.L > 0, hy = b.L < t.n;
    if (hx && (!hy || b.x >= b.y)) return Carry{(uint32_t)ix.SA[b.L - 1], b.x, true};
    if (hy) return Carry{(uint32_t)ix.SA[b.L], b.y, false};
    return Carry{0, 0, false};  // n == 0
}

// what Diff.Search returns for this bracket (Diff.cs:271-286), including the I[n] == 0 leaf
__device__ __forceinline__ void reference_result(const Texts &t, const Index &ix, uint32_t j, const Bracket &b,
                                                 int32_t *pos, int32_t *len)
{
    const uint32_t n = t.n;
    if (n == 0) {
        *pos = 0;
        *len = 0;
        return;
    }
    uint32_t ps, pe, x, y;
    if (b.L == 0) {  // leaf (0, 1)
        ps = (uint32_t)ix.SA[0];
        x = b.y;
        if (n == 1) {
            pe = 0;  // I[1] == I[n] == 0, the same suffix again
            y = x;
        } else {
            pe = (uint32_t)ix.SA[1];
            const uint32_t g = ix.LCP[1];
            if (g < x)
                y = g;
            else if (g > x)
                y = x;
            else {
                bool ls;
                y = match_from(t, pe, j, x, &ls);
            }
        }
    } else if (b.L == n) {  // leaf (n-1, n): I[n] == 0 -> the whole of `old`
        ps = (uint32_t)ix.SA[n - 1];
        x = b.x;
        pe = 0;
        bool ls;
        y = match_from(t, 0, j, 0, &ls);
    } else {
        ps = (uint32_t)ix.SA[b.L - 1];
        pe = (uint32_t)ix.SA[b.L];
        x = b.x;
        y = b.y;
    }
    if (x > y) {
        *pos = (int32_t)ps;
        *len = (int32_t)x;
    } else {
        *pos = (int32_t)pe;
        *len = (int32_t)y;
    }
}

// ---- kernels ---------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(256) invert_sa_kernel(const int32_t *__restrict__ SA, uint32_t n,
                                                         uint32_t *__restrict__ ISA)
{
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (uint64_t)gridDim.x * blockDim.x)
        ISA[SA[r]] = (uint32_t)r;
}

// LCP array, level A: one thread per kSuper text positions walks the chunk heads with stride kChunk.
// head_l[i / kChunk] = PLCP[i] for i % kChunk == 0.
__global__ void __launch_bounds__(kThreads)
lcp_heads_kernel(const uint8_t *__restrict__ T, uint32_t n, const int32_t *__restrict__ SA,
                 const uint32_t *__restrict__ ISA, uint32_t *__restrict__ head_l)
{
    const uint64_t sc = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t i0 = sc * kSuper;
    if (i0 >= n) return;
    uint32_t l = 0;
    for (int k = 0; k < kChunk; ++k) {
        const uint64_t i64 = i0 + (uint64_t)k * kChunk;
        if (i64 >= n) break;
        const uint32_t i = (uint32_t)i64;
        const uint32_t r = ISA[i];
        if (r == 0) {
            l = 0;
        } else {
            const uint32_t q = (uint32_t)SA[r - 1];
            const uint32_t known = l > (uint32_t)kChunk ? l - kChunk : 0;
            l = known + common_prefix(T + i + known, n - i - known, T + q + known, n - q - known);
        }
        head_l[i / kChunk] = l;
    }
}

// LCP array, level B: one thread per chunk, stride 1.  LCP[ISA[i]] = PLCP[i].
__global__ void __launch_bounds__(kThreads)
lcp_chain_kernel(const uint8_t *__restrict__ T, uint32_t n, const int32_t *__restrict__ SA,
                 const uint32_t *__restrict__ ISA, const uint32_t *__restrict__ head_l, uint32_t *__restrict__ LCP)
{
    const uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t i0 = c * kChunk;
    if (i0 >= n) return;
    uint32_t l = head_l[c];
    LCP[ISA[i0]] = l;
    for (int k = 1; k < kChunk; ++k) {
        const uint64_t i64 = i0 + k;
        if (i64 >= n) break;
        const uint32_t i = (uint32_t)i64;
        const uint32_t r = ISA[i];
        if (r == 0) {
            l = 0;
        } else {
            const uint32_t q = (uint32_t)SA[r - 1];
            const uint32_t known = l > 0 ? l - 1 : 0;
            l = known + common_prefix(T + i + known, n - i - known, T + q + known, n - q - known);
        }
        LCP[r] = l;
    }
}

// block minima: out[b] = min(in[b*W .. b*W+W)) ; one warp per block of W entries
__global__ void __launch_bounds__(256) block_min_kernel(const uint32_t *__restrict__ in, uint32_t count, uint32_t W,
                                                         uint32_t *__restrict__ out, uint32_t nblocks)
{
    const uint32_t warps_per_cta = blockDim.x / 32;
    for (uint64_t b = (uint64_t)blockIdx.x * warps_per_cta + warp_id(); b < nblocks;
         b += (uint64_t)gridDim.x * warps_per_cta) {
        uint32_t mn = 0xffffffffu;
        const uint64_t lo = b * W;
        for (uint64_t k = lo + lane_id(); k < lo + W && k < count; k += 32) mn = min(mn, in[k]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mn = min(mn, __shfl_xor_sync(kFullMask, mn, o));
        if (lane_id() == 0) out[b] = mn;
    }
}

// search, level A: chunk heads of [scan_begin, scan_begin+count), one thread per kSuper positions.
// head_p / head_l hold the carry of each head (bit 31 of head_l = less).
__global__ void __launch_bounds__(kThreads)
search_heads_kernel(Texts t, Index ix, uint32_t scan_begin, uint32_t count, uint32_t *__restrict__ head_p,
                    uint32_t *__restrict__ head_l)
{
    const uint64_t sc = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t k0 = sc * kSuper;
    if (k0 >= count || t.n == 0) return;
    Carry cy{0, 0, false};
    bool have = false;
    for (int k = 0; k < kChunk; ++k) {
        const uint64_t kk = k0 + (uint64_t)k * kChunk;
        if (kk >= count) break;
        const uint32_t j = scan_begin + (uint32_t)kk;
        const Bracket b = locate_step(t, ix, j, cy, kChunk, have);
        cy = carry_of(t, ix, b);
        have = true;
        head_p[kk / kChunk] = cy.p;
        head_l[kk / kChunk] = cy.l | (cy.less ? 0x80000000u : 0u);
    }
}

// search, level B: one thread per chunk; writes the reference's (pos, len) for every position.
__global__ void __launch_bounds__(kThreads)
search_chain_kernel(Texts t, Index ix, uint32_t scan_begin, uint32_t count, const uint32_t *__restrict__ head_p,
                    const uint32_t *__restrict__ head_l, int32_t *__restrict__ pos_out, int32_t *__restrict__ len_out)
{
    const uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t k0 = c * kChunk;
    if (k0 >= count) return;
    if (t.n == 0) {
        for (int k = 0; k < kChunk && k0 + k < count; ++k) {
            pos_out[k0 + k] = 0;
            len_out[k0 + k] = 0;
        }
        return;
    }
    // the head's carry describes the head itself: re-derive its bracket from it (stride 0)
    const uint32_t hl = head_l[c];
    Carry cy{head_p[c], hl & 0x7fffffffu, (hl >> 31) != 0};
    for (int k = 0; k < kChunk; ++k) {
        const uint64_t kk = k0 + k;
        if (kk >= count) break;
        const uint32_t j = scan_begin + (uint32_t)kk;
        Bracket b;
        if (k == 0)
            b = locate_anchor(t, ix, j, ix.ISA[cy.p], cy.l, cy.less);
        else
            b = locate_step(t, ix, j, cy, 1, true);
        int32_t pos, len;
        reference_result(t, ix, j, b, &pos, &len);
        pos_out[kk] = pos;
        len_out[kk] = len;
        cy = carry_of(t, ix, b);
    }
}

}  // namespace search
}  // namespace dq
