// dq_segsort.cuh -- doubling rounds without the global radix passes, for the groups that fit one CTA (SURVEY.md hard
// part H9).
//
// In a round r >= 1 the unresolved set is already ordered by group: the key of position k is (rank << 32 | r2) and the
// ranks ascend with k.  Sorting by that key only permutes suffixes inside their group, yet the LSD radix sort moves
// every pair 7-8 times through HBM to do it (24 B per pair per pass).  Most groups are small.  Here one CTA loads a
// window of the key array into shared memory, sorts the groups that lie inside it with a bitonic network, and writes
// them back in place: one read and one write per pair.  Groups that do not fit a window stay untouched; they are
// compacted (order preserving), sorted by the ordinary passes, and scattered back to the positions they came from.
//
// Ownership: tile t owns the groups whose FIRST element lies in [kNominal*t, kNominal*(t+1)).  All of them but the last
// end before the next head inside that range, so they fit; the last one is taken if it ends inside the window of
// kWindow = 2*kNominal elements, else it is left to the global passes.  The decision is local to the owner, and the
// owner alone writes the group (also the part beyond its nominal range), so tiles never race.  Neighbouring tiles do read
// each other's keys to find group boundaries, but only the rank half, which sorting inside a group never changes.
#pragma once
#include "dq_common.cuh"

namespace dq {
namespace segsort {

constexpr int kNominal = 1024;
constexpr int kWindow = 2 * kNominal;
constexpr int kThreads = 256;
constexpr int kPer = kWindow / kThreads;  // 8 window slots per thread
constexpr uint32_t kNoIndex = 0xffffffffu;

__device__ __forceinline__ uint32_t rank_of(uint64_t key) { return (uint32_t)(key >> 32); }

// K/V: keys and values of the a unresolved positions, in group order; sorted in place inside every group this tile takes.
// done[k] = 1 for every position written here (zero before the launch).
__global__ void __launch_bounds__(kThreads)
tile_sort_kernel(uint64_t *__restrict__ K, uint32_t *__restrict__ V, uint32_t a, uint8_t *__restrict__ done)
{
    __shared__ uint64_t sk[kWindow];
    __shared__ uint32_t sv[kWindow];
    __shared__ uint32_t s_first, s_last, s_end;

    const unsigned tid = threadIdx.x;
    const uint64_t t0 = (uint64_t)blockIdx.x * kNominal;
    if (t0 >= a) return;
    const uint32_t avail = (uint32_t)min((uint64_t)kWindow, (uint64_t)a - t0);  // window slots that exist
    if (tid == 0) {
        s_first = kNoIndex;  // first head in the nominal range
        s_last = 0;          // last head in the nominal range (valid when s_first != kNoIndex)
        s_end = kNoIndex;    // first head after s_last inside the window (or the end of the array)
    }
    for (uint32_t i = tid; i < (uint32_t)kWindow; i += kThreads) sk[i] = i < avail ? K[t0 + i] : ~0ull;
    __syncthreads();
    // heads: position i starts a group when its rank differs from the rank before it
    const uint32_t rank_before = t0 > 0 ? rank_of(K[t0 - 1]) : 0u;
    bool head[kPer];
#pragma unroll
    for (int j = 0; j < kPer; ++j) {
        const uint32_t i = tid + j * kThreads;
        head[j] = false;
        if (i < avail) {
            const uint32_t r = rank_of(sk[i]);
            head[j] = (t0 + i == 0) || r != (i ? rank_of(sk[i - 1]) : rank_before);
            if (head[j] && i < (uint32_t)kNominal) {
                atomicMin(&s_first, i);
                atomicMax(&s_last, i);
            }
        }
    }
    __syncthreads();
    const uint32_t first = s_first;
    if (first == kNoIndex) return;  // every position of the nominal range belongs to a group that started earlier
    const uint32_t last = s_last;
#pragma unroll
    for (int j = 0; j < kPer; ++j) {
        const uint32_t i = tid + j * kThreads;
        if (i < avail && head[j] && i > last) atomicMin(&s_end, i);
    }
    if (tid == 0 && t0 + avail >= a) atomicMin(&s_end, avail);  // the array ends inside the window: so does the last group
    __syncthreads();
    // [first, x): the groups taken here.  The last group is taken only if its end is visible.
    const uint32_t x = s_end != kNoIndex ? s_end : last;
    if (x <= first) return;
    const uint32_t cnt = x - first;
    uint32_t N = 2;
    while (N < cnt) N <<= 1;

    // move the range to the front of the window, pad to a power of two with keys that sort last
    uint64_t tmp[kPer];
#pragma unroll
    for (int j = 0; j < kPer; ++j) {
        const uint32_t i = tid + j * kThreads;
        tmp[j] = i < cnt ? sk[first + i] : ~0ull;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kPer; ++j) {
        const uint32_t i = tid + j * kThreads;
        if (i < N) {
            sk[i] = tmp[j];
            sv[i] = i < cnt ? V[t0 + first + i] : 0u;
        }
    }
    __syncthreads();
    // bitonic sort of sk[0..N) ascending, values along
    for (uint32_t size = 2; size <= N; size <<= 1) {
        for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
            for (uint32_t t = tid; t < (N >> 1); t += kThreads) {
                const uint32_t lo = 2 * t - (t & (stride - 1));
                const uint32_t hi = lo + stride;
                const bool up = (lo & size) == 0;
                const uint64_t ka = sk[lo], kb = sk[hi];
                if (up ? (kb < ka) : (ka < kb)) {
                    sk[lo] = kb;
                    sk[hi] = ka;
                    const uint32_t va = sv[lo];
                    sv[lo] = sv[hi];
                    sv[hi] = va;
                }
            }
            __syncthreads();
        }
    }
    for (uint32_t i = tid; i < cnt; i += kThreads) {
        K[t0 + first + i] = sk[i];
        V[t0 + first + i] = sv[i];
        done[t0 + first + i] = 1;
    }
}

// ---- order-preserving compaction of the positions no tile took ---------------------------------------------------------
constexpr int kBlock = 2048;  // positions per block of the three kernels below (256 threads x 8)

__global__ void __launch_bounds__(kThreads)
count_left_kernel(const uint8_t *__restrict__ done, uint32_t a, uint32_t *__restrict__ counts)
{
    __shared__ uint32_t s_cnt;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    const uint64_t b0 = (uint64_t)blockIdx.x * kBlock;
    uint32_t c = 0;
#pragma unroll
    for (int j = 0; j < kBlock / kThreads; ++j) {
        const uint64_t k = b0 + threadIdx.x + (uint64_t)j * kThreads;
        c += (k < a && !done[k]) ? 1u : 0u;
    }
    c = __reduce_add_sync(kFullMask, c);
    if (lane_id() == 0 && c) atomicAdd(&s_cnt, c);
    __syncthreads();
    if (threadIdx.x == 0) counts[blockIdx.x] = s_cnt;
}

// exclusive scan of counts[0..nblk) in place, total to *total.  One block of 1024 threads.
__global__ void __launch_bounds__(1024) scan_counts_kernel(uint32_t *__restrict__ counts, uint32_t nblk,
                                                            uint32_t *__restrict__ total)
{
    __shared__ uint32_t s_part[1024];
    const uint32_t per = (nblk + 1023u) / 1024u;
    const uint32_t lo = min(nblk, threadIdx.x * per), hi = min(nblk, lo + per);
    uint32_t sum = 0;
    for (uint32_t i = lo; i < hi; ++i) sum += counts[i];
    s_part[threadIdx.x] = sum;
    __syncthreads();
    uint32_t before = 0;
    for (uint32_t w = 0; w < threadIdx.x; ++w) before += s_part[w];
    for (uint32_t i = lo; i < hi; ++i) {
        const uint32_t c = counts[i];
        counts[i] = before;
        before += c;
    }
    if (threadIdx.x == 1023) *total = before;
}

// the positions left, in order: keys, values and where they came from
__global__ void __launch_bounds__(kThreads)
compact_left_kernel(const uint64_t *__restrict__ K, const uint32_t *__restrict__ V, const uint8_t *__restrict__ done,
                    uint32_t a, const uint32_t *__restrict__ offs, uint64_t *__restrict__ Kc, uint32_t *__restrict__ Vc,
                    uint32_t *__restrict__ pos)
{
    __shared__ uint32_t s_warp[kThreads / 32];
    const uint64_t b0 = (uint64_t)blockIdx.x * kBlock;
    // thread t takes 8 consecutive positions, so the order inside the block is (thread, slot)
    const uint64_t k0 = b0 + (uint64_t)threadIdx.x * (kBlock / kThreads);
    bool left[kBlock / kThreads];
    uint32_t mine = 0;
#pragma unroll
    for (int j = 0; j < kBlock / kThreads; ++j) {
        const uint64_t k = k0 + j;
        left[j] = k < a && !done[k];
        mine += left[j] ? 1u : 0u;
    }
    uint32_t incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(kFullMask, incl, o);
        if (lane_id() >= (unsigned)o) incl += v;
    }
    if (lane_id() == 31) s_warp[warp_id()] = incl;
    __syncthreads();
    uint32_t o = offs[blockIdx.x] + incl - mine;
    for (unsigned w = 0; w < warp_id(); ++w) o += s_warp[w];
#pragma unroll
    for (int j = 0; j < kBlock / kThreads; ++j) {
        if (left[j]) {
            const uint64_t k = k0 + j;
            Kc[o] = K[k];
            Vc[o] = V[k];
            pos[o] = (uint32_t)k;
            ++o;
        }
    }
}

// the compacted pairs, sorted, go back to the positions they came from (both in ascending group order)
__global__ void __launch_bounds__(256)
scatter_back_kernel(const uint64_t *__restrict__ Kc, const uint32_t *__restrict__ Vc, const uint32_t *__restrict__ pos,
                    uint32_t count, uint64_t *__restrict__ K, uint32_t *__restrict__ V)
{
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t p = pos[i];
        K[p] = Kc[i];
        V[p] = Vc[i];
    }
}

}  // namespace segsort
}  // namespace dq
