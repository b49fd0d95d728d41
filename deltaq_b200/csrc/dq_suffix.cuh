// dq_suffix.cuh -- kernels of the prefix-doubling suffix sorter (everything except the radix passes).
//
// Data model (all device resident, n = text length, a = number of still-unresolved suffixes):
//   T[n(+pad)]      text, zero padded by >= 16 bytes
//   ISA[n]          rank of every suffix = SA slot of the head of its group (final once the group is a singleton)
//   SA[n]           output; slot s is written exactly once, when the suffix occupying it becomes a singleton
//   active set      three parallel arrays of length a, in SA order with resolved slots squeezed out:
//                     sa[k]    suffix start,  rank[k] current group rank (== ISA[sa[k]]),
//                     slot[k]  the SA slot that position k of the active set stands for
//   A group is a maximal run of equal ranks; its members are contiguous in the active set and own the
//   contiguous SA slots [rank, rank + size).  Sorting the active set by (rank, ISA[sa+h]+1) therefore only
//   permutes suffixes inside groups and `slot` stays put.
//
// Round 0 sorts all n suffixes by their first 8 bytes (big-endian packed, zero padded).  The elements enter
// the stable sort in DESCENDING suffix order, so inside a run of equal keys the (at most 7) suffixes shorter
// than 8 bytes -- which can only tie with longer ones through zero padding -- come first, shortest first,
// which is exactly their final order ("a proper prefix sorts first": Span.SequenceCompareTo,
// /root/reference/test/DeltaQ.SuffixSorting.LibDivSufSort.Tests/LibDivSufSortTests.cs:46-59).  The rank
// kernel makes each of them a singleton.  From then on every non-singleton group agrees on its first h
// bytes and all members are at least h long, so at most one member has sa + h == n and key half 0 is free
// for it.
//
// Replaces (by result) DivSufSort.sort_typeBstar / construct_SA
// (/root/reference/src/DeltaQ.SuffixSorting.LibDivSufSort/DivSufSort.cs:186-511, :44-153) and
// SAIS.sais_main (/root/reference/src/DeltaQ.SuffixSorting.SAIS/SAIS.cs:280-495).
#pragma once
#include "dq_common.cuh"
#include "dq_radix.cuh"

namespace dq {
namespace suffix {

constexpr int kPackThreads = 256;
constexpr int kPackItems = 8;

// big-endian value of T[i..i+8), reading three aligned little-endian words (T is zero padded)
__device__ __forceinline__ uint64_t load_key8(const uint8_t *__restrict__ T, uint32_t i)
{
    const uint32_t *w = reinterpret_cast<const uint32_t *>(T) + (i >> 2);
    const uint32_t w0 = __ldg(w), w1 = __ldg(w + 1), w2 = __ldg(w + 2);
    const unsigned sh = (i & 3u) * 8u;
    const uint32_t a = __funnelshift_r(w0, w1, sh);  // T[i..i+4) little-endian
    const uint32_t b = __funnelshift_r(w1, w2, sh);  // T[i+4..i+8)
    return ((uint64_t)__byte_perm(a, 0, 0x0123) << 32) | (uint64_t)__byte_perm(b, 0, 0x0123);
}

// K0: element k stands for suffix i = n-1-k.  keys[k] = first 8 bytes, vals[k] = i; digit histograms of all
// 8 passes are accumulated on the way.  grid-stride; dynamic smem = npass*256*4.
__global__ void __launch_bounds__(kPackThreads)
pack_keys_kernel(const uint8_t *__restrict__ T, uint32_t n, uint64_t *__restrict__ keys,
                 uint32_t *__restrict__ vals, radix::PassPlan plan, uint32_t *__restrict__ ghist)
{
    DQ_DYN_SMEM(smem);
    uint32_t *sh = reinterpret_cast<uint32_t *>(smem);
    for (int i = threadIdx.x; i < plan.npass * radix::kRadix; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    const uint32_t chunk = kPackThreads * kPackItems;
    for (uint64_t base = (uint64_t)blockIdx.x * chunk; base < n; base += (uint64_t)gridDim.x * chunk) {
#pragma unroll
        for (int j = 0; j < kPackItems; ++j) {
            uint64_t k = base + (uint64_t)j * kPackThreads + threadIdx.x;
            if (k < n) {
                const uint32_t i = n - 1u - (uint32_t)k;
                const uint64_t key = load_key8(T, i);
                keys[k] = key;
                vals[k] = i;
                radix::hist_accumulate(sh, plan, key);
            }
        }
    }
    __syncthreads();
    radix::hist_flush(sh, plan.npass, ghist);
}

// K6: key[k] = rank[k] << 32 | (sa[k]+h < n ? ISA[sa[k]+h] + 1 : 0), plus the digit histograms.
__global__ void __launch_bounds__(kPackThreads)
build_keys_kernel(const uint32_t *__restrict__ sa, const uint32_t *__restrict__ rank,
                  const uint32_t *__restrict__ ISA, uint32_t n, uint32_t a, uint64_t h,
                  uint64_t *__restrict__ keys, radix::PassPlan plan, uint32_t *__restrict__ ghist)
{
    DQ_DYN_SMEM(smem);
    uint32_t *sh = reinterpret_cast<uint32_t *>(smem);
    for (int i = threadIdx.x; i < plan.npass * radix::kRadix; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    const uint32_t chunk = kPackThreads * kPackItems;
    for (uint64_t base = (uint64_t)blockIdx.x * chunk; base < a; base += (uint64_t)gridDim.x * chunk) {
        uint32_t s[kPackItems], r[kPackItems], r2[kPackItems];
#pragma unroll
        for (int j = 0; j < kPackItems; ++j) {
            uint64_t k = base + (uint64_t)j * kPackThreads + threadIdx.x;
            s[j] = k < a ? sa[k] : 0u;
            r[j] = k < a ? rank[k] : 0u;
        }
#pragma unroll
        for (int j = 0; j < kPackItems; ++j) {
            uint64_t k = base + (uint64_t)j * kPackThreads + threadIdx.x;
            uint64_t p = (uint64_t)s[j] + h;
            r2[j] = (k < a && p < n) ? __ldg(ISA + p) + 1u : 0u;
        }
#pragma unroll
        for (int j = 0; j < kPackItems; ++j) {
            uint64_t k = base + (uint64_t)j * kPackThreads + threadIdx.x;
            if (k < a) {
                const uint64_t key = ((uint64_t)r[j] << 32) | r2[j];
                keys[k] = key;
                radix::hist_accumulate(sh, plan, key);
            }
        }
    }
    __syncthreads();
    radix::hist_flush(sh, plan.npass, ghist);
}

// K6 for the multi-GPU path: the second key half was fetched from the position owners (r2[k], already +1 / 0)
__global__ void __launch_bounds__(kPackThreads)
build_keys_r2_kernel(const uint32_t *__restrict__ rank, const uint32_t *__restrict__ r2, uint32_t a,
                     uint64_t *__restrict__ keys, radix::PassPlan plan, uint32_t *__restrict__ ghist)
{
    DQ_DYN_SMEM(smem);
    uint32_t *sh = reinterpret_cast<uint32_t *>(smem);
    for (int i = threadIdx.x; i < plan.npass * radix::kRadix; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < a; k += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t key = ((uint64_t)rank[k] << 32) | r2[k];
        keys[k] = key;
        radix::hist_accumulate(sh, plan, key);
    }
    __syncthreads();
    radix::hist_flush(sh, plan.npass, ghist);
}

// K0 for a slice of text positions [pos_begin, pos_begin + count): element k stands for suffix
// pos_begin + count - 1 - k (descending, as in pack_keys_kernel).  T points at the slice (T[0] is text position
// pos_begin), zero padded.  hist16 (65536 counters, may be null) receives the histogram of the keys' top 16 bits
// (warp-aggregated atomics), from which the ranks agree on bucket splitters.
__global__ void __launch_bounds__(kPackThreads)
pack_slice_kernel(const uint8_t *__restrict__ T, uint32_t pos_begin, uint32_t count, uint64_t *__restrict__ keys,
                  uint32_t *__restrict__ vals, unsigned long long *__restrict__ hist16)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t rounds = (count + stride - 1) / stride;
    for (uint64_t r = 0; r < rounds; ++r) {
        const uint64_t k = r * stride + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
        const bool valid = k < count;
        uint32_t bin = 0xffffffffu;
        if (valid) {
            const uint32_t off = count - 1u - (uint32_t)k;
            const uint64_t key = load_key8(T, off);
            keys[k] = key;
            vals[k] = pos_begin + off;
            bin = (uint32_t)(key >> 48);
        }
        if (hist16) {
            const unsigned peers = __match_any_sync(kFullMask, bin);
            if (valid && (peers & lanemask_lt()) == 0) atomicAdd(&hist16[bin], (unsigned long long)__popc(peers));
        }
    }
}

// partition keys for the bucket exchange: dest[k] = lut[key >> 48] as a 64-bit radix key, idx[k] = k
__global__ void __launch_bounds__(256)
dest_keys_kernel(const uint64_t *__restrict__ keys, uint32_t count, const uint8_t *__restrict__ lut,
                 uint64_t *__restrict__ dest, uint32_t *__restrict__ idx)
{
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < count; k += (uint64_t)gridDim.x * blockDim.x) {
        dest[k] = lut[keys[k] >> 48];
        idx[k] = (uint32_t)k;
    }
}

// out[i] = in[perm[i]] for the (key, value) pairs
__global__ void __launch_bounds__(256)
gather_pairs_kernel(const uint64_t *__restrict__ kin, const uint32_t *__restrict__ vin, const uint32_t *__restrict__ perm,
                    uint32_t count, uint64_t *__restrict__ kout, uint32_t *__restrict__ vout)
{
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t p = perm[i];
        kout[i] = kin[p];
        vout[i] = vin[p];
    }
}

// ISA requests of the unresolved set: q[k] = sa[k] + h as a 64-bit radix key, idx[k] = k
__global__ void __launch_bounds__(256)
requests_kernel(const uint32_t *__restrict__ sa, uint32_t a, uint64_t h, uint64_t *__restrict__ q, uint32_t *__restrict__ idx)
{
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < a; k += (uint64_t)gridDim.x * blockDim.x) {
        q[k] = (uint64_t)sa[k] + h;
        idx[k] = (uint32_t)k;
    }
}

// digit histograms of an existing key array (used by dq_cuda_radix_sort_pairs only; the suffix sorter's
// producers accumulate their histograms while they write the keys)
__global__ void __launch_bounds__(kPackThreads)
hist_only_kernel(const uint64_t *__restrict__ keys, uint32_t count, radix::PassPlan plan, uint32_t *__restrict__ ghist)
{
    DQ_DYN_SMEM(smem);
    uint32_t *sh = reinterpret_cast<uint32_t *>(smem);
    for (int i = threadIdx.x; i < plan.npass * radix::kRadix; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < count; k += (uint64_t)gridDim.x * blockDim.x)
        radix::hist_accumulate(sh, plan, keys[k]);
    __syncthreads();
    radix::hist_flush(sh, plan.npass, ghist);
}

// ---- K4+K5: group heads, new ranks, singleton retirement, compaction -- one pass, decoupled look-back ----
constexpr int kRankThreads = 512;
constexpr int kRankItems = 8;
constexpr int kRankTile = kRankThreads * kRankItems;
constexpr int kRankWarps = kRankThreads / 32;

// descriptor: status(2) | last head's slot (31) | survivors (31)
constexpr uint64_t kField31 = 0x7fffffffull;
__host__ __device__ __forceinline__ uint64_t rk_pack(uint32_t mx, uint32_t sm) { return ((uint64_t)mx << 31) | sm; }
__host__ __device__ __forceinline__ uint32_t rk_max(uint64_t v) { return (uint32_t)((v >> 31) & kField31); }
__host__ __device__ __forceinline__ uint32_t rk_sum(uint64_t v) { return (uint32_t)(v & kField31); }

// keys/sa: the sorted active set (a entries).  slot_in == nullptr in round 0 (slot[k] = k).
// Outputs: ISA, SA, the compacted next active set (sa_out, rank_out, slot_out) and *count_out = its size.
//
// Warp-striped: warp w of the tile owns 32*kRankItems consecutive positions, row j of lane l is position
// base + j*32 + l, so every load and the compacted stores are coalesced.  Head flags and survivor flags live
// as one ballot per row (uniform registers), which turns both scans into bit arithmetic:
//   "slot of the last head at or before me" = shfl from the highest head bit at or below my lane,
//   "survivors before me"                   = popc of the survivor ballot below my lane.
//
// DIST (multi-GPU, deltaq_b200/parallel.py): this GPU sorts one key bucket that owns the SA slots
// [slot_base, slot_base + bucket size); ISA is partitioned by text position across the GPUs, so instead of
// writing ISA the kernel emits one (position, rank) update per element (upd_pos/upd_rank, dense, in sorted
// order) for the host side to route to the owners, and SA is the bucket-local array (index slot - slot_base).
template <bool ROUND0, bool DIST = false>
__global__ void __launch_bounds__(kRankThreads)
rank_compact_kernel(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ sa,
                    const uint32_t *__restrict__ slot_in, uint32_t a, uint32_t n, uint32_t *__restrict__ ISA,
                    int32_t *__restrict__ SA, uint32_t *__restrict__ sa_out, uint32_t *__restrict__ rank_out,
                    uint32_t *__restrict__ slot_out, uint64_t *__restrict__ lb, uint32_t *__restrict__ tile_ticket,
                    uint32_t *__restrict__ count_out, uint32_t slot_base = 0, uint64_t *__restrict__ upd_pos = nullptr,
                    uint32_t *__restrict__ upd_rank = nullptr)
{
    __shared__ uint32_t s_tile;
    __shared__ uint32_t s_wmax[kRankWarps], s_wsum[kRankWarps];
    __shared__ uint64_t s_excl;

    const unsigned tid = threadIdx.x, lane = lane_id(), warp = warp_id();
    if (tid == 0) s_tile = atomicAdd(tile_ticket, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint32_t wbase = tile * (uint32_t)kRankTile + warp * (32u * kRankItems);

    uint64_t key[kRankItems];
    uint32_t s[kRankItems], sl[kRankItems];
    unsigned hb[kRankItems + 1];  // hb[j] = ballot of "a group starts here" over row j; positions >= a count as heads
    unsigned vb[kRankItems];      // ballot of k < a
#pragma unroll
    for (int j = 0; j < kRankItems; ++j) {
        const uint32_t k = wbase + j * 32 + lane;
        const bool valid = k < a;
        key[j] = valid ? keys[k] : 0ull;
        s[j] = valid ? sa[k] : 0u;
        sl[j] = ROUND0 ? slot_base + k : (valid ? slot_in[k] : 0u);
    }
#pragma unroll
    for (int j = 0; j < kRankItems; ++j) {
        const uint32_t k = wbase + j * 32 + lane;
        const bool valid = k < a;
        // previous element: the lane to the left, or (lane 0) a reload of the element before the row
        uint64_t pk = __shfl_up_sync(kFullMask, key[j], 1);
        uint32_t ps = __shfl_up_sync(kFullMask, s[j], 1);
        bool hd = true;
        if (valid && k > 0) {
            if (lane == 0) {
                pk = keys[k - 1];
                if (ROUND0) ps = sa[k - 1];
            }
            hd = key[j] != pk;
            if (ROUND0) hd = hd || (n - ps < 8u);  // the previous suffix is shorter than the key: it stands alone
        }
        hb[j] = __ballot_sync(kFullMask, hd);
        vb[j] = __ballot_sync(kFullMask, valid);
    }
    {   // does a group start right after this warp's last position?
        const uint32_t k = wbase + 32u * kRankItems;
        bool hd = true;
        if (lane == 31 && k < a) {
            hd = keys[k] != key[kRankItems - 1];
            if (ROUND0) hd = hd || (n - s[kRankItems - 1] < 8u);
        }
        hb[kRankItems] = __shfl_sync(kFullMask, hd ? 1u : 0u, 31);
    }

    // survivor ballots and the warp's aggregates
    unsigned sb[kRankItems];
    uint32_t wsum = 0, wmax = 0;
#pragma unroll
    for (int j = 0; j < kRankItems; ++j) {
        const unsigned next = (hb[j] >> 1) | ((hb[j + 1] & 1u) << 31);  // head flag of the element after each lane
        sb[j] = vb[j] & ~(hb[j] & next);
        wsum += __popc(sb[j]);
        const unsigned hm = hb[j] & vb[j];
        const uint32_t last = __shfl_sync(kFullMask, sl[j], hm ? 31 - __clz((int)hm) : 0);
        if (hm) wmax = last;  // slots increase with position: the last head of the last row with one wins
    }
    if (lane == 0) {
        s_wmax[warp] = wmax;
        s_wsum[warp] = wsum;
    }
    __syncthreads();
    uint32_t emax = 0, esum = 0, bmax = 0, bsum = 0;
#pragma unroll
    for (int w = 0; w < kRankWarps; ++w) {
        if ((unsigned)w < warp) {
            emax = max(emax, s_wmax[w]);
            esum += s_wsum[w];
        }
        bmax = max(bmax, s_wmax[w]);
        bsum += s_wsum[w];
    }

    // tile prefix by decoupled look-back: warp 0 inspects 32 predecessors per round trip
    if (warp == 0) {
        const uint64_t agg = rk_pack(bmax, bsum);
        uint32_t xm = 0, xs = 0;
        if (tile > 0) {
            if (lane == 0) st_desc(lb + tile, ((uint64_t)radix::kStatusAggregate << 62) | agg);
            int64_t base = (int64_t)tile - 1;
            for (;;) {
                const int64_t t = base - lane;
                // before tile 0 there is nothing: a virtual inclusive descriptor of zero ends the walk
                const uint64_t v = t >= 0 ? ld_desc(lb + t) : ((uint64_t)radix::kStatusInclusive << 62);
                const unsigned st = (unsigned)(v >> 62);
                const unsigned ready = __ballot_sync(kFullMask, st != 0);
                const unsigned incl = __ballot_sync(kFullMask, st == radix::kStatusInclusive);
                // lanes up to (and including) the nearest inclusive predecessor must all be published
                const unsigned need = incl ? (0xffffffffu >> (31 - (__ffs((int)incl) - 1))) : 0xffffffffu;
                if ((ready & need) != need) {
                    DQ_SPIN_HINT();
                    continue;
                }
                const bool mine = (need >> lane) & 1u;
                xm = max(xm, __reduce_max_sync(kFullMask, mine ? rk_max(v) : 0u));
                xs += __reduce_add_sync(kFullMask, mine ? rk_sum(v) : 0u);
                if (incl) break;
                base -= 32;
            }
        }
        if (lane == 0) {
            const uint64_t excl = rk_pack(xm, xs);
            const uint64_t inc = rk_pack(max(xm, bmax), xs + bsum);
            st_desc(lb + tile, ((uint64_t)radix::kStatusInclusive << 62) | inc);
            s_excl = excl;
            if ((uint64_t)(tile + 1) * kRankTile >= a) *count_out = rk_sum(inc);
        }
    }
    __syncthreads();
    uint32_t carry = max(emax, rk_max(s_excl));  // slot of the last head before this warp's chunk
    uint32_t out = esum + rk_sum(s_excl);        // survivors before this warp's chunk

#pragma unroll
    for (int j = 0; j < kRankItems; ++j) {
        const unsigned hm = hb[j] & vb[j];
        const unsigned le = hm & (0xffffffffu >> (31 - lane));  // heads at or below my lane
        const uint32_t mine = __shfl_sync(kFullMask, sl[j], le ? 31 - __clz((int)le) : 0);
        const uint32_t nr = le ? mine : carry;
        const bool valid = (vb[j] >> lane) & 1u;
        if (valid) {
            if (DIST) {
                const uint32_t k = wbase + j * 32 + lane;
                upd_pos[k] = s[j];
                upd_rank[k] = nr;
            }
            if ((sb[j] >> lane) & 1u) {
                const uint32_t o = out + __popc(sb[j] & lanemask_lt());
                if (!DIST && (ROUND0 || nr != (uint32_t)(key[j] >> 32))) ISA[s[j]] = nr;
                sa_out[o] = s[j];
                rank_out[o] = nr;
                slot_out[o] = sl[j];
            } else {
                SA[sl[j] - slot_base] = (int32_t)s[j];
                if (!DIST) ISA[s[j]] = nr;
            }
        }
        out += __popc(sb[j]);
        if (hm) carry = __shfl_sync(kFullMask, sl[j], 31 - __clz((int)hm));
    }
}

}  // namespace suffix
}  // namespace dq
