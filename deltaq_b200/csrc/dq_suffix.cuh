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
constexpr int kRankThreads = 256;
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
template <bool ROUND0>
__global__ void __launch_bounds__(kRankThreads)
rank_compact_kernel(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ sa,
                    const uint32_t *__restrict__ slot_in, uint32_t a, uint32_t n, uint32_t *__restrict__ ISA,
                    int32_t *__restrict__ SA, uint32_t *__restrict__ sa_out, uint32_t *__restrict__ rank_out,
                    uint32_t *__restrict__ slot_out, uint64_t *__restrict__ lb, uint32_t *__restrict__ tile_ticket,
                    uint32_t *__restrict__ count_out)
{
    __shared__ uint32_t s_tile;
    __shared__ uint32_t s_wmax[kRankWarps], s_wsum[kRankWarps];
    __shared__ uint64_t s_excl;

    const unsigned tid = threadIdx.x, lane = lane_id(), warp = warp_id();
    if (tid == 0) s_tile = atomicAdd(tile_ticket, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint32_t k0 = tile * (uint32_t)kRankTile + tid * kRankItems;

    // head[j] for j in [0, kRankItems]: does a group start at k0 + j?  (positions >= a count as heads)
    uint64_t key[kRankItems + 1];
    uint32_t s[kRankItems], sl[kRankItems];
    bool head[kRankItems + 1];
    uint64_t prev_key = 0;
    uint32_t prev_sa = 0;
    if (k0 > 0 && k0 <= a) {
        prev_key = keys[k0 - 1];
        if (ROUND0) prev_sa = sa[k0 - 1];
    }
#pragma unroll
    for (int j = 0; j <= kRankItems; ++j) key[j] = (k0 + j < a) ? keys[k0 + j] : 0ull;
#pragma unroll
    for (int j = 0; j < kRankItems; ++j) {
        s[j] = (k0 + j < a) ? sa[k0 + j] : 0u;
        sl[j] = ROUND0 ? (k0 + j) : ((k0 + j < a) ? slot_in[k0 + j] : 0u);
    }
#pragma unroll
    for (int j = 0; j <= kRankItems; ++j) {
        const uint32_t k = k0 + j;
        bool hd;
        if (k >= a || k == 0) {
            hd = true;
        } else {
            const uint64_t pk = j == 0 ? prev_key : key[j - 1];
            hd = key[j] != pk;
            if (ROUND0) {
                const uint32_t ps = j == 0 ? prev_sa : s[j - 1];
                hd = hd || (n - ps < 8u);  // the previous suffix is shorter than the key: it stands alone
            }
        }
        head[j] = hd;
    }

    // thread-local: running "slot of the last head" and survivor count
    uint32_t tmax = 0, tsum = 0;
#pragma unroll
    for (int j = 0; j < kRankItems; ++j) {
        if (k0 + j < a) {
            if (head[j]) tmax = sl[j];
            if (!(head[j] && head[j + 1])) tsum++;
        }
    }
    // block scan (max, sum)
    uint32_t imax = tmax, isum = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t m = __shfl_up_sync(kFullMask, imax, o);
        uint32_t c = __shfl_up_sync(kFullMask, isum, o);
        if (lane >= (unsigned)o) {
            imax = max(imax, m);
            isum += c;
        }
    }
    if (lane == 31) {
        s_wmax[warp] = imax;
        s_wsum[warp] = isum;
    }
    __syncthreads();
    uint32_t wmax = 0, wsum = 0, bmax = 0, bsum = 0;
#pragma unroll
    for (int w = 0; w < kRankWarps; ++w) {
        if ((unsigned)w < warp) {
            wmax = max(wmax, s_wmax[w]);
            wsum += s_wsum[w];
        }
        bmax = max(bmax, s_wmax[w]);
        bsum += s_wsum[w];
    }
    // exclusive prefix of this thread inside the tile
    uint32_t emax = __shfl_up_sync(kFullMask, imax, 1);
    uint32_t esum = __shfl_up_sync(kFullMask, isum, 1);
    if (lane == 0) {
        emax = 0;
        esum = 0;
    }
    emax = max(emax, wmax);
    esum += wsum;

    // tile prefix by decoupled look-back (thread 0)
    if (tid == 0) {
        const uint64_t agg = rk_pack(bmax, bsum);
        uint64_t excl = 0;
        if (tile > 0) {
            st_desc(lb + tile, ((uint64_t)radix::kStatusAggregate << 62) | agg);
            uint32_t t = tile - 1;
            uint32_t xm = 0, xs = 0;
            for (;;) {
                uint64_t v = ld_desc(lb + t);
                unsigned st = (unsigned)(v >> 62);
                if (st == 0) {
                    DQ_SPIN_HINT();
                    continue;
                }
                xm = max(xm, rk_max(v));
                xs += rk_sum(v);
                if (st == radix::kStatusInclusive) break;
                --t;
            }
            excl = rk_pack(xm, xs);
        }
        const uint64_t incl = rk_pack(max(rk_max(excl), bmax), rk_sum(excl) + bsum);
        st_desc(lb + tile, ((uint64_t)radix::kStatusInclusive << 62) | incl);
        s_excl = excl;
        if ((uint64_t)(tile + 1) * kRankTile >= a) *count_out = rk_sum(incl);
    }
    __syncthreads();
    uint32_t run_max = max(emax, rk_max(s_excl));
    uint32_t out = esum + rk_sum(s_excl);

#pragma unroll
    for (int j = 0; j < kRankItems; ++j) {
        if (k0 + j < a) {
            if (head[j]) run_max = sl[j];
            const uint32_t nr = run_max;
            if (head[j] && head[j + 1]) {
                SA[sl[j]] = (int32_t)s[j];
                ISA[s[j]] = nr;
            } else {
                if (ROUND0 || nr != (uint32_t)(key[j] >> 32)) ISA[s[j]] = nr;
                sa_out[out] = s[j];
                rank_out[out] = nr;
                slot_out[out] = sl[j];
                out++;
            }
        }
    }
}

}  // namespace suffix
}  // namespace dq
