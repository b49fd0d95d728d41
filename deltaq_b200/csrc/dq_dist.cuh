// dq_dist.cuh -- device side of the multi-GPU ("group") suffix sort: one text sorted by all the GPUs of a context
// created with ndev > 1 (include/deltaq_cuda.h, dq_cuda_create).  Host side: dq_group.inl.
//
// The reference has no analogue (it is single threaded, /root/reference/src/DeltaQ.SuffixSorting.LibDivSufSort/
// DivSufSort.cs:18-42); this is SURVEY.md section 8(e): distributed prefix doubling.
//   * SA is partitioned by KEY: shard d owns the suffixes whose 8-byte key falls in its splitter interval, i.e. a
//     contiguous range of SA slots.  Groups never cross shards, so every sort after the first exchange is local.
//   * ISA is partitioned by text POSITION: shard o owns ISA[o << kb, (o+1) << kb).
// Every exchange is a stable partition by destination whose scatter writes straight into the destination GPU's
// memory over NVLink (peer pointers in the policy of the onesweep pass, dq_radix.cuh): partition and all-to-all-v are
// ONE kernel, the receive offsets are known before it starts, and nothing is staged or re-packed on either side.
//   round 0   (key, position) tuples  -> the bucket owner's sort input                     BucketPolicy
//   round r   requests ISA[sa+h]      -> the position owner's inbox (4 B each); the owner answers into the requester's
//             reply array in request order, so the requester never gathers                 RequestPolicy / reply_kernel
//   round r   (position, new rank)    -> the position owner's inbox (8 B each), applied there  UpdatePolicy / apply_kernel
#pragma once
#include "dq_common.cuh"
#include "dq_radix.cuh"
#include "dq_suffix.cuh"

namespace dq {
namespace dist {

constexpr int kMaxShards = 16;
constexpr uint32_t kPastEnd = 0xffffffffu;  // request marker: sa + h is past the end of the text, the answer is 0

struct Splitters {
    uint64_t s[kMaxShards];  // s[0..n): ascending; shard d owns the keys in [s[d-1], s[d])
    int n;
};

__device__ __forceinline__ uint32_t dest_of(const Splitters &sp, uint64_t key)
{
    uint32_t d = 0;
#pragma unroll
    for (int j = 0; j < kMaxShards - 1; ++j)
        if (j < sp.n) d += key >= sp.s[j] ? 1u : 0u;
    return d;
}

// ---- policies of the partition + exchange passes ------------------------------------------------------------
struct BucketPolicy {
    static constexpr bool kHasVal = true;
    static constexpr bool kDigitFromVal = false;
    Splitters sp;
    uint64_t *kout[kMaxShards];  // shard d's round-0 sort input (peer memory)
    uint32_t *vout[kMaxShards];
    int bits;
    __device__ __forceinline__ uint32_t digit(uint64_t key) const { return dest_of(sp, key); }
    __device__ __forceinline__ int nbits() const { return bits; }
    // gbase[d] of this pass = where this source's run starts inside shard d's input, so dst indexes it directly
    __device__ __forceinline__ void store(uint32_t d, uint32_t dst, uint64_t key, uint32_t val) const
    {
        kout[d][dst] = key;
        vout[d][dst] = val;
    }
};

// The exchanges by POSITION (requests and updates) partition finer than by owner: the digit is the owner followed by
// the next `sb` bits of the position, so that every owner receives its run ordered by sub-range of its ISA slice.  The
// owner then gathers / scatters one sub-range at a time -- a window that fits the L2 -- instead of hitting its whole
// slice at random (at 256 MiB of text per GPU that is the difference between DRAM-sector-bound and streaming).  The
// pass costs the same whatever the number of digits.
//
// ---- key skew (SURVEY hard part H7) -------------------------------------------------------------------------------
// Equal keys travel together, so one heavy key unbalances the buckets -- and an executable's zero padding makes the
// key 00^8 fifteen per cent of all suffixes, all of which would then go through the run-aware rounds on ONE GPU.  For a
// run-heavy text the buckets are therefore cut on the composite (key, run code): the code is the round-1 key of a
// suffix inside an equal-byte run (class and run length, dq_suffix.cuh), which is its true order among the suffixes
// with the same repeated-byte key, and 0 for every other suffix.  A shard then holds one slice of a heavy key's
// suffixes by run length; its round-0 group of them is a refinement consistent with the final order, and the groups of
// round 1 (same class, same run length) still never cross shards.
struct RunSplitters {
    uint64_t key[kMaxShards];
    uint32_t code[kMaxShards];
    int n;
};

__device__ __forceinline__ uint32_t run_code(const uint8_t *__restrict__ T, const uint32_t *__restrict__ run_end,
                                             uint32_t n, uint64_t key, uint32_t pos)
{
    if (key != (key & 0xffull) * 0x0101010101010101ull || (uint64_t)pos + 8 > n) return 0u;  // (short suffixes sort first)
    const uint32_t end = run_end[pos];
    const uint32_t R = end - pos;
    const bool below = end >= n || T[end] < (uint32_t)(key & 0xffu);
    return below ? R : (0x80000000u | (0x7fffffffu - R));
}

struct RunBucketPolicy {
    static constexpr bool kHasVal = true;
    static constexpr bool kDigitFromVal = true;
    RunSplitters sp;
    const uint8_t *T;         // the whole text and its run ends, on this shard
    const uint32_t *run_end;
    uint32_t n;
    uint64_t *kout[kMaxShards];
    uint32_t *vout[kMaxShards];
    int bits;
    __device__ __forceinline__ uint32_t digit(uint64_t key, uint32_t pos) const
    {
        const uint32_t code = run_code(T, run_end, n, key, pos);
        uint32_t d = 0;
#pragma unroll
        for (int j = 0; j < kMaxShards - 1; ++j)
            if (j < sp.n) d += (key > sp.key[j] || (key == sp.key[j] && code >= sp.code[j])) ? 1u : 0u;
        return d;
    }
    __device__ __forceinline__ int nbits() const { return bits; }
    __device__ __forceinline__ void store(uint32_t d, uint32_t dst, uint64_t key, uint32_t val) const
    {
        kout[d][dst] = key;
        vout[d][dst] = val;
    }
};

// active element = (rank << 32 | sa)
struct RequestPolicy {
    static constexpr bool kHasVal = false;
    static constexpr bool kDigitFromVal = false;
    uint64_t *kout;              // local: the active set regrouped by position owner (and sub-range)
    uint32_t *qout[kMaxShards];  // this source's region of owner d's request inbox (peer memory)
    const uint32_t *gbase;       // local exclusive digit offsets (the pass's own gbase)
    uint64_t h;
    uint32_t n;
    int kb;                      // owner = position >> kb
    int sb;                      // sub-range bits: digit = position >> (kb - sb)
    uint32_t self;
    int bits;
    __device__ __forceinline__ uint32_t digit(uint64_t key) const
    {
        const uint64_t q = (uint64_t)(uint32_t)key + h;
        return q < n ? (uint32_t)(q >> (kb - sb)) : (self << sb);
    }
    __device__ __forceinline__ int nbits() const { return bits; }
    __device__ __forceinline__ void store(uint32_t d, uint32_t dst, uint64_t key, uint32_t) const
    {
        kout[dst] = key;
        const uint64_t q = (uint64_t)(uint32_t)key + h;
        const uint32_t o = d >> sb;
        qout[o][dst - gbase[o << sb]] = q < n ? (uint32_t)(q - ((uint64_t)o << kb)) : kPastEnd;
    }
};

// update = (new rank << 32 | position); the owner receives (new rank << 32 | position - first owned position)
struct UpdatePolicy {
    static constexpr bool kHasVal = false;
    static constexpr bool kDigitFromVal = false;
    uint64_t *uout[kMaxShards];  // this source's region of owner d's update inbox (peer memory)
    const uint32_t *gbase;
    int kb;
    int sb;
    int bits;
    __device__ __forceinline__ uint32_t digit(uint64_t key) const { return (uint32_t)key >> (kb - sb); }
    __device__ __forceinline__ int nbits() const { return bits; }
    __device__ __forceinline__ void store(uint32_t d, uint32_t dst, uint64_t key, uint32_t) const
    {
        const uint32_t o = d >> sb;
        uout[o][dst - gbase[o << sb]] = key - ((uint64_t)o << kb);
    }
};

// counts of the policy's digits over keys[0..count) -> ghist[0..256) (zero before the launch).  Neighbouring keys
// often share a digit, so equal digits inside a warp are merged before they touch the shared counters.
template <typename Policy>
__global__ void __launch_bounds__(256)
hist_policy_kernel(const uint64_t *__restrict__ keys, uint32_t count, const Policy pol, uint32_t *__restrict__ ghist,
                   const uint32_t *__restrict__ vals = nullptr)
{
    __shared__ uint32_t sh[radix::kRadix];
    sh[threadIdx.x] = 0;
    __syncthreads();
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t rounds = (count + stride - 1) / stride;
    for (uint64_t r = 0; r < rounds; ++r) {
        const uint64_t k = r * stride + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
        const bool valid = k < count;
        const uint32_t d = valid ? radix::digit_of(pol, keys[k], vals ? vals[k] : 0u) : 0xffffu;
        const unsigned peers = __match_any_sync(kFullMask, d);
        if (valid && (peers & lanemask_lt()) == 0) atomicAdd(&sh[d], (uint32_t)__popc(peers));
    }
    __syncthreads();
    if (sh[threadIdx.x]) atomicAdd(&ghist[threadIdx.x], sh[threadIdx.x]);
}

// what a destination needs to know about one source's run in its inbox
struct RunMeta {
    uint32_t count;  // entries this source sent
    uint32_t base;   // where the run sits in the source's own regrouped array (requests: where the answers go)
};

struct MetaPtrs {
    RunMeta *p[kMaxShards];  // shard d's meta array (peer memory), indexed by source
};
struct ReplyPtrs {
    uint32_t *p[kMaxShards];  // shard s's reply array (peer memory)
};

// after the digit scan: tell every destination d how many entries this source sends it (the digits of owner d are
// [d << sb, (d + 1) << sb); total = all entries of the pass).  One warp.
__global__ void publish_meta_kernel(const uint32_t *__restrict__ gbase, uint32_t total, const MetaPtrs meta_of,
                                    uint32_t self, uint32_t shards, int sb)
{
    const unsigned d = threadIdx.x;
    if (d < shards) {
        const uint32_t lo = gbase[d << sb];
        const uint32_t hi = ((d + 1u) << sb) < (uint32_t)radix::kRadix ? gbase[(d + 1u) << sb] : total;
        meta_of.p[d][self] = RunMeta{hi - lo, lo};
    }
}

// sampled keys of a slice, for the splitters: element = a fixed odd multiplier walk over the slice's key array
__global__ void __launch_bounds__(256)
sample_keys_kernel(const uint64_t *__restrict__ keys, uint32_t count, uint32_t nsamples, uint64_t *__restrict__ out)
{
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nsamples) return;
    out[j] = keys[(uint32_t)(((uint64_t)j * 0x9E3779B1ull + 12345u) % count)];
}

// the same samples with their run codes (RunBucketPolicy)
__global__ void __launch_bounds__(256)
sample_pairs_kernel(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ vals, uint32_t count,
                    uint32_t nsamples, const uint8_t *__restrict__ T, const uint32_t *__restrict__ run_end, uint32_t n,
                    uint64_t *__restrict__ out_keys, uint32_t *__restrict__ out_codes)
{
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nsamples) return;
    const uint32_t k = (uint32_t)(((uint64_t)j * 0x9E3779B1ull + 12345u) % count);
    const uint64_t key = keys[k];
    out_keys[j] = key;
    out_codes[j] = run_code(T, run_end, n, key, vals[k]);
}

// owner side of the ISA fetch: for every source s, answer its requests in order, straight into the requester's
// reply array (peer memory): reply_of[s][meta[s].base + j] = ISA[q_j] + 1, or 0 past the end of the text.
__global__ void __launch_bounds__(256)
reply_kernel(const uint32_t *__restrict__ inbox, uint32_t cap, const RunMeta *__restrict__ meta,
             const uint32_t *__restrict__ ISA, const ReplyPtrs reply_of, uint32_t shards)
{
    for (uint32_t s = 0; s < shards; ++s) {
        const RunMeta m = meta[s];
        const uint32_t *q = inbox + (size_t)s * cap;
        uint32_t *out = reply_of.p[s] + m.base;
        for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < m.count; j += (uint64_t)gridDim.x * blockDim.x) {
            const uint32_t p = ld_stream(q + j);
            out[j] = p == kPastEnd ? 0u : __ldg(ISA + p) + 1u;
        }
    }
}

// owner side of the rank updates: ISA[position] = rank for every entry of every source's run
__global__ void __launch_bounds__(256)
apply_kernel(const uint64_t *__restrict__ inbox, uint32_t cap, const RunMeta *__restrict__ meta,
             uint32_t *__restrict__ ISA, uint32_t shards)
{
    for (uint32_t s = 0; s < shards; ++s) {
        const uint32_t cnt = meta[s].count;
        const uint64_t *u = inbox + (size_t)s * cap;
        for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < cnt; j += (uint64_t)gridDim.x * blockDim.x) {
            const uint64_t v = ld_stream(u + j);
            ISA[(uint32_t)v] = (uint32_t)(v >> 32);
        }
    }
}

// ---- small rounds: straight peer accesses ------------------------------------------------------------------
// Once few suffixes are unresolved, a round's exchanges cost more in launches and stream ordering than in bytes.  Such
// rounds read ISA[sa + h] with ordinary loads from the owner's slice (peer memory, NVLink) while building the keys, and
// the rank kernel stores changed ranks into the owner's slice the same way (rank_compact_kernel, DIST with isa_parts).
struct IsaParts {
    uint32_t *p[kMaxShards];  // shard o's ISA slice (peer memory): ISA[pos] = p[pos >> kb][pos & ((1 << kb) - 1)]
    int kb;
};

// depth: per position of the unresolved set, the bytes its group shares (equal-byte runs, dq_suffix.cuh); null: all h
__global__ void __launch_bounds__(suffix::kPackThreads)
build_keys_peer_kernel(const uint64_t *__restrict__ act, uint32_t a, const IsaParts isa, uint32_t n, uint64_t h,
                       uint64_t *__restrict__ keys, uint32_t *__restrict__ vals, radix::PassPlan plan,
                       uint32_t *__restrict__ ghist, const uint32_t *__restrict__ depth = nullptr)
{
    DQ_DYN_SMEM(smem);
    uint32_t *sh = reinterpret_cast<uint32_t *>(smem);
    for (int i = threadIdx.x; i < plan.npass * radix::kRadix; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    const uint32_t mask = (1u << isa.kb) - 1u;
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < a; k += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t e = act[k];
        const uint64_t q = (uint64_t)(uint32_t)e + (depth ? (uint64_t)depth[k] : h);
        uint32_t r2 = 0;
        if (q < n) r2 = isa.p[(uint32_t)q >> isa.kb][(uint32_t)q & mask] + 1u;
        const uint64_t key = (e & 0xffffffff00000000ull) | r2;
        keys[k] = key;
        vals[k] = (uint32_t)e;
        radix::hist_accumulate(sh, plan, key);
    }
    __syncthreads();
    radix::hist_flush(sh, plan.npass, ghist);
}

// Round 1 of a text full of equal-byte runs (build_keys_round1_kernel of the one-GPU path, for a device group): groups
// whose 8-byte key is one repeated byte are refined by run length -- from this shard's own copy of the text and its
// run ends, no rank is fetched -- every other group by ISA[sa + 8] + 1 read from the owner's slice.
__global__ void __launch_bounds__(suffix::kPackThreads)
build_keys_round1_peer_kernel(const uint64_t *__restrict__ act, uint32_t a, const IsaParts isa,
                              const uint8_t *__restrict__ T, const uint32_t *__restrict__ run_end, uint32_t n,
                              uint64_t *__restrict__ keys, uint32_t *__restrict__ vals, uint32_t *__restrict__ depth,
                              radix::PassPlan plan, uint32_t *__restrict__ ghist)
{
    DQ_DYN_SMEM(smem);
    uint32_t *sh = reinterpret_cast<uint32_t *>(smem);
    for (int i = threadIdx.x; i < plan.npass * radix::kRadix; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    const uint32_t mask = (1u << isa.kb) - 1u;
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < a; k += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t e = act[k];
        const uint32_t s = (uint32_t)e;
        const uint64_t k8 = suffix::load_key8(T, s);
        uint32_t r2, d;
        if ((uint64_t)s + 8 <= n && k8 == (k8 & 0xffull) * 0x0101010101010101ull) {
            const uint32_t end = run_end[s];
            const uint32_t R = end - s;
            const uint32_t bb = (uint32_t)(k8 & 0xffu);
            const bool below = end >= n || T[end] < bb;
            r2 = below ? R : (0x80000000u | (0x7fffffffu - R));
            d = suffix::kDepthFromKey;
        } else {
            const uint64_t q = (uint64_t)s + 8;
            r2 = q < n ? isa.p[(uint32_t)q >> isa.kb][(uint32_t)q & mask] + 1u : 0u;
            d = 8;
        }
        const uint64_t key = (e & 0xffffffff00000000ull) | r2;
        keys[k] = key;
        vals[k] = s;
        depth[k] = d;
        radix::hist_accumulate(sh, plan, key);
    }
    __syncthreads();
    radix::hist_flush(sh, plan.npass, ghist);
}

// key[k] = rank << 32 | reply[k], val[k] = sa, from the regrouped active set and the owners' answers; histograms of
// the round's digits on the way (as build_keys_kernel does on one GPU)
__global__ void __launch_bounds__(suffix::kPackThreads)
build_keys_reply_kernel(const uint64_t *__restrict__ act, const uint32_t *__restrict__ reply, uint32_t a,
                        uint64_t *__restrict__ keys, uint32_t *__restrict__ vals, radix::PassPlan plan,
                        uint32_t *__restrict__ ghist)
{
    DQ_DYN_SMEM(smem);
    uint32_t *sh = reinterpret_cast<uint32_t *>(smem);
    for (int i = threadIdx.x; i < plan.npass * radix::kRadix; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < a; k += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t e = act[k];
        const uint64_t key = (e & 0xffffffff00000000ull) | reply[k];
        keys[k] = key;
        vals[k] = (uint32_t)e;
        radix::hist_accumulate(sh, plan, key);
    }
    __syncthreads();
    radix::hist_flush(sh, plan.npass, ghist);
}

}  // namespace dist
}  // namespace dq
