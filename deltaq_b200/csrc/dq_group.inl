// dq_group.inl -- host side of the multi-GPU ("group") paths: one text sorted by all the GPUs of a context created
// with ndev > 1, and the match search sharded by new-data range over the replicated index.
// Included inside deltaq_cuda.cu's anonymous namespace.  Device side: dq_dist.cuh.
//
// One process drives every GPU of the group: shard s has its own dq_ctx (device, stream, scratch); one host thread
// enqueues on all streams.  GPUs exchange data only through kernels that write into peer memory (the partition
// passes of dq_dist.cuh) and are ordered by events recorded on one stream and waited for on the others -- no host
// round trip inside a phase, one host synchronisation per doubling round (the unresolved counts).
// The same device may be listed more than once: the shards are then logical, which is how a single-GPU box (and the
// CPU logic emulator of tests/emu) runs every line of this path.

namespace ds = dq::dist;

struct Shard {
    dq_ctx *c = nullptr;  // device, stream, scratch of this shard (shard 0: the group context itself)
    cudaEvent_t ev = nullptr;
    uint32_t own_begin = 0, own_cnt = 0;  // text positions whose ISA entries live here
    uint32_t cnt = 0, slot_base = 0;      // SA slots [slot_base, slot_base + cnt) are sorted here
    uint32_t a = 0;                       // unresolved suffixes entering the next round
    DevBuf slice, packed, isa_local, sa_local, upd, reply, inbox_req, inbox_upd, meta_req, meta_upd, samples;
    uint64_t *act = nullptr, *other = nullptr;  // packed unresolved set and the free 64-bit buffer
    uint32_t *slot_cur = nullptr, *slot_nxt = nullptr;
    uint32_t *depth_cur = nullptr, *depth_nxt = nullptr;  // run-aware rounds: bytes the group of each position shares
    uint32_t *h_small = nullptr;  // pinned: [0..16) digit counts, [16..18) rank kernel's counters
    uint64_t *h_samples = nullptr;  // pinned
};

struct GroupPhase {
    const char *name;
    double ms;
};

struct Group {
    std::vector<Shard> sh;
    int kb = 0;
    uint32_t n = 0;          // length of the text whose buckets are resident on the shards (0: none)
    bool replicated = false; // text/sa/isa of that text are complete on every shard's context
    uint32_t runend_n = 0;   // every shard's context holds the whole text of this length and its run ends (run-aware sort)
    bool trace = false;
    uint32_t shard_min = 128u << 20;  // inputs below this stay on shard 0 (DQ_SHARD_MIN overrides)
    uint64_t direct_max = 8u << 20;   // rounds with at most this many unresolved suffixes in all use straight peer
                                      // accesses instead of the request / reply / update exchanges (DQ_DIRECT_MAX)
    std::vector<GroupPhase> phases;
    std::chrono::steady_clock::time_point t_phase;
    // one host thread per shard for the per-shard stretches of a round (launches are cheap for the GPU and dear for one
    // host thread driving eight of them); null: the calling thread does everything (DQ_GROUP_THREADS=0, the emulator)
    std::unique_ptr<dq::diffhost::Crew> crew;
};

// fn(i) for every shard i, side by side when the group has its crew; the first failure is reported through `top`
template <typename F> int for_shards(dq_ctx *top, F &&fn)
{
    Group &g = *top->group;
    const size_t G = g.sh.size();
    std::vector<int> rc(G, DQ_OK);
    auto body = [&](int part) {
        const size_t i = (size_t)part;
        if (cudaSetDevice(g.sh[i].c->device) != cudaSuccess) {
            g.sh[i].c->err = "cudaSetDevice failed";
            rc[i] = DQ_ERR_CUDA;
            return;
        }
        rc[i] = fn(i);
    };
    if (g.crew)
        g.crew->run(body);
    else
        for (size_t i = 0; i < G; ++i) body((int)i);
    for (size_t i = 0; i < G; ++i)
        if (rc[i] != DQ_OK) {
            if (g.sh[i].c != top) top->err = g.sh[i].c->err;
            return rc[i];
        }
    return DQ_OK;
}

constexpr uint32_t kSamplesPerShard = 2048;

// a failing call on a shard's context reports through the group context
#define DQ_SUB(top, c, expr)                           \
    do {                                               \
        int rc_ = (expr);                              \
        if (rc_ != DQ_OK) {                            \
            if ((c) != (top)) (top)->err = (c)->err;   \
            return rc_;                                \
        }                                              \
    } while (0)

void destroy_group(dq_ctx *top)
{
    Group *g = top->group;
    if (!g) return;
    for (size_t i = 0; i < g->sh.size(); ++i) {
        Shard &s = g->sh[i];
        if (!s.c) continue;
        cudaSetDevice(s.c->device);
        cudaStreamSynchronize(s.c->stream);
        DevBuf *bufs[] = {&s.slice, &s.packed, &s.isa_local, &s.sa_local, &s.upd, &s.reply, &s.inbox_req, &s.inbox_upd,
                          &s.meta_req, &s.meta_upd, &s.samples};
        for (DevBuf *b : bufs)
            if (b->p) cudaFree(b->p);
        if (s.ev) cudaEventDestroy(s.ev);
        if (s.h_small) cudaFreeHost(s.h_small);
        if (s.h_samples) cudaFreeHost(s.h_samples);
        if (i > 0) destroy_single(s.c);
    }
    g->crew.reset();
    delete g;
    top->group = nullptr;
}

// devices[0] is the group context's own device (shard 0)
int create_group(dq_ctx *top, const int *devices, int ndev)
{
    Group *g = new (std::nothrow) Group();
    if (!g) return DQ_ERR_OUT_OF_MEMORY;
    top->group = g;
    if (const char *e = getenv("DQ_SHARD_MIN")) g->shard_min = (uint32_t)strtoul(e, nullptr, 10);
    if (const char *e = getenv("DQ_DIRECT_MAX")) g->direct_max = strtoull(e, nullptr, 10);
#ifndef DQ_EMU
    {
        const char *e = getenv("DQ_GROUP_THREADS");
        if (!e || atoi(e) != 0) g->crew.reset(new dq::diffhost::Crew(ndev - 1));
    }
#endif
    g->sh.resize((size_t)ndev);
    g->sh[0].c = top;
    for (int i = 1; i < ndev; ++i) {
        int rc = create_single(&g->sh[(size_t)i].c, devices[i]);
        if (rc != DQ_OK) {
            top->err = g_create_error;
            return rc;
        }
    }
    for (int i = 0; i < ndev; ++i) {
        Shard &s = g->sh[(size_t)i];
        DQ_CK(top, cudaSetDevice(s.c->device));
        DQ_CK(top, cudaEventCreateWithFlags(&s.ev, cudaEventDisableTiming));
        DQ_CK(top, cudaHostAlloc((void **)&s.h_small, 256, cudaHostAllocDefault));
        DQ_CK(top, cudaHostAlloc((void **)&s.h_samples, (size_t)kSamplesPerShard * 16, cudaHostAllocDefault));  // keys, then run codes
        for (int j = 0; j < ndev; ++j) {
            if (devices[j] == devices[i]) continue;
            int can = 0;
            DQ_CK(top, cudaDeviceCanAccessPeer(&can, devices[i], devices[j]));
            if (!can) {
                top->err = "dq_cuda_create: device " + std::to_string(devices[i]) + " cannot access device " +
                           std::to_string(devices[j]) + " (the group paths write into peer memory)";
                return DQ_ERR_CUDA;
            }
            cudaError_t e = cudaDeviceEnablePeerAccess(devices[j], 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled)
                (void)cudaGetLastError();
            else
                DQ_CK(top, e);
        }
    }
    DQ_CK(top, cudaSetDevice(top->device));
    return DQ_OK;
}

int group_sync(dq_ctx *top)
{
    for (Shard &s : top->group->sh) {
        DQ_CK(top, cudaSetDevice(s.c->device));
        DQ_CK(top, cudaStreamSynchronize(s.c->stream));
    }
    return DQ_OK;
}

// every shard's stream waits for everything enqueued so far on every other shard's stream: shard 0 waits for all the
// others and the others wait for shard 0 (3 (G - 1) + 1 stream operations instead of G (G - 1))
int group_barrier(dq_ctx *top)
{
    Group &g = *top->group;
    Shard &hub = g.sh[0];
    for (size_t i = 1; i < g.sh.size(); ++i) {
        Shard &s = g.sh[i];
        DQ_CK(top, cudaSetDevice(s.c->device));
        DQ_CK(top, cudaEventRecord(s.ev, s.c->stream));
        DQ_CK(top, cudaStreamWaitEvent(hub.c->stream, s.ev, 0));
    }
    DQ_CK(top, cudaSetDevice(hub.c->device));
    DQ_CK(top, cudaEventRecord(hub.ev, hub.c->stream));
    for (size_t i = 1; i < g.sh.size(); ++i) DQ_CK(top, cudaStreamWaitEvent(g.sh[i].c->stream, hub.ev, 0));
    return DQ_OK;
}

int group_mark(dq_ctx *top, const char *name)
{
    Group &g = *top->group;
    if (!g.trace) return DQ_OK;
    DQ_TRY(group_sync(top));
    const auto now = std::chrono::steady_clock::now();
    const double ms = std::chrono::duration<double, std::milli>(now - g.t_phase).count();
    g.t_phase = now;
    for (GroupPhase &p : g.phases)
        if (!strcmp(p.name, name)) {
            p.ms += ms;
            return DQ_OK;
        }
    g.phases.push_back(GroupPhase{name, ms});
    return DQ_OK;
}

uint32_t light_grid(const dq_ctx *c, uint64_t items)
{
    return (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(div_up(items, 256 * 4), (uint64_t)c->sm_count * 8));
}

// counts of pol's digits over keys[0..count), their exclusive scan, and the (count, base) of every run published to
// the destinations' meta arrays.  Leaves gbase/use_match in c->hist as run_passes does for pass 0.
template <typename Policy>
int digit_counts(Group &g, Shard &s, const uint64_t *keys, uint32_t count, const Policy &pol, bool publish,
                 bool to_requests, int sb = 0, const uint32_t *vals = nullptr)
{
    dq_ctx *c = s.c;
    DQ_TRY(zero_hist(c));
    uint32_t *ghist = c->hist.as<uint32_t>();
    uint32_t *gbase = ghist + rx::kMaxPasses * rx::kRadix;
    uint32_t *use_match = gbase + rx::kMaxPasses * rx::kRadix;
    if (count) {
        auto k = ds::hist_policy_kernel<Policy>;
        DQ_LAUNCH(k, light_grid(c, count), 256, 0, c->stream, keys, count, pol, ghist, vals);
    }
    // few digits (bucket partition): rank with MATCH; position digits spread over all 256 values: let the scan decide
    auto scan = rx::scan_hist_kernel;
    DQ_LAUNCH(scan, 1, rx::kRadix, 0, c->stream, ghist, gbase, use_match, count, sb > 0 ? 0u : 1u);
    c->stats.kernel_launches += 2;
    if (publish) {
        ds::MetaPtrs mp{};
        for (size_t d = 0; d < g.sh.size(); ++d)
            mp.p[d] = (to_requests ? g.sh[d].meta_req : g.sh[d].meta_upd).as<ds::RunMeta>();
        auto k = ds::publish_meta_kernel;
        DQ_LAUNCH(k, 1, 32, 0, c->stream, gbase, count, mp, (uint32_t)(&s - &g.sh[0]), (uint32_t)g.sh.size(), sb);
        c->stats.kernel_launches++;
    }
    DQ_CK(c, cudaGetLastError());
    return DQ_OK;
}

// one partition + exchange pass (dq_radix.cuh, onesweep_policy_kernel)
template <typename Policy>
int policy_pass(Shard &s, const uint64_t *kin, const uint32_t *vin, const Policy &pol, uint32_t count)
{
    if (count == 0) return DQ_OK;
    dq_ctx *c = s.c;
    uint32_t *gbase = c->hist.as<uint32_t>() + rx::kMaxPasses * rx::kRadix;
    uint32_t *use_match = gbase + rx::kMaxPasses * rx::kRadix;
    const uint32_t tiles = (uint32_t)div_up(count, rx::kTile);
    const size_t bytes = 256 + (size_t)tiles * rx::kRadix * 4;
    DQ_TRY(ensure(c, c->lb, bytes));
    uint8_t *lbp = c->lb.as<uint8_t>();
    DQ_CK(c, cudaMemsetAsync(lbp, 0, bytes, c->stream));
    auto k = rx::onesweep_policy_kernel<Policy>;
    DQ_LAUNCH(k, tiles, rx::kThreads, rx::pass_smem_bytes(), c->stream, kin, vin, pol, count, gbase,
              reinterpret_cast<uint32_t *>(lbp + 256), reinterpret_cast<uint32_t *>(lbp), use_match);
    c->stats.kernel_launches++;
    c->stats.radix_passes++;
    DQ_CK(c, cudaGetLastError());
    return DQ_OK;
}

int bits_for(size_t shards)
{
    int b = 1;
    while (((size_t)1 << b) < shards) ++b;
    return b;
}

// sub-range bits of the exchanges by position (dq_dist.cuh): owner digits are padded to a power of two, the rest of
// the 8 digit bits splits every owner's slice -- but never finer than 64 K positions (256 KiB of ISA) per sub-range
int sub_bits(size_t shards, int kb)
{
    const int sb = rx::kRadixBits - bits_for(shards);
    const char *e = getenv("DQ_SUB_MIN_LOG");  // tests: lets small texts use sub-ranges
    const int min_log = e ? atoi(e) : 16;
    return std::max(0, std::min(sb, kb - min_log));
}

// (rank << 32 | position) updates of every shard -> the position owners' ISA slices
int group_route_updates(dq_ctx *top, const std::vector<uint32_t> &counts)
{
    Group &g = *top->group;
    const size_t G = g.sh.size();
    uint32_t cap = 0;
    for (Shard &s : g.sh) cap = std::max(cap, s.cnt);
    const int sb = sub_bits(G, g.kb);
    DQ_TRY(for_shards(top, [&](size_t i) -> int {
        Shard &s = g.sh[i];
        ds::UpdatePolicy pol{};
        for (size_t d = 0; d < G; ++d) pol.uout[d] = g.sh[d].inbox_upd.as<uint64_t>() + (size_t)i * cap;
        pol.gbase = s.c->hist.as<uint32_t>() + rx::kMaxPasses * rx::kRadix;
        pol.kb = g.kb;
        pol.sb = sb;
        pol.bits = bits_for(G) + sb;
        DQ_TRY(digit_counts(g, s, s.upd.as<uint64_t>(), counts[i], pol, true, false, sb));
        return policy_pass(s, s.upd.as<uint64_t>(), nullptr, pol, counts[i]);
    }));
    DQ_TRY(group_barrier(top));
    return for_shards(top, [&](size_t i) -> int {
        Shard &s = g.sh[i];
        if (s.own_cnt == 0) return DQ_OK;
        auto k = ds::apply_kernel;
        DQ_LAUNCH(k, (uint32_t)s.c->sm_count * 8, 256, 0, s.c->stream, s.inbox_upd.as<uint64_t>(), cap,
                  s.meta_upd.as<ds::RunMeta>(), s.isa_local.as<uint32_t>(), (uint32_t)G);
        s.c->stats.kernel_launches++;
        DQ_CK(s.c, cudaGetLastError());
        return DQ_OK;
    });
}

// every shard gets the whole text (peer copies of the slices) and computes its run ends (dq_suffix.cuh, run_*_kernel)
int group_text_and_run_ends(dq_ctx *top, uint32_t n)
{
    Group &g = *top->group;
    DQ_TRY(group_barrier(top));
    DQ_TRY(for_shards(top, [&](size_t i) -> int {
        dq_ctx *c = g.sh[i].c;
        DQ_TRY(ensure(c, c->text, (size_t)n + 64));
        DQ_CK(c, cudaMemsetAsync(c->text.as<uint8_t>() + n, 0, 64, c->stream));
        for (Shard &s : g.sh)
            if (s.own_cnt)
                DQ_CK(c, cudaMemcpyAsync(c->text.as<uint8_t>() + s.own_begin, s.slice.p, s.own_cnt, cudaMemcpyDefault, c->stream));
        DQ_TRY(ensure(c, c->runend, (size_t)n * 4));
        const uint32_t ntiles = (uint32_t)div_up(n, sx::kRunTile);
        DQ_TRY(ensure(c, c->runtile, (size_t)ntiles * 8));
        uint32_t *tile_first = c->runtile.as<uint32_t>(), *next_after = tile_first + ntiles;
        auto k1 = sx::run_tile_first_kernel;
        DQ_LAUNCH(k1, ntiles, 256, 0, c->stream, c->text.as<uint8_t>(), n, tile_first);
        auto k2 = sx::run_tile_scan_kernel;
        DQ_LAUNCH(k2, 1, 1024, 0, c->stream, tile_first, ntiles, n, next_after);
        auto k3 = sx::run_end_kernel;
        DQ_LAUNCH(k3, ntiles, 256, 0, c->stream, c->text.as<uint8_t>(), n, next_after, c->runend.as<uint32_t>());
        c->stats.kernel_launches += 3;
        DQ_CK(c, cudaGetLastError());
        return DQ_OK;
    }));
    g.runend_n = n;
    return DQ_OK;
}

// Sorts the n-byte text at `text` (host memory, or device memory of any GPU of the group: the copies are
// cudaMemcpyDefault) with all shards.  On return shard s holds SA[slot_base, slot_base+cnt) in sa_local and
// ISA[own_begin, own_begin+own_cnt) in isa_local; sa_out (may be null) receives the whole suffix array.
int group_sort(dq_ctx *top, const uint8_t *text, uint32_t n, int32_t *sa_out)
{
    Group &g = *top->group;
    const size_t G = g.sh.size();
    g.n = 0;
    g.replicated = false;
    g.trace = getenv("DQ_TRACE") != nullptr;
    g.phases.clear();
    g.t_phase = std::chrono::steady_clock::now();
    const auto t_begin = g.t_phase;
    dq_stats st{};  // the group's figures; the shards' own counters are folded in at the end
    st.n = (int32_t)n;
    for (Shard &s : g.sh) {
        s.c->stats = dq_stats{};
        s.c->resident_n = -1;
        s.c->lcp_valid = false;
        s.c->pre3_valid = false;
        s.c->runend_valid_n = -1;
        s.c->pass_events_used = 0;
    }

    // ---- ownership of text positions: power-of-two slices, owner(pos) = pos >> kb
    int kb = 0;
    while (((uint64_t)G << kb) < n) ++kb;
    g.kb = kb;
    std::vector<uint32_t> sample_cnt(G, 0);
    uint32_t total_samples = 0;
    for (size_t i = 0; i < G; ++i) {
        Shard &s = g.sh[i];
        s.own_begin = (uint32_t)std::min<uint64_t>(n, (uint64_t)i << kb);
        s.own_cnt = (uint32_t)(std::min<uint64_t>(n, (uint64_t)(i + 1) << kb) - s.own_begin);
        sample_cnt[i] = s.own_cnt ? std::min<uint32_t>(kSamplesPerShard, s.own_cnt) : 0;
        total_samples += sample_cnt[i];
    }

    // ---- round 0a: text slices up, keys of every position, sampled keys for the splitters
    constexpr uint32_t kHalo = 64;  // a key reads at most 64 characters (small alphabets) past its position
    for (size_t i = 0; i < G; ++i) {
        Shard &s = g.sh[i];
        dq_ctx *c = s.c;
        DQ_CK(top, cudaSetDevice(c->device));
        DQ_SUB(top, c, ensure(c, s.slice, (size_t)s.own_cnt + kHalo + 64));
        if (s.own_cnt == 0) continue;
        // the halo comes from the text, the rest is zero
        const uint32_t halo = (uint32_t)(std::min<uint64_t>(n, (uint64_t)s.own_begin + s.own_cnt + kHalo) - s.own_begin);
        DQ_CK(top, cudaMemcpyAsync(s.slice.p, text + s.own_begin, halo, cudaMemcpyDefault, c->stream));
        DQ_CK(top, cudaMemsetAsync(s.slice.as<uint8_t>() + halo, 0, (size_t)s.own_cnt + kHalo + 64 - halo, c->stream));
    }
    // small alphabets (dq_suffix.cuh): every shard counts the byte values of its slice, the code is chosen once
    sx::AlphabetCode ac{};
    ac.bits = 8;
    if (n >= compact_min()) {
        uint64_t bh[256] = {};
        for (Shard &s : g.sh) {
            if (s.own_cnt == 0) continue;
            DQ_CK(top, cudaSetDevice(s.c->device));
            uint64_t part[256];
            DQ_SUB(top, s.c, byte_histogram(s.c, s.slice.as<uint8_t>(), s.own_cnt, part));
            for (int b = 0; b < 256; ++b) bh[b] += part[b];
        }
        ac = choose_code(bh);
    }
    const uint32_t key_chars = 64u / (uint32_t)ac.bits;
    for (size_t i = 0; i < G; ++i) {
        Shard &s = g.sh[i];
        dq_ctx *c = s.c;
        DQ_CK(top, cudaSetDevice(c->device));
        DQ_SUB(top, c, ensure(c, c->partK, (size_t)std::max<uint32_t>(s.own_cnt, 1) * 8));
        DQ_SUB(top, c, ensure(c, c->partV, (size_t)std::max<uint32_t>(s.own_cnt, 1) * 4));
        DQ_SUB(top, c, ensure(c, s.samples, (size_t)kSamplesPerShard * 16));
        DQ_SUB(top, c, ensure(c, s.isa_local, (size_t)std::max<uint32_t>(s.own_cnt, 1) * 4));
        DQ_SUB(top, c, ensure(c, s.meta_req, G * sizeof(ds::RunMeta)));
        DQ_SUB(top, c, ensure(c, s.meta_upd, G * sizeof(ds::RunMeta)));
        if (s.own_cnt == 0) continue;
        const uint8_t *P = nullptr;
        if (ac.bits < 8) {
            // the slice and its halo recoded: position p of the slice is character p of this shard's stream
            const uint32_t chars = (uint32_t)(std::min<uint64_t>(n, (uint64_t)s.own_begin + s.own_cnt + kHalo) - s.own_begin);
            DQ_SUB(top, c, encode_text(c, s.slice.as<uint8_t>(), chars, ac, s.packed));
            P = s.packed.as<uint8_t>();
        }
        // suffixes inside equal-byte runs are counted on the way (plain keys only): many of them => run-aware rounds
        DQ_SUB(top, c, ensure(c, c->hist, (size_t)2 * rx::kMaxPasses * rx::kRadix * 4 + 256));
        uint32_t *uniform_count = c->hist.as<uint32_t>() + 2 * rx::kMaxPasses * rx::kRadix + 32;
        DQ_CK(top, cudaMemsetAsync(uniform_count, 0, 4, c->stream));
        auto k = sx::pack_slice_kernel;
        DQ_LAUNCH(k, producer_grid(c, s.own_cnt), sx::kPackThreads, 0, c->stream, s.slice.as<uint8_t>(), s.own_begin,
                  s.own_cnt, c->partK.as<uint64_t>(), c->partV.as<uint32_t>(), ac.bits == 8 ? uniform_count : nullptr, P,
                  ac.bits);
        DQ_CK(top, cudaMemcpyAsync(s.h_small + 32, uniform_count, 4, cudaMemcpyDeviceToHost, c->stream));
        auto ks = ds::sample_keys_kernel;
        DQ_LAUNCH(ks, (uint32_t)div_up(sample_cnt[i], 256), 256, 0, c->stream, c->partK.as<uint64_t>(), s.own_cnt,
                  sample_cnt[i], s.samples.as<uint64_t>());
        c->stats.kernel_launches += 2;
        DQ_CK(top, cudaMemcpyAsync(s.h_samples, s.samples.p, (size_t)sample_cnt[i] * 8, cudaMemcpyDeviceToHost, c->stream));
    }
    DQ_CK(top, cudaGetLastError());
    DQ_TRY(group_sync(top));
    // equal-byte runs (dq_suffix.cuh): many of them => run-aware rounds, and buckets cut on (key, run code) so that the
    // heavy repeated-byte keys spread over the shards (dq_dist.cuh, "key skew")
    uint64_t uniform_total = 0;
    for (Shard &s : g.sh) uniform_total += s.own_cnt ? s.h_small[32] : 0;
    const bool run_heavy = ac.bits == 8 && uniform_total * 64 >= n && !getenv("DQ_GROUP_NO_RUNS");
    g.runend_n = 0;
    ds::Splitters sp{};
    ds::RunSplitters rsp{};
    if (run_heavy) {
        DQ_TRY(group_text_and_run_ends(top, n));
        for (size_t i = 0; i < G; ++i) {
            Shard &s = g.sh[i];
            dq_ctx *c = s.c;
            if (!sample_cnt[i]) continue;
            DQ_CK(top, cudaSetDevice(c->device));
            uint32_t *codes = reinterpret_cast<uint32_t *>(s.samples.as<uint64_t>() + kSamplesPerShard);
            auto ks = ds::sample_pairs_kernel;
            DQ_LAUNCH(ks, (uint32_t)div_up(sample_cnt[i], 256), 256, 0, c->stream, c->partK.as<uint64_t>(),
                      c->partV.as<uint32_t>(), s.own_cnt, sample_cnt[i], c->text.as<uint8_t>(), c->runend.as<uint32_t>(), n,
                      s.samples.as<uint64_t>(), codes);
            c->stats.kernel_launches++;
            DQ_CK(top, cudaMemcpyAsync(s.h_samples, s.samples.p, (size_t)kSamplesPerShard * 8 + (size_t)sample_cnt[i] * 4,
                                       cudaMemcpyDeviceToHost, c->stream));
        }
        DQ_CK(top, cudaGetLastError());
        DQ_TRY(group_sync(top));
        std::vector<std::pair<uint64_t, uint32_t>> all;
        all.reserve(total_samples);
        for (size_t i = 0; i < G; ++i) {
            const uint32_t *codes = reinterpret_cast<const uint32_t *>(g.sh[i].h_samples + kSamplesPerShard);
            for (uint32_t j = 0; j < sample_cnt[i]; ++j) all.emplace_back(g.sh[i].h_samples[j], codes[j]);
        }
        std::sort(all.begin(), all.end());
        rsp.n = (int)G - 1;
        for (int j = 0; j < rsp.n; ++j) {
            const auto &v = all.empty() ? std::pair<uint64_t, uint32_t>(0, 0) : all[(size_t)(j + 1) * all.size() / G];
            rsp.key[j] = v.first;
            rsp.code[j] = v.second;
        }
    } else {
        std::vector<uint64_t> all;
        all.reserve(total_samples);
        for (size_t i = 0; i < G; ++i) all.insert(all.end(), g.sh[i].h_samples, g.sh[i].h_samples + sample_cnt[i]);
        std::sort(all.begin(), all.end());
        sp.n = (int)G - 1;
        for (int j = 0; j < sp.n; ++j) sp.s[j] = all.empty() ? 0 : all[(size_t)(j + 1) * all.size() / G];
    }
    DQ_TRY(group_mark(top, "r0_upload_pack_sample"));

    // ---- round 0b: how many tuples go from every slice to every bucket
    ds::BucketPolicy bp{};
    bp.sp = sp;
    bp.bits = bits_for(G);
    ds::RunBucketPolicy rbp{};
    rbp.sp = rsp;
    rbp.n = n;
    rbp.bits = bits_for(G);
    for (Shard &s : g.sh) {
        DQ_CK(top, cudaSetDevice(s.c->device));
        if (run_heavy) {
            rbp.T = s.c->text.as<uint8_t>();
            rbp.run_end = s.c->runend.as<uint32_t>();
            DQ_SUB(top, s.c, digit_counts(g, s, s.c->partK.as<uint64_t>(), s.own_cnt, rbp, false, false, 0,
                                         s.c->partV.as<uint32_t>()));
        } else {
            DQ_SUB(top, s.c, digit_counts(g, s, s.c->partK.as<uint64_t>(), s.own_cnt, bp, false, false));
        }
        DQ_CK(top, cudaMemcpyAsync(s.h_small, s.c->hist.p, ds::kMaxShards * 4, cudaMemcpyDeviceToHost, s.c->stream));
    }
    DQ_TRY(group_sync(top));
    uint32_t cap = 0;
    {
        uint64_t base = 0;
        for (size_t d = 0; d < G; ++d) {
            uint64_t c = 0;
            for (size_t i = 0; i < G; ++i) c += g.sh[i].h_small[d];
            g.sh[d].cnt = (uint32_t)c;
            g.sh[d].slot_base = (uint32_t)base;
            base += c;
            cap = std::max(cap, (uint32_t)c);
        }
        if (base != n) {
            top->err = "internal: bucket counts do not add up";
            return DQ_ERR_INTERNAL;
        }
    }
    for (Shard &s : g.sh) {
        dq_ctx *c = s.c;
        DQ_CK(top, cudaSetDevice(c->device));
        DQ_SUB(top, c, dist_reserve(c, s.cnt));
        DQ_SUB(top, c, ensure(c, s.sa_local, (size_t)std::max<uint32_t>(s.cnt, 1) * 4));
        DQ_SUB(top, c, ensure(c, s.upd, (size_t)std::max<uint32_t>(s.cnt, 1) * 8));
        DQ_SUB(top, c, ensure(c, s.reply, (size_t)std::max<uint32_t>(s.cnt, 1) * 4));
        DQ_SUB(top, c, ensure(c, s.inbox_req, (size_t)std::max<uint32_t>(cap, 1) * 4 * G));
        DQ_SUB(top, c, ensure(c, s.inbox_upd, (size_t)std::max<uint32_t>(cap, 1) * 8 * G));
        bp.kout[&s - &g.sh[0]] = rbp.kout[&s - &g.sh[0]] = c->keyA.as<uint64_t>();
        bp.vout[&s - &g.sh[0]] = rbp.vout[&s - &g.sh[0]] = c->valA.as<uint32_t>();
    }
    // ---- round 0c: partition by bucket, scattered straight into the owners' sort inputs.  Equal keys must meet in
    // descending suffix order (dq_suffix.cuh, end-of-text rule): every slice is packed descending and the slices are
    // laid out from the last to the first.
    for (size_t i = 0; i < G; ++i) {
        Shard &s = g.sh[i];
        dq_ctx *c = s.c;
        DQ_CK(top, cudaSetDevice(c->device));
        uint32_t off[rx::kRadix] = {};
        for (size_t d = 0; d < G; ++d)
            for (size_t j = i + 1; j < G; ++j) off[d] += g.sh[j].h_small[d];
        uint32_t *gbase = c->hist.as<uint32_t>() + rx::kMaxPasses * rx::kRadix;
        DQ_CK(top, cudaMemcpyAsync(gbase, off, sizeof off, cudaMemcpyHostToDevice, c->stream));
        if (run_heavy) {
            rbp.T = c->text.as<uint8_t>();
            rbp.run_end = c->runend.as<uint32_t>();
            DQ_SUB(top, c, policy_pass(s, c->partK.as<uint64_t>(), c->partV.as<uint32_t>(), rbp, s.own_cnt));
        } else {
            DQ_SUB(top, c, policy_pass(s, c->partK.as<uint64_t>(), c->partV.as<uint32_t>(), bp, s.own_cnt));
        }
    }
    DQ_TRY(group_barrier(top));
    DQ_TRY(group_mark(top, "r0_partition_exchange"));

    // ---- round 0d: local sort of every bucket + first ranks
    sx::PeerIsa peers{};
    ds::IsaParts parts{};
    for (size_t d = 0; d < G; ++d) peers.p[d] = parts.p[d] = g.sh[d].isa_local.as<uint32_t>();
    peers.kb = parts.kb = kb;
    const bool direct0 = n <= g.direct_max;  // a small text: ranks go straight to the owners' ISA slices
    rx::PassPlan plan0{};
    rx::plan_add_field(plan0, 0, 64);
    std::vector<uint32_t> entered(G, 0);
    for (size_t i = 0; i < G; ++i) {
        Shard &s = g.sh[i];
        dq_ctx *c = s.c;
        entered[i] = s.cnt;
        s.a = 0;
        if (s.cnt == 0) continue;
        DQ_CK(top, cudaSetDevice(c->device));
        DQ_SUB(top, c, zero_hist(c));
        auto k = sx::hist_only_kernel;
        DQ_LAUNCH(k, producer_grid(c, s.cnt), sx::kPackThreads, plan0.npass * rx::kRadix * 4, c->stream,
                  c->keyA.as<uint64_t>(), s.cnt, plan0, c->hist.as<uint32_t>());
        c->stats.kernel_launches++;
        SortBufs b{c->keyA.as<uint64_t>(), c->keyB.as<uint64_t>(), c->valA.as<uint32_t>(), c->valB.as<uint32_t>()};
        DQ_SUB(top, c, run_passes(c, b, s.cnt, plan0, false));
        s.slot_cur = c->slotA.as<uint32_t>();
        s.slot_nxt = c->slotB.as<uint32_t>();
        DQ_SUB(top, c, (enqueue_rank<true, true>(c, b.kin, b.vin, nullptr, s.cnt, n, nullptr, nullptr, s.slot_cur,
                                         s.sa_local.as<int32_t>(), s.slot_base,
                                         direct0 ? nullptr : s.upd.as<uint64_t>(), b.kout, nullptr, nullptr, key_chars,
                                         peers)));
        s.act = b.kout;
        s.other = b.kin;
    }
    uint64_t total_active = 0;
    for (Shard &s : g.sh) {
        if (s.cnt) DQ_SUB(top, s.c, finish_rank(s.c, &s.a, nullptr));
        total_active += s.a;
    }
    // early copy (deltaq_cuda.cu, EarlyCopy): every bucket starts towards its part of the host array as soon as few of
    // its suffixes are unresolved; the slots resolved later are patched in at the end
    std::vector<EarlyCopy> ec(G);
    int32_t *host_sa = device_visible_host(sa_out);
    for (EarlyCopy &e : ec) e.host_sa = host_sa;
    auto start_early_copies = [&]() -> int {
        for (size_t i = 0; i < G; ++i) {
            Shard &s = g.sh[i];
            if (!s.cnt || !sa_out) continue;
            DQ_CK(top, cudaSetDevice(s.c->device));
            DQ_SUB(top, s.c, early_copy_maybe_start(s.c, ec[i], s.sa_local.as<int32_t>(), sa_out + s.slot_base, s.cnt, s.a,
                                                    s.cnt));
        }
        return DQ_OK;
    };
    DQ_TRY(start_early_copies());
    st.rounds = 1;
    st.active_sum = n;
    st.algorithmic_bytes = (int64_t)n * (41 + 24 * plan0.npass);
    DQ_TRY(group_mark(top, "r0_local_sort_rank"));
    if (!direct0) DQ_TRY(group_route_updates(top, entered));
    DQ_TRY(group_mark(top, "r0_route_updates"));

    // ---- equal-byte runs (dq_suffix.cuh): a text full of them (zero padding of executables) is refined by run length
    // in round 1 and carries a depth per group from then on, like the one-GPU path.  Every shard has the whole text and
    // its run ends by now (group_text_and_run_ends), and the rounds read ISA through peer pointers whatever their size:
    // only the suffixes outside runs fetch a rank, and they fetch it at their own depth.
    const bool run_aware = run_heavy && total_active > 0;
    if (run_aware) {
        DQ_TRY(for_shards(top, [&](size_t i) -> int {
            Shard &t = g.sh[i];
            dq_ctx *c = t.c;
            DQ_TRY(ensure(c, c->depthA, (size_t)std::max<uint32_t>(t.cnt, 1) * 4));
            DQ_TRY(ensure(c, c->depthB, (size_t)std::max<uint32_t>(t.cnt, 1) * 4));
            t.depth_cur = c->depthA.as<uint32_t>();
            t.depth_nxt = c->depthB.as<uint32_t>();
            return DQ_OK;
        }));
    }
    bool first_round = true;

    // ---- doubling rounds
    const int bits_r2 = bit_length(n), bits_rank = bit_length(n > 1 ? n - 1 : 1);
    rx::PassPlan rp{};
    rx::plan_add_field(rp, 0, bits_r2);
    rx::plan_add_field(rp, 32, bits_rank);
    uint64_t h = key_chars;
    while (total_active > 0) {
        DQ_TRY(start_early_copies());
        if (run_aware || total_active <= g.direct_max) {
            // ---- a small round (or any round of a run-aware sort): ISA read and written through peer pointers
            // (dq_dist.cuh, "small rounds")
            DQ_TRY(group_barrier(top));  // every rank written so far is in place
            std::vector<SortBufs> sorted(G);
            std::vector<uint32_t> min_depth(G, 0xffffffffu);
            const bool runs_round = run_aware && first_round;
            rx::PassPlan rp1{};  // the run-length keys of round 1 use all 32 low bits
            rx::plan_add_field(rp1, 0, 32);
            rx::plan_add_field(rp1, 32, bits_rank);
            const rx::PassPlan &plan_now = runs_round ? rp1 : rp;
            DQ_TRY(for_shards(top, [&](size_t i) -> int {
                Shard &s = g.sh[i];
                dq_ctx *c = s.c;
                entered[i] = s.a;
                if (s.a == 0) return DQ_OK;
                DQ_TRY(zero_hist(c));
                if (runs_round) {
                    auto k = ds::build_keys_round1_peer_kernel;
                    DQ_LAUNCH(k, producer_grid(c, s.a), sx::kPackThreads, plan_now.npass * rx::kRadix * 4, c->stream, s.act,
                              s.a, parts, c->text.as<uint8_t>(), c->runend.as<uint32_t>(), n, s.other, c->valA.as<uint32_t>(),
                              s.depth_cur, plan_now, c->hist.as<uint32_t>());
                } else {
                    auto k = ds::build_keys_peer_kernel;
                    DQ_LAUNCH(k, producer_grid(c, s.a), sx::kPackThreads, plan_now.npass * rx::kRadix * 4, c->stream, s.act,
                              s.a, parts, n, h, s.other, c->valA.as<uint32_t>(), plan_now, c->hist.as<uint32_t>(),
                              run_aware ? (const uint32_t *)s.depth_cur : (const uint32_t *)nullptr);
                }
                c->stats.kernel_launches++;
                SortBufs b{s.other, s.act, c->valA.as<uint32_t>(), c->valB.as<uint32_t>()};
                DQ_TRY(run_passes(c, b, s.a, plan_now, true));
                sorted[i] = b;
                return DQ_OK;
            }));
            DQ_TRY(group_barrier(top));  // every read of this round is done before any rank changes
            DQ_TRY(for_shards(top, [&](size_t i) -> int {
                Shard &s = g.sh[i];
                dq_ctx *c = s.c;
                if (s.a == 0) return DQ_OK;
                SortBufs &b = sorted[i];
                DQ_TRY((enqueue_rank<false, true>(c, b.kin, b.vin, s.slot_cur, s.a, n, nullptr, nullptr, s.slot_nxt,
                                                  s.sa_local.as<int32_t>(), s.slot_base, nullptr, b.kout,
                                                  run_aware ? s.depth_cur : nullptr, run_aware ? s.depth_nxt : nullptr,
                                                  (uint32_t)h, peers, ec[i].late, ec[i].late_count)));
                std::swap(s.slot_cur, s.slot_nxt);
                std::swap(s.depth_cur, s.depth_nxt);
                s.act = b.kout;
                s.other = b.kin;
                uint32_t next_a = 0;
                DQ_TRY(finish_rank(c, &next_a, run_aware ? &min_depth[i] : nullptr));
                if (next_a > s.a) {
                    c->err = "internal: active set grew";
                    return DQ_ERR_INTERNAL;
                }
                s.a = next_a;
                return DQ_OK;
            }));
            st.rounds++;
            st.active_sum += (int64_t)total_active;
            st.algorithmic_bytes += (int64_t)total_active * (52 + 24 * plan_now.npass);
            total_active = 0;
            for (Shard &s : g.sh) total_active += s.a;
            first_round = false;
            if (run_aware) {
                // every rank is now consistent to the smallest depth of an unresolved group, on any shard
                uint32_t md = 0xffffffffu;
                for (size_t i = 0; i < G; ++i)
                    if (g.sh[i].a) md = std::min(md, min_depth[i]);
                DQ_TRY(group_mark(top, "small_rounds"));
                if (total_active > 0) {
                    if (md <= h && st.rounds > 2) {
                        top->err = "internal: group depth did not grow";
                        return DQ_ERR_INTERNAL;
                    }
                    h = md;
                }
                if (st.rounds > 200) {
                    top->err = "internal: doubling did not converge";
                    return DQ_ERR_INTERNAL;
                }
                continue;
            }
            DQ_TRY(group_mark(top, "small_rounds"));
            h *= 2;
            if (h > ((uint64_t)1 << 31)) h = (uint64_t)1 << 31;
            if (st.rounds > 200) {
                top->err = "internal: doubling did not converge";
                return DQ_ERR_INTERNAL;
            }
            continue;
        }
        // requests: regroup the unresolved set by the owner of sa + h; the positions go to the owners' inboxes
        const int sb = sub_bits(G, kb);
        DQ_TRY(for_shards(top, [&](size_t i) -> int {
            Shard &s = g.sh[i];
            dq_ctx *c = s.c;
            ds::RequestPolicy pol{};
            pol.kout = s.other;
            for (size_t d = 0; d < G; ++d) pol.qout[d] = g.sh[d].inbox_req.as<uint32_t>() + (size_t)i * cap;
            pol.gbase = c->hist.as<uint32_t>() + rx::kMaxPasses * rx::kRadix;
            pol.h = h;
            pol.n = n;
            pol.kb = kb;
            pol.sb = sb;
            pol.self = (uint32_t)i;
            pol.bits = bits_for(G) + sb;
            DQ_TRY(digit_counts(g, s, s.act, s.a, pol, true, true, sb));
            return policy_pass(s, s.act, nullptr, pol, s.a);
        }));
        DQ_TRY(group_barrier(top));
        // owners answer, in request order, into the requesters' reply arrays
        {
            ds::ReplyPtrs rp_{};
            for (size_t d = 0; d < G; ++d) rp_.p[d] = g.sh[d].reply.as<uint32_t>();
            DQ_TRY(for_shards(top, [&](size_t i) -> int {
                Shard &s = g.sh[i];
                auto k = ds::reply_kernel;
                DQ_LAUNCH(k, (uint32_t)s.c->sm_count * 8, 256, 0, s.c->stream, s.inbox_req.as<uint32_t>(), cap,
                          s.meta_req.as<ds::RunMeta>(), s.isa_local.as<uint32_t>(), rp_, (uint32_t)G);
                s.c->stats.kernel_launches++;
                DQ_CK(s.c, cudaGetLastError());
                return DQ_OK;
            }));
        }
        DQ_TRY(group_barrier(top));
        DQ_TRY(group_mark(top, "rounds_fetch_isa"));
        // local: keys, sort, ranks
        DQ_TRY(for_shards(top, [&](size_t i) -> int {
            Shard &s = g.sh[i];
            dq_ctx *c = s.c;
            entered[i] = s.a;
            if (s.a == 0) return DQ_OK;
            DQ_TRY(zero_hist(c));
            // regrouped set is in s.other; keys go to s.act's buffer, values to valA
            uint64_t *keys = s.act;
            auto k = ds::build_keys_reply_kernel;
            DQ_LAUNCH(k, producer_grid(c, s.a), sx::kPackThreads, rp.npass * rx::kRadix * 4, c->stream, s.other,
                      s.reply.as<uint32_t>(), s.a, keys, c->valA.as<uint32_t>(), rp, c->hist.as<uint32_t>());
            c->stats.kernel_launches++;
            SortBufs b{keys, s.other, c->valA.as<uint32_t>(), c->valB.as<uint32_t>()};
            DQ_TRY(run_passes(c, b, s.a, rp, true));
            DQ_TRY((enqueue_rank<false, true>(c, b.kin, b.vin, s.slot_cur, s.a, n, nullptr, nullptr, s.slot_nxt,
                                              s.sa_local.as<int32_t>(), s.slot_base, s.upd.as<uint64_t>(), b.kout, nullptr,
                                              nullptr, 0, sx::PeerIsa{}, ec[i].late, ec[i].late_count)));
            std::swap(s.slot_cur, s.slot_nxt);
            s.act = b.kout;
            s.other = b.kin;
            uint32_t next_a = 0;
            DQ_TRY(finish_rank(c, &next_a, nullptr));
            if (next_a > s.a) {
                c->err = "internal: active set grew";
                return DQ_ERR_INTERNAL;
            }
            s.a = next_a;
            return DQ_OK;
        }));
        st.rounds++;
        st.active_sum += (int64_t)total_active;
        st.algorithmic_bytes += (int64_t)total_active * (52 + 24 * rp.npass);
        total_active = 0;
        for (Shard &s : g.sh) total_active += s.a;
        first_round = false;
        DQ_TRY(group_mark(top, "rounds_local_sort_rank"));
        DQ_TRY(group_route_updates(top, entered));
        DQ_TRY(group_mark(top, "rounds_route_updates"));
        h *= 2;
        if (h > ((uint64_t)1 << 31)) h = (uint64_t)1 << 31;
        if (st.rounds > 200) {
            top->err = "internal: doubling did not converge";
            return DQ_ERR_INTERNAL;
        }
    }
    st.algorithmic_bytes += (int64_t)n * 4;

    // ---- the buckets are the suffix array
    if (sa_out) {
        for (size_t i = 0; i < G; ++i) {
            Shard &s = g.sh[i];
            if (s.cnt == 0) continue;
            DQ_CK(top, cudaSetDevice(s.c->device));
            if (ec[i].started)
                DQ_SUB(top, s.c, early_copy_finish(s.c, ec[i]));
            else
                DQ_CK(top, cudaMemcpyAsync(sa_out + s.slot_base, s.sa_local.p, (size_t)s.cnt * 4, cudaMemcpyDefault, s.c->stream));
        }
    }
    DQ_TRY(group_sync(top));
    DQ_TRY(group_mark(top, "sa_out"));
    st.device_ms = (float)std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
    for (Shard &s : g.sh) {
        st.kernel_launches += s.c->stats.kernel_launches;
        st.radix_passes += s.c->stats.radix_passes;
    }
    top->stats = st;
    g.n = n;
    if (g.trace) {
        fprintf(stderr, "[dq trace] group sort n=%u shards=%zu rounds=%d:", n, G, st.rounds);
        for (const GroupPhase &p : g.phases) fprintf(stderr, " %s=%.2fms", p.name, p.ms);
        fprintf(stderr, "\n");
    }
    DQ_CK(top, cudaSetDevice(top->device));
    return DQ_OK;
}

// After group_sort: make text, SA and ISA complete in every shard's context (peer copies of the slices and buckets),
// which is what the match search reads.  `text` is the sorted text (host or device memory).
int group_replicate_index(dq_ctx *top)
{
    Group &g = *top->group;
    if (g.replicated) return DQ_OK;
    const uint32_t n = g.n;
    const int32_t rounds = top->stats.rounds;
    DQ_TRY(group_barrier(top));
    for (Shard &t : g.sh) {
        dq_ctx *c = t.c;
        DQ_CK(top, cudaSetDevice(c->device));
        DQ_SUB(top, c, ensure(c, c->text, (size_t)n + 64));
        DQ_SUB(top, c, ensure(c, c->sa, (size_t)std::max<uint32_t>(n, 1) * 4));
        DQ_SUB(top, c, ensure(c, c->isa, (size_t)std::max<uint32_t>(n, 1) * 4));
        DQ_CK(top, cudaMemsetAsync(c->text.as<uint8_t>() + n, 0, 64, c->stream));
        for (Shard &s : g.sh) {
            if (s.own_cnt) {
                DQ_CK(top, cudaMemcpyAsync(c->text.as<uint8_t>() + s.own_begin, s.slice.p, s.own_cnt, cudaMemcpyDefault, c->stream));
                DQ_CK(top, cudaMemcpyAsync(c->isa.as<uint32_t>() + s.own_begin, s.isa_local.p, (size_t)s.own_cnt * 4,
                                           cudaMemcpyDefault, c->stream));
            }
            if (s.cnt)
                DQ_CK(top, cudaMemcpyAsync(c->sa.as<int32_t>() + s.slot_base, s.sa_local.p, (size_t)s.cnt * 4,
                                           cudaMemcpyDefault, c->stream));
        }
        c->resident_n = (int32_t)n;
        c->resident_rounds = rounds;
        c->lcp_valid = false;
        c->pre3_valid = false;
        c->runend_valid_n = g.runend_n == n ? (int32_t)n : -1;  // a run-aware sort left the run ends on every shard
    }
    DQ_TRY(group_barrier(top));
    g.replicated = true;
    DQ_CK(top, cudaSetDevice(top->device));
    return DQ_OK;
}

// The search's index of `old` (LCP array, block minima, bucket and prefix tables) built by all shards: every shard
// computes the LCP entries (and counts the 3-byte prefixes) of one text range into its own zeroed copy, the copies are
// merged slice by slice through peer memory (dq_search.cuh, merge_copies_kernel), and every shard finishes its copy.
// Needs text / SA / ISA complete on every shard (group_replicate_index, or adopt_index on every shard).
int group_build_index(dq_ctx *top, uint32_t n)
{
    Group &g = *top->group;
    const size_t G = g.sh.size();
    bool all_valid = true;
    for (Shard &s : g.sh) all_valid = all_valid && s.c->lcp_valid;
    if (all_valid || n == 0) return DQ_OK;
    const bool pre3 = want_prefix3(n);
    // text ranges: whole supers, and whole warps of the seed level when that level runs (lcp_positions)
    const uint32_t seeds = seeds_per_warp(g.sh[0].c, n, (uint32_t)div_up(n, sr::kSuper));
    const uint64_t align = (uint64_t)sr::kSuper * std::max<uint32_t>(1, seeds);
    const uint64_t per = div_up(div_up((uint64_t)n, G), align) * align;
    DQ_TRY(for_shards(top, [&](size_t i) -> int {
        dq_ctx *c = g.sh[i].c;
        c->lcp_valid = false;
        c->pre3_valid = false;
        DQ_TRY(lcp_alloc(c, n));
        DQ_CK(c, cudaMemsetAsync(c->lcp.p, 0, (size_t)n * 4, c->stream));
        const uint64_t pb = std::min<uint64_t>(n, per * i), pe = std::min<uint64_t>(n, per * (i + 1));
        DQ_TRY(lcp_positions(c, n, pb, pe, true));
        if (pre3) DQ_TRY(prefix3_count(c, n, c->stream, pb, pe, true));
        return DQ_OK;
    }));
    DQ_TRY(group_barrier(top));
    sr::PeerArrays lcps{}, tabs{};
    lcps.count = tabs.count = (int)G;
    for (size_t d = 0; d < G; ++d) {
        lcps.p[d] = g.sh[d].c->lcp.as<uint32_t>();
        tabs.p[d] = g.sh[d].c->pre3.as<uint32_t>();
    }
    DQ_TRY(for_shards(top, [&](size_t i) -> int {
        dq_ctx *c = g.sh[i].c;
        const uint64_t sl = div_up((uint64_t)n, G);
        const uint64_t b = std::min<uint64_t>(n, sl * i), e = std::min<uint64_t>(n, sl * (i + 1));
        if (e > b) {
            auto k = sr::merge_copies_kernel<false>;
            DQ_LAUNCH(k, (uint32_t)c->sm_count * 8, 256, 0, c->stream, lcps, b, e);
            c->stats.kernel_launches++;
        }
        if (pre3) {
            const uint64_t ts = div_up((uint64_t)sr::kPrefix3Bins, G);
            const uint64_t tb = std::min<uint64_t>(sr::kPrefix3Bins, ts * i), te = std::min<uint64_t>(sr::kPrefix3Bins, ts * (i + 1));
            if (te > tb) {
                auto k = sr::merge_copies_kernel<true>;
                DQ_LAUNCH(k, (uint32_t)c->sm_count * 4, 256, 0, c->stream, tabs, tb, te);
                c->stats.kernel_launches++;
            }
        }
        DQ_CK(c, cudaGetLastError());
        return DQ_OK;
    }));
    DQ_TRY(group_barrier(top));
    return for_shards(top, [&](size_t i) -> int {
        dq_ctx *c = g.sh[i].c;
        if (pre3) DQ_TRY(prefix3_scan(c, n, c->stream));
        return lcp_finish(c, n);
    });
}

// Diff.Search for scan positions [scan_begin, scan_begin + count), sharded by new-data range: every shard answers a
// contiguous share against its own copy of the index.  Needs group_replicate_index (or adopt on every shard) first.
int group_search(dq_ctx *top, uint32_t n, const uint8_t *new_, uint32_t m, uint32_t scan_begin, uint32_t count,
                 int32_t *pos_out, int32_t *len_out)
{
    Group &g = *top->group;
    const size_t G = g.sh.size();
    const uint32_t base = count / (uint32_t)G, extra = count % (uint32_t)G;
    std::vector<uint32_t> begin(G + 1, 0);
    for (size_t i = 0; i < G; ++i) begin[i + 1] = begin[i] + base + (i < extra ? 1 : 0);
    // `new` goes up once: every shard copies its own share from the caller's buffer (its own PCIe link) and fetches
    // the other shares from its peers (NVLink); a query may read `new` to its end, so every shard needs all of it
    std::vector<uint32_t> nb(G + 1, 0);
    for (size_t i = 0; i < G; ++i) nb[i + 1] = (uint32_t)((uint64_t)m * (i + 1) / G);
    DQ_TRY(for_shards(top, [&](size_t i) -> int {
        dq_ctx *c = g.sh[i].c;
        c->search_seen = true;
        DQ_TRY(ensure(c, c->newtext, (size_t)m + 64));
        if (nb[i + 1] > nb[i])
            DQ_CK(c, cudaMemcpyAsync(c->newtext.as<uint8_t>() + nb[i], new_ + nb[i], nb[i + 1] - nb[i], cudaMemcpyDefault, c->stream));
        DQ_CK(c, cudaMemsetAsync(c->newtext.as<uint8_t>() + m, 0, 64, c->stream));
        return DQ_OK;
    }));
    DQ_TRY(group_barrier(top));
    DQ_TRY(for_shards(top, [&](size_t i) -> int {
        dq_ctx *c = g.sh[i].c;
        for (size_t j = 0; j < G; ++j)
            if (j != i && nb[j + 1] > nb[j] && g.sh[j].c->newtext.p != c->newtext.p)
                DQ_CK(c, cudaMemcpyAsync(c->newtext.as<uint8_t>() + nb[j], g.sh[j].c->newtext.as<uint8_t>() + nb[j],
                                         nb[j + 1] - nb[j], cudaMemcpyDefault, c->stream));
        return DQ_OK;
    }));
    DQ_TRY(group_barrier(top));
    DQ_TRY(for_shards(top, [&](size_t i) -> int {
        Shard &s = g.sh[i];
        dq_ctx *c = s.c;
        const uint32_t cnt = begin[i + 1] - begin[i];
        c->runend_new_m = -1;
        if (i) c->stats.kernel_launches = 0;
        DQ_TRY(search_resident(c, n, m, scan_begin + begin[i], cnt));
        if (cnt && (void *)(pos_out + begin[i]) != c->s_pos.p) {   // (a group's own table may be shard 0's buffer itself)
            DQ_CK(c, cudaMemcpyAsync(pos_out + begin[i], c->s_pos.p, (size_t)cnt * 4, cudaMemcpyDefault, c->stream));
            DQ_CK(c, cudaMemcpyAsync(len_out + begin[i], c->s_len.p, (size_t)cnt * 4, cudaMemcpyDefault, c->stream));
        }
        return DQ_OK;
    }));
    DQ_TRY(group_sync(top));
    float worst = 0.f, worst_index = 0.f;
    for (size_t i = 0; i < G; ++i) {
        Shard &s = g.sh[i];
        if (begin[i + 1] == begin[i]) continue;
        float ms = 0.f, ims = 0.f;
        DQ_CK(top, cudaSetDevice(s.c->device));
        DQ_CK(top, cudaEventElapsedTime(&ms, s.c->ev0, s.c->ev1));
        DQ_CK(top, cudaEventElapsedTime(&ims, s.c->ev0, s.c->ev_index));
        if (getenv("DQ_TRACE")) fprintf(stderr, "[dq trace] group search shard %zu: %u positions, %.2f ms (index %.2f ms)\n", i, begin[i + 1] - begin[i], ms, ims);
        worst = std::max(worst, ms);
        worst_index = std::max(worst_index, ims);
        if (i) top->stats.kernel_launches += s.c->stats.kernel_launches;
    }
    top->stats.search_index_ms = worst_index;
    top->stats.search_ms = worst;
    top->stats.search_queries = (int32_t)count;
    DQ_CK(top, cudaSetDevice(top->device));
    return DQ_OK;
}

// does this search go to all shards?  With I == NULL only if the resident index is the group's.
bool group_wants_search(dq_ctx *ctx, int32_t n, const int32_t *I, int32_t count)
{
    if (!ctx->group || n <= 0) return false;
    if (!I) return ctx->group->n == (uint32_t)n;
    return count > 0 && (uint32_t)count >= ctx->group->shard_min;
}

int group_search_common(dq_ctx *top, const uint8_t *old_, int32_t n, const int32_t *I, const uint8_t *new_, int32_t m,
                        int32_t scan_begin, int32_t count, int32_t *pos_out, int32_t *len_out)
{
    DQ_TRY(check_args(top, n >= 0 && m >= 0 && scan_begin >= 0 && count >= 0 && (int64_t)scan_begin + count <= m,
                      "bsdiff_search: bad lengths or scan range"));
    DQ_TRY(check_args(top, (n == 0 || old_ || !I) && (m == 0 || new_) && (count == 0 || (pos_out && len_out)),
                      "bsdiff_search: null buffer"));
    Group &g = *top->group;
    if (I) {
        // a caller-supplied suffix array: every shard takes its own copy (and inverts it)
        g.n = 0;
        g.replicated = false;
        for (Shard &s : g.sh) {
            DQ_CK(top, cudaSetDevice(s.c->device));
            s.c->stats = dq_stats{};
            DQ_SUB(top, s.c, adopt_index(s.c, old_, (uint32_t)n, I, cudaMemcpyDefault));
        }
    } else {
        DQ_TRY(group_replicate_index(top));
    }
    // the index build spans all GPUs: timed on the host between two synchronisations, and added to the search's figures
    DQ_TRY(group_sync(top));
    const auto t0 = std::chrono::steady_clock::now();
    DQ_TRY(group_build_index(top, (uint32_t)n));
    DQ_TRY(group_sync(top));
    const float index_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (getenv("DQ_TRACE")) fprintf(stderr, "[dq trace] group index build %.2f ms\n", index_ms);
    DQ_TRY(group_search(top, (uint32_t)n, new_, (uint32_t)m, (uint32_t)scan_begin, (uint32_t)count, pos_out, len_out));
    top->stats.search_index_ms += index_ms;
    top->stats.search_ms += index_ms;
    return DQ_OK;
}

// LCP array of a text the group has sorted (resident, sharded): index built by all GPUs, then one copy leaves shard 0
int group_lcp(dq_ctx *top, int32_t n, int32_t *lcp_out)
{
    Group &g = *top->group;
    DQ_TRY(group_replicate_index(top));
    DQ_TRY(group_sync(top));
    DQ_TRY(group_build_index(top, (uint32_t)n));
    DQ_TRY(group_sync(top));
    dq_ctx *c = g.sh[0].c;
    DQ_CK(top, cudaSetDevice(c->device));
    DQ_CK(top, cudaMemcpyAsync(lcp_out, c->lcp.p, (size_t)n * 4, cudaMemcpyDefault, c->stream));
    DQ_CK(top, cudaStreamSynchronize(c->stream));
    return DQ_OK;
}
