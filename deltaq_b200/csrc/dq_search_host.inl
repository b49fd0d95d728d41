// dq_search_host.inl -- host orchestration of the bsdiff match search (included inside deltaq_cuda.cu's
// anonymous namespace).

namespace sr = dq::search;

// level sizes of the block-minimum hierarchy over an n-entry LCP array
int lcp_levels(uint32_t n, uint32_t size[sr::kMaxLevels])
{
    int top = 0;
    size[0] = n;
    while (size[top] > 64 && top + 1 < sr::kMaxLevels) {
        size[top + 1] = (uint32_t)div_up(size[top], 64);
        ++top;
    }
    for (int l = top + 1; l < sr::kMaxLevels; ++l) size[l] = 0;
    return top;
}

// Supers per warp of the seed level (0 = no seed level).  Seeds pay off when repeats are long -- a warp of the head
// kernels that starts cold inside a repeat of R bytes compares R bytes, and without seeds every super starts cold --
// and cost a serial ~0.5 ms otherwise.  Long repeats show in the number of doubling rounds the sort needed
// (r rounds <=> some repeat of >= 8 * 2^(r-2) bytes); for a suffix array handed in by the caller only the size is
// known.  With seeds: enough supers per warp to keep the cold starts to about two thousand, at least 8 so short
// inputs still spread over many warps, at most 64 (the lanes of pass 1).
uint32_t seeds_per_warp(dq_ctx *ctx, uint32_t n, uint32_t supers)
{
    if (const char *e = getenv("DQ_SEEDS_PER")) return (uint32_t)atoi(e);  // tuning only
    const bool long_repeats = ctx->resident_rounds >= 0 ? ctx->resident_rounds >= 14 : n >= (32u << 20);
    if (!long_repeats) return 0;
    return (uint32_t)std::min<uint64_t>(64, std::max<uint64_t>(8, div_up(supers, 2048)));
}

// 3-byte prefix table of the resident text (Index::pre3): the from-scratch searches then start from a bucket of
// n / 2^24 suffixes instead of n / 2^16.  Measured: search 148 -> 129 ms at 512 MiB (C5 recipe), no gain at 16 MiB (C2,
// 3.04 vs 3.08 ms with the table's ~0.04 ms build), so it is built from 32 MiB of text up.
bool want_prefix3(uint32_t n)
{
    if (const char *e = getenv("DQ_PREFIX3")) return atoi(e) != 0;  // tests / tuning
    return n >= (32u << 20);
}

// 3-byte prefix table from the text, in two steps so that a device group can count by text range and add the counts
// up in between (dq_group.inl): counts of suffixes [p_begin, p_end) into the zeroed table, then the scans
int prefix3_count(dq_ctx *ctx, uint32_t n, cudaStream_t stream, uint64_t p_begin, uint64_t p_end, bool zero)
{
    DQ_TRY(ensure(ctx, ctx->pre3, ((size_t)sr::kPrefix3Bins + 4) * 4));
    DQ_TRY(ensure(ctx, ctx->pre3tile, (size_t)sr::kPrefix3Tiles * 4));
    uint32_t *table = ctx->pre3.as<uint32_t>();
    if (zero) DQ_CK(ctx, cudaMemsetAsync(table, 0, (size_t)sr::kPrefix3Bins * 4, stream));
    const uint64_t hi = std::min<uint64_t>(p_end, n);
    if (hi <= p_begin) return DQ_OK;
    const uint32_t g = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(div_up(hi - p_begin, 256), (uint64_t)ctx->sm_count * 8));
    auto k1 = sr::prefix3_hist_kernel;
    DQ_LAUNCH(k1, g, 256, 0, stream, ctx->text.as<uint8_t>(), n, table, p_begin, p_end);
    ctx->stats.kernel_launches++;
    return DQ_OK;
}

int prefix3_scan(dq_ctx *ctx, uint32_t n, cudaStream_t stream)
{
    uint32_t *table = ctx->pre3.as<uint32_t>(), *tiles = ctx->pre3tile.as<uint32_t>();
    auto k2 = sr::prefix3_tile_sum_kernel;
    DQ_LAUNCH(k2, sr::kPrefix3Tiles, 256, 0, stream, table, tiles);
    auto k3 = sr::prefix3_tile_scan_kernel;
    DQ_LAUNCH(k3, 1, 1024, 0, stream, tiles, ctx->text.as<uint8_t>(), n, table);
    auto k4 = sr::prefix3_apply_kernel;
    DQ_LAUNCH(k4, sr::kPrefix3Tiles, 256, 0, stream, table, tiles);
    ctx->stats.kernel_launches += 3;
    ctx->pre3_valid = true;
    return DQ_OK;
}

int build_prefix3(dq_ctx *ctx, uint32_t n, cudaStream_t stream)
{
    if (ctx->pre3_valid) return DQ_OK;
    DQ_TRY(prefix3_count(ctx, n, stream, 0, n, true));
    return prefix3_scan(ctx, n, stream);
}

// LCP array + block minima of the resident (ctx->text, ctx->sa, ctx->isa) of length n, in three steps (a device group
// runs the middle one by text range and merges the copies before the last one: dq_group.inl, group_build_index)
int lcp_alloc(dq_ctx *ctx, uint32_t n)
{
    const uint32_t chunks = (uint32_t)div_up(n, sr::kChunk), supers = (uint32_t)div_up(n, sr::kSuper);
    uint32_t size[sr::kMaxLevels];
    const int top = lcp_levels(n, size);
    size_t total = 0;
    for (int l = 0; l <= top; ++l) total += ((size_t)size[l] + 63) & ~(size_t)63;
    DQ_TRY(ensure(ctx, ctx->lcp, total * 4));
    DQ_TRY(ensure(ctx, ctx->headl, ((size_t)chunks + 1) * 4));
    DQ_TRY(ensure(ctx, ctx->seedl, ((size_t)supers + 1) * 4));
    DQ_TRY(ensure(ctx, ctx->seedp, ((size_t)supers + 1) * 4));
    DQ_TRY(ensure(ctx, ctx->bkt, (size_t)2 * 65536 * 4));
    DQ_TRY(ensure(ctx, ctx->phi, ((size_t)n + 64) * 4));
    DQ_TRY(ensure(ctx, ctx->plcp, ((size_t)n + 64) * 4));
    return DQ_OK;
}

// LCP entries of the text positions i in [pos_begin, pos_end) -- both multiples of kSuper * seeds_per_warp, or
// pos_end >= n: PLCP[i] along the text, then either LCP[ISA[i]] = PLCP[i] over the range (scatter: a device group, whose
// shards merge their copies afterwards) or LCP[r] = PLCP[SA[r]] over all ranks (whole text on one GPU)
int lcp_positions(dq_ctx *ctx, uint32_t n, uint64_t pos_begin, uint64_t pos_end, bool scatter)
{
    pos_end = std::min<uint64_t>(pos_end, n);
    if (pos_begin >= pos_end) return DQ_OK;
    const uint64_t chunks = div_up(n, sr::kChunk), supers = div_up(n, sr::kSuper);
    const uint8_t *T = ctx->text.as<uint8_t>();
    const int32_t *SA = ctx->sa.as<int32_t>();
    const uint32_t *ISA = ctx->isa.as<uint32_t>();
    uint32_t *PHI = ctx->phi.as<uint32_t>(), *PLCP = ctx->plcp.as<uint32_t>();
    // the chains of the range need the head after their last chunk: one more super of heads (and the seed it starts from)
    const uint64_t s0 = pos_begin / sr::kSuper, s1 = std::min<uint64_t>(div_up(pos_end, sr::kSuper) + 1, supers);
    const uint32_t per0 = seeds_per_warp(ctx, n, (uint32_t)supers);
    {
        // rank predecessors of every position the walks below touch (whole warps of the seed level included)
        const uint64_t pb = per0 ? (s0 / per0) * per0 * sr::kSuper : s0 * sr::kSuper;
        const uint64_t pe = std::min<uint64_t>(n, (per0 ? div_up(s1, per0) * per0 : s1) * sr::kSuper + sr::kChunk);
        auto k = sr::phi_kernel;
        const uint32_t g = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(div_up(pe - pb, 256 * 4), (uint64_t)ctx->sm_count * 16));
        DQ_LAUNCH(k, g, 256, 0, ctx->stream, SA, ISA, PHI, pb, pe);
        ctx->stats.kernel_launches++;
    }
    {
        // level S then level A (see lcp_heads_kernel)
        const uint32_t per = per0;
        const uint32_t *run = ctx->runend_valid_n == (int32_t)n ? ctx->runend.as<uint32_t>() : nullptr;
        auto k = sr::lcp_heads_kernel;
        if (per) {
            const uint64_t w0 = s0 / per, w1 = div_up(s1, per);
            DQ_LAUNCH(k, (uint32_t)div_up((w1 - w0) * 32, sr::kThreads), sr::kThreads, 0, ctx->stream, T, n, PHI,
                      ctx->seedl.as<uint32_t>(), run, (uint32_t)sr::kSuper, per, (const uint32_t *)nullptr, w0, w1);
            ctx->stats.kernel_launches++;
        }
        DQ_LAUNCH(k, (uint32_t)div_up((s1 - s0) * 32, sr::kThreads), sr::kThreads, 0, ctx->stream, T, n, PHI,
                  ctx->headl.as<uint32_t>(), run, (uint32_t)sr::kChunk, (uint32_t)sr::kHeads,
                  per ? (const uint32_t *)ctx->seedl.as<uint32_t>() : (const uint32_t *)nullptr, s0, s1);
        ctx->stats.kernel_launches++;
    }
    {
        const uint64_t c0 = pos_begin / sr::kChunk, c1 = std::min<uint64_t>(div_up(pos_end, sr::kChunk), chunks);
        auto k = sr::lcp_chain_kernel;
        DQ_LAUNCH(k, (uint32_t)div_up(c1 - c0, sr::kThreads), sr::kThreads, 0, ctx->stream, T, n, PHI,
                  ctx->headl.as<uint32_t>(), PLCP, c0, c1);
        ctx->stats.kernel_launches++;
    }
    {
        const uint64_t count = scatter ? pos_end - pos_begin : n;
        const uint32_t g = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(div_up(count, 256 * 4), (uint64_t)ctx->sm_count * 16));
        if (scatter) {
            auto k = sr::lcp_scatter_kernel;
            DQ_LAUNCH(k, g, 256, 0, ctx->stream, ISA, PLCP, ctx->lcp.as<uint32_t>(), pos_begin, pos_end);
        } else {
            auto k = sr::lcp_gather_kernel;
            DQ_LAUNCH(k, g, 256, 0, ctx->stream, SA, PLCP, ctx->lcp.as<uint32_t>(), n);
        }
        ctx->stats.kernel_launches++;
    }
    DQ_CK(ctx, cudaGetLastError());
    return DQ_OK;
}

// bucket bounds and the block-minimum levels over a complete level 0
int lcp_finish(dq_ctx *ctx, uint32_t n)
{
    {
        auto k = sr::bucket_bounds_kernel;
        DQ_LAUNCH(k, 256, 256, 0, ctx->stream, ctx->text.as<uint8_t>(), n, ctx->sa.as<int32_t>(), ctx->bkt.as<uint32_t>(),
                  ctx->bkt.as<uint32_t>() + 65536);
        ctx->stats.kernel_launches++;
    }
    uint32_t size[sr::kMaxLevels];
    const int top = lcp_levels(n, size);
    uint32_t *lv = ctx->lcp.as<uint32_t>();
    for (int l = 1; l <= top; ++l) {
        uint32_t *nxt = lv + (((size_t)size[l - 1] + 63) & ~(size_t)63);
        auto k = sr::block_min_kernel;
        const uint32_t g = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(div_up(size[l], 8), (uint64_t)ctx->sm_count * 16));
        DQ_LAUNCH(k, g, 256, 0, ctx->stream, lv, size[l - 1], 64u, nxt, size[l]);
        ctx->stats.kernel_launches++;
        lv = nxt;
    }
    DQ_CK(ctx, cudaGetLastError());
    ctx->lcp_valid = true;
    return DQ_OK;
}

int build_lcp(dq_ctx *ctx, uint32_t n)
{
    if (ctx->lcp_valid || n == 0) return DQ_OK;
    DQ_TRY(lcp_alloc(ctx, n));
    DQ_TRY(lcp_positions(ctx, n, 0, n, false));
    if (want_prefix3(n)) DQ_TRY(build_prefix3(ctx, n, ctx->stream));
    return lcp_finish(ctx, n);
}

int ensure_pinned(dq_ctx *ctx, PinBuf &b, size_t bytes)
{
    if (b.cap >= bytes && b.p) return DQ_OK;
    if (b.p) {
        DQ_CK(ctx, cudaFreeHost(b.p));
        b.p = nullptr;
        b.cap = 0;
    }
    size_t want = std::max<size_t>(bytes, 4096);
    DQ_CK(ctx, cudaHostAlloc(&b.p, want, cudaHostAllocDefault));
    b.cap = want;
    return DQ_OK;
}

// run_end[] of ctx->newtext (m bytes) into ctx->runend_new, on `stream`
int run_ends_of_new(dq_ctx *ctx, uint32_t m, cudaStream_t stream)
{
    const uint32_t ntiles = (uint32_t)div_up(m, sx::kRunTile);
    DQ_TRY(ensure(ctx, ctx->runend_new, (size_t)m * 4));
    DQ_TRY(ensure(ctx, ctx->runtile_new, (size_t)ntiles * 8));
    uint32_t *tile_first = ctx->runtile_new.as<uint32_t>(), *next_after = tile_first + ntiles;
    auto k1 = sx::run_tile_first_kernel;
    DQ_LAUNCH(k1, ntiles, 256, 0, stream, ctx->newtext.as<uint8_t>(), m, tile_first);
    auto k2 = sx::run_tile_scan_kernel;
    DQ_LAUNCH(k2, 1, 1024, 0, stream, tile_first, ntiles, m, next_after);
    auto k3 = sx::run_end_kernel;
    DQ_LAUNCH(k3, ntiles, 256, 0, stream, ctx->newtext.as<uint8_t>(), m, next_after, ctx->runend_new.as<uint32_t>());
    ctx->runend_new_m = (int32_t)m;
    return DQ_OK;
}

// (ctx->text, ctx->sa, ctx->isa) describe `old` (n bytes); ctx->newtext holds `new` (m bytes, padded).
// Fills ctx->s_pos / ctx->s_len [0, count).
// coded: pipeline mode for dq_cuda_bsdiff_streams -- the chain kernel runs in kSlices launches, each followed by
// encode_table_kernel over its tiles and an asynchronous D2H of the slice's code bytes on the copy stream (heads and
// tile entries are written by the kernel straight into pinned host memory); slice_end[] / slice_done[] tell the
// host loop when a prefix of the coded table is usable.
constexpr int kSlices = 8;

// table_ready (with coded): ctx->s_pos / ctx->s_len already hold the answers for [0, count) -- a device group filled
// them from all its shards (dq_group.inl) -- and only the coding and the copies to the host run here.
int search_resident(dq_ctx *ctx, uint32_t n, uint32_t m, uint32_t scan_begin, uint32_t count, bool coded = false,
                    bool table_ready = false)
{
    ctx->stats.search_queries = (int32_t)count;
    ctx->slices_used = 0;
    if (count == 0) return DQ_OK;
    DQ_CK(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    if (!table_ready) DQ_TRY(build_lcp(ctx, n));
    DQ_CK(ctx, cudaEventRecord(ctx->ev_index, ctx->stream));
    const uint32_t chunks = (uint32_t)div_up(count, sr::kChunk), supers = (uint32_t)div_up(count, sr::kSuper);
    DQ_TRY(ensure(ctx, ctx->s_pos, (size_t)count * 4));
    DQ_TRY(ensure(ctx, ctx->s_len, (size_t)count * 4));
    DQ_TRY(ensure(ctx, ctx->headp, (size_t)chunks * 4));
    DQ_TRY(ensure(ctx, ctx->headl, (size_t)std::max<uint64_t>(chunks, div_up(n, sr::kChunk)) * 4));
    sr::Texts t{ctx->text.as<uint8_t>(), ctx->newtext.as<uint8_t>(), n, m};
    if (ctx->runend_valid_n == (int32_t)n && m > 0) {
        // the sort found `old` full of equal-byte runs: give the comparisons run ends of `new` as well
        if (ctx->runend_new_m != (int32_t)m) {
            DQ_TRY(run_ends_of_new(ctx, m, ctx->stream));
            ctx->stats.kernel_launches += 3;
        }
        t.run_old = ctx->runend.as<uint32_t>();
        t.run_new = ctx->runend_new.as<uint32_t>();
    }
    sr::Index ix{};
    ix.SA = ctx->sa.as<int32_t>();
    ix.ISA = ctx->isa.as<uint32_t>();
    ix.top = lcp_levels(n, ix.size);
    ix.bkt_lo = ctx->bkt.as<uint32_t>();
    ix.bkt_hi = ctx->bkt.as<uint32_t>() + 65536;
    ix.pre3 = ctx->pre3_valid ? ctx->pre3.as<uint32_t>() : nullptr;
    {
        const uint32_t *lv = ctx->lcp.as<uint32_t>();
        for (int l = 0; l <= ix.top; ++l) {
            ix.lv[l] = lv;
            lv += ((size_t)ix.size[l] + 63) & ~(size_t)63;
        }
    }
    const int slices = coded ? kSlices : 1;
    if (coded) {
        const uint32_t tiles = (uint32_t)div_up(count, sr::kCodeTile);
        ctx->heads_cap = ctx->heads_cap_override ? ctx->heads_cap_override : count / 4 + 4096;
        DQ_TRY(ensure(ctx, ctx->d_code, (size_t)count + 64));
        DQ_TRY(ensure(ctx, ctx->d_headcount, 256));
        DQ_TRY(ensure_pinned(ctx, ctx->h_code, (size_t)count + 64));
        DQ_TRY(ensure_pinned(ctx, ctx->h_heads, (size_t)ctx->heads_cap * sizeof(sr::MatchHead)));
        DQ_TRY(ensure_pinned(ctx, ctx->h_tiles, (size_t)tiles * sizeof(uint2)));
        DQ_CK(ctx, cudaMemsetAsync(ctx->d_headcount.p, 0, 4, ctx->stream));
    }
    // All heads in one launch (sliced head launches are tail-bound each: measured slower end to end), then the
    // chain kernel in slices of whole super-chunks.
    const uint32_t supers_per = (uint32_t)div_up(supers, slices);
    ctx->slices_used = 0;
    if (!table_ready) {
        // level S then level A (see search_heads_kernel)
        const uint32_t per = seeds_per_warp(ctx, n, supers);
        DQ_TRY(ensure(ctx, ctx->seedl, (size_t)std::max<uint64_t>(supers, div_up(n, sr::kSuper)) * 4));
        DQ_TRY(ensure(ctx, ctx->seedp, (size_t)std::max<uint64_t>(supers, div_up(n, sr::kSuper)) * 4));
        auto k = sr::search_heads_kernel;
        if (per)
            DQ_LAUNCH(k, (uint32_t)div_up(div_up(supers, per) * 32, sr::kThreads), sr::kThreads, 0, ctx->stream, t, ix, scan_begin,
                      count, ctx->seedp.as<uint32_t>(), ctx->seedl.as<uint32_t>(), (uint32_t)sr::kSuper, per,
                      (const uint32_t *)nullptr, (const uint32_t *)nullptr);
        DQ_LAUNCH(k, (uint32_t)div_up((uint64_t)supers * 32, sr::kThreads), sr::kThreads, 0, ctx->stream, t, ix, scan_begin,
                  count, ctx->headp.as<uint32_t>(), ctx->headl.as<uint32_t>(), (uint32_t)sr::kChunk, (uint32_t)sr::kHeads,
                  per ? (const uint32_t *)ctx->seedp.as<uint32_t>() : (const uint32_t *)nullptr,
                  per ? (const uint32_t *)ctx->seedl.as<uint32_t>() : (const uint32_t *)nullptr);
        ctx->stats.kernel_launches += 2;
    }
    if (coded) DQ_CK(ctx, cudaEventRecord(ctx->heads_done, ctx->stream));
    for (int sl = 0; sl < slices; ++sl) {
        const uint32_t sb = std::min<uint64_t>((uint64_t)sl * supers_per, supers), se = std::min<uint64_t>((uint64_t)(sl + 1) * supers_per, supers);
        if (sb >= se) break;
        const uint32_t cb = sb * sr::kHeads, ce = std::min<uint64_t>((uint64_t)se * sr::kHeads, chunks);
        // a slice is less than one wave of chains of very uneven length: each slice runs on its own stream so
        // the next one fills the SMs as this one drains
        cudaStream_t st = coded ? ctx->slice_stream[sl] : ctx->stream;
        if (coded) DQ_CK(ctx, cudaStreamWaitEvent(st, ctx->heads_done, 0));
        if (!table_ready) {
            auto k = sr::search_chain_kernel;
            DQ_LAUNCH(k, (uint32_t)div_up(ce - cb, sr::kThreads), sr::kThreads, 0, st, t, ix, scan_begin, count,
                      ctx->headp.as<uint32_t>(), ctx->headl.as<uint32_t>(), ctx->s_pos.as<int32_t>(),
                      ctx->s_len.as<int32_t>(), cb, ce, chunks);
        }
        ctx->stats.kernel_launches++;
        if (coded) {
            const uint64_t b = (uint64_t)cb * sr::kChunk, e = std::min<uint64_t>((uint64_t)ce * sr::kChunk, count);
            static_assert(sr::kSuper % sr::kCodeTile == 0, "slices must cover whole code tiles");
            const uint32_t tb = (uint32_t)(b / sr::kCodeTile), te = (uint32_t)div_up(e, sr::kCodeTile);
            // the first position of the slice is compared with the last one of the slice before
            DQ_CK(ctx, cudaEventRecord(ctx->slice_ready[sl], st));
            if (sl > 0) DQ_CK(ctx, cudaStreamWaitEvent(st, ctx->slice_ready[sl - 1], 0));
            auto k = sr::encode_table_kernel;
            DQ_LAUNCH(k, te - tb, 256, 0, st, ctx->s_pos.as<int32_t>(), ctx->s_len.as<int32_t>(), count, tb,
                      ctx->d_code.as<uint8_t>(), static_cast<sr::MatchHead *>(ctx->h_heads.p), ctx->heads_cap,
                      ctx->d_headcount.as<uint32_t>(), static_cast<uint2 *>(ctx->h_tiles.p));
            ctx->stats.kernel_launches++;
            DQ_CK(ctx, cudaMemcpyAsync(static_cast<uint8_t *>(ctx->h_code.p) + b, ctx->d_code.as<uint8_t>() + b, e - b,
                                       cudaMemcpyDeviceToHost, st));
            DQ_CK(ctx, cudaEventRecord(ctx->slice_done[sl], st));
            DQ_CK(ctx, cudaStreamWaitEvent(ctx->stream, ctx->slice_done[sl], 0));
            ctx->slice_end[sl] = (int32_t)e;
            ctx->slices_used = sl + 1;
        }
    }
    DQ_CK(ctx, cudaGetLastError());
    DQ_CK(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    return DQ_OK;
}

// makes (text, sa, isa) resident from a caller-supplied suffix array
int adopt_index(dq_ctx *ctx, const uint8_t *old_, uint32_t n, const int32_t *I, cudaMemcpyKind kind)
{
    ctx->resident_n = -1;
    ctx->resident_rounds = -1;
    ctx->lcp_valid = false;
    ctx->pre3_valid = false;
    ctx->runend_valid_n = -1;
    DQ_TRY(upload_text(ctx, ctx->text, old_, n, kind));
    DQ_TRY(ensure(ctx, ctx->sa, (size_t)n * 4));
    DQ_TRY(ensure(ctx, ctx->isa, (size_t)n * 4));
    if (n) {
        // the caller's array is checked before anything follows its entries: it must be a permutation of [0, n)
        DQ_TRY(ensure(ctx, ctx->d_headcount, 256));
        uint32_t *bad = ctx->d_headcount.as<uint32_t>() + 8;
        DQ_CK(ctx, cudaMemsetAsync(bad, 0, 4, ctx->stream));
        DQ_CK(ctx, cudaMemsetAsync(ctx->isa.p, 0xff, (size_t)n * 4, ctx->stream));
        DQ_CK(ctx, cudaMemcpyAsync(ctx->sa.p, I, (size_t)n * 4, kind, ctx->stream));
        const uint32_t g = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(div_up(n, 256), (uint64_t)ctx->sm_count * 16));
        auto k = sr::invert_sa_kernel;
        DQ_LAUNCH(k, g, 256, 0, ctx->stream, ctx->sa.as<int32_t>(), n, ctx->isa.as<uint32_t>(), bad);
        auto kc = sr::check_inverse_kernel;
        DQ_LAUNCH(kc, g, 256, 0, ctx->stream, ctx->isa.as<uint32_t>(), n, bad);
        ctx->stats.kernel_launches += 2;
        DQ_CK(ctx, cudaGetLastError());
        DQ_CK(ctx, cudaMemcpyAsync(ctx->h_count + 8, bad, 4, cudaMemcpyDeviceToHost, ctx->stream));
        DQ_CK(ctx, cudaStreamSynchronize(ctx->stream));
        if (ctx->h_count[8] != 0) {
            ctx->err = "bsdiff_search: the suffix array passed as I is not a permutation of [0, n) (" +
                       std::to_string(ctx->h_count[8]) + " entries out of range or positions never named)";
            return DQ_ERR_INVALID_ARGUMENT;
        }
    }
    ctx->resident_n = (int32_t)n;
    return DQ_OK;
}

int search_common(dq_ctx *ctx, const uint8_t *old_, int32_t n, const int32_t *I, const uint8_t *new_, int32_t m,
                  int32_t scan_begin, int32_t count, int32_t *pos_out, int32_t *len_out, bool device_ptrs)
{
    DQ_TRY(check_args(ctx, n >= 0 && m >= 0 && scan_begin >= 0 && count >= 0 && (int64_t)scan_begin + count <= m,
                      "bsdiff_search: bad lengths or scan range"));
    DQ_TRY(check_args(ctx, (n == 0 || old_ || !I) && (m == 0 || new_) && (count == 0 || (pos_out && len_out)),
                      "bsdiff_search: null buffer"));
    DQ_CK(ctx, cudaSetDevice(ctx->device));
    ctx->search_seen = true;
    const cudaMemcpyKind in = device_ptrs ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    const cudaMemcpyKind out = device_ptrs ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    const int32_t launches_before = ctx->stats.kernel_launches;
    if (I) {
        DQ_TRY(check_args(ctx, n == 0 || old_, "bsdiff_search: old is null"));
        ctx->stats = dq_stats{};
        DQ_TRY(adopt_index(ctx, old_, (uint32_t)n, I, in));
    } else {
        DQ_TRY(check_args(ctx, ctx->resident_n == n,
                          "bsdiff_search: I is NULL but no suffix array of this length is resident on the device"));
        ctx->stats.kernel_launches = launches_before;
    }
    DQ_TRY(upload_text(ctx, ctx->newtext, new_, (uint32_t)m, in));
    ctx->runend_new_m = -1;
    DQ_TRY(search_resident(ctx, (uint32_t)n, (uint32_t)m, (uint32_t)scan_begin, (uint32_t)count));
    if (count) {
        DQ_CK(ctx, cudaMemcpyAsync(pos_out, ctx->s_pos.p, (size_t)count * 4, out, ctx->stream));
        DQ_CK(ctx, cudaMemcpyAsync(len_out, ctx->s_len.p, (size_t)count * 4, out, ctx->stream));
    }
    DQ_CK(ctx, cudaStreamSynchronize(ctx->stream));
    if (count) DQ_CK(ctx, cudaEventElapsedTime(&ctx->stats.search_ms, ctx->ev0, ctx->ev1));
    if (count) DQ_CK(ctx, cudaEventElapsedTime(&ctx->stats.search_index_ms, ctx->ev0, ctx->ev_index));
    return DQ_OK;
}

// dq_cuda_lcp / dq_cuda_lcp_device: the LCP array the search is anchored on, for the caller.  lcp_out[r] = length of the
// longest common prefix of the suffixes SA[r-1] and SA[r]; lcp_out[0] = 0.
int lcp_common(dq_ctx *ctx, const uint8_t *text, int32_t n, const int32_t *I, int32_t *lcp_out, bool device_ptrs)
{
    DQ_TRY(check_args(ctx, n >= 0 && (n == 0 || lcp_out) && (!I || n == 0 || text), "lcp: bad arguments"));
    if (n == 0) return DQ_OK;
    DQ_CK(ctx, cudaSetDevice(ctx->device));
    if (I) {
        ctx->stats = dq_stats{};
        DQ_TRY(adopt_index(ctx, text, (uint32_t)n, I, device_ptrs ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice));
    } else {
        DQ_TRY(check_args(ctx, ctx->resident_n == n,
                          "lcp: I is NULL but no suffix array of this length is resident on the device"));
    }
    DQ_TRY(build_lcp(ctx, (uint32_t)n));
    DQ_CK(ctx, cudaMemcpyAsync(lcp_out, ctx->lcp.p, (size_t)n * 4,
                               device_ptrs ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, ctx->stream));
    DQ_CK(ctx, cudaStreamSynchronize(ctx->stream));
    return DQ_OK;
}
