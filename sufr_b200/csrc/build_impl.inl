// The build pipeline (class Build), generic in the position type.  Included twice by builder.cu together with
// pos_kernels.inl: namespace sufr::p32 (pos_t = uint32_t) and sufr::p64 (pos_t = uint64_t).  No include guard on purpose.
// ---------------------------------------------------------------------------------------------
class Build {
   public:
    Build(Ctx& c, const SufrB200Args& a, uint32_t index_bits, int text_memory, int result_memory)
        : ctx(c), args(a), index_bits_(index_bits), text_memory_(text_memory), result_memory_(result_memory),
          timer(c.stream) {}

    void run(SufrB200Result* out);

   private:
    Ctx& ctx;
    const SufrB200Args& args;
    uint32_t index_bits_;
    int text_memory_, result_memory_;
    EventTimer timer;
    cudaStream_t st() const { return ctx.stream; }

    uint64_t n = 0;       // text length
    uint64_t s = 0;       // number of elements being sorted on this rank
    KeySpec ks{};
    SeedMaskInfo mask;
    bool has_mask = false;
    bool filter_active = false;
    uint32_t alphabet = 0;
    uint32_t code_n_ = 0;  // packed-text code of 'N' (0 = the text has none)
    uint32_t refine_rounds = 0, doubling_rounds = 0;
    uint64_t doubling_depth_ = 0;  // h of the last prefix-doubling round

    DevBuf<uint8_t> d_text;     // transformed text
    DevBuf<uint64_t> d_words;   // packed text
    DevBuf<uint32_t> d_maskpos;
    DevBuf<uint64_t> d_packed2, d_irr;  // 2-bit fast path (keys.cuh: first_key_fast2)
    DevBuf<uint8_t> d_cls;
    bool sentinel_ = false;             // filtered suffixes ride through the sort with key ~0
    uint64_t sort_n_ = 0;               // elements handed to the main sort (>= s when sentinel_)
    uint64_t indexed_count_ = 0;        // bytes in ACGT$ (counted by the transform kernel)
    uint64_t h2d_bytes_ = 0, d2h_bytes_ = 0;
    DevBuf<uint64_t> d_nstarts, d_nends;
    std::vector<uint64_t> n_ranges_host;
    DevBuf<pos_t> d_sa;
    DevBuf<uint32_t> d_lcp;
    // 64-bit device results of the fast path: round 0 writes them directly, the (few) entries the refinement
    // changes afterwards are patched in at the end, so the full-array widening pass disappears.  Dropped (and the
    // widening pass used) as soon as something rewrites the arrays wholesale: prefix doubling, the post-sort
    // filter, the N-run rule, shard slicing.
    DevBuf<unsigned long long> wide_sa_, wide_lcp_;
    DevBuf<uint32_t> wide_slots_;
    uint64_t wide_m_ = 0;
    bool wide_ok_ = false;
    void drop_wide() {
        wide_ok_ = false;
        wide_sa_.reset();
        wide_lcp_.reset();
        wide_slots_.reset();
        wide_m_ = 0;
    }
    uint64_t sort_kmask_ = ~0ull;  // general path: key bits covered by the first sort (see make_keys_and_sort)
    bool partial_sort_ = false;
    DevBuf<uint64_t> keys_spare_;  // fast path: the radix sort's ping-pong partners, output of the fused round 0
    DevBuf<pos_t> pos_spare_;
    DevBuf<uint32_t> d_isa;     // inverse suffix array (only when prefix doubling ran)
    DevBuf<uint32_t> d_counts;  // radix sort count matrix
    uint64_t shard_offset = 0, shard_count = 0, total_suffixes = 0;
    bool layout_exact_ = true;  // shard_offset / total_suffixes are known without an exchange
    bool cuts_ready_ = false;   // the shard's key range has been fixed (first histogram of this build)
    uint32_t cut_b0_ = 0, cut_b1_ = 0;
    bool full_set_ = true;  // every text position is being sorted on this rank (prefix doubling needs that)
    int t_keys_mark = -1, t_sorted_mark = -1;
    rsort::EventPairs downsweep_events;
    uint64_t sorted_elements = 0;

    template <typename T>
    DevBuf<T> dalloc(size_t count) { return DevBuf<T>(ctx.pool, count ? count : 1); }
    void launched(uint64_t k = 1) { ctx.launches += k; }

    template <typename Op, typename In, typename Out>
    typename Op::T scan_total(uint64_t count, In in, Op op, Out out) {
        auto partials = dalloc<typename Op::T>(scan::partials_count(count));
        scan::inclusive_scan(count, in, op, out, partials.get(), st());
        launched(count ? 3 : 0);
        typename Op::T total;
        size_t idx = count ? (size_t)div_up(count, scan::CHUNK) : 0;
        SUFR_CUDA_CHECK(cudaMemcpyAsync(&total, partials.get() + idx, sizeof(total), cudaMemcpyDeviceToHost, st()));
        SUFR_CUDA_CHECK(cudaStreamSynchronize(st()));
        return total;
    }

    // two-step scan: pass 1 (reduce + spine) returns the grand total, pass 2 runs the output functor
    template <typename Op, typename In>
    typename Op::T scan_begin(uint64_t count, In in, Op op, DevBuf<typename Op::T>& partials) {
        partials = dalloc<typename Op::T>(scan::partials_count(count));
        scan::scan_reduce(count, in, op, partials.get(), st());
        launched(count ? 2 : 0);
        typename Op::T total;
        size_t idx = count ? (size_t)div_up(count, scan::CHUNK) : 0;
        SUFR_CUDA_CHECK(cudaMemcpyAsync(&total, partials.get() + idx, sizeof(total), cudaMemcpyDeviceToHost, st()));
        SUFR_CUDA_CHECK(cudaStreamSynchronize(st()));
        return total;
    }
    template <typename Op, typename In, typename Out>
    void scan_finish(uint64_t count, In in, Op op, Out out, DevBuf<typename Op::T>& partials) {
        scan::scan_apply(count, in, op, out, partials.get(), st());
        launched(count ? 1 : 0);
    }

    bool looks_repetitive();
    void encode(const uint8_t* d_raw);
    void find_n_runs();
    void make_keys_and_sort(DevBuf<uint64_t>& keys_sorted, bool prefilter, bool sharded);
    void sort_phase(bool prefilter, bool sharded);
    void refine(DevBuf<uint64_t>& keys_sorted);
    void doubling(DevBuf<uint32_t>& slot, DevBuf<pos_t>& pos, DevBuf<uint32_t>& seg, uint64_t m, uint64_t nseg,
                  uint64_t h);
    void isa_init(uint32_t* isa);
    void n_run_rule();
    void apply_filter();
    void segmented_sort_u64key(DevBuf<uint64_t>& ck, DevBuf<pos_t>& pos, uint64_t m, int key_bits);
    void sort_groups(DevBuf<uint64_t>& keys, DevBuf<pos_t>& pos, const uint32_t* seg, const uint32_t* slot, uint64_t m,
                     uint64_t nseg, int key_lo, int key_hi, bool rank_keys);
};

static int bits_for(uint64_t v) {  // number of bits needed to represent values 0..v
    int b = 0;
    while (v) { b++; v >>= 1; }
    return b ? b : 1;
}

void Build::encode(const uint8_t* d_raw) {
    d_text = dalloc<uint8_t>(n + 16);
    auto d_present = dalloc<uint32_t>(256);
    auto d_sample = dalloc<unsigned long long>(257);  // [256] = number of indexed suffix starts
    SUFR_CUDA_CHECK(cudaMemsetAsync(d_present.get(), 0, 256 * sizeof(uint32_t), st()));
    SUFR_CUDA_CHECK(cudaMemsetAsync(d_sample.get(), 0, 257 * sizeof(unsigned long long), st()));
    SUFR_CUDA_CHECK(cudaMemsetAsync(d_text.get() + n, 0, 16, st()));
    if (n) {
        transform_kernel<<<grid_for(n, 16), kBlock, 0, st()>>>(d_raw, d_text.get(), n, args.ignore_softmask,
                                                               d_present.get(), d_sample.get(), d_sample.get() + 256);
        SUFR_KERNEL_CHECK();
        launched();
    }
    uint32_t present[256];
    unsigned long long sample[257];
    SUFR_CUDA_CHECK(cudaMemcpyAsync(present, d_present.get(), sizeof(present), cudaMemcpyDeviceToHost, st()));
    SUFR_CUDA_CHECK(cudaMemcpyAsync(sample, d_sample.get(), sizeof(sample), cudaMemcpyDeviceToHost, st()));
    SUFR_CUDA_CHECK(cudaStreamSynchronize(st()));
    indexed_count_ = sample[256];
    uint8_t lut[256];
    alphabet = 0;
    for (int b = 0; b < 256; b++) lut[b] = present[b] ? (uint8_t)(++alphabet) : 0;
    code_n_ = lut['N'];
    uint32_t bits = (uint32_t)bits_for(alphabet);
    PackedText& pt = ks.pt;
    pt.n = n;
    pt.bits = bits;
    pt.K = 64 / bits;
    uint32_t used = pt.K * bits;
    pt.keep_mask = used == 64 ? ~0ull : ~((1ull << (64 - used)) - 1ull);
    pt.sym_mask = (1u << bits) - 1u;
    uint64_t num_words = div_up(n, pt.K);
    d_words = dalloc<uint64_t>(num_words + 2);
    SUFR_CUDA_CHECK(cudaMemsetAsync(d_words.get() + num_words, 0, 2 * sizeof(uint64_t), st()));
    auto d_lut = dalloc<uint8_t>(256);
    SUFR_CUDA_CHECK(cudaMemcpyAsync(d_lut.get(), lut, 256, cudaMemcpyHostToDevice, st()));
    pt.words = d_words.get();
    ks.text = d_text.get();

    // 2-bit fast path for the first sort: full sort of a text dominated by four byte values (DNA).  The four
    // most frequent bytes of a 1/64 sample are the "regular" symbols; any choice is correct, it only decides
    // how many keys contain fill.
    ks.fast2 = 0;
    DevBuf<uint8_t> d_cls2;
    uint64_t words2 = 0;
    if (ks.mode == kModeFull && n >= 4096 && !getenv("SUFR_B200_DEBUG_NO_FAST2")) {
        int order[256];
        for (int b = 0; b < 256; b++) order[b] = b;
        std::sort(order, order + 256, [&](int a, int b) { return sample[a] != sample[b] ? sample[a] > sample[b] : a < b; });
        unsigned long long total = 0, top4 = 0;
        for (int b = 0; b < 256; b++) total += sample[b];
        for (int k = 0; k < 4; k++) top4 += sample[order[k]];
        if (total > 0 && sample[order[3]] > 0 && top4 * 16 >= total * 15) {
            int reg[4] = {order[0], order[1], order[2], order[3]};
            std::sort(reg, reg + 4);
            uint8_t cls[256], cls2[256];
            for (int b = 0; b < 256; b++) {
                int c = 0, rank = -1;
                for (int k = 0; k < 4; k++) {
                    if (reg[k] < b) c++;
                    if (reg[k] == b) rank = k;
                }
                cls[b] = (uint8_t)(rank >= 0 ? rank : c);
                cls2[b] = (uint8_t)(rank >= 0 ? rank : ((c > 3 ? 3 : c) | 4));
            }
            words2 = (div_up(n, 32) + 3) & ~1ull;  // even, with padding words generated by the kernel
            d_packed2 = dalloc<uint64_t>(words2 + 2);
            d_irr = dalloc<uint64_t>(words2 / 2 + 2);
            d_cls = dalloc<uint8_t>(256);
            d_cls2 = dalloc<uint8_t>(256);
            SUFR_CUDA_CHECK(cudaMemcpyAsync(d_cls.get(), cls, 256, cudaMemcpyHostToDevice, st()));
            SUFR_CUDA_CHECK(cudaMemcpyAsync(d_cls2.get(), cls2, 256, cudaMemcpyHostToDevice, st()));
            SUFR_CUDA_CHECK(cudaMemsetAsync(d_packed2.get() + words2, 0, 2 * 8, st()));
            SUFR_CUDA_CHECK(cudaMemsetAsync(d_irr.get() + words2 / 2, 0xFF, 2 * 8, st()));
            ks.fast2 = 1;
            ks.reg_indexed = 1;
            for (int k = 0; k < 4; k++)
                if (!strchr("ACGT$", reg[k]) || reg[k] == 0) ks.reg_indexed = 0;
            ks.packed2_words = words2 + 2;
            ks.irr_words = words2 / 2 + 2;
            ks.packed2 = d_packed2.get();
            ks.irr = d_irr.get();
            ks.cls = d_cls.get();
        }
    }
    if (num_words) {
        // grid: a multiple of the SM count; every block walks its tiles with a two-deep cp.async pipeline
        uint32_t grid = (uint32_t)std::min<uint64_t>(div_up(num_words, kBlock), (uint64_t)num_sms() * 6);
        pack_kernel<<<grid, kBlock, 0, st()>>>(d_text.get(), n, d_lut.get(), bits, pt.K, num_words, d_words.get(),
                                               ks.fast2 ? d_cls2.get() : nullptr, words2, d_packed2.get(),
                                               reinterpret_cast<uint32_t*>(d_irr.get()));
        SUFR_KERNEL_CHECK();
        launched();
    }
    SUFR_CUDA_CHECK(cudaStreamSynchronize(st()));  // lut / present are freed on return
}

// Cheap probe (4 M sampled keys, one small sort): does the text have so many long repeats that the build will
// need prefix doubling?  Decides whether the suffix filter is applied before the sort (cheap, but a text that
// then needs doubling has to be redone over all positions) or after it.
bool Build::looks_repetitive() {
    if (ks.mode != kModeFull || n < (1u << 24)) return false;
    const uint64_t count = 1u << 22;
    const uint64_t stride_pos = n / count;
    auto keys = dalloc<uint64_t>(count), keys_b = dalloc<uint64_t>(count);
    auto pos = dalloc<pos_t>(count), pos_b = dalloc<pos_t>(count);
    auto counts = dalloc<uint32_t>(rsort::counts_words());
    auto d_eq = dalloc<unsigned long long>(1);
    SUFR_CUDA_CHECK(cudaMemsetAsync(d_eq.get(), 0, 8, st()));
    sample_keys_kernel<<<grid_for(count, 2), kBlock, 0, st()>>>(ks, stride_pos, count, keys.get(), pos.get());
    SUFR_KERNEL_CHECK();
    launched();
    const int used = (int)(ks.pt.K * ks.pt.bits);
    const int begin_bit = ks.fast2 ? 64 - kFast2ProbeBits : 64 - used;
    bool in_b = rsort::sort_pairs<uint64_t, pos_t>(keys.get(), keys_b.get(), pos.get(), pos_b.get(), count, begin_bit, 64,
                                                      counts.get(), st(), &ctx.launches);
    count_equal_neighbours_kernel<<<grid_for(count, 4), kBlock, 0, st()>>>(in_b ? keys_b.get() : keys.get(), count,
                                                                          ks.fast2 ? ~0ull << (64 - kFast2ProbeBits) : ~0ull, d_eq.get());
    SUFR_KERNEL_CHECK();
    launched();
    unsigned long long eq = 0;
    SUFR_CUDA_CHECK(cudaMemcpyAsync(&eq, d_eq.get(), 8, cudaMemcpyDeviceToHost, st()));
    SUFR_CUDA_CHECK(cudaStreamSynchronize(st()));
    // a random text gives count^2 / (2 * 4^20) ~ 8 equal neighbours among 4 M samples of 20 symbols
    return eq > 512;
}

void Build::find_n_runs() {
    // sufr_builder.rs:174-195: maximal runs of 'N' of length >= 1000 that are terminated by another byte
    ks.n_starts = nullptr;
    ks.n_ends = nullptr;
    ks.num_n_ranges = 0;
    if (!args.allow_ambiguity || n == 0) return;
    const uint8_t* t = d_text.get();
    uint32_t nstart = scan_total(n, NRunStartIn{t}, scan::SumU32{}, CountOnly{});
    if (nstart == 0) return;
    auto starts = dalloc<uint64_t>(nstart);
    auto ends = dalloc<uint64_t>(nstart);
    scan_total(n, NRunStartIn{t}, scan::SumU32{}, IndexOut{starts.get()});
    uint32_t nend = scan_total(n, NRunEndIn{t}, scan::SumU32{}, IndexOut{ends.get()});
    // the k-th start pairs with the k-th end; a run that reaches the end of the text has no end
    uint32_t npairs = nend < nstart ? nend : nstart;
    if (npairs == 0) return;
    NRunLongIn lin{starts.get(), ends.get(), 1000};
    uint32_t nlong = scan_total(npairs, lin, scan::SumU32{}, CountOnly{});
    if (nlong == 0) return;
    d_nstarts = dalloc<uint64_t>(nlong);
    d_nends = dalloc<uint64_t>(nlong);
    scan_total(npairs, lin, scan::SumU32{}, NRunLongOut{starts.get(), ends.get(), d_nstarts.get(), d_nends.get()});
    n_ranges_host.resize(2 * (size_t)nlong);
    std::vector<uint64_t> hs(nlong), he(nlong);
    SUFR_CUDA_CHECK(cudaMemcpyAsync(hs.data(), d_nstarts.get(), nlong * 8, cudaMemcpyDeviceToHost, st()));
    SUFR_CUDA_CHECK(cudaMemcpyAsync(he.data(), d_nends.get(), nlong * 8, cudaMemcpyDeviceToHost, st()));
    SUFR_CUDA_CHECK(cudaStreamSynchronize(st()));
    for (uint32_t i = 0; i < nlong; i++) {
        n_ranges_host[2 * i] = hs[i];
        n_ranges_host[2 * i + 1] = he[i];
    }
    if (!has_mask) {  // find_lcp consults the runs only in the MaxQueryLen branch (sufr_builder.rs:301-307)
        ks.n_starts = d_nstarts.get();
        ks.n_ends = d_nends.get();
        ks.num_n_ranges = nlong;
    }
}

void Build::make_keys_and_sort(DevBuf<uint64_t>& keys_sorted, bool prefilter, bool sharded) {
    const int descending = ks.mode != kModeFull;
    const int world = args.world_size > 1 ? args.world_size : 1;
    uint64_t lo = 0, hi = 0;
    shard_offset = 0;
    layout_exact_ = world == 1;
    uint64_t shard_estimate = 0;
    if (world > 1) {
        // Splitters from a histogram of the top 12 key bits: every rank computes the same histogram of the
        // replicated text, so ranks agree on the key ranges without communicating.  The full sort only needs
        // balanced ranges, so it histograms every 64th position; the modes whose tie order depends on the
        // input order (mask / max-query-len) use the exact histogram, which also yields the exact shard
        // offsets.  The cut points are fixed by the FIRST histogram of a build: the full-sort fallback (which
        // only some ranks may take) re-counts exactly but keeps the same cuts.
        const uint32_t hbits = kShardHistBits, bins = 1u << hbits;
        const bool exact = ks.mode != kModeFull || cuts_ready_;
        const uint32_t sample_shift = exact ? 0 : 6;  // every 64th position: 48 M samples of a 3.1 Gbp text
        auto d_hist = dalloc<unsigned long long>(bins);
        SUFR_CUDA_CHECK(cudaMemsetAsync(d_hist.get(), 0, bins * 8, st()));
        if (n) {
            key_hist_kernel<<<grid_for(n >> sample_shift, 8), kBlock, bins * sizeof(uint32_t), st()>>>(
                ks, n, hbits, d_text.get(), filter_active ? 1 : 0, sample_shift, d_hist.get());
            SUFR_KERNEL_CHECK();
            launched();
        }
        std::vector<unsigned long long> hist(bins);
        SUFR_CUDA_CHECK(cudaMemcpyAsync(hist.data(), d_hist.get(), bins * 8, cudaMemcpyDeviceToHost, st()));
        SUFR_CUDA_CHECK(cudaStreamSynchronize(st()));
        unsigned long long total = 0;
        for (uint32_t b = 0; b < bins; b++) total += hist[b];
        if (!cuts_ready_) {
            // bin boundaries b_0 = 0 <= b_1 <= ... <= b_world = bins with balanced counts
            // (counts are of INDEXED suffixes, so shard offsets refer to the final suffix array)
            std::vector<uint32_t> cut(world + 1, bins);
            cut[0] = 0;
            unsigned long long acc = 0;
            int g = 1;
            for (uint32_t b = 0; b < bins && g < world; b++) {
                acc += hist[b];
                while (g < world && acc * world >= total * g) cut[g++] = b + 1;
            }
            // N-run rule (sufr_builder.rs:305-307, :701-712): suffixes inside recorded runs with equal (run length,
            // next byte) form tie chains that are re-ordered by position AFTER the sort.  A chain N^r X.. with r < 4
            // spans several histogram bins, so no cut may fall inside the range of N-prefixed keys: every chain then
            // lives on one rank (also when a rank falls back to the unsharded build and slices its range out).
            // On the 2-bit fast path all N-prefixed suffixes share one key, hence one bin.
            if (ks.num_n_ranges && !ks.fast2 && code_n_ && ks.pt.bits <= hbits) {
                const uint32_t nlo = code_n_ << (hbits - ks.pt.bits), nhi = (code_n_ + 1) << (hbits - ks.pt.bits);
                for (int k = 1; k < world; k++)
                    if (cut[k] > nlo && cut[k] < nhi) cut[k] = nlo;
            }
            cut_b0_ = cut[args.rank];
            cut_b1_ = cut[args.rank + 1];
            cuts_ready_ = true;
        }
        const uint32_t b0 = cut_b0_, b1 = cut_b1_;
        unsigned long long before = 0, mine = 0;
        for (uint32_t b = 0; b < b0; b++) before += hist[b];
        for (uint32_t b = b0; b < b1; b++) mine += hist[b];
        if (exact) {
            total_suffixes = total;
            shard_offset = before;
            shard_count = mine;
            layout_exact_ = true;
        } else {
            shard_estimate = mine << sample_shift;
        }
        lo = (uint64_t)b0 << (64 - hbits);
        hi = b1 >= bins ? 0 : (uint64_t)b1 << (64 - hbits);
        if (b0 >= b1) { lo = ~0ull; hi = ~0ull; }  // empty shard: [max, max) selects nothing
    }

    DevBuf<uint64_t> keys_a, keys_b;
    DevBuf<pos_t> pos_a, pos_b;
    const int used_bits = (int)(ks.pt.K * ks.pt.bits);
    bool first_digit_done = false;  // fast path: the records come out of key generation sorted by the first digit
    uint64_t kept = n;     // suffixes that survive the filter (all ranks' ranges together)
    uint64_t sort_n = n;   // elements handed to the sort
    if (prefilter && n) kept = indexed_count_;
    // Few filtered suffixes (the common case: delimiters, sparse N): no compaction at all, they get the
    // key ~0 and drop off the end of the sorted array.  Needs an unused low bit in the packed word.
    const bool sentinel = prefilter && !sharded && used_bits < 64 && kept < n && (n - kept) * 16 <= n;
    sentinel_ = sentinel;
    // Measured and NOT the default: selection FUSED with key generation and the first radix pass, as on one GPU -- the
    // shard's records are written once, already in first-digit order, and the sort proper has three passes left.  The
    // fused kernels pay their per-POSITION cost (two key computations, tile staging) over the whole text, the selection
    // kernel only a 32-bit test: one shard of the 3.1 Gbp text takes 84.5 -> 83.7 ms of 2, 46.3 -> 54.1 ms of 4,
    // 27.1 -> 39.1 ms of 8.  Kept behind SUFR_B200_DEBUG_SHARD_FUSE=1 (tested: test_shard_selection_fused_...).
    const char* fuse_env = getenv("SUFR_B200_DEBUG_SHARD_FUSE");
    const bool fuse_shard = sharded && ks.mode == kModeFull && ks.fast2 && !descending && n > 0 && sizeof(pos_t) == 4 &&
                            (fuse_env && atoi(fuse_env) != 0);
    if (fuse_shard) {
        static bool attr_set[64] = {};
        allow_dynamic_smem(fast2_keygen_scatter_kernel, kKsSmem, attr_set);
        const uint64_t tiles = div_up(n, kKsTile);
        const uint32_t grid = (uint32_t)std::min<uint64_t>(tiles, (uint64_t)num_sms() * 4);
        const uint64_t chunk = div_up(tiles, grid) * kKsTile;
        const uint32_t used_grid = (uint32_t)div_up(n, chunk);
        const int shift = 64 - kFast2SortBits;
        const uint32_t bins = 1u << kShardHistBits;
        const uint32_t bin0 = (uint32_t)(lo >> (64 - kShardHistBits));
        const uint32_t bin1 = hi == 0 ? bins : (uint32_t)(hi >> (64 - kShardHistBits));
        // an empty shard ([max, max)) keeps span = 0 out of the kernels' "everything" meaning: nothing to do at all
        KeyRange range{bin0, bin1 > bin0 ? bin1 - bin0 : 0u};
        unsigned long long c = 0;
        if (range.span) {
            d_counts = dalloc<uint32_t>(rsort::counts_words());
            auto d_cnt = dalloc<unsigned long long>(1);
            SUFR_CUDA_CHECK(cudaMemsetAsync(d_cnt.get(), 0, 8, st()));
            fast2_first_digit_hist_kernel<<<used_grid, kBlock, 0, st()>>>(ks, n, prefilter ? 1 : 0, chunk, shift, d_counts.get(),
                                                                         range, d_cnt.get());
            SUFR_KERNEL_CHECK();
            rsort::scan_counts_kernel<<<1, 1024, 0, st()>>>(d_counts.get(), (uint32_t)rsort::RADIX * used_grid);
            SUFR_KERNEL_CHECK();
            SUFR_CUDA_CHECK(cudaMemcpyAsync(&c, d_cnt.get(), 8, cudaMemcpyDeviceToHost, st()));
            SUFR_CUDA_CHECK(cudaStreamSynchronize(st()));
            if (c > 0xFFFFFFFFull) throw Error(SUFR_B200_ERR_UNSUPPORTED, "a shard holds 2^32 suffixes or more: use more GPUs");
        }
        s = c;
        sort_n = s;
        keys_a = dalloc<uint64_t>(s);
        pos_a = dalloc<pos_t>(s);
        if (s) {
            fast2_keygen_scatter_kernel<<<used_grid, kBlock, kKsSmem, st()>>>(ks, n, prefilter ? 1 : 0, keys_a.get(), pos_a.get(),
                                                                            chunk, shift, d_counts.get(), range);
            SUFR_KERNEL_CHECK();
            launched(3);
            first_digit_done = true;
        }
    } else if (sharded && ks.mode == kModeFull) {
        // unordered selection: one key computation per position, capacity from the sampled histogram
        uint64_t capacity = shard_estimate + shard_estimate / 16 + (1u << 20);
        auto d_cnt = dalloc<unsigned long long>(1);
        for (int attempt = 0;; attempt++) {
            keys_a = dalloc<uint64_t>(capacity);
            pos_a = dalloc<pos_t>(capacity);
            SUFR_CUDA_CHECK(cudaMemsetAsync(d_cnt.get(), 0, 8, st()));
            if (n) {
                if (ks.fast2)
                    select_fast2_kernel<<<grid_for(n, 32), kBlock, 0, st()>>>(ks, n, lo, hi, prefilter ? 1 : 0, keys_a.get(),
                                                                            pos_a.get(), d_cnt.get(), capacity);
                else
                    select_append_kernel<<<grid_for(n, 32), kBlock, 0, st()>>>(ks, n, lo, hi, prefilter ? 1 : 0, keys_a.get(),
                                                                             pos_a.get(), d_cnt.get(), capacity);
                SUFR_KERNEL_CHECK();
                launched();
            }
            unsigned long long c = 0;
            SUFR_CUDA_CHECK(cudaMemcpyAsync(&c, d_cnt.get(), 8, cudaMemcpyDeviceToHost, st()));
            SUFR_CUDA_CHECK(cudaStreamSynchronize(st()));
            s = c;
            if (c <= capacity) break;
            if (attempt) throw Error(SUFR_B200_ERR_INTERNAL, "shard selection overflowed twice");
            keys_a.reset();
            pos_a.reset();
            capacity = c;
        }
        sort_n = s;
    } else if (sharded || (prefilter && !sentinel && kept < n)) {
        SelectIn in{ks, n, descending, sharded ? 1 : 0, lo, hi, d_text.get(), prefilter ? 1 : 0};
        s = sharded ? shard_count : kept;  // both exact: histogram of indexed suffixes / indexed count
        keys_a = dalloc<uint64_t>(s);
        pos_a = dalloc<pos_t>(s);
        if (n) {
            uint32_t got = scan_total(n, in, scan::SumU32{}, SelectOut{ks, n, descending, keys_a.get(), pos_a.get()});
            if (got != s) throw Error(SUFR_B200_ERR_INTERNAL, "selection count mismatch");
        }
        sort_n = s;
    } else {
        s = sentinel ? kept : n;
        sort_n = n;
        keys_a = dalloc<uint64_t>(n);
        pos_a = dalloc<pos_t>(n);
        if (n) {
            if (ks.fast2 && !descending) {
                // key generation fused with the first radix pass (kernels.cuh): histogram of the first digit per
                // block, scan, then generate + scatter; the sort proper starts at the second digit
                static bool attr_set[64] = {};
                allow_dynamic_smem(fast2_keygen_scatter_kernel, kKsSmem, attr_set);
                const uint64_t tiles = div_up(n, kKsTile);
                const uint32_t grid = (uint32_t)std::min<uint64_t>(tiles, (uint64_t)num_sms() * 4);
                const uint64_t chunk = div_up(tiles, grid) * kKsTile;
                const uint32_t used_grid = (uint32_t)div_up(n, chunk);
                const int shift = 64 - kFast2SortBits;
                d_counts = dalloc<uint32_t>(rsort::counts_words());
                fast2_first_digit_hist_kernel<<<used_grid, kBlock, 0, st()>>>(ks, n, sentinel ? 1 : 0, chunk, shift,
                                                                             d_counts.get(), KeyRange{0, 0}, nullptr);
                SUFR_KERNEL_CHECK();
                rsort::scan_counts_kernel<<<1, 1024, 0, st()>>>(d_counts.get(), (uint32_t)rsort::RADIX * used_grid);
                SUFR_KERNEL_CHECK();
                fast2_keygen_scatter_kernel<<<used_grid, kBlock, kKsSmem, st()>>>(ks, n, sentinel ? 1 : 0, keys_a.get(),
                                                                                pos_a.get(), chunk, shift, d_counts.get(),
                                                                                KeyRange{0, 0});
                launched(2);
                first_digit_done = true;
            } else
                keygen_kernel<<<grid_for(n, 4), kBlock, 0, st()>>>(ks, n, descending, d_text.get(), sentinel ? 1 : 0,
                                                                  keys_a.get(), pos_a.get());
            SUFR_KERNEL_CHECK();
            launched();
        }
    }
    t_keys_mark = timer.mark();
    keys_b = dalloc<uint64_t>(sort_n);
    pos_b = dalloc<pos_t>(sort_n);
    if (!first_digit_done) d_counts = dalloc<uint32_t>(rsort::counts_words());
    // 3-bit keys: all used bits (sentinel keys have the unused low bits set, so those join the sort then).
    // 2-bit fast path: only the top kFast2SortBits; ties go to the exact refinement.
    // General path: about log2(n) + 8 bits (rounded up to whole passes) separate all but ~1/256 of the
    // neighbours; the elements that still agree on them are refined from key word 0 like any other tie.
    // (Not with sentinel keys: their order relies on the low bits.)
    // A capped key (seed mask of weight W, --max-query-len Q) has only cap * bits meaningful bits; the rest is 0.
    int key_bits = used_bits;
    if (ks.mode != kModeFull && ks.cap < (uint64_t)ks.pt.K) key_bits = (int)ks.cap * (int)ks.pt.bits;
    int begin_bit = ks.fast2 ? 64 - kFast2SortBits : (sentinel ? 0 : 64 - key_bits);
    partial_sort_ = false;
    sort_kmask_ = ~0ull;
    if (!ks.fast2 && !sentinel && !getenv("SUFR_B200_DEBUG_FULL_WORD_SORT")) {
        const int want = ((bits_for(sort_n ? sort_n - 1 : 0) + 8 + rsort::RADIX_BITS - 1) / rsort::RADIX_BITS) * rsort::RADIX_BITS;
        if (want < key_bits) {
            begin_bit = 64 - want;
            partial_sort_ = true;
            sort_kmask_ = ~0ull << begin_bit;
        }
    }
    bool in_b = rsort::sort_pairs<uint64_t, pos_t>(keys_a.get(), keys_b.get(), pos_a.get(), pos_b.get(), sort_n,
                                                      begin_bit + (first_digit_done ? rsort::RADIX_BITS : 0), 64, d_counts.get(),
                                                      st(), &ctx.launches, &downsweep_events);
    sorted_elements = sort_n;
    sort_n_ = sort_n;
    if (in_b) {
        keys_sorted = std::move(keys_b);
        d_sa = std::move(pos_b);
        if (ks.fast2) { keys_spare_ = std::move(keys_a); pos_spare_ = std::move(pos_a); }
    } else {
        keys_sorted = std::move(keys_a);
        d_sa = std::move(pos_a);
        if (ks.fast2) { keys_spare_ = std::move(keys_b); pos_spare_ = std::move(pos_b); }
    }
    // the ping-pong partners are released here (end of scope), except on the fast path (see refine)
}

// Stable sort of (ck, pos) on the low `key_bits` bits of the composite key.
void Build::segmented_sort_u64key(DevBuf<uint64_t>& ck, DevBuf<pos_t>& pos, uint64_t m, int key_bits) {
    auto ck_b = dalloc<uint64_t>(m);
    auto pos_b = dalloc<pos_t>(m);
    bool in_b = rsort::sort_pairs<uint64_t, pos_t>(ck.get(), ck_b.get(), pos.get(), pos_b.get(), m, 0, key_bits,
                                                      d_counts.get(), st(), &ctx.launches);
    if (in_b) {
        ck = std::move(ck_b);
        pos = std::move(pos_b);
    }
}

// Sorts every unresolved group by key bits [key_lo, key_hi) of its members' keys (keys, positions, SA slots).
// rank_keys: prefix-doubling keys (group << 32 | rank); else key words.
void Build::sort_groups(DevBuf<uint64_t>& keys, DevBuf<pos_t>& pos, const uint32_t* seg, const uint32_t* slot, uint64_t m,
                        uint64_t nseg, int key_lo, int key_hi, bool rank_keys) {
    auto is_large = dalloc<uint8_t>(nseg ? nseg : 1);
    auto d_any = dalloc<unsigned long long>(1);
    SUFR_CUDA_CHECK(cudaMemsetAsync(is_large.get(), 0, nseg ? nseg : 1, st()));
    SUFR_CUDA_CHECK(cudaMemsetAsync(d_any.get(), 0, 8, st()));
    // small groups: sorted in registers by the thread at the group start
    small_groups_kernel<<<grid_for(m, 1), kBlock, 0, st()>>>(m, seg, slot, pos.get(), keys.get(), d_sa.get(), is_large.get(),
                                                            d_any.get());
    SUFR_KERNEL_CHECK();
    launched();
    unsigned long long any_large = 0;
    SUFR_CUDA_CHECK(cudaMemcpyAsync(&any_large, d_any.get(), 8, cudaMemcpyDeviceToHost, st()));
    SUFR_CUDA_CHECK(cudaStreamSynchronize(st()));
    if (!any_large) return;
    // larger groups: stable radix sort by (group, key): LSD, the key bits first, then the group bits
    DevBuf<unsigned long long> part;
    LargeIn lin{seg, is_large.get()};
    const unsigned long long lt = scan_begin(m, lin, scan::SumU64{}, part);
    const uint64_t ml = (uint32_t)lt, nlseg = lt >> 32;
    if (ml == 0) return;
    if (rank_keys && ml * 4 > m) {
        // most of the round sits in large groups (tandem repeats): sort everything in place of a compaction -- the lean
        // path, 12 bytes of temporary per element; the small groups are simply sorted again
        part.reset();
        segmented_sort_u64key(keys, pos, m, key_hi + (nseg > 1 ? bits_for(nseg - 1) : 0));
        writeback_pos_kernel<<<grid_for(m, 2), kBlock, 0, st()>>>(m, pos.get(), slot, d_sa.get());
        SUFR_KERNEL_CHECK();
        launched();
        return;
    }
    const int sb = nlseg > 1 ? bits_for(nlseg - 1) : 0;
    auto idx = dalloc<uint32_t>(ml);
    if (rank_keys) {
        auto lck = dalloc<uint64_t>(ml), lck_b = dalloc<uint64_t>(ml);
        auto lpos = dalloc<pos_t>(ml), lpos_b = dalloc<pos_t>(ml);
        scan_finish(m, lin, scan::SumU64{}, LargeOutRank{pos.get(), keys.get(), idx.get(), lck.get(), lpos.get(), key_hi}, part);
        // (compact group << rank bits | rank): one sort over rank and group bits together
        bool in_b = rsort::sort_pairs<uint64_t, pos_t>(lck.get(), lck_b.get(), lpos.get(), lpos_b.get(), ml, key_lo,
                                                          key_hi + sb, d_counts.get(), st(), &ctx.launches);
        if (in_b) { std::swap(lck, lck_b); std::swap(lpos, lpos_b); }
        scatter_large_rank_kernel<<<grid_for(ml, 2), kBlock, 0, st()>>>(ml, lck.get(), lpos.get(), idx.get(), slot, keys.get(),
                                                                       pos.get(), d_sa.get(), key_hi);
    } else {
        auto lkeys = dalloc<uint64_t>(ml), keys_b = dalloc<uint64_t>(ml);
        auto segpos = dalloc<uint64_t>(ml), segpos_b = dalloc<uint64_t>(ml);  // group << 32 | compact index
        auto lpos = dalloc<pos_t>(ml);
        scan_finish(m, lin, scan::SumU64{}, LargeOut{pos.get(), keys.get(), idx.get(), lkeys.get(), segpos.get(), lpos.get()},
                    part);
        bool in_b = rsort::sort_pairs<uint64_t, uint64_t>(lkeys.get(), keys_b.get(), segpos.get(), segpos_b.get(), ml, key_lo,
                                                          key_hi, d_counts.get(), st(), &ctx.launches);
        if (in_b) { std::swap(lkeys, keys_b); std::swap(segpos, segpos_b); }
        if (sb) {
            in_b = rsort::sort_pairs<uint64_t, uint64_t>(segpos.get(), segpos_b.get(), lkeys.get(), keys_b.get(), ml, 32, 32 + sb,
                                                         d_counts.get(), st(), &ctx.launches);
            if (in_b) { std::swap(lkeys, keys_b); std::swap(segpos, segpos_b); }
        }
        scatter_large_kernel<<<grid_for(ml, 2), kBlock, 0, st()>>>(ml, lkeys.get(), segpos.get(), lpos.get(), idx.get(), slot,
                                                                  keys.get(), pos.get(), d_sa.get());
    }
    SUFR_KERNEL_CHECK();
    launched();
}

void Build::refine(DevBuf<uint64_t>& keys_sorted) {
    const uint32_t K = ks.pt.K;
    const int used = (int)(K * ks.pt.bits);
    // In fast2 + sentinel mode the filtered suffixes sit inside the last group until the refinement has
    // pushed them to the very end, so round 0 looks at all sort_n_ elements there.
    const bool fast2 = ks.fast2 != 0;
    const uint64_t r0n = (fast2 && sentinel_) ? sort_n_ : s;
    d_lcp = dalloc<uint32_t>(r0n);
    if (r0n == 0) return;

    // round 0: boundaries of the initial sort, fused with the collection of the unresolved elements.
    // `word` is the last key word (3-bit packing) the groups are known to agree on; the fast path has only
    // sorted a 2-bit approximation of the first symbols, so its refinement starts with word 0.
    int word = (fast2 || partial_sort_) ? -1 : 0;
    int final_word = (!fast2 && (uint64_t)(word + 1) * K >= ks.cap) ? 1 : 0;
    // the fast path's round 0 is out of place: the ordered positions land in the sort's ping-pong partner; the
    // ordered keys are not written (the LCP marks carry the group structure), so the key partner is free already
    if (fast2) keys_spare_.reset();
    ViewAll v0{keys_sorted.get(), d_sa.get(), nullptr, sort_kmask_};  // general path only
    // Full sort on the fast path: groups whose members agree on all 31 symbols of the 2-bit key (no fill, not a large
    // run) start their exact refinement at key word 1 -- word 0 (21 symbols) could not split them.  kSegDeep.
    const int deep_marks = (fast2 && ks.mode == kModeFull && !sentinel_ && !getenv("SUFR_B200_DEBUG_NO_DEEP_GROUPS")) ? 1 : 0;
    uint64_t m = 0, nseg = 0;
    DevBuf<uint32_t> slot, seg;
    DevBuf<pos_t> pos;
    {
        uint64_t capacity = final_word ? 1 : std::max<uint64_t>(1u << 20, r0n / 8);
        if (const char* dbg = getenv("SUFR_B200_DEBUG_SPARSE_CAP")) capacity = std::max<uint64_t>(1, strtoull(dbg, nullptr, 10));
        auto act_slot = dalloc<uint32_t>(capacity);
        auto act_pos = dalloc<pos_t>(capacity);
        auto d_cnt = dalloc<unsigned long long>(1);
        SUFR_CUDA_CHECK(cudaMemsetAsync(d_cnt.get(), 0, 8, st()));
        if (fast2) {
            drop_wide();
            if (sizeof(pos_t) == 4 && index_bits_ == 64 && result_memory_ == SUFR_B200_MEM_DEVICE && !args.allow_ambiguity &&
                !getenv("SUFR_B200_DEBUG_NO_EARLY_WIDE")) {
                // only when the two extra arrays fit beside the sort buffers that are still alive
                size_t free_b = 0, total_b = 0;
                SUFR_CUDA_CHECK(cudaMemGetInfo(&free_b, &total_b));
                const size_t fits = std::min<size_t>(2, ctx.pool.count_fits(8 * r0n));
                if (free_b > (2 - fits) * 8 * r0n + (4ull << 30)) {
                    wide_sa_ = dalloc<unsigned long long>(r0n);
                    wide_lcp_ = dalloc<unsigned long long>(r0n);
                    wide_ok_ = true;
                }
            }
            auto round0 = deep_marks ? round0_fast2_kernel<true> : round0_fast2_kernel<false>;
            round0<<<grid_for(r0n, 4), kBlock, 0, st()>>>(keys_sorted.get(), d_sa.get(), pos_spare_.get(), r0n, d_lcp.get(),
                                                         act_slot.get(), act_pos.get(), d_cnt.get(), capacity, wide_sa_.get(),
                                                         wide_lcp_.get());
        } else {
            resolve0_append_kernel<<<grid_for(r0n, 4), kBlock, 0, st()>>>(keys_sorted.get(), d_sa.get(), r0n, ks, final_word,
                                                                         sort_kmask_, d_lcp.get(), act_slot.get(), act_pos.get(),
                                                                         d_cnt.get(), capacity);
        }
        SUFR_KERNEL_CHECK();
        launched();
        if (fast2) {  // the ordered positions are in the partner now
            keys_sorted.reset();
            d_sa = std::move(pos_spare_);
        }
        if (final_word) {
            keys_sorted.reset();
            return;
        }
        unsigned long long cnt = 0;
        SUFR_CUDA_CHECK(cudaMemcpyAsync(&cnt, d_cnt.get(), 8, cudaMemcpyDeviceToHost, st()));
        SUFR_CUDA_CHECK(cudaStreamSynchronize(st()));
        if (cnt == 0) {
            keys_sorted.reset();
            return;
        }
        if (cnt <= capacity) {
            // sparse: order the list by SA slot, then number the segments
            m = cnt;
            auto slot_b = dalloc<uint32_t>(m);
            auto pos_b = dalloc<pos_t>(m);
            bool in_b = rsort::sort_pairs<uint32_t, pos_t>(act_slot.get(), slot_b.get(), act_pos.get(), pos_b.get(), m, 0,
                                                              bits_for(r0n - 1), d_counts.get(), st(), &ctx.launches);
            if (in_b) { std::swap(act_slot, slot_b); std::swap(act_pos, pos_b); }
            slot = std::move(act_slot);
            pos = std::move(act_pos);
            if (wide_ok_) {  // every later change of SA / LCP happens at one of these slots (or is an LCP fix-up)
                wide_slots_ = dalloc<uint32_t>(m);
                SUFR_CUDA_CHECK(cudaMemcpyAsync(wide_slots_.get(), slot.get(), m * 4, cudaMemcpyDeviceToDevice, st()));
                wide_m_ = m;
            }
            seg = dalloc<uint32_t>(m);
            if (fast2) nseg = scan_total(m, LcpSegIn{d_lcp.get(), slot.get()}, scan::SumU32{},
                                         SparseSegOut{seg.get(), d_lcp.get(), slot.get(), r0n});
            else nseg = scan_total(m, SparseSegIn{v0, slot.get()}, scan::SumU32{}, SparseSegOut{seg.get()});
        } else {
            // dense (repetitive text): order-preserving compaction by scan
            drop_wide();
            act_slot.reset();
            act_pos.reset();
            DevBuf<unsigned long long> part;
            if (fast2) {
                LcpActiveIn ain{d_lcp.get(), r0n};
                unsigned long long tot = scan_begin(r0n, ain, scan::SumU64{}, part);
                m = (uint32_t)tot;
                nseg = tot >> 32;
                slot = dalloc<uint32_t>(m);
                pos = dalloc<pos_t>(m);
                seg = dalloc<uint32_t>(m);
                scan_finish(r0n, ain, scan::SumU64{}, LcpActiveOut{d_sa.get(), slot.get(), pos.get(), seg.get(), d_lcp.get(), r0n}, part);
            } else {
                ActiveIn<ViewAll> ain{v0, r0n, 0, sentinel_ ? 1 : 0};
                unsigned long long tot = scan_begin(r0n, ain, scan::SumU64{}, part);
                m = (uint32_t)tot;
                nseg = tot >> 32;
                slot = dalloc<uint32_t>(m);
                pos = dalloc<pos_t>(m);
                seg = dalloc<uint32_t>(m);
                scan_finish(r0n, ain, scan::SumU64{}, ActiveOut<ViewAll>{v0, slot.get(), pos.get(), seg.get()}, part);
            }
        }
    }
    keys_sorted.reset();
    unsigned long long tot = 0;
    // An unresolved MAJORITY after the first sort means a text of tandem repeats: it will need prefix doubling,
    // which needs every position.  Give up this (filtered / sharded) attempt now rather than after the
    // patient word rounds.
    if (!full_set_ && ks.mode == kModeFull && m * 2 >= r0n) throw NeedFullSort{};

    // Full sort: after kMaxWordRounds words switch to prefix doubling (depth doubles per round).
    // When this attempt sorts only a subset of the positions (filter applied up front / one shard), doubling
    // means redoing the build over all positions -- on every rank of a multi-GPU build.  Dispersed repeats (copies of
    // earlier segments with some divergence, as in real genomes and BASELINE's repetitive variant) resolve a good
    // share of what is left with every further word, and the cost of a round falls with it: such an attempt stays
    // with word rounds (up to 320 words = 6720 bases) while they pay, so a shard finishes on its own and N GPUs
    // divide the work.  Tandem repeats make no progress per word and leave for prefix doubling at once.
    // (With deep groups the bulk of a DNA text is one word ahead: three rounds take it to 84 symbols, the depth four
    // rounds used to reach; prefix doubling then starts from h = 63, which every group has reached.)
    const int kMaxWordRounds = getenv("SUFR_B200_DEBUG_WORD_ROUNDS") ? atoi(getenv("SUFR_B200_DEBUG_WORD_ROUNDS")) : (deep_marks ? 2 : 3);
    const int kPatientWordRounds = 320;
    uint64_t m_last = m;  // unresolved elements at the start of the previous round
    int direct_tail_word = -1;
    while (m > 0) {
        // (a round that resolved less than an eighth of what it was given: tandem repeats, further words will not help)
        const bool stalled = full_set_ && word >= 1 && m * 8 > m_last * 7 && !getenv("SUFR_B200_DEBUG_WORD_ROUNDS");
        if (ks.mode == kModeFull && (word >= kMaxWordRounds || stalled)) {
            bool patient = !full_set_ && word < kPatientWordRounds && (m * 16 < s || m * 16 <= m_last * 15);
            if (!patient) {
                doubling(slot, pos, seg, m, nseg, (uint64_t)(word + 1) * K);
                return;
            }
        }
        // The deep tail of a subset attempt: few elements left, mostly pairs (a segment and its copy) that would take
        // one launch-bound round per further key word.  Finish the small groups by direct comparison.
        // (Not in every round: groups of more than eight members keep going through word rounds and shed small groups as
        // they split; a sweep every 16 words collects those.)
        if (ks.mode == kModeFull && !full_set_ && word >= kMaxWordRounds && m * 32 <= s && !sentinel_ &&
            (direct_tail_word < 0 || word - direct_tail_word >= 16) && !getenv("SUFR_B200_DEBUG_NO_DIRECT_TAIL")) {
            direct_tail_word = word;
            auto is_large = dalloc<uint8_t>(nseg ? nseg : 1);
            auto d_left = dalloc<unsigned long long>(1);
            SUFR_CUDA_CHECK(cudaMemsetAsync(is_large.get(), 0, nseg ? nseg : 1, st()));
            SUFR_CUDA_CHECK(cudaMemsetAsync(d_left.get(), 0, 8, st()));
            finish_small_groups_kernel<<<grid_for(m, 1), kBlock, 0, st()>>>(ks, m, (uint64_t)(word + 1) * K, seg.get(), slot.get(),
                                                                           pos.get(), d_sa.get(), d_lcp.get(), is_large.get(),
                                                                           d_left.get());
            SUFR_KERNEL_CHECK();
            launched();
            unsigned long long left = 0;
            SUFR_CUDA_CHECK(cudaMemcpyAsync(&left, d_left.get(), 8, cudaMemcpyDeviceToHost, st()));
            SUFR_CUDA_CHECK(cudaStreamSynchronize(st()));
            if (left == 0) return;
            DevBuf<unsigned long long> part;
            LargeIn lin{seg.get(), is_large.get()};
            const unsigned long long lt = scan_begin(m, lin, scan::SumU64{}, part);
            const uint64_t m2 = (uint32_t)lt, nseg2 = lt >> 32;
            auto slot2 = dalloc<uint32_t>(m2);
            auto pos2 = dalloc<pos_t>(m2);
            auto seg2 = dalloc<uint32_t>(m2);
            scan_finish(m, lin, scan::SumU64{}, LeftoverOut{slot.get(), pos.get(), slot2.get(), pos2.get(), seg2.get(), seg.get()}, part);
            slot = std::move(slot2);
            pos = std::move(pos2);
            seg = std::move(seg2);
            m = m2;
            nseg = nseg2;
            if (m == 0) return;
            // ... and the larger ones, one block per group (finish_large_groups_kernel)
            if (!getenv("SUFR_B200_DEBUG_NO_BLOCK_TAIL")) {
                auto starts = dalloc<uint64_t>(nseg);
                const uint32_t heads = scan_total(m, SegHeadIn{seg.get()}, scan::SumU32{}, IndexOut{starts.get()});
                if (heads != nseg) throw Error(SUFR_B200_ERR_INTERNAL, "group count mismatch in the direct-comparison tail");
                SUFR_CUDA_CHECK(cudaMemsetAsync(is_large.get(), 0, nseg, st()));
                SUFR_CUDA_CHECK(cudaMemsetAsync(d_left.get(), 0, 8, st()));
                const uint32_t blocks = (uint32_t)std::min<uint64_t>(nseg, (uint64_t)num_sms() * 16);
                finish_large_groups_kernel<<<blocks, kBlock, 0, st()>>>(ks, m, (uint64_t)(word + 1) * K, starts.get(), nseg, seg.get(),
                                                                       slot.get(), pos.get(), d_sa.get(), d_lcp.get(), is_large.get(),
                                                                       d_left.get());
                SUFR_KERNEL_CHECK();
                launched();
                SUFR_CUDA_CHECK(cudaMemcpyAsync(&left, d_left.get(), 8, cudaMemcpyDeviceToHost, st()));
                SUFR_CUDA_CHECK(cudaStreamSynchronize(st()));
                if (left == 0) return;
                if (left < m) {  // keep what the blocks did not finish
                    DevBuf<unsigned long long> part3;
                    LargeIn lin3{seg.get(), is_large.get()};
                    const unsigned long long lt3 = scan_begin(m, lin3, scan::SumU64{}, part3);
                    const uint64_t m3 = (uint32_t)lt3, nseg3 = lt3 >> 32;
                    auto slot3 = dalloc<uint32_t>(m3);
                    auto pos3 = dalloc<pos_t>(m3);
                    auto seg3 = dalloc<uint32_t>(m3);
                    scan_finish(m, lin3, scan::SumU64{}, LeftoverOut{slot.get(), pos.get(), slot3.get(), pos3.get(), seg3.get(), seg.get()},
                                part3);
                    slot = std::move(slot3);
                    pos = std::move(pos3);
                    seg = std::move(seg3);
                    m = m3;
                    nseg = nseg3;
                    if (m == 0) return;
                }
            }
        }
        word++;
        refine_rounds++;
        m_last = m;
        if (getenv("SUFR_B200_LOG_ROUNDS")) {
            SUFR_CUDA_CHECK(cudaStreamSynchronize(st()));
            fprintf(stderr, "[sufr_b200] word round %u (key word %d): unresolved = %llu in %llu groups, t = %.3f s\n",
                    refine_rounds, word, (unsigned long long)m, (unsigned long long)nseg,
                    std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count());
        }
        final_word = ((uint64_t)(word + 1) * K >= ks.cap) ? 1 : 0;
        auto keys = dalloc<uint64_t>(m);
        round_keys_kernel<<<grid_for(m, 2), kBlock, 0, st()>>>(ks, m, (uint32_t)word, sentinel_ ? 1 : 0, pos.get(), seg.get(),
                                                               keys.get());
        SUFR_KERNEL_CHECK();
        launched();
        sort_groups(keys, pos, seg.get(), slot.get(), m, nseg, 64 - used, 64, false);
        ViewActive va{keys.get(), pos.get(), seg.get(), slot.get()};
        resolve_kernel<ViewActive><<<grid_for(m, 2), kBlock, 0, st()>>>(va, m, ks, (uint32_t)word, final_word, 0,
                                                                       d_lcp.get());
        SUFR_KERNEL_CHECK();
        launched();
        if (final_word) return;
        DevBuf<unsigned long long> part;
        ActiveIn<ViewActive> ain{va, m, 0, sentinel_ ? 1 : 0};
        tot = scan_begin(m, ain, scan::SumU64{}, part);
        uint64_t m2 = (uint32_t)tot, nseg2 = tot >> 32;
        if (m2 == 0) return;
        auto slot2 = dalloc<uint32_t>(m2);
        auto pos2 = dalloc<pos_t>(m2);
        auto seg2 = dalloc<uint32_t>(m2);
        scan_finish(m, ain, scan::SumU64{}, ActiveOut<ViewActive>{va, slot2.get(), pos2.get(), seg2.get()}, part);
        slot = std::move(slot2);
        pos = std::move(pos2);
        seg = std::move(seg2);
        m = m2;
        nseg = nseg2;
    }
}

// isa[sa[j]] = j for the whole suffix array.  A direct scatter of 4-byte ranks runs at ~22 G stores/s on a B200 (a
// 32-byte sector read and written per rank; ~50 G/s even when the targets stay inside the L2:
// profiles/r2_scatter_window.txt).  Large arrays therefore SORT the records (position, rank) on the top bits of the
// position with the radix pass of the build, far enough that 2^chunk_bits consecutive records hold exactly the positions
// [c << chunk_bits, (c + 1) << chunk_bits) -- the suffix array is a permutation, so the chunk boundaries need no
// histogram -- and a block places one chunk in shared memory and writes it out as whole lines.
void Build::isa_init(uint32_t* isa) {
    // SUFR_B200_DEBUG_SORT_ISA=1 takes the sorting route on small texts too (tests)
    const bool force = getenv("SUFR_B200_DEBUG_SORT_ISA") != nullptr;
    const bool direct = sizeof(pos_t) != 4 || n < 8192 || (!force && n < (1ull << 26)) || getenv("SUFR_B200_DEBUG_DIRECT_ISA");
    if (direct) {
        isa_init_kernel<<<grid_for(n, 4), kBlock, 0, st()>>>(n, d_sa.get(), isa);
        SUFR_KERNEL_CHECK();
        launched();
        return;
    }
    if constexpr (sizeof(pos_t) == 4) {
        const int nbits = bits_for(n - 1);
        const int max_chunk = force ? 10 : kIsaChunkBitsMax;
        const int passes = nbits > max_chunk ? (nbits - max_chunk + rsort::RADIX_BITS - 1) / rsort::RADIX_BITS : 0;
        const int chunk_bits = std::min(max_chunk, std::max(nbits - passes * rsort::RADIX_BITS, force ? 6 : 12));
        const int low = nbits - passes * rsort::RADIX_BITS > 0 ? nbits - passes * rsort::RADIX_BITS : 0;  // sorted bits: [low, nbits)
        auto va = dalloc<uint32_t>(n), kb = dalloc<uint32_t>(n), vb = dalloc<uint32_t>(n);
        DevBuf<uint32_t> ka;
        if (passes > 1) ka = dalloc<uint32_t>(n);  // the suffix array itself is only ever read
        if (!d_counts) d_counts = dalloc<uint32_t>(rsort::counts_words());
        iota_kernel<<<grid_for(n, 4), kBlock, 0, st()>>>(va.get(), n, 0u);
        SUFR_KERNEL_CHECK();
        const uint32_t* rk = reinterpret_cast<const uint32_t*>(d_sa.get());
        const uint32_t* rv = va.get();
        for (int pass = 0; pass < passes; pass++) {  // one sort_pairs call = ONE pass: in -> out
            const int b0 = low + pass * rsort::RADIX_BITS, b1 = std::min(nbits, b0 + rsort::RADIX_BITS);
            uint32_t* ko = (pass & 1) ? ka.get() : kb.get();
            uint32_t* vo = (pass & 1) ? va.get() : vb.get();
            rsort::sort_pairs<uint32_t, uint32_t>(const_cast<uint32_t*>(rk), ko, const_cast<uint32_t*>(rv), vo, n, b0, b1,
                                                  d_counts.get(), st(), &ctx.launches);
            rk = ko;
            rv = vo;
        }
        static bool attr_set[64] = {};
        allow_dynamic_smem(isa_chunk_kernel, (size_t)4 << kIsaChunkBitsMax, attr_set);
        const uint64_t chunks = div_up(n, 1ull << chunk_bits);
        isa_chunk_kernel<<<(uint32_t)std::min<uint64_t>(chunks, (uint64_t)num_sms() * 8), 512, (size_t)4 << chunk_bits, st()>>>(
            n, chunk_bits, rk, rv, isa);
        SUFR_KERNEL_CHECK();
        launched(2);
    }
}

// Prefix doubling on the still-unresolved groups (Larsson-Sadakane style: only active groups are sorted).
// Needs the rank of EVERY text position, hence only valid when all positions were sorted on this rank.
void Build::doubling(DevBuf<uint32_t>& slot, DevBuf<pos_t>& pos, DevBuf<uint32_t>& seg, uint64_t m, uint64_t nseg,
                     uint64_t h) {
    if (!full_set_) throw NeedFullSort{};
    d_isa = dalloc<uint32_t>(n);
    uint32_t* const isa_ptr = d_isa.get();
    isa_init(isa_ptr);
    scan_total(m, GroupStartIn{seg.get()}, scan::MaxU32{}, GroupRankOut{slot.get(), pos.get(), isa_ptr});
    const bool log_rounds = getenv("SUFR_B200_LOG_ROUNDS") != nullptr;
    while (m > 0) {
        if (doubling_rounds > 64) throw Error(SUFR_B200_ERR_INTERNAL, "prefix doubling did not converge");
        // The working LCP array is 32 bits wide with bit 31 and the two top values reserved (common.cuh): a group that is
        // still unresolved at h >= 2^30 can have an LCP of 2^31 or more.  Refuse rather than mis-encode.
        if (h >= (1ull << 30))
            throw Error(SUFR_B200_ERR_UNSUPPORTED, "suffixes that share 2^30 symbols or more: longest common prefixes of 2^31 "
                                                   "and above do not fit the 32-bit working LCP array");
        doubling_rounds++;
        doubling_depth_ = h;
        if (log_rounds) {
            SUFR_CUDA_CHECK(cudaStreamSynchronize(st()));
            fprintf(stderr, "[sufr_b200] doubling round %u: h = %llu, unresolved = %llu in %llu groups, t = %.3f s\n",
                    doubling_rounds, (unsigned long long)h, (unsigned long long)m, (unsigned long long)nseg,
                    std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count());
        }
        auto ck = dalloc<uint64_t>(m);
        const int rank_bits = bits_for(n);  // ranks are 0 .. n
        doubling_keys_kernel<<<grid_for(m, 2), kBlock, 0, st()>>>(m, n, h, pos.get(), seg.get(), isa_ptr, ck.get(), rank_bits);
        SUFR_KERNEL_CHECK();
        launched();
        // the group number sits in the high half of every key and is equal inside a group: sort on the rank bits
        sort_groups(ck, pos, seg.get(), slot.get(), m, nseg, 0, rank_bits, true);
        uint32_t mark = kLcpLowerBound | (uint32_t)(h < 0x7FFFFFFFull ? h : 0x7FFFFFFFull);
        scan_total(m, DoublingStartIn{ck.get()}, scan::MaxU32{},
                   DoublingRankOut{ck.get(), slot.get(), pos.get(), isa_ptr, d_lcp.get(), mark, rank_bits});
        DevBuf<unsigned long long> part;
        DoublingActiveIn din{ck.get(), m};
        unsigned long long tot = scan_begin(m, din, scan::SumU64{}, part);
        uint64_t m2 = (uint32_t)tot, nseg2 = tot >> 32;
        if (m2 == 0) break;
        auto slot2 = dalloc<uint32_t>(m2);
        auto pos2 = dalloc<pos_t>(m2);
        auto seg2 = dalloc<uint32_t>(m2);
        scan_finish(m, din, scan::SumU64{}, DoublingActiveOut{slot.get(), pos.get(), slot2.get(), pos2.get(), seg2.get()},
                    part);
        slot = std::move(slot2);
        pos = std::move(pos2);
        seg = std::move(seg2);
        m = m2;
        nseg = nseg2;
        h *= 2;
    }
}

void Build::n_run_rule() {
    if (ks.num_n_ranges == 0 || s < 2) return;
    n_rule_lcp_kernel<<<grid_for(s, 2), kBlock, 0, st()>>>(ks, d_text.get(), s, d_sa.get(), d_lcp.get());
    SUFR_KERNEL_CHECK();
    launched();
    NTieIn in{ks, d_text.get(), d_sa.get(), s};
    unsigned long long tot = scan_total(s, in, scan::SumU64{}, CountOnlyU64{});
    uint64_t m = (uint32_t)tot, nseg = tot >> 32;
    if (m == 0) return;
    auto slot = dalloc<uint32_t>(m);
    auto pos = dalloc<pos_t>(m);
    auto ck = dalloc<uint64_t>(m);
    const int pos_bits = sizeof(pos_t) == 4 ? 32 : bits_for(n);
    if (pos_bits + (nseg > 1 ? bits_for(nseg - 1) : 0) > 64)
        throw Error(SUFR_B200_ERR_UNSUPPORTED, "N-run tie rule: chain number and position do not fit one 64-bit sort key");
    scan_total(s, in, scan::SumU64{}, NTieOut{d_sa.get(), slot.get(), pos.get(), ck.get(), pos_bits});
    segmented_sort_u64key(ck, pos, m, pos_bits + (nseg > 1 ? bits_for(nseg - 1) : 0));
    writeback_pos_kernel<<<grid_for(m, 2), kBlock, 0, st()>>>(m, pos.get(), slot.get(), d_sa.get());
    SUFR_KERNEL_CHECK();
    launched();
}

void Build::apply_filter() {
    if (!filter_active || s == 0) return;
    if (!getenv("SUFR_B200_DEBUG_SLOW_FILTER")) {
        const uint64_t per_block = (uint64_t)kBlock * kFilterRows;
        const uint32_t nblocks = div_up_u32(s, per_block);
        auto flags = dalloc<uint32_t>((size_t)nblocks * per_block / 32 + 1);
        auto counts = dalloc<uint32_t>(nblocks);
        auto tail_min = dalloc<uint32_t>(nblocks);
        auto offsets = dalloc<uint32_t>(nblocks);
        KeepRanges ranges{};
        if (s == n && !getenv("SUFR_B200_DEBUG_GATHER_FILTER")) {
            // the whole text was sorted: rank ranges of the indexed first bytes from a byte histogram (KeepRanges)
            auto d_hist = dalloc<unsigned long long>(256);
            SUFR_CUDA_CHECK(cudaMemsetAsync(d_hist.get(), 0, 256 * 8, st()));
            byte_hist_kernel<<<grid_for(n, 64), kBlock, 0, st()>>>(d_text.get(), n, d_hist.get());
            SUFR_KERNEL_CHECK();
            launched();
            unsigned long long hist[256];
            SUFR_CUDA_CHECK(cudaMemcpyAsync(hist, d_hist.get(), sizeof(hist), cudaMemcpyDeviceToHost, st()));
            SUFR_CUDA_CHECK(cudaStreamSynchronize(st()));
            uint64_t start = 0;
            for (int c = 0; c < 256; c++) {
                if (hist[c] && (c == '$' || c == 'A' || c == 'C' || c == 'G' || c == 'T')) {
                    ranges.lo[ranges.count] = start;
                    ranges.hi[ranges.count] = start + hist[c];
                    ranges.count++;
                }
                start += hist[c];
            }
            if (start != n) throw Error(SUFR_B200_ERR_INTERNAL, "byte histogram does not add up to the text length");
        }
        filter_flags_kernel<<<nblocks, kBlock, 0, st()>>>(d_text.get(), d_sa.get(), ranges, d_lcp.get(), s, flags.get(),
                                                         counts.get(), tail_min.get());
        SUFR_KERNEL_CHECK();
        launched();
        uint32_t kept = scan_total(nblocks, BlockCountIn{counts.get()}, scan::SumU32{}, BlockOffsetOut{offsets.get()});
        if (kept == s) return;
        auto sa2 = dalloc<pos_t>(kept);
        auto lcp2 = dalloc<uint32_t>(kept);
        filter_compact_kernel<<<nblocks, kBlock, 0, st()>>>(d_sa.get(), d_lcp.get(), s, flags.get(), offsets.get(),
                                                           counts.get(), tail_min.get(), sa2.get(), lcp2.get());
        SUFR_KERNEL_CHECK();
        launched();
        SUFR_CUDA_CHECK(cudaStreamSynchronize(st()));
        d_sa = std::move(sa2);
        d_lcp = std::move(lcp2);
        s = kept;
        return;
    }
    // reference implementation of the same filter by segmented-min scans (debug knob above)
    FilterCountIn cin{d_text.get(), d_sa.get()};
    uint32_t kept = scan_total(s, cin, scan::SumU32{}, CountOnly{});
    if (kept == s) return;
    auto scanned = dalloc<uint32_t>(s);
    scan_total(s, FilterLcpIn{d_text.get(), d_sa.get(), d_lcp.get()}, scan::SegMinU64{},
               FilterLcpOut{d_text.get(), d_sa.get(), nullptr, scanned.get()});
    auto sa2 = dalloc<pos_t>(kept);
    auto idx = dalloc<uint32_t>(kept);
    scan_total(s, cin, scan::SumU32{}, FilterSaOut{d_sa.get(), sa2.get(), idx.get()});
    auto lcp2 = dalloc<uint32_t>(kept);
    if (kept) {
        gather_u32_kernel<<<grid_for(kept, 2), kBlock, 0, st()>>>(kept, idx.get(), scanned.get(), lcp2.get());
        SUFR_KERNEL_CHECK();
        launched();
    }
    d_sa = std::move(sa2);
    d_lcp = std::move(lcp2);
    s = kept;
}

void Build::sort_phase(bool prefilter, bool sharded) {
    full_set_ = !prefilter && !sharded;
    DevBuf<uint64_t> keys_sorted;
    make_keys_and_sort(keys_sorted, prefilter, sharded);
    t_sorted_mark = timer.mark();
    refine(keys_sorted);
    if (ks.fast2 && s) {
        lcp_fixup_kernel<<<grid_for(s, 4), kBlock, 0, st()>>>(ks, s, d_sa.get(), d_lcp.get(),
                                                              wide_ok_ ? wide_lcp_.get() : nullptr);
        SUFR_KERNEL_CHECK();
        launched();
    }
}

void Build::run(SufrB200Result* out) {
    n = args.text_len;
    constexpr bool kPos64 = sizeof(pos_t) == 8;
    if (!kPos64 && n >= 0xFFFFFFFFull) throw Error(SUFR_B200_ERR_INTERNAL, "32-bit positions for a text of 2^32 - 1 bytes or more");
    if (index_bits_ == 0) index_bits_ = n < 0xFFFFFFFFull ? 32 : 64;  // suffix_array.rs:461
    if (index_bits_ != 32 && index_bits_ != 64) throw Error(SUFR_B200_ERR_ARGUMENT, "index_bits must be 0, 32 or 64");
    if (index_bits_ == 32 && n >= 0xFFFFFFFFull)
        throw Error(SUFR_B200_ERR_ARGUMENT, "a text of 2^32 - 1 bytes or more needs index_bits 64 (suffix_array.rs:460-470)");
    // A rank sorts fewer than 2^32 suffixes (SA slots, group numbers and list lengths are 32-bit): longer texts are
    // built as key-range shards over several GPUs, which they need for capacity anyway (16 bytes per sort record).
    if (n >= 0xFFFFFFFFull && args.world_size < 2)
        throw Error(SUFR_B200_ERR_UNSUPPORTED,
                    "a text of 2^32 - 1 bytes or more must be built as key-range shards (world_size >= 2, "
                    "sufr_b200_create_multi / --devices): one rank sorts fewer than 2^32 suffixes");
    if (args.world_size > 1 && (args.rank < 0 || args.rank >= args.world_size))
        throw Error(SUFR_B200_ERR_ARGUMENT, "rank out of range");

    // sufr_builder.rs:163-172
    if (args.seed_mask && args.has_max_query_len)
        throw Error(SUFR_B200_ERR_ARGUMENT, "Cannot use max_query_len and seed_mask together");
    if (args.seed_mask) {
        if (!parse_seed_mask(args.seed_mask, mask))
            throw Error(SUFR_B200_ERR_ARGUMENT, std::string("Invalid seed mask '") + args.seed_mask + "'");
        has_mask = true;
    }
    ks.mode = has_mask ? kModeMask : (args.has_max_query_len && args.max_query_len > 0 ? kModeMaxQueryLen : kModeFull);
    ks.cap = has_mask ? mask.weight : (ks.mode == kModeMaxQueryLen ? args.max_query_len : ~0ull);
    ks.weight = has_mask ? (uint32_t)mask.weight : 0;
    ks.mask_len = has_mask ? (uint32_t)mask.bytes.size() : 0;
    ks.mask_pos = nullptr;
    filter_active = args.is_dna && !args.allow_ambiguity;  // sufr_builder.rs:446-449

    SUFR_CUDA_CHECK(cudaSetDevice(ctx.device));
    ctx.pool.reset_peak();
    const uint64_t launches0 = ctx.launches;
    // working set: text n, packed <= n, keys 2x8n, positions 2x4n, lcp 4n (+ output widening 16n for u64)
    {
        uint64_t per = 30;  // the u64 widening happens after the key buffers are gone
        uint64_t shard_n = args.world_size > 1 ? n / args.world_size + n / 8 : n;
        ctx.pool.reserve((size_t)(2 * n + per * shard_n + (64ull << 20)));
    }

    SufrB200Timings tm{};
    // ---- text on the device
    DevBuf<uint8_t> d_raw_owned;
    const uint8_t* d_raw = args.text;
    if (text_memory_ == SUFR_B200_MEM_HOST) {
        d_raw_owned = dalloc<uint8_t>(n + 16);
        int e0 = timer.mark();
        if (n) SUFR_CUDA_CHECK(cudaMemcpyAsync(d_raw_owned.get(), args.text, n, cudaMemcpyHostToDevice, st()));
        h2d_bytes_ += n;
        int e1 = timer.mark();
        SUFR_CUDA_CHECK(cudaStreamSynchronize(st()));
        tm.h2d_ms = timer.ms(e0, e1);
        d_raw = d_raw_owned.get();
    }

    int t0 = timer.mark();
    encode(d_raw);
    d_raw_owned.reset();
    if (has_mask) {
        std::vector<uint32_t> mp(mask.positions.begin(), mask.positions.end());
        d_maskpos = dalloc<uint32_t>(mp.size());
        SUFR_CUDA_CHECK(cudaMemcpyAsync(d_maskpos.get(), mp.data(), mp.size() * 4, cudaMemcpyHostToDevice, st()));
        SUFR_CUDA_CHECK(cudaStreamSynchronize(st()));
        ks.mask_pos = d_maskpos.get();
    }
    find_n_runs();
    int t1 = timer.mark();

    // First attempt sorts only what this rank outputs (indexed suffixes of its key range).  Texts with
    // repeats deeper than the word-refinement limit need the ranks of ALL positions: redo unfiltered and
    // unsharded, filter afterwards, and cut this rank's slice out of the global result.
    bool prefilter = filter_active, sharded = args.world_size > 1, sliced = false;
    if (prefilter && !sharded && !getenv("SUFR_B200_DEBUG_NO_PROBE") && looks_repetitive()) prefilter = false;
    try {
        sort_phase(prefilter, sharded);
    } catch (const NeedFullSort&) {
        SUFR_CUDA_CHECK(cudaStreamSynchronize(st()));
        d_sa.reset();
        d_lcp.reset();
        d_isa.reset();
        drop_wide();
        keys_spare_.reset();
        pos_spare_.reset();
        refine_rounds = doubling_rounds = 0;
        doubling_depth_ = 0;
        for (auto& ev : downsweep_events) { cudaEventDestroy(ev.first); cudaEventDestroy(ev.second); }
        downsweep_events.clear();
        sliced = sharded;
        prefilter = false;
        sharded = false;
        sort_phase(false, false);  // the pool was sized for an unsharded build of n positions up front
    }
    int t2 = t_sorted_mark;
    int t3 = timer.mark();

    // finish: lower-bound LCP marks left by prefix doubling (text order, on the unfiltered arrays), then the
    // suffix filter when it was not applied up front, then the N-run rule
    if (doubling_rounds || (filter_active && !prefilter) || sliced) drop_wide();
    if (doubling_rounds && s) {
        if (!d_isa || s != n) throw Error(SUFR_B200_ERR_INTERNAL, "prefix doubling ran on a partial suffix set");
        if (doubling_depth_ <= 8192 && !getenv("SUFR_B200_DEBUG_PLCP")) {
            // shallow repeats: extend every marked pair from its lower bound (at most 2h symbols each)
            lcp_bounds_direct_kernel<<<grid_for(s, 4), kBlock, 0, st()>>>(ks, s, d_sa.get(), d_lcp.get());
        } else {
            uint64_t chunks = div_up(n, kPlcpChunk);
            plcp_complete_kernel<<<grid_for(chunks, 1), kBlock, 0, st()>>>(ks, n, d_sa.get(), d_isa.get(), d_lcp.get());
        }
        SUFR_KERNEL_CHECK();
        launched();
    }
    d_isa.reset();
    int t4 = timer.mark();
    if (!prefilter) apply_filter();
    int t5 = timer.mark();
    n_run_rule();
    if (sliced) {
        if (shard_offset + shard_count > s) throw Error(SUFR_B200_ERR_INTERNAL, "shard slice out of range");
        auto sa2 = dalloc<pos_t>(shard_count);
        auto lcp2 = dalloc<uint32_t>(shard_count);
        if (shard_count) {
            SUFR_CUDA_CHECK(cudaMemcpyAsync(sa2.get(), d_sa.get() + shard_offset, shard_count * sizeof(pos_t), cudaMemcpyDeviceToDevice, st()));
            SUFR_CUDA_CHECK(cudaMemcpyAsync(lcp2.get(), d_lcp.get() + shard_offset, shard_count * 4, cudaMemcpyDeviceToDevice, st()));
        }
        d_sa = std::move(sa2);
        d_lcp = std::move(lcp2);
        s = shard_count;
    }

    // shard bookkeeping: without an exact histogram the caller fills shard_offset / total_suffixes after the
    // (count, first, last) exchange it needs for the seam repair anyway
    if (args.world_size <= 1) total_suffixes = s;
    else if (!layout_exact_) { total_suffixes = s; shard_offset = 0; }
    uint64_t first = 0, last = 0;
    if (s) {
        pos_t fl[2];
        SUFR_CUDA_CHECK(cudaMemcpyAsync(&fl[0], d_sa.get(), sizeof(pos_t), cudaMemcpyDeviceToHost, st()));
        SUFR_CUDA_CHECK(cudaMemcpyAsync(&fl[1], d_sa.get() + (s - 1), sizeof(pos_t), cudaMemcpyDeviceToHost, st()));
        SUFR_CUDA_CHECK(cudaStreamSynchronize(st()));
        first = fl[0];
        last = fl[1];
    }

    // outputs
    auto owner = std::make_unique<ResultOwner>();
    owner->ctx = &ctx;
    owner->memory = result_memory_;
    const size_t w = index_bits_ / 8;
    DevBuf<unsigned long long> sa64, lcp64;
    void* d_sa_out = d_sa.get();
    void* d_lcp_out = d_lcp.get();
    // Host results of large builds travel compactly (LCP as bytes + exceptions, 64-bit SA as u32) and are
    // widened by host threads while the next copy is in flight: PCIe, not the GPU, bounds the end-to-end time.
    uint64_t compact_min = 1u << 22;
    if (const char* dbg = getenv("SUFR_B200_DEBUG_COMPACT_MIN")) compact_min = strtoull(dbg, nullptr, 10);
    // Transfer model: one GPU pushes everything through one PCIe link (~50 GB/s), so fewer bytes on the wire win even
    // though host threads then have to widen them (2 * s * w bytes of stores); from four ranks on, every rank has its own
    // link and little to send, while the ranks would compete for the same host cores: plain transfer, no widening.
    bool compact = sizeof(pos_t) == 4 && result_memory_ == SUFR_B200_MEM_HOST && s >= compact_min && s > 0 &&
                   (args.world_size < 4 || getenv("SUFR_B200_DEBUG_COMPACT_MIN")) &&
                   !getenv("SUFR_B200_DEBUG_NO_COMPACT_D2H");
    DevBuf<uint8_t> d_lcp8;
    DevBuf<uint32_t> d_exc_idx, d_exc_val;
    uint64_t exc_count = 0;
    if (compact) {
        // LCP values >= 255 travel as (index, value) pairs and cost the host a random write each: worth it only
        // while they are rare (a repetitive text takes the plain path)
        const uint64_t capacity = s / 2048 + 1024;
        d_lcp8 = dalloc<uint8_t>(s + 16);
        d_exc_idx = dalloc<uint32_t>(capacity);
        d_exc_val = dalloc<uint32_t>(capacity);
        auto d_cnt = dalloc<unsigned long long>(1);
        SUFR_CUDA_CHECK(cudaMemsetAsync(d_cnt.get(), 0, 8, st()));
        lcp_to_u8_kernel<<<grid_for(s, 16), kBlock, 0, st()>>>(d_lcp.get(), s, d_lcp8.get(), d_exc_idx.get(),
                                                              d_exc_val.get(), d_cnt.get(), capacity);
        SUFR_KERNEL_CHECK();
        launched();
        unsigned long long c = 0;
        SUFR_CUDA_CHECK(cudaMemcpyAsync(&c, d_cnt.get(), 8, cudaMemcpyDeviceToHost, st()));
        SUFR_CUDA_CHECK(cudaStreamSynchronize(st()));
        exc_count = c;
        if (c > capacity) {  // repetitive text: most LCP values do not fit a byte
            compact = false;
            d_lcp8.reset();
            d_exc_idx.reset();
            d_exc_val.reset();
        }
    }
    DevBuf<uint32_t> sa32;  // 64-bit positions narrowed for a u32 result (short texts built with SUFR_B200_DEBUG_POS64)
    if (sizeof(pos_t) == 8) {
        drop_wide();
        if (index_bits_ == 64) {  // positions are 64-bit already; LCP values widen
            lcp64 = dalloc<unsigned long long>(s);
            if (s) {
                convert_kernel<<<grid_for(s, 4), kBlock, 0, st()>>>(s, d_lcp.get(), lcp64.get());
                SUFR_KERNEL_CHECK();
                launched();
            }
            d_lcp.reset();
            d_sa_out = d_sa.get();
            d_lcp_out = lcp64.get();
        } else {
            sa32 = dalloc<uint32_t>(s);
            if (s) {
                convert_kernel<<<grid_for(s, 4), kBlock, 0, st()>>>(s, d_sa.get(), sa32.get());
                SUFR_KERNEL_CHECK();
                launched();
            }
            d_sa.reset();
            d_sa_out = sa32.get();
            d_lcp_out = d_lcp.get();
        }
    } else if (index_bits_ == 64 && !compact && wide_ok_) {
        if (wide_m_) {
            wide_patch_kernel<<<grid_for(wide_m_, 2), kBlock, 0, st()>>>(wide_m_, wide_slots_.get(), d_sa.get(), d_lcp.get(),
                                                                       wide_sa_.get(), wide_lcp_.get());
            SUFR_KERNEL_CHECK();
            launched();
        }
        sa64 = std::move(wide_sa_);
        lcp64 = std::move(wide_lcp_);
        wide_slots_.reset();
        d_sa.reset();
        d_lcp.reset();
        d_sa_out = sa64.get();
        d_lcp_out = lcp64.get();
    } else if (index_bits_ == 64 && !compact) {
        drop_wide();
        sa64 = dalloc<unsigned long long>(s);
        lcp64 = dalloc<unsigned long long>(s);
        if (s) {
            widen_both(s, d_sa.get(), d_lcp.get(), sa64.get(), lcp64.get(), st());
            launched();
        }
        d_sa.reset();
        d_lcp.reset();
        d_sa_out = sa64.get();
        d_lcp_out = lcp64.get();
    }
    int t6 = timer.mark();
    SUFR_CUDA_CHECK(cudaStreamSynchronize(st()));
    tm.encode_ms = timer.ms(t0, t1);
    tm.keys_ms = timer.ms(t1, t_keys_mark);
    tm.sort_ms = timer.ms(t_keys_mark, t2);
    tm.refine_ms = timer.ms(t2, t3);
    tm.finish_ms = timer.ms(t4, t6);
    tm.lcp_ms = timer.ms(t3, t4);
    tm.total_ms = timer.ms(t0, t6);
    for (auto& ev : downsweep_events) {
        float t = 0;
        SUFR_CUDA_CHECK(cudaEventElapsedTime(&t, ev.first, ev.second));
        tm.dominant_kernel_ms += t;
        cudaEventDestroy(ev.first);
        cudaEventDestroy(ev.second);
    }
    tm.dominant_kernel_launches = downsweep_events.size();
    tm.dominant_kernel_bytes = sorted_elements * 2 * (sizeof(uint64_t) + sizeof(uint32_t));
    downsweep_events.clear();

    if (result_memory_ == SUFR_B200_MEM_DEVICE) {
        owner->text = d_text.release();
        // whichever buffers hold the output arrays now belong to the result
        owner->sa = d_sa_out == (void*)sa64.get() ? (void*)sa64.release()
                    : d_sa_out == (void*)sa32.get() ? (void*)sa32.release() : (void*)d_sa.release();
        owner->lcp = d_lcp_out == (void*)lcp64.get() ? (void*)lcp64.release() : (void*)d_lcp.release();

    } else {
        int e0 = timer.mark();
        // sharded builds: only rank 0 returns the transformed text (it writes the text section of the file)
        const bool want_text = args.world_size <= 1 || args.rank == 0;
        // a CUDA error below must not strand the page-locked buffers in the cache's lent-out list
        struct PinnedGuard {
            PinnedCache& cache;
            std::vector<void*> held;
            std::vector<cudaEvent_t> events;
            bool keep = false;
            void* get(size_t bytes) { void* p = cache.get(bytes); held.push_back(p); return p; }
            ~PinnedGuard() {
                for (auto e : events) cudaEventDestroy(e);
                if (!keep) for (void* p : held) cache.put(p);
            }
        } pinned{ctx.pinned};
        owner->text = want_text ? pinned.get(n) : nullptr;
        owner->sa = pinned.get(s * w);
        owner->lcp = pinned.get(s * w);
        struct OwnerReset {  // the buffers go back through the guard, not through a half-built owner
            ResultOwner* o; bool keep = false;
            ~OwnerReset() { if (!keep) { o->text = o->sa = o->lcp = nullptr; } }
        } owner_reset{owner.get()};
        int e1 = timer.mark();
        const bool log_e2e = getenv("SUFR_B200_LOG_E2E") != nullptr;
        const auto wall0 = std::chrono::steady_clock::now();
        auto since = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - wall0).count(); };
        if (compact) {
            // the widening is store-bandwidth bound on the host: all cores of this rank's share (one is left for the
            // thread that feeds the copy engine)
            int hw = (int)std::thread::hardware_concurrency();
            int threads = std::max(2, std::min(48, hw / std::max(1, (int)args.world_size) - 1));
            if (const char* dbg = getenv("SUFR_B200_WIDEN_THREADS")) threads = std::max(1, atoi(dbg));
            // 1. LCP bytes + exceptions, widened on the host while the suffix array is in flight
            d2h_bytes_ += s + exc_count * 8 + s * 4 + (want_text ? n : 0);
            uint8_t* h8 = (uint8_t*)ctx.pinned.get(s);
            struct Scratch8 { PinnedCache& c; void* p; ~Scratch8() { if (p) c.put(p); } } h8_guard{ctx.pinned, h8};
            std::vector<uint32_t> eidx(exc_count), eval(exc_count);
            SUFR_CUDA_CHECK(cudaMemcpyAsync(h8, d_lcp8.get(), s, cudaMemcpyDeviceToHost, st()));
            if (exc_count) {
                SUFR_CUDA_CHECK(cudaMemcpyAsync(eidx.data(), d_exc_idx.get(), exc_count * 4, cudaMemcpyDeviceToHost, st()));
                SUFR_CUDA_CHECK(cudaMemcpyAsync(eval.data(), d_exc_val.get(), exc_count * 4, cudaMemcpyDeviceToHost, st()));
            }
            SUFR_CUDA_CHECK(cudaStreamSynchronize(st()));
            if (log_e2e) fprintf(stderr, "[sufr_b200] e2e: LCP bytes on the host after %.1f ms (%d widening threads)\n", since(), threads);
            void* lcp_out = owner->lcp;
            const uint32_t bits = index_bits_;
            std::thread lcp_worker([=, &eidx, &eval]() {
                if (bits == 64) host_widen(h8, (uint64_t*)lcp_out, s, threads);
                else host_widen(h8, (uint32_t*)lcp_out, s, threads);
                // the exceptions arrive unordered (random writes into the result): spread them over the threads
                std::vector<std::thread> pool;
                for (int t = 0; t < threads; t++) {
                    pool.emplace_back([=, &eidx, &eval]() {
                        const uint64_t lo = exc_count * (uint64_t)t / threads, hi = exc_count * (uint64_t)(t + 1) / threads;
                        for (uint64_t e = lo; e < hi; e++) {
                            if (bits == 64) ((uint64_t*)lcp_out)[eidx[e]] = eval[e];
                            else ((uint32_t*)lcp_out)[eidx[e]] = eval[e];
                        }
                    });
                }
                for (auto& th : pool) th.join();
                if (log_e2e) fprintf(stderr, "[sufr_b200] e2e: LCP widened after %.1f ms\n", since());
            });
            struct Joiner {  // a CUDA error below must not leave the worker running on freed buffers
                std::thread& t;
                ~Joiner() { if (t.joinable()) t.join(); }
            } joiner{lcp_worker};
            // 2. text and suffix array (u32 on the wire, in chunks that are widened while the next ones arrive)
            if (n && want_text) SUFR_CUDA_CHECK(cudaMemcpyAsync(owner->text, d_text.get(), n, cudaMemcpyDeviceToHost, st()));
            if (index_bits_ == 64) {
                uint32_t* h32 = (uint32_t*)ctx.pinned.get(s * 4);
                Scratch8 h32_guard{ctx.pinned, h32};
                constexpr int kChunks = 8;
                for (int c = 0; c < kChunks; c++) {
                    uint64_t lo = s * (uint64_t)c / kChunks, hi = s * (uint64_t)(c + 1) / kChunks;
                    if (hi > lo)
                        SUFR_CUDA_CHECK(cudaMemcpyAsync(h32 + lo, d_sa.get() + lo, (hi - lo) * 4, cudaMemcpyDeviceToHost, st()));
                    cudaEvent_t ev = nullptr;
                    SUFR_CUDA_CHECK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
                    pinned.events.push_back(ev);
                    SUFR_CUDA_CHECK(cudaEventRecord(ev, st()));
                }
                for (int c = 0; c < kChunks; c++) {
                    uint64_t lo = s * (uint64_t)c / kChunks, hi = s * (uint64_t)(c + 1) / kChunks;
                    SUFR_CUDA_CHECK(cudaEventSynchronize(pinned.events[c]));
                    if (hi > lo) host_widen(h32 + lo, (uint64_t*)owner->sa + lo, hi - lo, threads);
                }
            } else {
                SUFR_CUDA_CHECK(cudaMemcpyAsync(owner->sa, d_sa.get(), s * 4, cudaMemcpyDeviceToHost, st()));
            }
            SUFR_CUDA_CHECK(cudaStreamSynchronize(st()));
            if (log_e2e) fprintf(stderr, "[sufr_b200] e2e: text + SA transferred and widened after %.1f ms\n", since());
            lcp_worker.join();
            if (log_e2e) fprintf(stderr, "[sufr_b200] e2e: done after %.1f ms\n", since());
        } else {
            d2h_bytes_ += 2 * s * w + (want_text ? n : 0);
            if (n && want_text) SUFR_CUDA_CHECK(cudaMemcpyAsync(owner->text, d_text.get(), n, cudaMemcpyDeviceToHost, st()));
            if (s) {
                SUFR_CUDA_CHECK(cudaMemcpyAsync(owner->sa, d_sa_out, s * w, cudaMemcpyDeviceToHost, st()));
                SUFR_CUDA_CHECK(cudaMemcpyAsync(owner->lcp, d_lcp_out, s * w, cudaMemcpyDeviceToHost, st()));
            }
        }
        int e2 = timer.mark();
        SUFR_CUDA_CHECK(cudaStreamSynchronize(st()));
        (void)e0;
        tm.d2h_ms = timer.ms(e1, e2);
        pinned.keep = true;  // ownership moves to the returned result
        owner_reset.keep = true;
    }
    if (!n_ranges_host.empty()) {
        owner->n_ranges = (uint64_t*)malloc(n_ranges_host.size() * 8);
        memcpy(owner->n_ranges, n_ranges_host.data(), n_ranges_host.size() * 8);
    }

    memset(out, 0, sizeof(*out));
    out->index_bits = index_bits_;
    out->memory = (uint32_t)result_memory_;
    out->text_len = n;
    out->num_suffixes = s;
    out->total_suffixes = total_suffixes;
    out->shard_offset = shard_offset;
    out->first_suffix = first;
    out->last_suffix = last;
    out->text = (uint8_t*)owner->text;
    out->sa = owner->sa;
    out->lcp = owner->lcp;
    out->n_ranges = owner->n_ranges;
    out->num_n_ranges = n_ranges_host.size() / 2;
    out->timings = tm;
    out->kernel_launches = ctx.launches - launches0;
    out->peak_device_bytes = ctx.pool.peak();
    out->alphabet_size = alphabet;
    out->bits_per_symbol = ks.pt.bits;
    out->refine_rounds = refine_rounds;
    out->doubling_rounds = doubling_rounds;
    out->h2d_bytes = h2d_bytes_;
    out->d2h_bytes = d2h_bytes_;
    out->position_bits = 8 * sizeof(pos_t);
    out->owner = owner.release();
}

