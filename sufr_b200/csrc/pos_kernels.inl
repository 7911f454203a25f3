// Kernels and scan functors that carry TEXT POSITIONS (sort payloads, suffix array, unresolved lists).
// Included twice by builder.cu, inside namespace sufr::p32 (pos_t = uint32_t, texts below 2^32 - 1 bytes: what the
// reference's SufrBuilder::<u32> handles) and sufr::p64 (pos_t = uint64_t: SufrBuilder::<u64>, suffix_array.rs:460-470).
// Everything that counts ELEMENTS of one rank (SA slots, group numbers, list lengths) stays 32-bit: a rank sorts fewer
// than 2^32 suffixes even when the text is longer (key-range shards).  LCP working values stay 32-bit as well.
// No include guard on purpose.

// Element e of the sort input is suffix e (full sort) or n-1-e (mask / max-query-len: the stable sort
// then leaves equal keys in position-descending order, the reference's tie rule, sufr_builder.rs:701-703).
// With `filter`, suffixes the reference does not index (sufr_builder.rs:446-449) get the key ~0, which no
// real key equals when the packed word has unused low bits: they sort behind everything and are dropped.
__global__ void __launch_bounds__(kBlock) keygen_kernel(KeySpec ks, uint64_t n, int descending,
                                                        const uint8_t* __restrict__ text, int filter,
                                                        uint64_t* __restrict__ keys, pos_t* __restrict__ pos) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride) {
        uint64_t p = descending ? n - 1 - e : e;
        keys[e] = (filter && !indexed_byte(text[p])) ? ~0ull : first_key(ks, p);
        pos[e] = (pos_t)p;
    }
}

// Fast path, unsharded: key generation FUSED with the first radix pass.  The first pass of an LSD sort need not be
// stable (there is no earlier order to preserve; the full sort has no true ties, so the order among records that
// agree on the sorted bits never reaches the result), so its ranking is one shared-memory atomic per record
// instead of eight ballots, and the 37 GB of records are written once, already in first-digit order, instead of
// written in text order and read back.  A warp owns 1024 consecutive positions and lane l the 32 positions of
// packed2 word l, so a key is two funnel shifts of registers the lane already holds.
//   fast2_first_digit_hist_kernel: [digit][block] histogram of the first digit (block b = the positions of block b
//                                  of the scatter kernel), scanned by rsort::scan_counts_kernel;
//   fast2_keygen_scatter_kernel:   per tile of 8192 positions: rank of every record inside its (tile, digit) by
//                                  atomicAdd, digit starts by a 256-wide scan, records placed in shared memory at
//                                  start + rank, coalesced copy to the digit's cursor in the output.
struct Fast2Lane {  // what a lane needs to form the 32 keys of its packed2 word
    uint64_t w, wn, M, P0;
};
struct Fast2LaneRaw {  // the global loads of fast2_lane_load, issued one tile ahead
    uint64_t w, wext, ir;
};
__device__ __forceinline__ Fast2LaneRaw fast2_lane_issue(const KeySpec& ks, uint64_t W0, int lane) {
    Fast2LaneRaw r;
    const uint64_t q = (W0 >> 5) + lane;
    r.w = q < ks.packed2_words ? __ldg(ks.packed2 + q) : 0ull;
    r.wext = (lane == 31 && q + 1 < ks.packed2_words) ? __ldg(ks.packed2 + q + 1) : 0ull;
    const uint64_t qi = (W0 >> 6) + lane;
    r.ir = (lane <= 16 && qi < ks.irr_words) ? __ldg(ks.irr + qi) : ~0ull;
    return r;
}
__device__ __forceinline__ Fast2Lane fast2_lane_finish(const KeySpec& ks, const Fast2LaneRaw& r, uint64_t W0, int lane,
                                                       int filter) {
    Fast2Lane L;
    L.w = r.w;
    L.wn = __shfl_down_sync(0xffffffffu, r.w, 1);
    if (lane == 31) L.wn = r.wext;
    const uint64_t i0 = __shfl_sync(0xffffffffu, r.ir, lane >> 1);
    const uint64_t i1 = __shfl_sync(0xffffffffu, r.ir, (lane >> 1) + 1);
    L.M = (lane & 1) ? ((i0 << 32) | (i1 >> 32)) : i0;  // irregular bits of positions P0 .. P0+63
    if (filter && !ks.reg_indexed) L.M = ~0ull;         // a regular byte may be filtered: exact path everywhere
    L.P0 = W0 + 32u * lane;
    return L;
}
__device__ __forceinline__ Fast2Lane fast2_lane_load(const KeySpec& ks, uint64_t W0, int lane, int filter) {
    return fast2_lane_finish(ks, fast2_lane_issue(ks, W0, lane), W0, lane, filter);
}
// key of position P0 + j (j is a compile-time constant in the unrolled callers); filtered suffixes get ~0
__device__ __forceinline__ uint64_t fast2_lane_key(const KeySpec& ks, const Fast2Lane& L, int j, uint64_t n, int filter) {
    constexpr uint64_t kWin = ~0ull << (64 - kFast2Symbols);
    uint64_t key = (j ? ((L.w << (2 * j)) | (L.wn >> ((64 - 2 * j) & 63))) : L.w) & ~3ull;
    if ((L.M << j) & kWin) {  // irregular symbol in the window (rare): exact key, and the suffix filter
        const uint64_t p = L.P0 + j;
        key = 0;
        if (p < n) key = (filter && !indexed_byte(ks.text[p])) ? ~0ull : first_key_fast2_slow(ks, p);
    }
    return key;
}
constexpr int kKsTile = (kBlock / 32) * 1024;           // positions per block iteration
constexpr int kKsSmem = kKsTile * (8 + 2 + 2);          // keys, local positions, ranks
// Key-range shard (multi-GPU): only the suffixes whose key falls into histogram bins [bin0, bin0 + span) are counted /
// generated; span == 0 means "everything".  A filtered suffix (key ~0) belongs to no shard.
struct KeyRange {
    uint32_t bin0, span;
    __device__ __forceinline__ bool take(uint64_t key) const {
        return span == 0 || (key != ~0ull && ((uint32_t)(key >> (64 - kShardHistBits)) - bin0) < span);
    }
};

__global__ void __launch_bounds__(kBlock) fast2_first_digit_hist_kernel(KeySpec ks, uint64_t n, int filter,
                                                                        uint64_t chunk_elems, int shift,
                                                                        uint32_t* __restrict__ counts, KeyRange range,
                                                                        unsigned long long* __restrict__ total) {
    constexpr int WARPS = kBlock / 32;
    __shared__ uint32_t hist[WARPS][256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < WARPS * 256; i += kBlock) (&hist[0][0])[i] = 0;
    __syncthreads();
    const uint64_t begin = (uint64_t)blockIdx.x * chunk_elems;
    const uint64_t end = begin + chunk_elems < n ? begin + chunk_elems : n;
    for (uint64_t W0 = begin + (uint64_t)warp * 1024; W0 < end; W0 += (uint64_t)WARPS * 1024) {
        const Fast2Lane L = fast2_lane_load(ks, W0, lane, filter);
#pragma unroll
        for (int j = 0; j < 32; j++) {
            const uint64_t key = fast2_lane_key(ks, L, j, n, filter);
            if (L.P0 + j < end && range.take(key)) atomicAdd(&hist[warp][(uint32_t)(key >> shift) & 255u], 1u);
        }
    }
    __syncthreads();
    uint32_t acc = 0;
#pragma unroll
    for (int w2 = 0; w2 < WARPS; w2++) acc += hist[w2][threadIdx.x];
    counts[(uint64_t)threadIdx.x * gridDim.x + blockIdx.x] = acc;
    if (total) {  // number of records of this shard (the caller sizes the record arrays with it)
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, off);
        if (lane == 0 && acc) atomicAdd(total, (unsigned long long)acc);
    }
}

__global__ void __launch_bounds__(kBlock, 2) fast2_keygen_scatter_kernel(KeySpec ks, uint64_t n, int filter,
                                                                         uint64_t* __restrict__ keys_out,
                                                                         pos_t* __restrict__ pos_out,
                                                                         uint64_t chunk_elems, int shift,
                                                                         const uint32_t* __restrict__ bases, KeyRange range) {
    constexpr int WARPS = kBlock / 32;
    extern __shared__ __align__(16) unsigned char ks_smem[];
    uint64_t* exk = reinterpret_cast<uint64_t*>(ks_smem);                  // records in digit order
    uint16_t* exl = reinterpret_cast<uint16_t*>(ks_smem + kKsTile * 8);    // their positions, relative to the tile
    uint16_t* rk = reinterpret_cast<uint16_t*>(ks_smem + kKsTile * 10);    // rank inside (tile, digit), by (j, lane)
    __shared__ uint32_t cnt[256];      // records per digit in this tile, then the digit's start in the tile
    __shared__ uint32_t running[256];  // global write cursor of each digit for this block
    __shared__ uint32_t goff[256];     // global index = goff[d] + tile-local slot (mod 2^32)
    __shared__ uint32_t warp_tot[WARPS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    running[tid] = bases[(uint64_t)tid * gridDim.x + blockIdx.x];
    const uint64_t begin = (uint64_t)blockIdx.x * chunk_elems;
    const uint64_t end = begin + chunk_elems < n ? begin + chunk_elems : n;
    Fast2LaneRaw raw = fast2_lane_issue(ks, begin + (uint64_t)warp * 1024, lane);
    for (uint64_t tile0 = begin; tile0 < end; tile0 += kKsTile) {
        cnt[tid] = 0;
        __syncthreads();
        const Fast2Lane L = fast2_lane_finish(ks, raw, tile0 + (uint64_t)warp * 1024, lane, filter);
        // the next tile's packed words are in flight while this one is ranked, placed and written
        if (tile0 + kKsTile < end) raw = fast2_lane_issue(ks, tile0 + kKsTile + (uint64_t)warp * 1024, lane);
        uint16_t* rkw = rk + warp * 1024;
        // phase 1: rank inside (tile, digit); any order will do
#pragma unroll
        for (int j = 0; j < 32; j++) {
            const uint64_t key = fast2_lane_key(ks, L, j, n, filter);
            if (L.P0 + j < end && range.take(key))
                rkw[j * 32 + lane] = (uint16_t)atomicAdd(&cnt[(uint32_t)(key >> shift) & 255u], 1u);
        }
        __syncthreads();
        // thread tid owns digit tid: exclusive scan over the digits
        const uint32_t c = cnt[tid];
        uint32_t incl = c;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            uint32_t o = __shfl_up_sync(0xffffffffu, incl, off);
            if (lane >= off) incl += o;
        }
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        uint32_t wprefix = 0;
#pragma unroll
        for (int w2 = 0; w2 < WARPS; w2++)
            if (w2 < warp) wprefix += warp_tot[w2];
        uint32_t count = 0;  // records of this tile (all its positions, or those inside the shard's key range)
#pragma unroll
        for (int w2 = 0; w2 < WARPS; w2++) count += warp_tot[w2];
        const uint32_t start = wprefix + incl - c;
        cnt[tid] = start;
        goff[tid] = running[tid] - start;
        running[tid] += c;
        __syncthreads();
        // phase 2: the keys again (two shifts), placed at start + rank
#pragma unroll
        for (int j = 0; j < 32; j++) {
            if (L.P0 + j < end) {
                const uint64_t key = fast2_lane_key(ks, L, j, n, filter);
                if (!range.take(key)) continue;
                const uint32_t slot = cnt[(uint32_t)(key >> shift) & 255u] + rkw[j * 32 + lane];
                exk[slot] = key;
                exl[slot] = (uint16_t)(warp * 1024 + lane * 32 + j);
            }
        }
        __syncthreads();
#pragma unroll 4
        for (int k = 0; k < kKsTile / kBlock; k++) {
            const uint32_t sl = k * kBlock + tid;
            if (sl < count) {
                const uint64_t key = exk[sl];
                const uint32_t dst = goff[(uint32_t)(key >> shift) & 255u] + sl;
                keys_out[dst] = key;
                pos_out[dst] = (pos_t)(tile0 + exl[sl]);
            }
        }
        __syncthreads();
    }
}

// Repetitiveness probe: first keys of every `stride`-th position; after sorting them, the number of adjacent
// equal keys tells a random-like text (a handful) from one with long repeats (thousands).
__global__ void __launch_bounds__(kBlock) sample_keys_kernel(KeySpec ks, uint64_t stride_pos, uint64_t count,
                                                             uint64_t* __restrict__ keys, pos_t* __restrict__ pos) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
        uint64_t p = i * stride_pos;
        keys[i] = first_key(ks, p);
        pos[i] = (pos_t)p;
    }
}
__global__ void __launch_bounds__(kBlock) count_equal_neighbours_kernel(const uint64_t* __restrict__ keys, uint64_t count,
                                                                        uint64_t mask, unsigned long long* __restrict__ out) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    unsigned long long c = 0;
    for (uint64_t i = 1 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride)
        c += (((keys[i] ^ keys[i - 1]) & mask) == 0) ? 1 : 0;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) c += __shfl_down_sync(0xffffffffu, c, off);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}

// Number of indexed suffixes (16 bytes per load).
__global__ void __launch_bounds__(kBlock) count_indexed_kernel(const uint8_t* __restrict__ text, uint64_t n,
                                                               unsigned long long* __restrict__ out) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t nvec = (((uintptr_t)text) & 15) == 0 ? n / 16 : 0;
    unsigned long long c = 0;
    for (uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += stride) {
        uint4 x = reinterpret_cast<const uint4*>(text)[v];
        uint32_t w[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int k = 0; k < 4; k++)
#pragma unroll
            for (int b = 0; b < 4; b++) c += indexed_byte((uint8_t)(w[k] >> (8 * b))) ? 1 : 0;
    }
    for (uint64_t i = nvec * 16 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        c += indexed_byte(text[i]) ? 1 : 0;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) c += __shfl_down_sync(0xffffffffu, c, off);
    __shared__ unsigned long long part[kBlock / 32];
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int w = 0; w < kBlock / 32; w++) t += part[w];
        if (t) atomicAdd(out, t);
    }
}


// Selection of the suffixes this build sorts: the suffix filter (sufr_builder.rs:446-449) applied up front,
// and / or the key range [lo, hi) of this rank's shard (multi-GPU).  Order-preserving compaction.
struct SelectIn {
    static constexpr bool kFlags = true;  // scan.cuh: scanned with warp votes
    KeySpec ks;
    uint64_t n;
    int descending;
    int use_range;
    uint64_t lo, hi;  // hi == 0 means "no upper bound"
    const uint8_t* text;
    int filter;
    __device__ uint32_t operator()(uint64_t e) const {
        uint64_t p = descending ? n - 1 - e : e;
        if (filter && !indexed_byte(text[p])) return 0u;
        if (!use_range) return 1u;
        uint64_t k = first_key(ks, p);
        return (k >= lo && (hi == 0 || k < hi)) ? 1u : 0u;
    }
};
struct SelectOut {
    KeySpec ks;
    uint64_t n;
    int descending;
    uint64_t* keys;
    pos_t* pos;
    __device__ void operator()(uint64_t e, uint32_t v, uint32_t incl) const {
        if (v) {
            uint64_t p = descending ? n - 1 - e : e;
            keys[incl - 1] = first_key(ks, p);
            pos[incl - 1] = (pos_t)p;
        }
    }
};


// Full sort only (the order of equal keys is irrelevant there): UNORDERED selection of this rank's suffixes,
// one key computation per position and no device-wide scan.  A block compacts a chunk of 2048 positions with
// warp ballots and reserves its output range with ONE global atomic.  `count` keeps counting past `capacity`.
constexpr int kSelectRows = 8;
__global__ void __launch_bounds__(kBlock) select_append_kernel(KeySpec ks, uint64_t n, uint64_t lo, uint64_t hi,
                                                               int filter, uint64_t* __restrict__ keys,
                                                               pos_t* __restrict__ pos,
                                                               unsigned long long* __restrict__ count,
                                                               uint64_t capacity) {
    constexpr int WARPS = kBlock / 32;
    __shared__ uint32_t wcount[kSelectRows * WARPS];
    __shared__ unsigned long long gbase;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    // a block iteration covers WARPS x 1024 positions: warp w owns [c*8192 + w*1024, +1024), 32 rows of 32
    const uint64_t chunk = (uint64_t)WARPS * 1024;
    const uint64_t chunks = (n + chunk - 1) / chunk;
    for (uint64_t c = blockIdx.x; c < chunks; c += gridDim.x) {
        const uint64_t W0 = c * chunk + (uint64_t)warp * 1024;
        for (int g = 0; g < 32 / kSelectRows; g++) {
            uint64_t k[kSelectRows];
            uint32_t lidx[kSelectRows];
            uint32_t takes = 0;
#pragma unroll
            for (int r = 0; r < kSelectRows; r++) {
                const int row = g * kSelectRows + r;
                const uint64_t p = W0 + (uint64_t)row * 32 + lane;
                uint64_t key = p < n ? key_word(ks, p, 0) : 0ull;
                bool take = p < n && (!filter || indexed_byte(ks.text[p])) && key >= lo && (hi == 0 || key < hi);
                k[r] = key;
                unsigned m = __ballot_sync(0xffffffffu, take);
                lidx[r] = __popc(m & lt_mask);
                if (lane == 0) wcount[r * WARPS + warp] = __popc(m);
                takes |= (take ? 1u : 0u) << r;
            }
            __syncthreads();
            if (warp == 0) {
                static_assert(kSelectRows * WARPS == 64, "warp_scan64");
                uint32_t acc = warp_scan64(wcount);
                if (lane == 0) gbase = acc ? atomicAdd(count, (unsigned long long)acc) : 0ull;
            }
            __syncthreads();
#pragma unroll
            for (int r = 0; r < kSelectRows; r++) {
                if (takes & (1u << r)) {
                    unsigned long long idx = gbase + wcount[r * WARPS + warp] + lidx[r];
                    if (idx < capacity) {
                        keys[idx] = k[r];
                        pos[idx] = (pos_t)(W0 + (uint64_t)(g * kSelectRows + r) * 32 + lane);
                    }
                }
            }
            __syncthreads();
        }
    }
}

// Fast-path variant of select_append_kernel (which serves the other alphabets).  A warp owns 1024 consecutive
// positions and lane l the 32 positions of packed2 word l, so a key is two funnel shifts of registers the lane
// already holds (no per-row shuffles or ballots).  Phase 1 leaves a 32-bit take mask per lane; phase 2 enumerates the taken positions densely (prefix
// sums over the lanes, k-th set bit of the owner's mask) so that the records leave the warp as coalesced stores.
__device__ __forceinline__ uint32_t select_bit(uint32_t m, uint32_t k) {  // position of the k-th (0-based) set bit
    uint32_t pos = 0, c;
    c = __popc(m & 0xFFFFu); if (k >= c) { k -= c; pos += 16; m >>= 16; }
    c = __popc(m & 0xFFu);   if (k >= c) { k -= c; pos += 8;  m >>= 8; }
    c = __popc(m & 0xFu);    if (k >= c) { k -= c; pos += 4;  m >>= 4; }
    c = __popc(m & 0x3u);    if (k >= c) { k -= c; pos += 2;  m >>= 2; }
    c = m & 1u;              if (k >= c) pos += 1;
    return pos;
}
__global__ void __launch_bounds__(kBlock) select_fast2_kernel(KeySpec ks, uint64_t n, uint64_t lo, uint64_t hi,
                                                              int filter, uint64_t* __restrict__ keys,
                                                              pos_t* __restrict__ pos,
                                                              unsigned long long* __restrict__ count,
                                                              uint64_t capacity) {
    constexpr int WARPS = kBlock / 32;
    constexpr uint64_t kWin = ~0ull << (64 - kFast2Symbols);
    __shared__ uint32_t incl_s[WARPS][32];
    __shared__ uint32_t wbase[WARPS];
    __shared__ unsigned long long gbase;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t chunk = (uint64_t)WARPS * 1024;
    const uint64_t chunks = (n + chunk - 1) / chunk;
    const uint32_t bin0 = (uint32_t)(lo >> (64 - kShardHistBits));
    const uint32_t span = (hi == 0 ? 1u << kShardHistBits : (uint32_t)(hi >> (64 - kShardHistBits))) - bin0;
    for (uint64_t c = blockIdx.x; c < chunks; c += gridDim.x) {
        const uint64_t W0 = c * chunk + (uint64_t)warp * 1024;
        const uint64_t q = (W0 >> 5) + lane;
        const uint64_t w = q < ks.packed2_words ? __ldg(ks.packed2 + q) : 0ull;
        uint64_t wn = __shfl_down_sync(0xffffffffu, w, 1);
        if (lane == 31) wn = q + 1 < ks.packed2_words ? __ldg(ks.packed2 + q + 1) : 0ull;
        const uint64_t qi = (W0 >> 6) + lane;
        const uint64_t ir = (lane <= 16 && qi < ks.irr_words) ? __ldg(ks.irr + qi) : ~0ull;
        const uint64_t i0 = __shfl_sync(0xffffffffu, ir, lane >> 1);
        const uint64_t i1 = __shfl_sync(0xffffffffu, ir, (lane >> 1) + 1);
        const uint64_t M = (lane & 1) ? ((i0 << 32) | (i1 >> 32)) : i0;  // irregular bits of positions P0 .. P0+63
        const uint64_t P0 = W0 + 32u * lane;
        // Fast test (no irregular symbol among the 32 that follow, which also means "inside the text"): the
        // shard is a range of 12-bit histogram bins, so the top 32 key bits decide.  Everything else -- and every
        // position when a regular byte can be filtered out -- takes the exact path.
        const uint32_t w_hi = (uint32_t)(w >> 32), w_lo = (uint32_t)w, wn_hi = (uint32_t)(wn >> 32);
        uint32_t m_hi = (uint32_t)(M >> 32), m_lo = (uint32_t)M;
        if (filter && !ks.reg_indexed) m_hi = m_lo = ~0u;
        uint32_t T = 0;
#pragma unroll
        for (int j = 0; j < 32; j++) {
            const uint32_t khi = j < 16 ? __funnelshift_l(w_lo, w_hi, 2 * j) : __funnelshift_l(wn_hi, w_lo, 2 * j - 32);
            const uint32_t mj = __funnelshift_l(m_lo, m_hi, j);
            bool take;
            if (mj == 0) {
                take = ((khi >> (32 - kShardHistBits)) - bin0) < span;
            } else {
                const uint64_t p = P0 + j;
                take = false;
                if (p < n && (!filter || indexed_byte(ks.text[p]))) {
                    const uint64_t key = first_key_fast2_slow(ks, p);
                    take = key >= lo && (hi == 0 || key < hi);
                }
            }
            T |= (take ? 1u : 0u) << j;
        }
        // dense enumeration of the taken positions
        uint32_t incl = __popc(T);
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            uint32_t o = __shfl_up_sync(0xffffffffu, incl, off);
            if (lane >= off) incl += o;
        }
        incl_s[warp][lane] = incl;
        const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
        if (lane == 31) wbase[warp] = total;
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t acc = 0;
            for (int k = 0; k < WARPS; k++) { uint32_t t = wbase[k]; wbase[k] = acc; acc += t; }
            gbase = acc ? atomicAdd(count, (unsigned long long)acc) : 0ull;
        }
        __syncthreads();
        const unsigned long long base = gbase + wbase[warp];
        for (uint32_t s0 = 0; s0 < total; s0 += 32) {
            const uint32_t sidx = s0 + lane;
            const bool act = sidx < total;
            uint32_t l = 0;
            if (act) {
#pragma unroll
                for (int step = 16; step; step >>= 1)
                    if (incl_s[warp][l + step - 1] <= sidx) l += step;
            }
            const uint32_t Tl = __shfl_sync(0xffffffffu, T, l);
            const uint64_t wl = __shfl_sync(0xffffffffu, w, l);
            const uint64_t wnl = __shfl_sync(0xffffffffu, wn, l);
            const uint64_t Ml = __shfl_sync(0xffffffffu, M, l);
            if (act) {
                const uint32_t j = select_bit(Tl, sidx - (incl_s[warp][l] - __popc(Tl)));
                const uint64_t p = W0 + 32u * l + j;
                uint64_t key = (j ? ((wl << (2 * j)) | (wnl >> (64 - 2 * j))) : wl) & ~3ull;
                if ((Ml << j) & kWin) key = first_key_fast2_slow(ks, p);
                const unsigned long long idx = base + sidx;
                if (idx < capacity) {
                    keys[idx] = key;
                    pos[idx] = (pos_t)p;
                }
            }
        }
        __syncthreads();
    }
}

// Histogram of the top `hbits` bits of the first key word over the indexed suffixes (splitter selection).
// With sample_shift > 0 only every 2^sample_shift-th position is counted (enough to balance the shards).
__global__ void __launch_bounds__(kBlock) key_hist_kernel(KeySpec ks, uint64_t n, uint32_t hbits,
                                                          const uint8_t* __restrict__ text, int filter,
                                                          uint32_t sample_shift,
                                                          unsigned long long* __restrict__ hist) {
    extern __shared__ uint32_t sh[];
    const uint32_t bins = 1u << hbits;
    for (uint32_t i = threadIdx.x; i < bins; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t ns = (n + ((1ull << sample_shift) - 1)) >> sample_shift;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < ns; i += stride) {
        uint64_t p = i << sample_shift;
        if (!filter || indexed_byte(text[p])) atomicAdd(&sh[(uint32_t)(first_key(ks, p) >> (64 - hbits))], 1u);
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < bins; i += blockDim.x)
        if (sh[i]) atomicAdd(&hist[i], (unsigned long long)sh[i]);
}

// ------------------------------------------------------------------ views of the group structure
// Round 0 works on the whole sorted array (one segment); later rounds on the compacted "active"
// elements (members of groups whose keys are still equal), with their segment ids and SA slots.
struct ViewAll {
    const uint64_t* key;
    const pos_t* pos;
    const uint32_t* large;  // fast path: bitmap of the members of large, only partially sorted groups (else NULL)
    uint64_t kmask;         // general path: the key bits the first sort covered (a group = equal on those bits)
    __device__ uint64_t k(uint64_t i) const {
        return large ? fast2_canon(key[i], (large[i >> 5] >> (i & 31)) & 1u) : (key[i] & kmask);
    }
    __device__ uint64_t raw(uint64_t i) const { return key[i]; }
    __device__ pos_t p(uint64_t i) const { return pos[i]; }
    __device__ bool same_seg(uint64_t i) const { return i > 0; }
    __device__ uint32_t slot(uint64_t i) const { return (uint32_t)i; }
    __device__ uint32_t deep(uint64_t) const { return 0u; }
};
struct ViewActive {
    const uint64_t* key;
    const pos_t* pos;
    const uint32_t* seg;
    const uint32_t* slot_;
    __device__ uint64_t k(uint64_t i) const { return key[i]; }
    __device__ uint64_t raw(uint64_t i) const { return key[i]; }
    __device__ pos_t p(uint64_t i) const { return pos[i]; }
    __device__ bool same_seg(uint64_t i) const { return i > 0 && seg[i] == seg[i - 1]; }
    __device__ uint32_t slot(uint64_t i) const { return slot_[i]; }
    __device__ uint32_t deep(uint64_t i) const { return seg[i] & kSegDeep; }  // see kSegDeep
};

// After sorting by key word `word`: write the LCP of every newly created group boundary, mark the
// still-unresolved elements, or (last word of a capped key) close ties.
template <typename View>
__global__ void __launch_bounds__(kBlock) resolve_kernel(View v, uint64_t m, KeySpec ks, uint32_t word,
                                                         int final_word, int is_round0,
                                                         uint32_t* __restrict__ lcp) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += stride) {
        if (!v.same_seg(i)) {
            if (is_round0) lcp[v.slot(i)] = 0;  // i == 0
            continue;                            // segment heads keep the LCP of an earlier round
        }
        uint64_t ki = v.k(i), kp = v.k(i - 1);
        if (ki != kp) {
            lcp[v.slot(i)] = lcp_from_words(ks, kp, ki, (uint64_t)(word + (v.deep(i) ? 1u : 0u)) * ks.pt.K, v.p(i - 1), v.p(i));
        } else if (final_word) {
            uint64_t la = key_len(ks, v.p(i - 1)), lb = key_len(ks, v.p(i));
            lcp[v.slot(i)] = (uint32_t)(la < lb ? la : lb);
        } else {
            lcp[v.slot(i)] = kLcpPending;
        }
    }
}

// Unordered append of the flagged elements of a 256-element block row: ballots per warp, one global atomic
// per row (a per-warp atomic on a single counter serialises when most warps have something to append).
__device__ __forceinline__ void block_append(bool active, uint32_t slot_val, pos_t pos_val,
                                             uint32_t* __restrict__ act_slot, pos_t* __restrict__ act_pos,
                                             unsigned long long* __restrict__ act_count, uint64_t capacity,
                                             uint32_t* wcount, unsigned long long* gbase) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned m = __ballot_sync(0xffffffffu, active);
    if (lane == 0) wcount[warp] = __popc(m);
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t acc = 0;
        for (int w = 0; w < kBlock / 32; w++) {
            uint32_t t = wcount[w];
            wcount[w] = acc;
            acc += t;
        }
        *gbase = acc ? atomicAdd(act_count, (unsigned long long)acc) : 0ull;
    }
    __syncthreads();
    if (active) {
        unsigned long long idx = *gbase + wcount[warp] + __popc(m & ((1u << lane) - 1u));
        if (idx < capacity) {
            act_slot[idx] = slot_val;
            act_pos[idx] = pos_val;
        }
    }
    __syncthreads();
}

// Round 0 in one pass over the sorted keys: LCP of every boundary (as resolve_kernel) and, because the
// unresolved elements are normally a tiny fraction, an UNORDERED warp-aggregated append of (slot, position)
// of every element that is still in a group of size > 1.  The short list is then sorted by slot.
// `kmask` = the key bits the first sort covered: when it stopped short of the whole word (Build::sort_bits),
// elements that agree on those bits form the groups, and the refinement starts over with key word 0.
__global__ void __launch_bounds__(kBlock) resolve0_append_kernel(const uint64_t* __restrict__ keys,
                                                                 const pos_t* __restrict__ pos, uint64_t s,
                                                                 KeySpec ks, int final_word, uint64_t kmask,
                                                                 uint32_t* __restrict__ lcp,
                                                                 uint32_t* __restrict__ act_slot,
                                                                 pos_t* __restrict__ act_pos,
                                                                 unsigned long long* __restrict__ act_count,
                                                                 uint64_t capacity) {
    __shared__ uint32_t wcount[kBlock / 32];
    __shared__ unsigned long long gbase;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t base = (uint64_t)blockIdx.x * blockDim.x; base < s; base += stride) {
        const uint64_t j = base + threadIdx.x;
        bool active = false;
        pos_t p = 0;
        if (j < s) {
            uint64_t kj = keys[j];
            p = pos[j];
            bool head = true;
            if (j == 0) {
                lcp[0] = 0;
            } else {
                uint64_t kp = keys[j - 1];
                head = ((kp ^ kj) & kmask) != 0;  // they differ inside the sorted bits: clz(kp ^ kj) is exact
                if (head) {
                    lcp[j] = lcp_from_words(ks, kp, kj, 0, pos[j - 1], p);
                } else if (final_word) {
                    uint64_t la = key_len(ks, pos[j - 1]), lb = key_len(ks, p);
                    lcp[j] = (uint32_t)(la < lb ? la : lb);
                } else {
                    lcp[j] = kLcpPending;
                }
            }
            if (!final_word) active = !head || (j + 1 < s && ((keys[j + 1] ^ kj) & kmask) == 0);
        }
        block_append(active, (uint32_t)j, p, act_slot, act_pos, act_count, capacity, wcount, &gbase);
    }
}

// Fast path, after the 4-pass radix sort on the top kFast2SortBits: group ordering and round 0 in ONE pass over the
// radix-sorted records.  A run is a maximal sequence of records that tie on the sorted bits.  Runs of 2..8 are
// ordered by their full 31-symbol keys; longer runs ("large") are left to the refinement as one group.  A group is then a run of equal canonical keys (fast2_canon): exact ties on all 31 symbols, or
// the members of a large run; everything in a group of size > 1 is collected for the exact refinement (which
// starts at key word 0).  Boundary LCP = clz(x ^ y) / 2 when neither key contains fill and both neighbours are
// final (singleton groups), else it is recomputed from the final order (kLcpFixup).  Filtered suffixes (key ~0)
// sort behind everything.
// Per tile of 1024 records (+ 9 before, + 8 behind) staged in shared memory:
//   phase 0: bitmap of run heads (record whose sorted bits differ from its predecessor's); from it every record
//            gets the extent of its run with two shifts and a clz / ffs.  A run of >= 9 records is "large";
//   phase 1: records of runs of 2..8 compute their rank inside the run by counting (every thread does the same
//            bounded work, no per-group serial sorting) -> the tile in final order in a second buffer;
//   phase 2: per record this tile owns (its run's head lies in the tile, or the run is large and the record itself
//            does): boundary LCP / pending / fix-up mark and the unresolved flag; unresolved records are appended
//            with one global atomic per tile.
// Runs that are cut off by the staging window are either long enough to be known large or are not adjacent to
// anything this tile emits.  OUT OF PLACE (the radix sort's ping-pong partner receives the ordered positions): a tile
// reads records that a neighbouring tile orders, so an in-place update would race.  The ordered KEYS are not written
// at all: the LCP marks carry the group structure the refinement needs (kLcpPending <=> the record continues the
// group of its predecessor), see LcpSegIn / LcpActiveIn.
constexpr int kR0Tile = 1024, kR0Back = kFast2SmallGroup + 1, kR0Fwd = kFast2SmallGroup;
constexpr int kR0N = kR0Tile + kR0Back + kR0Fwd;
constexpr int kR0Steps = (kR0N + 1 + kBlock - 1) / kBlock;
struct RunExtent {
    uint32_t back, fwd;  // records of the same run before / behind (32 = "32 or more")
    uint32_t pback;      // the same `back` for record a-1 (31 = "31 or more")
};
__device__ __forceinline__ RunExtent run_extent(const uint32_t* hbm, uint32_t a) {  // hbm[-1] and hbm[+1] exist
    const uint32_t wi = a >> 5, pos = a & 31;
    const uint32_t w0 = hbm[(int)wi - 1], w1 = hbm[wi], w2 = hbm[wi + 1];
    const uint32_t L = __funnelshift_l(w0, w1, 31 - pos);                           // bit 31 = head(a), 30 = head(a-1) ..
    const uint32_t R = (uint32_t)((((uint64_t)w2 << 32) | w1) >> (pos + 1));        // bit 0 = head(a+1) ..
    RunExtent e;
    e.back = L ? (uint32_t)__clz((int)L) : 32u;
    e.fwd = R ? (uint32_t)__ffs((int)R) - 1u : 32u;
    e.pback = (L << 1) ? (uint32_t)__clz((int)(L << 1)) : 31u;
    return e;
}
template <bool kDeepMarks>  // a compile-time switch: the marking code costs the hot path 2 % even when it never runs
__global__ void __launch_bounds__(kBlock) round0_fast2_kernel(const uint64_t* __restrict__ keys,
                                                              const pos_t* __restrict__ pos,
                                                              pos_t* __restrict__ pos_out,
                                                              uint64_t s, uint32_t* __restrict__ lcp,
                                                              uint32_t* __restrict__ act_slot,
                                                              pos_t* __restrict__ act_pos,
                                                              unsigned long long* __restrict__ act_count,
                                                              uint64_t capacity,
                                                              unsigned long long* __restrict__ sa64,
                                                              unsigned long long* __restrict__ lcp64) {
    constexpr uint32_t RL = kFast2SmallGroup + 1;  // a run of RL records or more is "large"
    __shared__ uint64_t ka[kR0N], kb[kR0N];
    __shared__ pos_t pa[kR0N], pb[kR0N];
    __shared__ uint32_t hbm_raw[kR0Steps * (kBlock / 32) + 2];
    __shared__ uint32_t wcount[kBlock / 32];
    __shared__ unsigned long long gbase;
    uint32_t* hbm = hbm_raw + 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        hbm_raw[0] = 0;
        hbm_raw[kR0Steps * (kBlock / 32) + 1] = 0;
    }
    const uint64_t tiles = (s + kR0Tile - 1) / kR0Tile;
    for (uint64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const uint64_t t0 = tile * kR0Tile;
        const uint32_t off = t0 ? (uint32_t)kR0Back : 0u;  // local index of record t0
        const uint64_t g0 = t0 - off;                       // global index of local record 0
        const uint32_t cnt = (uint32_t)((s - g0) < (uint64_t)(off + kR0Tile + kR0Fwd) ? (s - g0)
                                                                                       : (uint64_t)(off + kR0Tile + kR0Fwd));
        const uint32_t own_end = off + (uint32_t)((s - t0) < (uint64_t)kR0Tile ? (s - t0) : (uint64_t)kR0Tile);
        __syncthreads();
        for (uint32_t a = threadIdx.x; a < cnt; a += kBlock) {
            ka[a] = keys[g0 + a];
            pa[a] = pos[g0 + a];
        }
        __syncthreads();
        // phase 0: run heads; local record 0 and the end of the staged window count as heads
#pragma unroll
        for (int k = 0; k < kR0Steps; k++) {
            const uint32_t a = threadIdx.x + (uint32_t)k * kBlock;
            bool hd = a == cnt;
            if (a < cnt) hd = a == 0 || ((ka[a] ^ ka[a - 1]) & kFast2TopMask) != 0;
            const uint32_t w = __ballot_sync(0xffffffffu, hd);
            if (lane == 0) hbm[a >> 5] = w;
        }
        __syncthreads();
        // phase 1: final order of the tile.  A thread owns the same local indices in phases 1 and 2 and runs keep
        // their index ranges, so the extents are computed once.
        uint32_t ext[kR0Steps];
#pragma unroll
        for (int k = 0; k < kR0Steps; k++) {
            const uint32_t a = threadIdx.x + (uint32_t)k * kBlock;
            ext[k] = 0;
            if (a >= cnt) continue;
            const RunExtent e = run_extent(hbm, a);
            ext[k] = e.back | (e.fwd << 8) | (e.pback << 16);
            const uint64_t key = ka[a];
            uint32_t na = a;
            if (e.back + e.fwd + 1 < RL && e.back + e.fwd > 0) {
                const uint32_t h = a - e.back;
                uint32_t rank = 0;
                for (uint32_t b = h; b <= a + e.fwd; b++) {
                    const uint64_t o = ka[b];
                    rank += (o < key || (o == key && b < a)) ? 1u : 0u;
                }
                na = h + rank;
            }
            kb[na] = key;
            pb[na] = pa[a];
        }
        __syncthreads();
        // phase 2: LCP / flags of the records this tile owns
        uint32_t act = 0;
        pos_t apos[kR0Steps];
        pos_t* const pos_w = pos_out + g0;
        uint32_t* const lcp_t = lcp + g0;
#pragma unroll
        for (int k = 0; k < kR0Steps; k++) {
            const uint32_t d = threadIdx.x + (uint32_t)k * kBlock;
            apos[k] = 0;
            if (d >= cnt) continue;
            const uint32_t back = ext[k] & 0xFFu, fwd = (ext[k] >> 8) & 0xFFu, pback = ext[k] >> 16;
            const bool is_large = back + fwd + 1 >= RL;
            const uint32_t h = d - back;
            const bool emit = is_large ? (d >= off && d < own_end) : (h >= off && h < own_end);
            if (!emit) continue;
            const bool first = g0 == 0 && d == 0;  // global record 0
            const uint64_t k0 = kb[d];
            bool head, next_same;
            uint32_t out;
            if (is_large) {  // canonical key = the sorted bits: the whole run is one group
                head = back == 0;
                next_same = fwd > 0;
                out = first ? 0u : (head ? kLcpFixup : kLcpPending);
            } else {
                const uint64_t c0 = k0 & ~3ull;
                next_same = fwd > 0 && (kb[d + 1] & ~3ull) == c0;
                head = true;
                out = 0;
                if (!first) {
                    const uint64_t km1 = kb[d - 1];
                    bool prev_multi;
                    if (back > 0) {  // predecessor in the same (small) run
                        head = (km1 & ~3ull) != c0;
                        prev_multi = back >= 2 && ((kb[d - 2] ^ km1) & ~3ull) == 0;
                    } else {         // predecessor = last record of the previous run
                        prev_multi = pback > 0 && d >= 2 && (pback + 1 >= RL || ((kb[d - 2] ^ km1) & ~3ull) == 0);
                    }
                    if (!head) {
                        out = kLcpPending;
                        if (kDeepMarks) {  // no member of the group (equal canonical keys inside this small run) has fill
                            uint64_t any = 0;
                            for (uint32_t b = h; b <= d + fwd; b++) any |= ((kb[b] & ~3ull) == c0) ? kb[b] : 0ull;
                            if (!(any & 1ull)) out = kLcpPendingDeep;
                        }
                    } else if (((km1 | k0) & 1ull) == 0 && !next_same && !prev_multi)
                        out = (uint32_t)__clzll((long long)(km1 ^ k0)) >> 1;
                    else
                        out = kLcpFixup;
                }
            }
            lcp_t[d] = out;
            apos[k] = pb[d];
            pos_w[d] = apos[k];
            if (sa64) {  // 64-bit device results: written here, later changes are patched in (Build::refine)
                sa64[g0 + d] = apos[k];
                lcp64[g0 + d] = out;
            }
            if (!head || next_same) act |= 1u << k;
        }
        // append: per-thread count -> warp prefix -> block prefix -> one atomic
        const uint32_t mine = __popc(act);
        uint32_t incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        if (lane == 31) wcount[warp] = incl;
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t acc = 0;
            for (int w = 0; w < kBlock / 32; w++) {
                uint32_t tt = wcount[w];
                wcount[w] = acc;
                acc += tt;
            }
            gbase = acc ? atomicAdd(act_count, (unsigned long long)acc) : 0ull;
        }
        __syncthreads();
        unsigned long long idx = gbase + wcount[warp] + incl - mine;
#pragma unroll
        for (int k = 0; k < kR0Steps; k++) {
            if (act & (1u << k)) {
                if (idx < capacity) {
                    act_slot[idx] = (uint32_t)(g0 + threadIdx.x + (uint32_t)k * kBlock);
                    act_pos[idx] = apos[k];
                }
                idx++;
            }
        }
    }
}

// Exact LCP of the boundaries the fast path could not read off the keys, from the FINAL suffix order.
__global__ void __launch_bounds__(kBlock) lcp_fixup_kernel(KeySpec ks, uint64_t s, const pos_t* __restrict__ sa,
                                                           uint32_t* __restrict__ lcp,
                                                           unsigned long long* __restrict__ lcp64) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x * 4;
    for (uint64_t j0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; j0 < s; j0 += stride) {
        uint32_t v[4];
        if (j0 + 3 < s) {
            uint4 x = *reinterpret_cast<const uint4*>(lcp + j0);
            v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w;
        } else {
#pragma unroll
            for (int u = 0; u < 4; u++) v[u] = j0 + u < s ? lcp[j0 + u] : 0u;
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const uint64_t j = j0 + u;
            if (j < s && v[u] == kLcpFixup) {
                const uint32_t l = j ? (uint32_t)lcp_direct(ks, sa[j - 1], sa[j], 0) : 0u;
                lcp[j] = l;
                if (lcp64) lcp64[j] = l;  // 64-bit results written early by round 0 (see Build::refine)
            }
        }
    }
}

// Group structure of the fast path after round 0, read off the LCP marks: record j continues the group of record
// j-1 iff lcp[j] == kLcpPending; it is unresolved iff it continues a group or its successor does.
struct LcpSegIn {
    static constexpr bool kFlags = true;  // scan.cuh: scanned with warp votes  // segment ids of the slot-sorted unresolved list
    const uint32_t* lcp;
    const uint32_t* slot;
    __device__ uint32_t operator()(uint64_t a) const { return lcp_is_pending(lcp[slot[a]]) ? 0u : 1u; }
};
struct LcpActiveIn {
    static constexpr bool kFlags = true;  // scan.cuh: scanned with warp votes  // dense variant: order-preserving compaction of all unresolved records
    const uint32_t* lcp;
    uint64_t m;
    __device__ unsigned long long operator()(uint64_t i) const {
        const bool head = !lcp_is_pending(lcp[i]);
        const bool next_same = i + 1 < m && lcp_is_pending(lcp[i + 1]);
        return (!head || next_same) ? (1ull | ((unsigned long long)head << 32)) : 0ull;
    }
};
struct LcpActiveOut {
    const pos_t* pos;
    uint32_t* new_slot;
    pos_t* new_pos;
    uint32_t* new_seg;
    const uint32_t* lcp;  // the marks: a group is deep iff its continuing members carry kLcpPendingDeep
    uint64_t s;
    struct Staged { pos_t pos; uint32_t deep; };
    __device__ Staged load(uint64_t i, unsigned long long val, unsigned long long) const {
        Staged st{0, 0};
        if (val & 1ull) {
            st.pos = pos[i];
            const bool head = (val >> 32) & 1ull;  // a head's own mark is its boundary LCP: look at its successor
            st.deep = lcp[head && i + 1 < s ? i + 1 : i] == kLcpPendingDeep ? kSegDeep : 0u;
        }
        return st;
    }
    __device__ void store(uint64_t i, unsigned long long val, unsigned long long incl, const Staged& st) const {
        if (val & 1ull) {
            const uint32_t a = (uint32_t)incl - 1;
            new_slot[a] = (uint32_t)i;
            new_pos[a] = st.pos;
            new_seg[a] = ((uint32_t)(incl >> 32) - 1) | st.deep;
        }
    }
};

// Boundaries created by prefix doubling carry a lower bound (kLcpLowerBound | h, true LCP in [h, 2h)).  While the
// doubling stayed shallow the exact values are cheapest by extending each marked pair from its bound; deep repeats
// take the text-order walk of plcp_complete_kernel instead (O(n + chunks * LCP) rather than O(sum of LCP)).
__global__ void __launch_bounds__(kBlock) lcp_bounds_direct_kernel(KeySpec ks, uint64_t s, const pos_t* __restrict__ sa,
                                                                   uint32_t* __restrict__ lcp) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x * 4;
    for (uint64_t j0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; j0 < s; j0 += stride) {
        uint32_t v[4];
        if (j0 + 3 < s) {
            uint4 x = *reinterpret_cast<const uint4*>(lcp + j0);
            v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w;
        } else {
#pragma unroll
            for (int u = 0; u < 4; u++) v[u] = j0 + u < s ? lcp[j0 + u] : 0u;
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const uint64_t j = j0 + u;
            if (j == 0 || j >= s || v[u] >= kLcpFixup || !(v[u] & kLcpLowerBound)) continue;
            const uint64_t pa = sa[j - 1], pb = sa[j];
            uint64_t ea, eb;
            if (ks.num_n_ranges && n_run_end(ks, pa, ea) && n_run_end(ks, pb, eb)) {
                const uint64_t ra = ea - pa, rb = eb - pb;
                lcp[j] = (uint32_t)(ra < rb ? ra : rb);  // sufr_builder.rs:305-307
            } else {
                lcp[j] = (uint32_t)lcp_direct(ks, pa, pb, v[u] & ~kLcpLowerBound);
            }
        }
    }
}

// segment ids of the slot-sorted active list: a new segment starts where the key differs from the
// previous SA slot's key
struct SparseSegIn {
    static constexpr bool kFlags = true;  // scan.cuh: scanned with warp votes
    ViewAll v;  // group keys of the sorted array
    const uint32_t* slot;
    __device__ uint32_t operator()(uint64_t a) const {
        uint32_t j = slot[a];
        return (j == 0 || v.k(j) != v.k(j - 1)) ? 1u : 0u;
    }
};
struct SparseSegOut {
    uint32_t* seg;
    const uint32_t* lcp = nullptr;   // fast path: the marks (deep groups, see LcpActiveOut); else NULL
    const uint32_t* slot = nullptr;
    uint64_t s = 0;
    __device__ void operator()(uint64_t a, uint32_t v, uint32_t incl) const {
        uint32_t deep = 0;
        if (lcp) {
            const uint64_t j = slot[a];  // v = 1: head of its group
            deep = lcp[v && j + 1 < s ? j + 1 : j] == kLcpPendingDeep ? kSegDeep : 0u;
        }
        seg[a] = (incl - 1) | deep;
    }
};

// Compaction of the elements that are still in a group of size > 1.  Sum-scan input: bit 0 = active,
// bit 32 = active and first of its group.
template <typename View>
struct ActiveIn {
    static constexpr bool kFlags = true;  // scan.cuh: scanned with warp votes
    View v;
    uint64_t m;
    int final_word;
    int sentinel;  // filtered suffixes ride along with key ~0 (only with packings that leave a low bit unused)
    __device__ unsigned long long operator()(uint64_t i) const {
        if (final_word) return 0;
        // every load is issued unconditionally (clamped indices): a load behind a branch on another load's value
        // would serialise the memory round trips of a thread's rows
        const uint64_t ip = i > 0 ? i - 1 : 0, in = i + 1 < m ? i + 1 : i;
        const uint64_t ki = v.k(i), kp = v.k(ip), kn = v.k(in), raw = v.raw(i);
        const bool seg_i = v.same_seg(i), seg_n = v.same_seg(in);
        const bool head = !seg_i | (kp != ki);
        const bool next_same = (i + 1 < m) & seg_n & (kn == ki);
        const bool active = (!head | next_same) & !(sentinel && raw == ~0ull);  // sentinel keys are never refined
        return active ? (1ull | ((unsigned long long)head << 32)) : 0ull;
    }
};
template <typename View>
struct ActiveOut {
    View v;
    uint32_t* new_slot;
    pos_t* new_pos;
    uint32_t* new_seg;
    struct Staged { uint32_t slot, deep; pos_t pos; };
    __device__ Staged load(uint64_t i, unsigned long long val, unsigned long long) const {
        Staged st{0, 0, 0};
        if (val & 1ull) { st.slot = v.slot(i); st.pos = v.p(i); st.deep = v.deep(i); }
        return st;
    }
    __device__ void store(uint64_t, unsigned long long val, unsigned long long incl, const Staged& st) const {
        if (val & 1ull) {
            uint32_t a = (uint32_t)incl - 1;
            new_slot[a] = st.slot;
            new_pos[a] = st.pos;
            new_seg[a] = ((uint32_t)(incl >> 32) - 1) | st.deep;
        }
    }
};

// ---- one refinement round = sort every unresolved group by a 64-bit key of its members
// (word rounds: the next key word; prefix doubling: group << 32 | rank of the suffix h symbols further on).
//   round_keys_kernel        the keys of ALL unresolved elements, one thread per element (independent gathers)
//   small_groups_kernel      groups of at most kSmallSeg members: the thread at the group start rank-sorts them in
//                            registers (stable) and writes keys, positions and SA slots in place.  On repetitive
//                            texts most unresolved groups are pairs (a segment and its copy).
//   LargeIn / LargeOut*      the members of the larger groups, compacted for the radix sort;
//   scatter_large*_kernel    and put back.
// (Measured and dropped: skipping large groups whose members are already in order.  The groups of a tandem-repeat
// text -- ~10^3..10^4 suffixes each, unresolved for log2(LCP) doubling rounds -- mix the arrays that share a unit, and
// do get re-ordered in almost every round: profiles/r2_round_log_config5.txt.)
constexpr int kSmallSeg = 8;
__global__ void __launch_bounds__(kBlock) round_keys_kernel(KeySpec ks, uint64_t m, uint32_t word, int filter,
                                                            const pos_t* __restrict__ pos, const uint32_t* __restrict__ seg,
                                                            uint64_t* __restrict__ keys) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t a = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; a < m; a += stride) {
        const pos_t p = pos[a];
        const uint32_t w = word + ((seg[a] & kSegDeep) ? 1u : 0u);  // deep groups are one key word ahead
        keys[a] = (filter && !indexed_byte(ks.text[p])) ? ~0ull : key_word(ks, p, w);
    }
}
__global__ void __launch_bounds__(kBlock) small_groups_kernel(uint64_t m, const uint32_t* __restrict__ seg,
                                                              const uint32_t* __restrict__ slot, pos_t* __restrict__ pos,
                                                              uint64_t* __restrict__ keys, pos_t* __restrict__ sa,
                                                              uint8_t* __restrict__ is_large,
                                                              unsigned long long* __restrict__ any_large) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t a = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; a < m; a += stride) {
        const uint32_t g = seg[a];
        if (a > 0 && seg[a - 1] == g) continue;  // not the first element of its group
        int len = 1;
        while (len <= kSmallSeg && a + len < m && seg[a + len] == g) len++;
        if (len > kSmallSeg) {
            is_large[g & ~kSegDeep] = 1;
            *any_large = 1;  // "some group is large" (benign race: every writer stores 1)
            continue;
        }
        pos_t p[kSmallSeg];
        uint64_t k[kSmallSeg];
        bool sorted = true;
#pragma unroll
        for (int i = 0; i < kSmallSeg; i++) {
            if (i < len) {
                p[i] = pos[a + i];
                k[i] = keys[a + i];
                if (i > 0 && k[i] < k[i - 1]) sorted = false;
            } else {
                p[i] = 0;
                k[i] = ~0ull;
            }
        }
        if (sorted) continue;  // already in order: nothing to write
#pragma unroll
        for (int i = 0; i < kSmallSeg; i++) {
            if (i < len) {
                int r = 0;
#pragma unroll
                for (int j = 0; j < kSmallSeg; j++)
                    if (j < len && (k[j] < k[i] || (k[j] == k[i] && j < i))) r++;
                keys[a + r] = k[i];
                pos[a + r] = p[i];
                sa[slot[a + r]] = p[i];
            }
        }
    }
}
// Full sort, few unresolved elements left (the deep tail of a repetitive text): groups of at most kSmallSeg members
// are finished in ONE step by comparing their suffixes directly from the depth they are known to share -- instead of
// one launch-bound round per key word (a 4 kb repeat is 200 words deep).  The thread at the group start insertion-
// sorts the members, writes the final order into the suffix array and the exact LCP of every new boundary, and
// clears the members' "continues the group" marks; larger groups are left to further rounds (flag in is_large).
__device__ __forceinline__ bool suffix_less(const KeySpec& ks, uint64_t pa, uint64_t pb, uint64_t depth, uint64_t& lcp_out) {
    const uint64_t l = lcp_direct(ks, pa, pb, depth);
    lcp_out = l;
    const uint64_t n = ks.pt.n;
    if (pa + l >= n || pb + l >= n) return pa > pb;  // one suffix is a prefix of the other: the shorter one first
    return sym_at(ks.pt, pa + l) < sym_at(ks.pt, pb + l);
}
__global__ void __launch_bounds__(kBlock) finish_small_groups_kernel(KeySpec ks, uint64_t m, uint64_t depth,
                                                                     const uint32_t* __restrict__ seg,
                                                                     const uint32_t* __restrict__ slot,
                                                                     const pos_t* __restrict__ pos, pos_t* __restrict__ sa,
                                                                     uint32_t* __restrict__ lcp, uint8_t* __restrict__ is_large,
                                                                     unsigned long long* __restrict__ large_elems) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t a = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; a < m; a += stride) {
        const uint32_t g = seg[a];
        if (a > 0 && seg[a - 1] == g) continue;
        int len = 1;
        while (len <= kSmallSeg && a + len < m && seg[a + len] == g) len++;
        if (len > kSmallSeg) {
            is_large[g & ~kSegDeep] = 1;
            atomicAdd(large_elems, 1ull);
            continue;
        }
        pos_t p[kSmallSeg];
#pragma unroll
        for (int i = 0; i < kSmallSeg; i++) p[i] = i < len ? pos[a + i] : (pos_t)0;
        // insertion sort by direct comparison (len <= 8)
        for (int i = 1; i < len; i++) {
            const pos_t x = p[i];
            int j = i;
            uint64_t l;
            while (j > 0 && suffix_less(ks, x, p[j - 1], depth, l)) {
                p[j] = p[j - 1];
                j--;
            }
            p[j] = x;
        }
        sa[slot[a]] = p[0];
        for (int i = 1; i < len; i++) {
            uint64_t l;
            suffix_less(ks, p[i - 1], p[i], depth, l);
            sa[slot[a + i]] = p[i];
            lcp[slot[a + i]] = (uint32_t)l;
        }
    }
}
// The larger groups of the same deep tail: ONE BLOCK per group orders up to kFinishCap members in shared memory with a
// bitonic network whose comparator is the direct suffix comparison, then writes the final order and the exact LCP of
// every new boundary.  A shard cannot run prefix doubling, and the families of a repetitive genome (hundreds of copies of
// a segment) otherwise take one launch-bound word round per 21 symbols of their depth -- ~190 rounds on BASELINE's
// repetitive variant.  The comparisons of a group are budgeted (kFinishBudget key-word loads): a group that exceeds it
// (a tandem array: thousands of members that share 10^5 symbols) is left as it was, flagged for the word rounds / the
// full-sort fallback, and so is a group with more than kFinishCap members.
constexpr uint32_t kFinishCap = 32768 / sizeof(pos_t);
constexpr unsigned long long kFinishBudget = 1ull << 25;
struct SegHeadIn {
    static constexpr bool kFlags = true;  // scan.cuh: scanned with warp votes
    const uint32_t* seg;
    __device__ uint32_t operator()(uint64_t a) const { return (a == 0 || seg[a] != seg[a - 1]) ? 1u : 0u; }
};
__global__ void __launch_bounds__(kBlock) finish_large_groups_kernel(KeySpec ks, uint64_t m, uint64_t depth,
                                                                     const uint64_t* __restrict__ group_start,
                                                                     uint64_t num_groups, const uint32_t* __restrict__ seg,
                                                                     const uint32_t* __restrict__ slot,
                                                                     const pos_t* __restrict__ pos, pos_t* __restrict__ sa,
                                                                     uint32_t* __restrict__ lcp, uint8_t* __restrict__ is_large,
                                                                     unsigned long long* __restrict__ left_elems) {
    __shared__ pos_t sp[kFinishCap];
    __shared__ unsigned long long work;
    constexpr pos_t kInf = ~(pos_t)0;  // padding up to the power of two: sorts behind every suffix
    const uint64_t K = ks.pt.K;
    for (uint64_t g = blockIdx.x; g < num_groups; g += gridDim.x) {
        const uint64_t a0 = group_start[g];
        const uint64_t len = (g + 1 < num_groups ? group_start[g + 1] : m) - a0;
        __syncthreads();  // the previous group's shared memory is no longer read
        if (len > kFinishCap) {
            if (threadIdx.x == 0) {
                is_large[seg[a0] & ~kSegDeep] = 1;
                atomicAdd(left_elems, (unsigned long long)len);
            }
            continue;
        }
        uint32_t n2 = 2;
        while (n2 < len) n2 <<= 1;
        for (uint32_t i = threadIdx.x; i < n2; i += kBlock) sp[i] = i < len ? pos[a0 + i] : kInf;
        if (threadIdx.x == 0) work = 0;
        __syncthreads();
        bool over = false;
        for (uint32_t k = 2; k <= n2 && !over; k <<= 1) {
            for (uint32_t j = k >> 1; j > 0; j >>= 1) {
                unsigned long long mine = 0;
                for (uint32_t t = threadIdx.x; t < n2 / 2; t += kBlock) {
                    const uint32_t i = 2 * t - (t & (j - 1)), ixj = i + j;
                    const pos_t x = sp[i], y = sp[ixj];
                    bool y_less = false;  // y < x ?
                    if (y != kInf) {
                        if (x == kInf) y_less = true;
                        else {
                            uint64_t l;
                            y_less = suffix_less(ks, y, x, depth, l);
                            mine += (l - depth) / K + 1;
                        }
                    }
                    const bool ascending = (i & k) == 0;
                    if (y_less == ascending) { sp[i] = y; sp[ixj] = x; }
                }
                if (mine) atomicAdd(&work, mine);
                __syncthreads();
                over = work > kFinishBudget;
                __syncthreads();
                if (over) break;
            }
        }
        if (over) {  // nothing has been written: the group stays as it was
            if (threadIdx.x == 0) {
                is_large[seg[a0] & ~kSegDeep] = 1;
                atomicAdd(left_elems, (unsigned long long)len);
            }
            continue;
        }
        for (uint32_t i = threadIdx.x; i < len; i += kBlock) {
            const pos_t p = sp[i];
            sa[slot[a0 + i]] = p;
            if (i > 0) {  // the head keeps the LCP of an earlier round
                uint64_t l;
                suffix_less(ks, sp[i - 1], p, depth, l);
                lcp[slot[a0 + i]] = (uint32_t)l;
            }
        }
    }
}
// the members of the groups finish_small_groups_kernel left over, with new group numbers
struct LeftoverOut {
    const uint32_t* slot;
    const pos_t* pos;
    uint32_t* new_slot;
    pos_t* new_pos;
    uint32_t* new_seg;
    const uint32_t* seg;
    struct Staged { uint32_t slot, deep; pos_t pos; };
    __device__ Staged load(uint64_t a, unsigned long long val, unsigned long long) const {
        Staged st{0, 0, 0};
        if (val & 1ull) { st.slot = slot[a]; st.pos = pos[a]; st.deep = seg[a] & kSegDeep; }
        return st;
    }
    __device__ void store(uint64_t, unsigned long long val, unsigned long long incl, const Staged& st) const {
        if (val & 1ull) {
            const uint32_t b = (uint32_t)incl - 1;
            new_slot[b] = st.slot;
            new_pos[b] = st.pos;
            new_seg[b] = ((uint32_t)(incl >> 32) - 1) | st.deep;
        }
    }
};

// elements of the flagged groups, in order
struct LargeIn {
    static constexpr bool kFlags = true;  // scan.cuh: scanned with warp votes
    const uint32_t* seg;
    const uint8_t* flagged;
    __device__ unsigned long long operator()(uint64_t a) const {
        const uint32_t g = seg[a], gp = seg[a > 0 ? a - 1 : 0];
        const bool head = a == 0 || gp != g;
        return flagged[g & ~kSegDeep] ? (1ull | ((unsigned long long)head << 32)) : 0ull;
    }
};
struct LargeOut {
    const pos_t* pos;
    const uint64_t* keys;
    uint32_t* idx;      // index in the active arrays
    uint64_t* lkeys;
    uint64_t* segidx;   // sort payload: compact group number << 32 | compact index (positions may need 64 bits)
    pos_t* lpos;        // position of compact element b
    struct Staged { uint64_t key; pos_t pos; };
    __device__ Staged load(uint64_t a, unsigned long long val, unsigned long long) const {
        Staged st{0, 0};
        if (val & 1ull) { st.key = keys[a]; st.pos = pos[a]; }
        return st;
    }
    __device__ void store(uint64_t a, unsigned long long val, unsigned long long incl, const Staged& st) const {
        if (val & 1ull) {
            uint32_t b = (uint32_t)incl - 1;
            idx[b] = (uint32_t)a;
            lkeys[b] = st.key;
            segidx[b] = (((incl >> 32) - 1) << 32) | b;
            lpos[b] = st.pos;
        }
    }
};
// prefix doubling: composite key (compact group number << rank_bits | rank), the position as the payload
struct LargeOutRank {
    const pos_t* pos;
    const uint64_t* keys;
    uint32_t* idx;
    uint64_t* lck;
    pos_t* lpos;
    int rank_bits;
    struct Staged { uint64_t key; pos_t pos; };
    __device__ Staged load(uint64_t a, unsigned long long val, unsigned long long) const {
        Staged st{0, 0};
        if (val & 1ull) { st.key = keys[a]; st.pos = pos[a]; }
        return st;
    }
    __device__ void store(uint64_t a, unsigned long long val, unsigned long long incl, const Staged& st) const {
        if (val & 1ull) {
            uint32_t b = (uint32_t)incl - 1;
            idx[b] = (uint32_t)a;
            lck[b] = (((incl >> 32) - 1) << rank_bits) | (st.key & ((1ull << rank_bits) - 1ull));
            lpos[b] = st.pos;
        }
    }
};
__global__ void __launch_bounds__(kBlock) scatter_large_rank_kernel(uint64_t ml, const uint64_t* __restrict__ lck,
                                                                    const pos_t* __restrict__ lpos,
                                                                    const uint32_t* __restrict__ idx,
                                                                    const uint32_t* __restrict__ slot,
                                                                    uint64_t* __restrict__ keys, pos_t* __restrict__ pos,
                                                                    pos_t* __restrict__ sa, int rank_bits) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t rmask = (1ull << rank_bits) - 1ull;
    for (uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; b < ml; b += stride) {
        const uint32_t a = idx[b];
        const pos_t p = lpos[b];
        keys[a] = (keys[a] & ~rmask) | (lck[b] & rmask);  // the group number stays
        pos[a] = p;
        sa[slot[a]] = p;
    }
}
__global__ void __launch_bounds__(kBlock) scatter_large_kernel(uint64_t ml, const uint64_t* __restrict__ lkeys,
                                                               const uint64_t* __restrict__ lsegidx,
                                                               const pos_t* __restrict__ lpos,
                                                               const uint32_t* __restrict__ idx,
                                                               const uint32_t* __restrict__ slot,
                                                               uint64_t* __restrict__ keys, pos_t* __restrict__ pos,
                                                               pos_t* __restrict__ sa) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; b < ml; b += stride) {
        uint32_t a = idx[b];
        const pos_t p = lpos[(uint32_t)lsegidx[b]];  // the element that was sorted into compact slot b
        keys[a] = lkeys[b];
        pos[a] = p;
        sa[slot[a]] = p;
    }
}

// After a sort of the unresolved list: positions go back to their SA slots (the slots of a group are unchanged).
__global__ void __launch_bounds__(kBlock) writeback_pos_kernel(uint64_t m, const pos_t* __restrict__ pos,
                                                               const uint32_t* __restrict__ slot,
                                                               pos_t* __restrict__ sa) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t a = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; a < m; a += stride) sa[slot[a]] = pos[a];
}

// ------------------------------------------------------------------ prefix doubling (deep repeats)
__global__ void __launch_bounds__(kBlock) isa_init_kernel(uint64_t n, const pos_t* __restrict__ sa,
                                                          uint32_t* __restrict__ isa) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) isa[sa[j]] = (uint32_t)j;
}
// Records (position, rank) sorted on the position bits above `chunk_bits`: chunk c = records [c << chunk_bits, ...) holds
// exactly the positions [c << chunk_bits, (c + 1) << chunk_bits) in some order.  A block orders one chunk in shared
// memory and writes the ranks as whole lines (build_impl.inl, isa_init).
constexpr int kIsaChunkBitsMax = 14;
__global__ void __launch_bounds__(kBlock) iota_kernel(uint32_t* __restrict__ out, uint64_t count, uint32_t base) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < count; j += stride) out[j] = base + (uint32_t)j;
}
__global__ void __launch_bounds__(512) isa_chunk_kernel(uint64_t n, int chunk_bits, const uint32_t* __restrict__ p,
                                                        const uint32_t* __restrict__ rank, uint32_t* __restrict__ isa) {
    extern __shared__ uint32_t isa_local[];
    const uint32_t size = 1u << chunk_bits, mask = size - 1u;
    const uint64_t chunks = (n + size - 1) >> chunk_bits;
    for (uint64_t c = blockIdx.x; c < chunks; c += gridDim.x) {
        const uint64_t base = c << chunk_bits;
        const uint32_t count = (n - base) < (uint64_t)size ? (uint32_t)(n - base) : size;
        for (uint32_t t = threadIdx.x; t < count; t += blockDim.x) isa_local[p[base + t] & mask] = rank[base + t];
        __syncthreads();
        for (uint32_t t = threadIdx.x; t < count; t += blockDim.x) isa[base + t] = isa_local[t];
        __syncthreads();
    }
}

// rank of an active element = SA slot of the first element of its group
struct GroupStartIn {
    static constexpr bool kLastIndex = true;  // scan.cuh: scanned with warp votes
    const uint32_t* seg;
    __device__ uint32_t operator()(uint64_t a) const { return (a == 0 || seg[a] != seg[a - 1]) ? (uint32_t)a : 0u; }
};
struct GroupRankOut {
    const uint32_t* slot;
    const pos_t* pos;
    uint32_t* isa;
    struct Staged { uint32_t rank; pos_t pos; };
    __device__ Staged load(uint64_t a, uint32_t, uint32_t first) const { return Staged{slot[first], pos[a]}; }
    __device__ void store(uint64_t, uint32_t, uint32_t, const Staged& st) const { isa[st.pos] = st.rank; }
};

// composite key (segment << rank_bits | rank of suffix p+h, 0 = beyond the end); rank_bits = bits of n, so that
// the radix sort of a round covers as few digits as possible
__global__ void __launch_bounds__(kBlock) doubling_keys_kernel(uint64_t m, uint64_t n, uint64_t h,
                                                               const pos_t* __restrict__ pos,
                                                               const uint32_t* __restrict__ seg,
                                                               const uint32_t* __restrict__ isa,
                                                               uint64_t* __restrict__ ck, int rank_bits) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t a = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; a < m; a += stride) {
        uint64_t q = (uint64_t)pos[a] + h;
        uint32_t r = q < n ? isa[q] + 1u : 0u;
        ck[a] = ((uint64_t)(seg[a] & ~kSegDeep) << rank_bits) | r;
    }
}

struct DoublingStartIn {
    static constexpr bool kLastIndex = true;  // scan.cuh: scanned with warp votes
    const uint64_t* ck;
    __device__ uint32_t operator()(uint64_t a) const { return (a == 0 || ck[a] != ck[a - 1]) ? (uint32_t)a : 0u; }
};
// new ranks + lower-bound LCP marks for the boundaries created by this round
struct DoublingRankOut {
    const uint64_t* ck;
    const uint32_t* slot;
    const pos_t* pos;
    uint32_t* isa;
    uint32_t* lcp;
    uint32_t mark;  // kLcpLowerBound | min(h, 2^31 - 1)
    int rank_bits;
    struct Staged { uint32_t rank, mslot; pos_t pos; bool write, marked; };
    __device__ Staged load(uint64_t a, uint32_t, uint32_t first) const {
        const uint64_t cf = ck[first], cfp = ck[first > 0 ? first - 1 : 0], ca = ck[a], cap = ck[a > 0 ? a - 1 : 0];
        Staged st;
        // the first subgroup of a group keeps the group's rank (the slot of its first element): nothing to write
        st.write = first > 0 && (cf >> rank_bits) == (cfp >> rank_bits);
        st.marked = a > 0 && ca != cap && (ca >> rank_bits) == (cap >> rank_bits);
        st.rank = slot[first];
        st.mslot = slot[a];
        st.pos = pos[a];
        return st;
    }
    __device__ void store(uint64_t, uint32_t, uint32_t, const Staged& st) const {
        if (st.write) isa[st.pos] = st.rank;
        if (st.marked) lcp[st.mslot] = mark;
    }
};
struct DoublingActiveIn {
    static constexpr bool kFlags = true;  // scan.cuh: scanned with warp votes
    const uint64_t* ck;
    uint64_t m;
    __device__ unsigned long long operator()(uint64_t a) const {
        uint64_t c = ck[a];
        bool head = a == 0 || ck[a - 1] != c;
        bool next_same = a + 1 < m && ck[a + 1] == c;
        bool active = !head || next_same;
        return active ? (1ull | ((unsigned long long)head << 32)) : 0ull;
    }
};
struct DoublingActiveOut {
    const uint32_t* slot;
    const pos_t* pos;
    uint32_t* new_slot;
    pos_t* new_pos;
    uint32_t* new_seg;
    struct Staged { uint32_t slot; pos_t pos; };
    __device__ Staged load(uint64_t a, unsigned long long val, unsigned long long) const {
        Staged st{0, 0};
        if (val & 1ull) { st.slot = slot[a]; st.pos = pos[a]; }
        return st;
    }
    __device__ void store(uint64_t, unsigned long long val, unsigned long long incl, const Staged& st) const {
        if (val & 1ull) {
            uint32_t b = (uint32_t)incl - 1;
            new_slot[b] = st.slot;
            new_pos[b] = st.pos;
            new_seg[b] = (uint32_t)(incl >> 32) - 1;
        }
    }
};

// ------------------------------------------------------------------ LCP completion
// Boundaries created by prefix doubling only carry a lower bound (kLcpLowerBound | h).  They are
// completed in TEXT order, Kasai / PLCP style: with phi(i) = the suffix preceding suffix i in the
// suffix array, lcp(i, phi(i)) >= lcp(i-1, phi(i-1)) - 1, so a thread that walks a chunk of consecutive
// text positions extends each match from where the previous one ended instead of from the lower
// bound.  Total work is O(n + chunks * LCP) instead of O(sum of LCP^2) for tandem repeats.
// Needs the inverse suffix array of ALL positions (available whenever doubling ran).
// Pairs that both start inside recorded N runs use the reference's shortcut (sufr_builder.rs:305-307).
constexpr uint32_t kPlcpChunk = 512;

__global__ void __launch_bounds__(kBlock) plcp_complete_kernel(KeySpec ks, uint64_t n, const pos_t* __restrict__ sa,
                                                               const uint32_t* __restrict__ isa,
                                                               uint32_t* __restrict__ lcp) {
    const uint64_t chunks = (n + kPlcpChunk - 1) / kPlcpChunk;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; c < chunks; c += stride) {
        uint64_t i0 = c * kPlcpChunk;
        uint64_t i1 = i0 + kPlcpChunk < n ? i0 + kPlcpChunk : n;
        uint64_t l = 0;  // lcp(i-1, phi(i-1)), a lower bound + 1 for position i
        for (uint64_t i = i0; i < i1; i++) {
            uint32_t j = isa[i];
            uint32_t v = lcp[j];
            if (lcp_is_pending(v) || !(v & kLcpLowerBound) || j == 0) {
                l = lcp_is_pending(v) ? 0 : v;  // exact value known from the key words
                continue;
            }
            uint64_t prev = sa[j - 1];
            uint64_t ea, eb;
            if (ks.num_n_ranges && n_run_end(ks, prev, ea) && n_run_end(ks, i, eb)) {
                uint64_t ra = ea - prev, rb = eb - i;
                l = ra < rb ? ra : rb;  // <= the true LCP, so still a valid bound for the next position
                lcp[j] = (uint32_t)l;
                continue;
            }
            uint64_t lower = v & ~kLcpLowerBound;
            if (l > 0 && l - 1 > lower) lower = l - 1;
            l = lcp_direct(ks, prev, i, lower);
            uint64_t out = l;
            if (ks.mode == kModeMaxQueryLen && out > ks.cap) out = ks.cap;
            lcp[j] = (uint32_t)out;
        }
    }
}

// ------------------------------------------------------------------ long runs of N (sufr_builder.rs:174-195)
struct NRunStartIn {
    static constexpr bool kFlags = true;  // scan.cuh: scanned with warp votes
    const uint8_t* t;
    __device__ uint32_t operator()(uint64_t i) const { return (t[i] == 'N' && (i == 0 || t[i - 1] != 'N')) ? 1u : 0u; }
};
struct NRunEndIn {
    static constexpr bool kFlags = true;  // scan.cuh: scanned with warp votes
    const uint8_t* t;
    __device__ uint32_t operator()(uint64_t i) const { return (i > 0 && t[i] != 'N' && t[i - 1] == 'N') ? 1u : 0u; }
};
struct IndexOut {
    uint64_t* out;
    __device__ void operator()(uint64_t i, uint32_t v, uint32_t incl) const {
        if (v) out[incl - 1] = i;
    }
};
struct NRunLongIn {
    static constexpr bool kFlags = true;  // scan.cuh: scanned with warp votes
    const uint64_t* starts;
    const uint64_t* ends;
    uint64_t min_len;
    __device__ uint32_t operator()(uint64_t k) const { return (ends[k] - starts[k] >= min_len) ? 1u : 0u; }
};
struct NRunLongOut {
    const uint64_t* starts;
    const uint64_t* ends;
    uint64_t* out_starts;
    uint64_t* out_ends;
    __device__ void operator()(uint64_t k, uint32_t v, uint32_t incl) const {
        if (v) {
            out_starts[incl - 1] = starts[k];
            out_ends[incl - 1] = ends[k];
        }
    }
};

// N-run rule applied to the finished order (full / mql sort with allow_ambiguity): two neighbours that
// both start in recorded runs have LCP min(r1, r2); if additionally r1 == r2 and the bytes after the
// runs are equal the reference calls them equal and emits the larger position first
// (sufr_builder.rs:305-307, 701-712).  Marks such "tie pairs".
__device__ __forceinline__ bool n_tie_pair(const KeySpec& ks, const uint8_t* text, uint64_t pa, uint64_t pb,
                                           uint64_t& lcp_out, bool& both) {
    uint64_t ea, eb;
    both = n_run_end(ks, pa, ea) && n_run_end(ks, pb, eb);
    if (!both) return false;
    uint64_t ra = ea - pa, rb = eb - pb;
    lcp_out = ra < rb ? ra : rb;
    return ra == rb && text[ea] == text[eb];
}
__global__ void __launch_bounds__(kBlock) n_rule_lcp_kernel(KeySpec ks, const uint8_t* __restrict__ text, uint64_t s,
                                                            const pos_t* __restrict__ sa,
                                                            uint32_t* __restrict__ lcp) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t j = 1 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < s; j += stride) {
        uint64_t l;
        bool both;
        n_tie_pair(ks, text, sa[j - 1], sa[j], l, both);
        if (both) {
            if (ks.mode == kModeMaxQueryLen) {
                // find_lcp ignores max_query_len on this branch (sufr_builder.rs:305-307)
            }
            lcp[j] = (uint32_t)l;
        }
    }
}
struct NTieIn {
    static constexpr bool kFlags = true;  // scan.cuh: scanned with warp votes
    KeySpec ks;
    const uint8_t* text;
    const pos_t* sa;
    uint64_t s;
    __device__ bool tie(uint64_t j) const {  // pair (j-1, j)
        if (j == 0 || j >= s) return false;
        uint64_t l;
        bool both;
        return n_tie_pair(ks, text, sa[j - 1], sa[j], l, both);
    }
    __device__ unsigned long long operator()(uint64_t j) const {
        bool t0 = tie(j), t1 = tie(j + 1);
        bool active = t0 || t1;
        bool head = !t0;
        return active ? (1ull | ((unsigned long long)head << 32)) : 0ull;
    }
};
struct NTieOut {
    const pos_t* sa;
    uint32_t* new_slot;
    pos_t* new_pos;
    uint64_t* ck;
    int pos_bits;  // bits of a text position: key = chain << pos_bits | (2^pos_bits - 1 - position)
    __device__ void operator()(uint64_t j, unsigned long long val, unsigned long long incl) const {
        if (val & 1ull) {
            uint32_t a = (uint32_t)incl - 1;
            pos_t p = sa[j];
            new_slot[a] = (uint32_t)j;
            new_pos[a] = p;
            ck[a] = (((incl >> 32) - 1) << pos_bits) | (((1ull << pos_bits) - 1ull) - (uint64_t)p);  // position descending
        }
    }
};

// ------------------------------------------------------------------ suffix filter (sufr_builder.rs:446-449)

struct FilterCountIn {
    static constexpr bool kFlags = true;  // scan.cuh: scanned with warp votes
    const uint8_t* text;
    const pos_t* sa;
    __device__ uint32_t operator()(uint64_t j) const { return indexed_byte(text[sa[j]]) ? 1u : 0u; }
};
struct FilterSaOut {
    const pos_t* sa;
    pos_t* out_sa;
    uint32_t* kept_index;  // compacted index -> original index
    __device__ void operator()(uint64_t j, uint32_t v, uint32_t incl) const {
        if (v) {
            out_sa[incl - 1] = sa[j];
            kept_index[incl - 1] = (uint32_t)j;
        }
    }
};
// LCP of two kept neighbours = min over the skipped stretch.  Values are ordered by (value, is-lower-bound)
// so that an exact value wins over an equal lower bound; segments restart after every kept element.
__device__ __forceinline__ uint32_t lcp_to_ord(uint32_t v) {
    return (v & kLcpLowerBound) ? (((v & ~kLcpLowerBound) << 1) | 1u) : (v << 1);
}
__device__ __forceinline__ uint32_t ord_to_lcp(uint32_t o) { return (o & 1u) ? ((o >> 1) | kLcpLowerBound) : (o >> 1); }
struct FilterLcpIn {
    const uint8_t* text;
    const pos_t* sa;
    const uint32_t* lcp;
    __device__ unsigned long long operator()(uint64_t j) const {
        unsigned long long restart = (j == 0 || indexed_byte(text[sa[j - 1]])) ? 1ull : 0ull;
        return (restart << 32) | lcp_to_ord(lcp[j]);
    }
};
struct FilterLcpOut {
    const uint8_t* text;
    const pos_t* sa;
    const uint32_t* excl_count;  // unused
    uint32_t* scanned;           // per original index: min over its stretch
    __device__ void operator()(uint64_t j, unsigned long long, unsigned long long incl) const {
        scanned[j] = ord_to_lcp((uint32_t)incl);
    }
};
// Suffix filter after a full sort: flags by ballot, ordered block-level compaction, and the LCP of a kept
// element = min over the run of dropped elements before it.  Dropped suffixes cluster by first symbol (all
// N-starts are adjacent), so runs are few but can be millions long: a kept element looks back inside its
// block, and past the block start it walks per-block summaries (min over a block's trailing dropped run).
constexpr int kFilterRows = 8;
// Ranks of the kept suffixes when the WHOLE text was sorted: the suffixes that start with one byte are adjacent in the
// suffix array, so a byte histogram of the text gives the rank ranges of '$', 'A', 'C', 'G', 'T' and no rank has to
// look its first byte up (a random read of the text per rank otherwise).  count == 0: look the byte up.
struct KeepRanges {
    uint64_t lo[5], hi[5];
    int count;
    __device__ bool contains(uint64_t j) const {
        bool k = false;
#pragma unroll
        for (int r = 0; r < 5; r++) k |= r < count && j >= lo[r] && j < hi[r];
        return k;
    }
};
__global__ void __launch_bounds__(kBlock) byte_hist_kernel(const uint8_t* __restrict__ text, uint64_t n,
                                                           unsigned long long* __restrict__ hist) {
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t nvec = (((uintptr_t)text) & 15) == 0 ? n / 16 : 0;
    uint32_t ca = 0, cc = 0, cg = 0, ct = 0, cn = 0;  // the common bytes of a DNA text are counted in registers
    for (uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += stride) {
        if ((ca | cc | cg | ct | cn) > 0x7FFFFF00u) {
            atomicAdd(&hist['A'], (unsigned long long)ca); atomicAdd(&hist['C'], (unsigned long long)cc);
            atomicAdd(&hist['G'], (unsigned long long)cg); atomicAdd(&hist['T'], (unsigned long long)ct);
            atomicAdd(&hist['N'], (unsigned long long)cn);
            ca = cc = cg = ct = cn = 0;
        }
        const uint4 x = reinterpret_cast<const uint4*>(text)[v];
        const uint32_t w[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint32_t ea = swar_eq(w[k], 0x41414141u), ec = swar_eq(w[k], 0x43434343u);
            const uint32_t eg = swar_eq(w[k], 0x47474747u), et = swar_eq(w[k], 0x54545454u);
            ca += __popc(ea); cc += __popc(ec); cg += __popc(eg); ct += __popc(et);
            uint32_t acgt = ea | ec | eg | et;
            if (acgt != 0x80808080u) {  // runs of N (masked regions) would serialise on one shared counter
                const uint32_t en = swar_eq(w[k], 0x4E4E4E4Eu);
                cn += __popc(en);
                acgt |= en;
            }
            if (acgt != 0x80808080u) {
#pragma unroll
                for (int b = 0; b < 4; b++)
                    if (!((acgt >> (8 * b + 7)) & 1u)) atomicAdd(&h[(w[k] >> (8 * b)) & 0xFFu], 1u);
            }
        }
    }
    for (uint64_t i = nvec * 16 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) atomicAdd(&h[text[i]], 1u);
    if (ca) atomicAdd(&hist['A'], (unsigned long long)ca);
    if (cc) atomicAdd(&hist['C'], (unsigned long long)cc);
    if (cg) atomicAdd(&hist['G'], (unsigned long long)cg);
    if (ct) atomicAdd(&hist['T'], (unsigned long long)ct);
    if (cn) atomicAdd(&hist['N'], (unsigned long long)cn);
    __syncthreads();
    if (h[threadIdx.x]) atomicAdd(&hist[threadIdx.x], (unsigned long long)h[threadIdx.x]);
}
__global__ void __launch_bounds__(kBlock) filter_flags_kernel(const uint8_t* __restrict__ text,
                                                              const pos_t* __restrict__ sa, KeepRanges ranges,
                                                              const uint32_t* __restrict__ lcp, uint64_t s,
                                                              uint32_t* __restrict__ flags32,
                                                              uint32_t* __restrict__ block_counts,
                                                              uint32_t* __restrict__ block_tail_min) {
    constexpr int WARPS = kBlock / 32;
    __shared__ uint32_t wsum[WARPS];
    __shared__ uint32_t wmask[kFilterRows * WARPS];
    __shared__ uint32_t wmin[WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t base = (uint64_t)blockIdx.x * kBlock * kFilterRows;
    uint32_t cnt = 0;
#pragma unroll
    for (int r = 0; r < kFilterRows; r++) {
        uint64_t j = base + (uint64_t)r * kBlock + threadIdx.x;
        bool keep = j < s && (ranges.count ? ranges.contains(j) : indexed_byte(text[sa[j]]));
        unsigned m = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) {
            flags32[(base + (uint64_t)r * kBlock) / 32 + warp] = m;
            wmask[r * WARPS + warp] = m;
            cnt += __popc(m);
        }
    }
    if (lane == 0) wsum[warp] = cnt;
    __syncthreads();
    // index (within the block) of the last kept element, -1 if none
    int last_kept = -1;
    for (int i = kFilterRows * WARPS - 1; i >= 0; i--) {
        uint32_t m = wmask[i];
        if (m) { last_kept = i * 32 + 31 - __clz((int)m); break; }
    }
    uint32_t tmin = 0xFFFFFFFFu;
#pragma unroll
    for (int r = 0; r < kFilterRows; r++) {
        int local = r * kBlock + threadIdx.x;
        uint64_t j = base + (uint64_t)local;
        if (local > last_kept && j < s) {
            uint32_t v = lcp[j];
            tmin = v < tmin ? v : tmin;
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        uint32_t o = __shfl_down_sync(0xffffffffu, tmin, off);
        tmin = o < tmin ? o : tmin;
    }
    if (lane == 0) wmin[warp] = tmin;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0, mn = 0xFFFFFFFFu;
        for (int w = 0; w < WARPS; w++) { t += wsum[w]; mn = wmin[w] < mn ? wmin[w] : mn; }
        block_counts[blockIdx.x] = t;
        block_tail_min[blockIdx.x] = mn;
    }
}
struct BlockCountIn {
    const uint32_t* counts;
    __device__ uint32_t operator()(uint64_t b) const { return counts[b]; }
};
struct BlockOffsetOut {
    uint32_t* offsets;
    __device__ void operator()(uint64_t b, uint32_t v, uint32_t incl) const { offsets[b] = incl - v; }
};
__global__ void __launch_bounds__(kBlock) filter_compact_kernel(const pos_t* __restrict__ sa,
                                                                const uint32_t* __restrict__ lcp, uint64_t s,
                                                                const uint32_t* __restrict__ flags32,
                                                                const uint32_t* __restrict__ block_offsets,
                                                                const uint32_t* __restrict__ block_counts,
                                                                const uint32_t* __restrict__ block_tail_min,
                                                                pos_t* __restrict__ out_sa,
                                                                uint32_t* __restrict__ out_lcp) {
    constexpr int WARPS = kBlock / 32;
    __shared__ uint32_t woff[kFilterRows * WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const uint64_t base = (uint64_t)blockIdx.x * kBlock * kFilterRows;
    uint32_t masks[kFilterRows];
#pragma unroll
    for (int r = 0; r < kFilterRows; r++) {
        uint64_t w = (base + (uint64_t)r * kBlock) / 32 + warp;
        masks[r] = (w * 32 < s) ? flags32[w] : 0u;
        if (lane == 0) woff[r * WARPS + warp] = __popc(masks[r]);
    }
    __syncthreads();
    if (warp == 0) {
        static_assert(kFilterRows * WARPS == 64, "warp_scan64");
        warp_scan64(woff);
    }
    __syncthreads();
    const uint32_t block_base = block_offsets[blockIdx.x];
#pragma unroll
    for (int r = 0; r < kFilterRows; r++) {
        if (masks[r] & (1u << lane)) {
            uint64_t j = base + (uint64_t)r * kBlock + threadIdx.x;
            uint32_t v = lcp[j];
            uint64_t jj = j;
            bool found = false;
            while (jj > base) {  // dropped elements before j inside this block
                jj--;
                if ((flags32[jj >> 5] >> (jj & 31)) & 1u) { found = true; break; }
                uint32_t x = lcp[jj];
                v = x < v ? x : v;
            }
            if (!found) {  // reached the block start: walk the summaries of the blocks before
                for (long long b = (long long)blockIdx.x - 1; b >= 0; b--) {
                    uint32_t x = block_tail_min[b];
                    v = x < v ? x : v;
                    if (block_counts[b]) break;  // that block holds a kept element: its trailing run ends the walk
                }
            }
            uint32_t dst = block_base + woff[r * WARPS + warp] + __popc(masks[r] & lt_mask);
            out_sa[dst] = sa[j];
            out_lcp[dst] = v;
        }
    }
}

__global__ void __launch_bounds__(kBlock) gather_u32_kernel(uint64_t m, const uint32_t* __restrict__ idx,
                                                            const uint32_t* __restrict__ src,
                                                            uint32_t* __restrict__ dst) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t a = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; a < m; a += stride) dst[a] = src[idx[a]];
}

// Compact device->host transfer of the LCP array: almost every value of a non-repetitive text fits a byte.
// out8[j] = min(lcp[j], 255); values >= 255 are also appended to an exception list (index, value).
__global__ void __launch_bounds__(kBlock) lcp_to_u8_kernel(const uint32_t* __restrict__ lcp, uint64_t s,
                                                           uint8_t* __restrict__ out8, uint32_t* __restrict__ exc_idx,
                                                           uint32_t* __restrict__ exc_val,
                                                           unsigned long long* __restrict__ exc_count, uint64_t capacity) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x * 16;
    for (uint64_t j0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 16; j0 < s; j0 += stride) {
        uint32_t v[16];
        if (j0 + 15 < s) {
#pragma unroll
            for (int q = 0; q < 4; q++) {
                uint4 x = *reinterpret_cast<const uint4*>(lcp + j0 + 4 * q);
                v[4 * q] = x.x; v[4 * q + 1] = x.y; v[4 * q + 2] = x.z; v[4 * q + 3] = x.w;
            }
        } else {
#pragma unroll
            for (int u = 0; u < 16; u++) v[u] = j0 + u < s ? lcp[j0 + u] : 0u;
        }
        uint32_t packed[4] = {0, 0, 0, 0};
#pragma unroll
        for (int u = 0; u < 16; u++) {
            uint32_t b = v[u] < 255u ? v[u] : 255u;
            packed[u >> 2] |= b << (8 * (u & 3));
            if (v[u] >= 255u && j0 + u < s) {
                unsigned long long e = atomicAdd(exc_count, 1ull);
                if (e < capacity) {
                    exc_idx[e] = (uint32_t)(j0 + u);
                    exc_val[e] = v[u];
                }
            }
        }
        if (j0 + 15 < s) {
            *reinterpret_cast<uint4*>(out8 + j0) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
        } else {
#pragma unroll
            for (int u = 0; u < 16; u++)
                if (j0 + u < s) out8[j0 + u] = (uint8_t)(packed[u >> 2] >> (8 * (u & 3)));
        }
    }
}

// 64-bit results written early by round 0: copy the entries the refinement changed afterwards.
__global__ void __launch_bounds__(kBlock) wide_patch_kernel(uint64_t m, const uint32_t* __restrict__ slots,
                                                            const pos_t* __restrict__ sa,
                                                            const uint32_t* __restrict__ lcp,
                                                            unsigned long long* __restrict__ sa64,
                                                            unsigned long long* __restrict__ lcp64) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += stride) {
        const uint32_t t = slots[i];
        sa64[t] = sa[t];
        lcp64[t] = lcp[t];
    }
}

// Element-wise width conversion of a result array (u32 -> u64 widening of LCP values, u64 -> u32 narrowing of
// positions when a short text was built with 64-bit positions).
template <typename Src, typename Dst>
__global__ void __launch_bounds__(kBlock) convert_kernel(uint64_t m, const Src* __restrict__ src, Dst* __restrict__ dst) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += stride) dst[i] = (Dst)src[i];
}

// Both result arrays in one launch: 16-byte loads, two 16-byte streaming stores per load (arrays 16-byte aligned).
__global__ void __launch_bounds__(kBlock) widen2_kernel(uint64_t m, const uint32_t* __restrict__ sa,
                                                        const uint32_t* __restrict__ lcp,
                                                        unsigned long long* __restrict__ sa64,
                                                        unsigned long long* __restrict__ lcp64) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t nvec = m / 4;
    for (uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += stride) {
        const uint4 a = __ldcs(reinterpret_cast<const uint4*>(sa) + v);
        const uint4 b = __ldcs(reinterpret_cast<const uint4*>(lcp) + v);
        __stcs(reinterpret_cast<ulonglong2*>(sa64) + 2 * v, make_ulonglong2(a.x, a.y));
        __stcs(reinterpret_cast<ulonglong2*>(sa64) + 2 * v + 1, make_ulonglong2(a.z, a.w));
        __stcs(reinterpret_cast<ulonglong2*>(lcp64) + 2 * v, make_ulonglong2(b.x, b.y));
        __stcs(reinterpret_cast<ulonglong2*>(lcp64) + 2 * v + 1, make_ulonglong2(b.z, b.w));
    }
    for (uint64_t i = nvec * 4 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += stride) {
        sa64[i] = sa[i];
        lcp64[i] = lcp[i];
    }
}


// u32 positions and LCP values -> u64 result arrays (widen2_kernel); with 64-bit positions only the LCP values widen
// (Build::run), so this overload is never reached.
inline void widen_both(uint64_t s, const uint32_t* sa, const uint32_t* lcp, unsigned long long* sa64, unsigned long long* lcp64,
                       cudaStream_t stream) {
    widen2_kernel<<<grid_for(s, 8), kBlock, 0, stream>>>(s, sa, lcp, sa64, lcp64);
    SUFR_KERNEL_CHECK();
}
inline void widen_both(uint64_t, const uint64_t*, const uint32_t*, unsigned long long*, unsigned long long*, cudaStream_t) {
    throw Error(3, "widen_both called with 64-bit positions");
}
