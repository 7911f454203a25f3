// The consumer of the LCP array: LCP-subsampled suffix array and batched search over a
// device-resident index.  Mirrors, for one batch of queries at a time,
//   SufrFile::subsample_suffix_array   libsufr/src/sufr_file.rs:429-456   keep (suffix, rank) where lcp < max_query_len
//   SufrSearch::search                 libsufr/src/sufr_search.rs:104-176 first / last occurrence, rank mapping
//   suffix_search_first / _last        :179-249                          binary search with LCP skipping
//   compare                            :261-350                          full / max-query-len / seed-mask comparison
// One thread per query; the reference runs one rayon task per query (sufr_file.rs:816-826).
#pragma once
#include "common.cuh"

namespace sufr {
namespace search {

struct Params {
    const uint8_t* text;
    uint64_t n;               // text length
    const void* sa;           // the array that is searched: full suffix array, or the subsampled one
    const void* rank;         // ranks of the subsampled entries in the full array (NULL: searching the full array)
    uint64_t len;             // entries of `sa`
    int wide;                 // entries are u64 (else u32)
    int is_mask;              // SuffixSortType::Mask (else MaxQueryLen(build_mql))
    uint64_t build_mql;       // max_query_len the index was built with (0 = none)
    int has_run_mql;          // a max_query_len was given at search time
    uint64_t run_mql;
    const uint32_t* mask_pos; // seed mask: offsets of the care positions
    uint32_t weight;
    uint32_t mask_len;
};

struct Comparison {
    uint64_t lcp;
    int cmp;  // query vs suffix: -1 Less, 0 Equal, 1 Greater
};

__device__ __forceinline__ uint64_t entry(const void* a, uint64_t i, int wide) {
    return wide ? reinterpret_cast<const unsigned long long*>(a)[i] : (uint64_t) reinterpret_cast<const uint32_t*>(a)[i];
}

// util.rs:19-37
__device__ __forceinline__ uint64_t full_offset(const Params& P, uint64_t lcp) {
    if (!P.is_mask || lcp == 0 || lcp > P.mask_len) return lcp;
    const uint64_t offset = P.mask_pos[lcp - 1];
    const uint64_t next_offset = lcp < P.weight ? P.mask_pos[lcp] : 0;
    if (next_offset > offset && next_offset - offset > 1) return next_offset;
    return offset + 1;
}

// sufr_search.rs:261-350
__device__ Comparison compare(const Params& P, const uint8_t* q, uint64_t qlen, uint64_t suffix_pos, uint64_t skip) {
    uint64_t lcp, mql;
    if (!P.is_mask) {
        if (P.build_mql > 0 && P.has_run_mql) mql = P.build_mql < P.run_mql ? P.build_mql : P.run_mql;
        else if (P.has_run_mql) mql = P.run_mql;
        else mql = P.build_mql;
        if (mql > 0 && skip >= mql) {
            lcp = skip;
        } else {
            const uint64_t text_start = suffix_pos + skip;
            uint64_t text_end = mql > 0 ? text_start + mql : text_start + qlen;
            if (text_end > P.n) text_end = P.n;
            uint64_t c = 0;
            while (skip + c < qlen && text_start + c < text_end && q[skip + c] == P.text[text_start + c]) c++;
            lcp = c + skip;
        }
    } else {
        mql = P.has_run_mql ? P.run_mql : 0;
        if (skip >= P.weight || (mql > 0 && skip >= mql)) {
            lcp = skip;
        } else {
            const uint64_t end = mql > 0 ? (mql < P.weight ? mql : P.weight) : P.weight;
            uint64_t query_len = 0, suffix_len = 0;
            for (uint64_t k = skip; k < end; k++) {
                if (P.mask_pos[k] < qlen) query_len++;
                if (suffix_pos + P.mask_pos[k] < P.n) suffix_len++;
            }
            const uint64_t len = query_len < suffix_len ? query_len : suffix_len;
            uint64_t c = 0;
            while (c < len) {
                const uint64_t off = P.mask_pos[skip + c];
                if (suffix_pos + off >= P.n || q[off] != P.text[suffix_pos + off]) break;
                c++;
            }
            lcp = skip + c;
        }
    }
    Comparison r;
    r.lcp = lcp;
    if (mql > 0 && lcp >= mql) {
        r.cmp = 0;  // seen enough
    } else {
        const uint64_t fo = full_offset(P, lcp);
        if (fo >= qlen) r.cmp = 0;                        // the entire query matched
        else if (suffix_pos + fo >= P.n) r.cmp = 1;       // (the reference panics here; the end of the text sorts first)
        else r.cmp = q[fo] < P.text[suffix_pos + fo] ? -1 : (q[fo] > P.text[suffix_pos + fo] ? 1 : 0);
    }
    return r;
}

// sufr_search.rs:179-213
__device__ long long search_first(const Params& P, const uint8_t* q, uint64_t qlen) {
    long long low = 0, high = (long long)P.len - 1;
    uint64_t left_lcp = 0, right_lcp = 0;
    while (high >= low) {
        const long long mid = low + (high - low) / 2;
        const uint64_t mid_val = entry(P.sa, mid, P.wide);
        const Comparison c = compare(P, q, qlen, mid_val, left_lcp < right_lcp ? left_lcp : right_lcp);
        const uint64_t before = mid > 0 ? entry(P.sa, mid - 1, P.wide) : mid_val;
        if (c.cmp == 0 && (mid == 0 || compare(P, q, qlen, before, 0).cmp == 1)) return mid;
        if (c.cmp == 1) { low = mid + 1; left_lcp = c.lcp; }
        else { high = mid - 1; right_lcp = c.lcp; }
    }
    return -1;
}

// sufr_search.rs:216-249
__device__ long long search_last(const Params& P, const uint8_t* q, uint64_t qlen, long long low) {
    const long long n = (long long)P.len;
    long long high = n - 1;
    uint64_t left_lcp = 0, right_lcp = 0;
    while (high >= low) {
        const long long mid = low + (high - low) / 2;
        const uint64_t mid_val = entry(P.sa, mid, P.wide);
        const Comparison c = compare(P, q, qlen, mid_val, left_lcp < right_lcp ? left_lcp : right_lcp);
        const uint64_t after = mid < n - 1 ? entry(P.sa, mid + 1, P.wide) : mid_val;
        if (c.cmp == 0 && (mid == n - 1 || compare(P, q, qlen, after, 0).cmp == -1)) return mid;
        if (c.cmp == -1) { high = mid - 1; right_lcp = c.lcp; }
        else { low = mid + 1; left_lcp = c.lcp; }
    }
    return -1;
}

// One thread per query.  rank_begin / rank_end = the half-open range of ranks in the FULL suffix array
// (SearchResultLocations::ranks), both ~0 when the query does not occur.
__global__ void __launch_bounds__(128) search_kernel(Params P, const uint8_t* __restrict__ queries,
                                                     const uint64_t* __restrict__ offsets, uint64_t num_queries,
                                                     unsigned long long* __restrict__ rank_begin,
                                                     unsigned long long* __restrict__ rank_end) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= num_queries) return;
    const uint8_t* q = queries + offsets[i];
    const uint64_t qlen = offsets[i + 1] - offsets[i];
    rank_begin[i] = ~0ull;
    rank_end[i] = ~0ull;
    if (P.len == 0) return;
    const long long start = search_first(P, q, qlen);
    if (start < 0) return;
    long long end = search_last(P, q, qlen, start);
    if (end < 0) end = start;
    if (!P.rank) {
        rank_begin[i] = (unsigned long long)start;
        rank_end[i] = (unsigned long long)end + 1;
    } else {  // compressed suffix array: sufr_search.rs:121-139
        rank_begin[i] = entry(P.rank, start, P.wide);
        if (start == end) rank_end[i] = (uint64_t)start == P.len - 1 ? P.len : entry(P.rank, start + 1, P.wide);
        else rank_end[i] = entry(P.rank, end, P.wide) + 1;
    }
}

// sufr_file.rs:443-453: the entries with lcp < max_query_len, and their ranks
struct SubsampleIn {
    static constexpr bool kFlags = true;  // scan.cuh: scanned with warp votes
    const void* lcp;
    int wide;
    uint64_t mql;
    __device__ uint32_t operator()(uint64_t i) const { return entry(lcp, i, wide) < mql ? 1u : 0u; }
};
struct SubsampleOut {
    const void* sa;
    int wide;
    void* out_sa;
    void* out_rank;
    __device__ void operator()(uint64_t i, uint32_t v, uint32_t incl) const {
        if (!v) return;
        if (wide) {
            reinterpret_cast<unsigned long long*>(out_sa)[incl - 1] = reinterpret_cast<const unsigned long long*>(sa)[i];
            reinterpret_cast<unsigned long long*>(out_rank)[incl - 1] = i;
        } else {
            reinterpret_cast<uint32_t*>(out_sa)[incl - 1] = reinterpret_cast<const uint32_t*>(sa)[i];
            reinterpret_cast<uint32_t*>(out_rank)[incl - 1] = (uint32_t)i;
        }
    }
};

}  // namespace search
}  // namespace sufr
