// Full on-device verification of a finished build, independent of the build's own data structures: it reads
// only the TRANSFORMED TEXT BYTES, the suffix array and the LCP array.
//
//   positions   every SA entry is a text position the reference indexes (sufr_builder.rs:446-449), and no position
//               occurs twice (bitmap); together with the expected count this makes SA a permutation of the indexed set
//   pairs       for EVERY rank j >= 1: (SA[j-1], SA[j]) are in the order the mode defines and LCP[j] is the mode's
//               value, by direct comparison of the two suffixes (eight text bytes per step):
//                 full sort      byte order, the shorter suffix first when one is a prefix of the other; exact LCP
//                 N-run rule     both suffixes inside recorded runs of N (sufr_builder.rs:305-307, :701-712):
//                                LCP = min(r1, r2); equal (r, next byte) is a tie -> larger position first
//                 max-query-len  order on the first Q bytes; LCP = min(true LCP, Q); Q-equal prefixes are ties ->
//                                larger position first (what this library emits)
//                 seed mask      order on the care positions that exist, ties -> larger position first; LCP in care
//                                units (types.rs:36-200, SURVEY 8a "Semantics distilled" 2)
//   LCP[0] == 0 (or, for shard > 0 of a sharded build, the seam value against the previous shard's last suffix)
//
// Cost control for texts with deep repeats (LCP values of 10^6 and more make "compare every pair directly"
// quadratic): the pair pass spends at most kPairBudget bytes per pair and DEFERS the others (bitmap).  Deferred
// pairs are then settled either by unbounded direct comparison (few of them), or -- when every text position is
// indexed and the sort is a full sort -- by the classic linear-time proof: the order of (a, b) with equal first
// bytes follows from the ranks of a+1 and b+1 (inverse suffix array), and with the order established the LCP
// values follow Kasai's bound LCP(i) >= LCP(i-1) - 1 in text order (chunks of kKasaiChunk positions).
#pragma once
#include "common.cuh"

namespace sufr {
namespace verify {

struct Params {
    const uint8_t* text;  // transformed text, readable up to text[n + 15]
    uint64_t n;
    const void* sa;
    const void* lcp;
    uint64_t s;
    int wide;             // elements are u64 (else u32)
    int filter;           // --dna without --allow-ambiguity: only ACGT$ positions are indexed
    int mode;             // 0 full, 1 max-query-len, 2 seed mask
    uint64_t q;           // max-query-len
    const uint32_t* mask_pos;
    uint32_t weight;
    const uint64_t* n_starts;
    const uint64_t* n_ends;
    uint32_t num_n_ranges;
    int has_prev;         // sharded: rank 0 of this array follows `prev_last` of the previous shard
    uint64_t prev_last;
};

constexpr uint64_t kPairBudget = 2048;   // bytes compared per pair before it is deferred
constexpr uint32_t kKasaiChunk = 4096;

struct Report {  // all counters are over this array (shard)
    unsigned long long pairs_checked;
    unsigned long long order_errors;
    unsigned long long lcp_errors;
    unsigned long long out_of_range;
    unsigned long long not_indexed;
    unsigned long long duplicates;
    unsigned long long first_bad_rank;  // smallest rank with an order / LCP error (~0 if none)
    unsigned long long max_lcp;
    unsigned long long lcp_sum;
    unsigned long long deferred;        // pairs whose comparison exceeded the budget
};

__device__ __forceinline__ uint64_t elem(const void* a, uint64_t i, int wide) {
    return wide ? reinterpret_cast<const unsigned long long*>(a)[i] : (uint64_t) reinterpret_cast<const uint32_t*>(a)[i];
}

__device__ __forceinline__ bool idx_byte(uint8_t c) { return c == '$' || c == 'A' || c == 'C' || c == 'G' || c == 'T'; }

// eight text bytes starting at p (little-endian: byte p in the low bits); the text is padded
__device__ __forceinline__ uint64_t load8(const uint8_t* t, uint64_t p) {
    const uint64_t* w = reinterpret_cast<const uint64_t*>(t - ((uintptr_t)t & 7));
    const uint64_t a = p + ((uintptr_t)t & 7);
    const uint64_t q = a >> 3;
    const uint32_t r = (uint32_t)(a & 7) * 8;
    const uint64_t lo = __ldg(w + q);
    if (r == 0) return lo;
    return (lo >> r) | (__ldg(w + q + 1) << (64 - r));
}

// common prefix of the suffixes at a and b, at most `limit` bytes, starting from a known common prefix `l`
__device__ __forceinline__ uint64_t common_prefix(const uint8_t* t, uint64_t a, uint64_t b, uint64_t limit, uint64_t l = 0) {
    while (l < limit) {
        const uint64_t x = load8(t, a + l) ^ load8(t, b + l);
        if (x) {
            l += (uint64_t)(__ffsll((long long)x) - 1) >> 3;
            break;
        }
        l += 8;
    }
    return l < limit ? l : limit;
}

__device__ __forceinline__ bool run_end(const Params& P, uint64_t p, uint64_t& end) {
    uint32_t lo = 0, hi = P.num_n_ranges;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        const uint64_t s = P.n_starts[mid], e = P.n_ends[mid];
        if (s <= p && p < e) { end = e; return true; }
        if (s < p) lo = mid + 1; else hi = mid;
    }
    return false;
}

// expected LCP of the ordered pair (a, b) and whether the order is right.  With budget > 0 a comparison that
// needs more than `budget` bytes sets `deferred` instead of finishing.
__device__ __forceinline__ bool check_pair(const Params& P, uint64_t a, uint64_t b, uint64_t& want, uint64_t budget,
                                           bool& deferred) {
    deferred = false;
    const uint8_t* t = P.text;
    const uint64_t n = P.n;
    if (P.mode == 2) {
        uint32_t c = 0;
        for (; c < P.weight; c++) {
            const uint64_t pa = a + P.mask_pos[c], pb = b + P.mask_pos[c];
            if (pa >= n || pb >= n) {  // one key ends here: the shorter key (larger position) comes first
                want = c;
                return a > b;
            }
            if (t[pa] != t[pb]) {
                want = c;
                return t[pa] < t[pb];
            }
        }
        want = P.weight;
        return a > b;
    }
    uint64_t ea, eb;
    if (P.num_n_ranges && run_end(P, a, ea) && run_end(P, b, eb)) {
        const uint64_t ra = ea - a, rb = eb - b;
        want = ra < rb ? ra : rb;
        if (ra != rb) return t[a + want] < t[b + want];
        if (t[ea] != t[eb]) return t[ea] < t[eb];
        return a > b;
    }
    uint64_t limit = n - (a > b ? a : b);
    const bool capped = P.mode == 1 && P.q < limit;
    if (capped) limit = P.q;
    uint64_t l;
    if (budget && limit > budget) {
        l = common_prefix(t, a, b, budget);
        if (l >= budget) { deferred = true; return true; }
    } else {
        l = common_prefix(t, a, b, limit);
    }
    want = l;
    if (l < limit) return t[a + l] < t[b + l];
    return a > b;  // Q-equal prefixes, or a is a proper prefix of b: the larger position first
}

__global__ void __launch_bounds__(256) positions_kernel(Params P, uint32_t* __restrict__ bitmap, Report* __restrict__ rep) {
    unsigned long long oor = 0, nidx = 0, dup = 0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < P.s; j += stride) {
        const uint64_t p = elem(P.sa, j, P.wide);
        if (p >= P.n) { oor++; continue; }
        if (P.filter && !idx_byte(P.text[p])) nidx++;
        const uint32_t bit = 1u << (p & 31);
        if (atomicOr(&bitmap[p >> 5], bit) & bit) dup++;
    }
    if (oor) atomicAdd(&rep->out_of_range, oor);
    if (nidx) atomicAdd(&rep->not_indexed, nidx);
    if (dup) atomicAdd(&rep->duplicates, dup);
}

// pass 1 (only_deferred == 0): every pair with a byte budget, deferred pairs marked in `defer_bits`;
// pass 2 (only_deferred == 1): the marked pairs, unbounded.
__global__ void __launch_bounds__(256) pairs_kernel(Params P, Report* __restrict__ rep, uint32_t* __restrict__ defer_bits,
                                                    int only_deferred) {
    unsigned long long ord = 0, bad = 0, cnt = 0, mx = 0, sum = 0, first = ~0ull, def = 0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < P.s; j += stride) {
        if (only_deferred && !((defer_bits[j >> 5] >> (j & 31)) & 1u)) continue;
        const uint64_t b = elem(P.sa, j, P.wide);
        const uint64_t got = elem(P.lcp, j, P.wide);
        if (!only_deferred) {
            if (got > mx) mx = got;
            sum += got;
        }
        if (j == 0 && !P.has_prev) {
            if (got != 0) { bad++; first = 0; }
            continue;
        }
        const uint64_t a = j ? elem(P.sa, j - 1, P.wide) : P.prev_last;
        if (a >= P.n || b >= P.n) continue;  // counted by positions_kernel
        uint64_t want = 0;
        bool deferred = false;
        const bool ok = a != b && check_pair(P, a, b, want, only_deferred ? 0 : kPairBudget, deferred);
        if (deferred) {
            def++;
            atomicOr(&defer_bits[j >> 5], 1u << (j & 31));
            continue;
        }
        cnt++;
        if (!ok) ord++;
        if (want != got) bad++;
        if ((!ok || want != got) && j < first) first = j;
    }
    if (cnt) atomicAdd(&rep->pairs_checked, cnt);
    if (ord) atomicAdd(&rep->order_errors, ord);
    if (bad) atomicAdd(&rep->lcp_errors, bad);
    if (sum) atomicAdd(&rep->lcp_sum, sum);
    if (def) atomicAdd(&rep->deferred, def);
    atomicMax(&rep->max_lcp, mx);
    if (first != ~0ull) atomicMin(&rep->first_bad_rank, first);
}

// ---- linear-time settlement of the deferred pairs (full sort, every position indexed, unsharded)
__global__ void __launch_bounds__(256) isa_kernel(Params P, uint32_t* __restrict__ isa) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < P.s; j += stride) {
        const uint64_t p = elem(P.sa, j, P.wide);
        if (p < P.n) isa[p] = (uint32_t)j;
    }
}
// Order by the ranks of the successors (Burkhardt & Karkkainen's suffix-array check): with SA a permutation, it is
// the suffix order iff for EVERY adjacent pair (a, b): t[a] < t[b], or t[a] == t[b] and (a is the last position or
// rank(a+1) < rank(b+1)).  Applied to every pair, so the criterion holds as a whole; only the deferred pairs still
// count as "checked" here (the others were also compared directly).  Pairs the N-run rule orders by position -- the
// pair itself or its successors inside recorded runs -- are compared by the rule instead.
__global__ void __launch_bounds__(256) order_by_rank_kernel(Params P, const uint32_t* __restrict__ isa,
                                                            const uint32_t* __restrict__ defer_bits, Report* __restrict__ rep) {
    unsigned long long ord = 0, cnt = 0, first = ~0ull;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t j = 1 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < P.s; j += stride) {
        const bool is_def = (defer_bits[j >> 5] >> (j & 31)) & 1u;
        const uint64_t a = elem(P.sa, j - 1, P.wide), b = elem(P.sa, j, P.wide);
        if (a >= P.n || b >= P.n) continue;
        bool ok;
        uint64_t ea, eb;
        if (P.num_n_ranges && run_end(P, a, ea) && run_end(P, b, eb)) {
            continue;  // ordered by the rule; checked by the direct pass (never deferred)
        } else if (P.text[a] != P.text[b]) {
            ok = P.text[a] < P.text[b];
        } else if (a + 1 >= P.n || b + 1 >= P.n) {
            ok = a + 1 >= P.n;  // the one-byte suffix is a prefix of the other
        } else if (P.num_n_ranges && run_end(P, a + 1, ea) && run_end(P, b + 1, eb)) {
            // the successors may be a tie that the reference orders by position: compare the pair itself
            if (!is_def) continue;  // done by the direct pass
            uint64_t want;
            bool d;
            ok = check_pair(P, a, b, want, 0, d);
        } else {
            ok = isa[a + 1] < isa[b + 1];
        }
        if (is_def) cnt++;
        if (!ok) { ord++; if (j < first) first = j; }
    }
    if (cnt) atomicAdd(&rep->pairs_checked, cnt);
    if (ord) atomicAdd(&rep->order_errors, ord);
    if (first != ~0ull) atomicMin(&rep->first_bad_rank, first);
}
// LCP of the deferred pairs by Kasai's bound, in text order
__global__ void __launch_bounds__(256) kasai_kernel(Params P, const uint32_t* __restrict__ isa,
                                                    const uint32_t* __restrict__ defer_bits, Report* __restrict__ rep) {
    unsigned long long bad = 0, first = ~0ull;
    const uint64_t chunks = (P.n + kKasaiChunk - 1) / kKasaiChunk;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; c < chunks; c += stride) {
        const uint64_t i0 = c * kKasaiChunk, i1 = i0 + kKasaiChunk < P.n ? i0 + kKasaiChunk : P.n;
        uint64_t l = 0;  // true LCP of the previous position with its predecessor in suffix order (0 = unknown)
        for (uint64_t i = i0; i < i1; i++) {
            const uint64_t j = isa[i];
            if (j == 0) { l = 0; continue; }
            const uint64_t a = elem(P.sa, j - 1, P.wide);
            uint64_t ea, eb;
            if (a >= P.n || (P.num_n_ranges && run_end(P, a, ea) && run_end(P, i, eb))) { l = 0; continue; }
            const bool is_def = (defer_bits[j >> 5] >> (j & 31)) & 1u;
            const uint64_t limit = P.n - (a > i ? a : i);
            if (!is_def) {
                // settled by the direct pass: its LCP (< budget) is exact and serves as the bound for i + 1
                const uint64_t got = elem(P.lcp, j, P.wide);
                l = got < limit ? got : limit;
                continue;
            }
            const uint64_t l0 = l > 0 ? l - 1 : 0;
            l = common_prefix(P.text, a, i, limit, l0 < limit ? l0 : limit);
            if (l != elem(P.lcp, j, P.wide)) { bad++; if (j < first) first = j; }
        }
    }
    if (bad) atomicAdd(&rep->lcp_errors, bad);
    if (first != ~0ull) atomicMin(&rep->first_bad_rank, first);
}

// number of text positions the reference indexes
__global__ void __launch_bounds__(256) count_indexed_kernel(const uint8_t* __restrict__ t, uint64_t n, unsigned long long* out) {
    unsigned long long c = 0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) c += idx_byte(t[i]) ? 1 : 0;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) c += __shfl_down_sync(0xffffffffu, c, off);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}

}  // namespace verify
}  // namespace sufr
