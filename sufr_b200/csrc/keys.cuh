// Packed text and key-word extraction.
//
// The transformed text (reference: sufr_builder.rs:144-160) is re-coded with an order-preserving dense
// alphabet (code = 1 + rank of the byte among the bytes present; 0 = "beyond the end of the text", which
// sorts before every byte -- the reference's prefix rule, sufr_builder.rs:376-378) and packed
// `K = floor(64 / bits)` symbols per 64-bit word, first symbol in the most significant bits, unused low
// bits zero.  Integer comparison of key words is then lexicographic comparison of K symbols, and the
// common-prefix length of two words is clz(x ^ y) / bits.
#pragma once
#include "common.cuh"

namespace sufr {

struct PackedText {
    const uint64_t* words;  // ceil(n / K) + 2 words, zero padded
    uint64_t n;             // text length in symbols
    uint32_t bits;          // bits per symbol
    uint32_t K;             // symbols per word
    uint64_t keep_mask;     // clears the unused low bits of a word
    uint32_t sym_mask;
};

enum KeyMode : int { kModeFull = 0, kModeMaxQueryLen = 1, kModeMask = 2 };

struct KeySpec {
    PackedText pt;
    int mode;
    uint64_t cap;              // key length cap in symbols: Q (mql), weight (mask), UINT64_MAX (full)
    const uint32_t* mask_pos;  // device: offsets of the care positions (types.rs:192-199)
    uint32_t weight;
    uint32_t mask_len;
    // long runs of N (sufr_builder.rs:174-195), device arrays sorted by start; only consulted in
    // full / max-query-len mode (find_lcp, sufr_builder.rs:301-307)
    const uint64_t* n_starts;
    const uint64_t* n_ends;
    uint32_t num_n_ranges;
    // 2-bit fast path of the first sort (full sort of DNA-like texts); see first_key()
    int fast2;
    const uint64_t* packed2;   // 2 bits per symbol, 32 per word: rank among the 4 regular bytes, or the class of an irregular byte
    const uint64_t* irr;       // 1 bit per symbol, 64 per word: symbol is not one of the 4 regular bytes (or is beyond the text)
    const uint8_t* text;       // transformed text
    const uint8_t* cls;        // [256] byte -> number of regular bytes smaller than it (0..4); regular bytes: their rank
    uint64_t packed2_words;    // allocation sizes (bounds of the per-lane word loads)
    uint64_t irr_words;
    int reg_indexed;           // every regular byte is one of ACGT$: only irregular positions can be filtered out
};

// Fast-path keys: bit 0 = the key contains fill (an irregular symbol inside its 31-symbol window),
// bits 63..2 = 31 symbols.  The radix sort covers only the top kFast2SortBits (4 passes); groups of up to
// kFast2SmallGroup keys that tie on those bits are then ordered by their full keys in registers
// (round0_fast2_kernel), larger groups and exact ties on all 31 symbols go to the 3-bit refinement.
constexpr int kShardHistBits = 12;   // multi-GPU key ranges are cut at multiples of 2^(64 - kShardHistBits)
constexpr int kFast2Symbols = 31;
constexpr int kFast2SortBits = 32;
constexpr int kFast2ProbeBits = 40;  // the repetitiveness probe sorts its sample deeper
constexpr int kFast2SmallGroup = 8;
constexpr uint64_t kFast2TopMask = ~0ull << (64 - kFast2SortBits);
// Group key of a sorted element: members of a large (unsorted) group agree on the sorted bits only, everything
// else is ordered by the full symbol bits.
__host__ __device__ __forceinline__ uint64_t fast2_canon(uint64_t key, bool large) {
    return large ? (key & kFast2TopMask) : (key & ~3ull);
}

__device__ __forceinline__ void split_pos(const PackedText& pt, uint64_t p, uint64_t& q, uint32_t& r) {
    if (pt.n <= 0xFFFFFFFFull) {
        uint32_t p32 = (uint32_t)p;
        uint32_t q32 = p32 / pt.K;
        q = q32;
        r = p32 - q32 * pt.K;
    } else {
        q = p / pt.K;
        r = (uint32_t)(p - q * pt.K);
    }
}

// K symbols starting at text position p, top-aligned; symbols beyond the text are 0.
__device__ __forceinline__ uint64_t load_key(const PackedText& pt, uint64_t p) {
    if (p >= pt.n) return 0;
    uint64_t q;
    uint32_t r;
    split_pos(pt, p, q, r);
    uint64_t w0 = __ldg(pt.words + q);
    if (r == 0) return w0;
    uint64_t w1 = __ldg(pt.words + q + 1);
    uint32_t sh = r * pt.bits;
    uint32_t used = pt.K * pt.bits;
    return ((w0 << sh) | (w1 >> (used - sh))) & pt.keep_mask;
}

__device__ __forceinline__ uint32_t sym_at(const PackedText& pt, uint64_t i) {
    if (i >= pt.n) return 0;
    uint64_t q;
    uint32_t r;
    split_pos(pt, i, q, r);
    return (uint32_t)(__ldg(pt.words + q) >> (64 - pt.bits * (r + 1))) & pt.sym_mask;
}

// Word `w` of the sort key of suffix p (K symbols of the key starting at key offset w*K).
__device__ __forceinline__ uint64_t key_word(const KeySpec& ks, uint64_t p, uint32_t w) {
    const uint32_t K = ks.pt.K;
    if (ks.mode == kModeFull) return load_key(ks.pt, p + (uint64_t)w * K);
    if (ks.mode == kModeMaxQueryLen) {
        uint64_t done = (uint64_t)w * K;
        if (done >= ks.cap) return 0;
        uint64_t x = load_key(ks.pt, p + done);
        uint64_t rem = ks.cap - done;
        if (rem < K) x &= ~0ull << (64 - (uint32_t)rem * ks.pt.bits);
        return x;
    }
    uint64_t x = 0;
    uint32_t k0 = w * K;
    for (uint32_t j = 0; j < K; j++) {
        uint32_t k = k0 + j;
        if (k >= ks.weight) break;
        uint64_t s = sym_at(ks.pt, p + ks.mask_pos[k]);
        x |= s << (64 - ks.pt.bits * (j + 1));
    }
    return x;
}

// First sort key of suffix p on the 2-bit fast path.  Let r0 < r1 < r2 < r3 be the regular bytes.  Regular
// symbols are coded by their rank.  The first irregular symbol x (or the end of the text) is coded
// min(c, 3) with c = #{regular bytes < x} and every later position is FILL: zeros when c <= 3 (x sorts before
// every continuation of r_c), ones when c == 4 (x sorts after every continuation of r3).  This keeps
// key(a) <= key(b) whenever suffix a < suffix b, so sorting by the key is consistent with the suffix order
// and only ties need the exact comparison.
__device__ __forceinline__ uint64_t first_key_fast2(const KeySpec& ks, uint64_t p) {
    const uint64_t n = ks.pt.n;
    uint64_t q = p >> 5;
    uint32_t r = (uint32_t)(p & 31);
    uint64_t w0 = __ldg(ks.packed2 + q);
    uint64_t w = r ? ((w0 << (2 * r)) | (__ldg(ks.packed2 + q + 1) >> (64 - 2 * r))) : w0;
    uint64_t q2 = p >> 6;
    uint32_t r2 = (uint32_t)(p & 63);
    uint64_t m0 = __ldg(ks.irr + q2);
    uint64_t m = r2 ? ((m0 << r2) | (__ldg(ks.irr + q2 + 1) >> (64 - r2))) : m0;
    m &= ~0ull << (64 - kFast2Symbols);
    w &= ~3ull;
    if (m == 0) return w;
    uint32_t j = (uint32_t)__clzll((long long)m);  // first irregular symbol of the window
    uint32_t c = (p + j < n) ? ks.cls[ks.text[p + j]] : 0u;
    uint64_t keep = ~0ull << (62 - 2 * j);           // symbols 0..j (symbol j already holds min(c, 3))
    uint64_t key = w & keep;
    if (c == 4) key |= ~keep & ~3ull;
    return key | 1ull;
}

__device__ __forceinline__ uint64_t first_key(const KeySpec& ks, uint64_t p);
// out-of-line copy for the rare irregular windows inside unrolled loops
__device__ __noinline__ uint64_t first_key_fast2_slow(const KeySpec& ks, uint64_t p) { return first_key_fast2(ks, p); }

// Number of symbols that exist in the key of suffix p.
__device__ __forceinline__ uint64_t key_len(const KeySpec& ks, uint64_t p) {
    uint64_t rest = ks.pt.n - p;
    if (ks.mode == kModeFull) return rest;
    if (ks.mode == kModeMaxQueryLen) return rest < ks.cap ? rest : ks.cap;
    if (rest >= ks.mask_len) return ks.weight;
    uint32_t c = 0;
    while (c < ks.weight && ks.mask_pos[c] < rest) c++;
    return c;
}

__device__ __forceinline__ uint32_t sym_lcp(const PackedText& pt, uint64_t x, uint64_t y) {
    if (x == y) return pt.K;
    return (uint32_t)__clzll((long long)(x ^ y)) / pt.bits;
}

// LCP (in key symbols) of two suffixes whose keys agree on the first `base` symbols and whose next
// key words are x and y.
__device__ __forceinline__ uint32_t lcp_from_words(const KeySpec& ks, uint64_t x, uint64_t y, uint64_t base,
                                                  uint64_t pa, uint64_t pb) {
    uint64_t l = base + sym_lcp(ks.pt, x, y);
    uint64_t la = key_len(ks, pa), lb = key_len(ks, pb);
    if (la < l) l = la;
    if (lb < l) l = lb;
    return (uint32_t)l;
}

__device__ __forceinline__ uint64_t first_key(const KeySpec& ks, uint64_t p) {
    return ks.fast2 ? first_key_fast2(ks, p) : key_word(ks, p, 0);
}

// If p lies in a recorded run of N, returns true and the run's end (find_n_run, sufr_builder.rs:241-254).
__device__ __forceinline__ bool n_run_end(const KeySpec& ks, uint64_t p, uint64_t& end) {
    uint32_t lo = 0, hi = ks.num_n_ranges;
    while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        uint64_t s = ks.n_starts[mid], e = ks.n_ends[mid];
        if (s <= p && p < e) { end = e; return true; }
        if (s < p) lo = mid + 1; else hi = mid;
    }
    return false;
}

// Exact LCP of suffixes pa != pb by word-wise comparison, starting from a known lower bound (full mode).
__device__ __forceinline__ uint64_t lcp_direct(const KeySpec& ks, uint64_t pa, uint64_t pb, uint64_t lower) {
    const uint64_t n = ks.pt.n;
    uint64_t limit = n - (pa > pb ? pa : pb);
    uint64_t l = lower;
    while (l < limit) {
        uint64_t x = load_key(ks.pt, pa + l), y = load_key(ks.pt, pb + l);
        if (x != y) {
            l += sym_lcp(ks.pt, x, y);
            break;
        }
        l += ks.pt.K;
    }
    return l < limit ? l : limit;
}

}  // namespace sufr
