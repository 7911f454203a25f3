// Build pipeline + C ABI of libsufr_b200.so.
//
// Replaces SufrBuilder::<T>::new up to write() (libsufr/src/sufr_builder.rs:143-217) on one B200:
//
//   encode   transform_kernel, pack_kernel            text transform, dense alphabet, packed text
//   keys     keygen_kernel (or shard compaction)      first key word of every suffix
//   sort     rsort::{upsweep,scan,downsweep}          stable LSD radix sort of (key word, position)
//   refine   resolve / active compaction / segmented  groups of equal key words: next key word, then
//            sort rounds, then prefix doubling        (full sort only) prefix doubling on ranks
//   lcp      resolve_kernel, lcp_complete_kernel      LCP from clz(x ^ y) of adjacent key words
//   finish   N-run tie rule, suffix filter, widen     sufr_builder.rs:305-307, :446-449
//
// The oracle (oracle/) is never linked or called from here.
#include <fcntl.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cerrno>
#include <chrono>
#include <condition_variable>
#include <cstdlib>
#include <deque>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/sufr_b200.h"
#include "common.cuh"
#include "host_io.hpp"
#include "kernels.cuh"
#include "keys.cuh"
#include "pool.hpp"
#include "radix_sort.cuh"
#include "scan.cuh"
#include "verify.cuh"

namespace sufr {

thread_local std::string g_last_error;

// Pinned host buffers are expensive to allocate (page-locking tens of GB takes seconds), so host results
// hand their buffers back to the context for the next build.
struct PinnedCache {
    std::vector<std::pair<void*, size_t>> free_list;
    void* get(size_t bytes) {
        if (bytes == 0) bytes = 1;
        size_t best = (size_t)-1;
        for (size_t i = 0; i < free_list.size(); i++)
            if (free_list[i].second >= bytes && (best == (size_t)-1 || free_list[i].second < free_list[best].second)) best = i;
        if (best != (size_t)-1 && free_list[best].second <= 2 * bytes + (1u << 20)) {
            void* p = free_list[best].first;
            sizes[p] = free_list[best].second;
            free_list.erase(free_list.begin() + best);
            return p;
        }
        void* p = nullptr;
        cudaError_t e = cudaMallocHost(&p, bytes);
        if (e != cudaSuccess) {
            cudaGetLastError();
            release();
            e = cudaMallocHost(&p, bytes);
            if (e != cudaSuccess) {
                cudaGetLastError();
                throw Error(SUFR_B200_ERR_OUT_OF_MEMORY, "cudaMallocHost(" + std::to_string(bytes) + " bytes) failed");
            }
        }
        sizes[p] = bytes;
        return p;
    }
    void put(void* p) {
        auto it = sizes.find(p);
        if (it == sizes.end()) { cudaFreeHost(p); return; }
        free_list.push_back({p, it->second});
        sizes.erase(it);
        while (free_list.size() > 6) {  // keep a couple of builds' worth
            cudaFreeHost(free_list.front().first);
            free_list.erase(free_list.begin());
        }
    }
    void release() {
        for (auto& f : free_list) cudaFreeHost(f.first);
        free_list.clear();
    }
    std::map<void*, size_t> sizes;  // buffers currently lent out
};

struct Ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    DevicePool pool;
    PinnedCache pinned;
    uint64_t launches = 0;
    std::mutex mu;
};

// Holds everything a returned result owns.
struct ResultOwner {
    Ctx* ctx = nullptr;          // device buffers go back to this pool
    bool owns_ctx = false;       // sufr_b200_create made a private context
    bool text_malloc = false;    // host text from malloc (sufr_b200_create), not from the pinned cache
    int memory = SUFR_B200_MEM_HOST;
    void* text = nullptr;
    void* sa = nullptr;
    void* lcp = nullptr;
    uint64_t* n_ranges = nullptr;
    // kept for sufr_b200_patch_seam on device results
};

struct EventTimer {
    cudaStream_t stream;
    std::vector<cudaEvent_t> ev;
    explicit EventTimer(cudaStream_t s) : stream(s) {}
    ~EventTimer() {
        for (auto e : ev) cudaEventDestroy(e);
    }
    int mark() {
        cudaEvent_t e;
        SUFR_CUDA_CHECK(cudaEventCreate(&e));
        SUFR_CUDA_CHECK(cudaEventRecord(e, stream));
        ev.push_back(e);
        return (int)ev.size() - 1;
    }
    double ms(int a, int b) {
        float t = 0;
        SUFR_CUDA_CHECK(cudaEventElapsedTime(&t, ev[a], ev[b]));
        return t;
    }
};

// Thrown when prefix doubling is needed but this attempt sorts only a subset of the positions (suffix
// filter applied up front, or one key-range shard): the build is redone over all positions.
struct NeedFullSort {};

// Host side of the compact device->host transfer: widen a staged array into the result with a few threads.
// The destination is written once and not read back, so the AVX2 path uses streaming (non-temporal) stores:
// no read-for-ownership traffic on the host memory bus, which the incoming DMA writes share.
}  // namespace sufr
#include <immintrin.h>
namespace sufr {

template <typename Src, typename Dst>
static void widen_range_scalar(const Src* src, Dst* dst, uint64_t lo, uint64_t hi) {
    for (uint64_t i = lo; i < hi; i++) dst[i] = (Dst)src[i];
}
__attribute__((target("avx2"))) static void widen_range_avx2(const uint8_t* src, uint64_t* dst, uint64_t lo, uint64_t hi) {
    uint64_t i = lo;
    for (; i < hi && ((uintptr_t)(dst + i) & 31); i++) dst[i] = src[i];
    for (; i + 16 <= hi; i += 16) {
        __m128i b = _mm_loadu_si128((const __m128i*)(src + i));
        _mm256_stream_si256((__m256i*)(dst + i), _mm256_cvtepu8_epi64(b));
        _mm256_stream_si256((__m256i*)(dst + i + 4), _mm256_cvtepu8_epi64(_mm_srli_si128(b, 4)));
        _mm256_stream_si256((__m256i*)(dst + i + 8), _mm256_cvtepu8_epi64(_mm_srli_si128(b, 8)));
        _mm256_stream_si256((__m256i*)(dst + i + 12), _mm256_cvtepu8_epi64(_mm_srli_si128(b, 12)));
    }
    for (; i < hi; i++) dst[i] = src[i];
    _mm_sfence();
}
__attribute__((target("avx2"))) static void widen_range_avx2(const uint8_t* src, uint32_t* dst, uint64_t lo, uint64_t hi) {
    uint64_t i = lo;
    for (; i < hi && ((uintptr_t)(dst + i) & 31); i++) dst[i] = src[i];
    for (; i + 16 <= hi; i += 16) {
        __m128i b = _mm_loadu_si128((const __m128i*)(src + i));
        _mm256_stream_si256((__m256i*)(dst + i), _mm256_cvtepu8_epi32(b));
        _mm256_stream_si256((__m256i*)(dst + i + 8), _mm256_cvtepu8_epi32(_mm_srli_si128(b, 8)));
    }
    for (; i < hi; i++) dst[i] = src[i];
    _mm_sfence();
}
__attribute__((target("avx2"))) static void widen_range_avx2(const uint32_t* src, uint64_t* dst, uint64_t lo, uint64_t hi) {
    uint64_t i = lo;
    for (; i < hi && ((uintptr_t)(dst + i) & 31); i++) dst[i] = src[i];
    for (; i + 8 <= hi; i += 8) {
        __m256i a = _mm256_loadu_si256((const __m256i*)(src + i));
        _mm256_stream_si256((__m256i*)(dst + i), _mm256_cvtepu32_epi64(_mm256_castsi256_si128(a)));
        _mm256_stream_si256((__m256i*)(dst + i + 4), _mm256_cvtepu32_epi64(_mm256_extracti128_si256(a, 1)));
    }
    for (; i < hi; i++) dst[i] = src[i];
    _mm_sfence();
}

template <typename Src, typename Dst>
static void host_widen(const Src* src, Dst* dst, uint64_t count, int threads) {
    if (threads < 1) threads = 1;
    static const bool avx2 = __builtin_cpu_supports("avx2");
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; t++) {
        pool.emplace_back([=]() {
            uint64_t lo = count * (uint64_t)t / threads, hi = count * (uint64_t)(t + 1) / threads;
            if (avx2) widen_range_avx2(src, dst, lo, hi);
            else widen_range_scalar(src, dst, lo, hi);
        });
    }
    for (auto& th : pool) th.join();
}

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
    DevBuf<uint32_t> d_sa, d_lcp;
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
    DevBuf<uint32_t> pos_spare_;
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
    void doubling(DevBuf<uint32_t>& slot, DevBuf<uint32_t>& pos, DevBuf<uint32_t>& seg, uint64_t m, uint64_t nseg,
                  uint64_t h);
    void n_run_rule();
    void apply_filter();
    void segmented_sort_u64key(DevBuf<uint64_t>& ck, DevBuf<uint32_t>& pos, uint64_t m, int key_bits);
    void sort_groups(DevBuf<uint64_t>& keys, DevBuf<uint32_t>& pos, const uint32_t* seg, const uint32_t* slot, uint64_t m,
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
    auto pos = dalloc<uint32_t>(count), pos_b = dalloc<uint32_t>(count);
    auto counts = dalloc<uint32_t>(rsort::counts_words());
    auto d_eq = dalloc<unsigned long long>(1);
    SUFR_CUDA_CHECK(cudaMemsetAsync(d_eq.get(), 0, 8, st()));
    sample_keys_kernel<<<grid_for(count, 2), kBlock, 0, st()>>>(ks, stride_pos, count, keys.get(), pos.get());
    SUFR_KERNEL_CHECK();
    launched();
    const int used = (int)(ks.pt.K * ks.pt.bits);
    const int begin_bit = ks.fast2 ? 64 - kFast2ProbeBits : 64 - used;
    bool in_b = rsort::sort_pairs<uint64_t, uint32_t>(keys.get(), keys_b.get(), pos.get(), pos_b.get(), count, begin_bit, 64,
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
        // balanced ranges, so it histograms every 16th position; the modes whose tie order depends on the
        // input order (mask / max-query-len) use the exact histogram, which also yields the exact shard
        // offsets.  The cut points are fixed by the FIRST histogram of a build: the full-sort fallback (which
        // only some ranks may take) re-counts exactly but keeps the same cuts.
        const uint32_t hbits = kShardHistBits, bins = 1u << hbits;
        const bool exact = ks.mode != kModeFull || cuts_ready_;
        const uint32_t sample_shift = exact ? 0 : 4;
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
    DevBuf<uint32_t> pos_a, pos_b;
    const int used_bits = (int)(ks.pt.K * ks.pt.bits);
    bool first_digit_done = false;  // fast path: the records come out of key generation sorted by the first digit
    uint64_t kept = n;     // suffixes that survive the filter (all ranks' ranges together)
    uint64_t sort_n = n;   // elements handed to the sort
    if (prefilter && n) kept = indexed_count_;
    // Few filtered suffixes (the common case: delimiters, sparse N): no compaction at all, they get the
    // key ~0 and drop off the end of the sorted array.  Needs an unused low bit in the packed word.
    const bool sentinel = prefilter && !sharded && used_bits < 64 && kept < n && (n - kept) * 16 <= n;
    sentinel_ = sentinel;
    if (sharded && ks.mode == kModeFull) {
        // unordered selection: one key computation per position, capacity from the sampled histogram
        uint64_t capacity = shard_estimate + shard_estimate / 16 + (1u << 20);
        auto d_cnt = dalloc<unsigned long long>(1);
        for (int attempt = 0;; attempt++) {
            keys_a = dalloc<uint64_t>(capacity);
            pos_a = dalloc<uint32_t>(capacity);
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
        pos_a = dalloc<uint32_t>(s);
        if (n) {
            uint32_t got = scan_total(n, in, scan::SumU32{}, SelectOut{ks, n, descending, keys_a.get(), pos_a.get()});
            if (got != s) throw Error(SUFR_B200_ERR_INTERNAL, "selection count mismatch");
        }
        sort_n = s;
    } else {
        s = sentinel ? kept : n;
        sort_n = n;
        keys_a = dalloc<uint64_t>(n);
        pos_a = dalloc<uint32_t>(n);
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
                                                                             d_counts.get());
                SUFR_KERNEL_CHECK();
                rsort::scan_counts_kernel<<<1, 1024, 0, st()>>>(d_counts.get(), (uint32_t)rsort::RADIX * used_grid);
                SUFR_KERNEL_CHECK();
                fast2_keygen_scatter_kernel<<<used_grid, kBlock, kKsSmem, st()>>>(ks, n, sentinel ? 1 : 0, keys_a.get(),
                                                                                pos_a.get(), chunk, shift, d_counts.get());
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
    pos_b = dalloc<uint32_t>(sort_n);
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
    bool in_b = rsort::sort_pairs<uint64_t, uint32_t>(keys_a.get(), keys_b.get(), pos_a.get(), pos_b.get(), sort_n,
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
void Build::segmented_sort_u64key(DevBuf<uint64_t>& ck, DevBuf<uint32_t>& pos, uint64_t m, int key_bits) {
    auto ck_b = dalloc<uint64_t>(m);
    auto pos_b = dalloc<uint32_t>(m);
    bool in_b = rsort::sort_pairs<uint64_t, uint32_t>(ck.get(), ck_b.get(), pos.get(), pos_b.get(), m, 0, key_bits,
                                                      d_counts.get(), st(), &ctx.launches);
    if (in_b) {
        ck = std::move(ck_b);
        pos = std::move(pos_b);
    }
}

// Sorts every unresolved group by key bits [key_lo, key_hi) of its members' keys (keys, positions, SA slots).
// rank_keys: prefix-doubling keys (group << 32 | rank); else key words.
void Build::sort_groups(DevBuf<uint64_t>& keys, DevBuf<uint32_t>& pos, const uint32_t* seg, const uint32_t* slot, uint64_t m,
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
        auto lpos = dalloc<uint32_t>(ml), lpos_b = dalloc<uint32_t>(ml);
        scan_finish(m, lin, scan::SumU64{}, LargeOutRank{pos.get(), keys.get(), idx.get(), lck.get(), lpos.get(), key_hi}, part);
        // (compact group << rank bits | rank): one sort over rank and group bits together
        bool in_b = rsort::sort_pairs<uint64_t, uint32_t>(lck.get(), lck_b.get(), lpos.get(), lpos_b.get(), ml, key_lo,
                                                          key_hi + sb, d_counts.get(), st(), &ctx.launches);
        if (in_b) { std::swap(lck, lck_b); std::swap(lpos, lpos_b); }
        scatter_large_rank_kernel<<<grid_for(ml, 2), kBlock, 0, st()>>>(ml, lck.get(), lpos.get(), idx.get(), slot, keys.get(),
                                                                       pos.get(), d_sa.get(), key_hi);
    } else {
        auto lkeys = dalloc<uint64_t>(ml), keys_b = dalloc<uint64_t>(ml);
        auto segpos = dalloc<uint64_t>(ml), segpos_b = dalloc<uint64_t>(ml);
        scan_finish(m, lin, scan::SumU64{}, LargeOut{pos.get(), keys.get(), idx.get(), lkeys.get(), segpos.get()}, part);
        bool in_b = rsort::sort_pairs<uint64_t, uint64_t>(lkeys.get(), keys_b.get(), segpos.get(), segpos_b.get(), ml, key_lo,
                                                          key_hi, d_counts.get(), st(), &ctx.launches);
        if (in_b) { std::swap(lkeys, keys_b); std::swap(segpos, segpos_b); }
        if (sb) {
            in_b = rsort::sort_pairs<uint64_t, uint64_t>(segpos.get(), segpos_b.get(), lkeys.get(), keys_b.get(), ml, 32, 32 + sb,
                                                         d_counts.get(), st(), &ctx.launches);
            if (in_b) { std::swap(lkeys, keys_b); std::swap(segpos, segpos_b); }
        }
        scatter_large_kernel<<<grid_for(ml, 2), kBlock, 0, st()>>>(ml, lkeys.get(), segpos.get(), idx.get(), slot, keys.get(),
                                                                  pos.get(), d_sa.get());
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
    uint64_t m = 0, nseg = 0;
    DevBuf<uint32_t> slot, pos, seg;
    {
        uint64_t capacity = final_word ? 1 : std::max<uint64_t>(1u << 20, r0n / 8);
        if (const char* dbg = getenv("SUFR_B200_DEBUG_SPARSE_CAP")) capacity = std::max<uint64_t>(1, strtoull(dbg, nullptr, 10));
        auto act_slot = dalloc<uint32_t>(capacity);
        auto act_pos = dalloc<uint32_t>(capacity);
        auto d_cnt = dalloc<unsigned long long>(1);
        SUFR_CUDA_CHECK(cudaMemsetAsync(d_cnt.get(), 0, 8, st()));
        if (fast2) {
            drop_wide();
            if (index_bits_ == 64 && result_memory_ == SUFR_B200_MEM_DEVICE && !args.allow_ambiguity &&
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
            round0_fast2_kernel<<<grid_for(r0n, 4), kBlock, 0, st()>>>(keys_sorted.get(), d_sa.get(), pos_spare_.get(), r0n,
                                                                      d_lcp.get(), act_slot.get(), act_pos.get(), d_cnt.get(),
                                                                      capacity, wide_sa_.get(), wide_lcp_.get());
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
            auto pos_b = dalloc<uint32_t>(m);
            bool in_b = rsort::sort_pairs<uint32_t, uint32_t>(act_slot.get(), slot_b.get(), act_pos.get(), pos_b.get(), m, 0,
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
            if (fast2) nseg = scan_total(m, LcpSegIn{d_lcp.get(), slot.get()}, scan::SumU32{}, SparseSegOut{seg.get()});
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
                pos = dalloc<uint32_t>(m);
                seg = dalloc<uint32_t>(m);
                scan_finish(r0n, ain, scan::SumU64{}, LcpActiveOut{d_sa.get(), slot.get(), pos.get(), seg.get()}, part);
            } else {
                ActiveIn<ViewAll> ain{v0, r0n, 0, sentinel_ ? 1 : 0};
                unsigned long long tot = scan_begin(r0n, ain, scan::SumU64{}, part);
                m = (uint32_t)tot;
                nseg = tot >> 32;
                slot = dalloc<uint32_t>(m);
                pos = dalloc<uint32_t>(m);
                seg = dalloc<uint32_t>(m);
                scan_finish(r0n, ain, scan::SumU64{}, ActiveOut<ViewAll>{v0, slot.get(), pos.get(), seg.get()}, part);
            }
        }
    }
    keys_sorted.reset();
    unsigned long long tot = 0;
    // A large unresolved fraction after the first sort means a repetitive text: it will need prefix doubling,
    // which needs every position.  Give up this (filtered / sharded) attempt now rather than after the
    // patient word rounds.
    if (!full_set_ && ks.mode == kModeFull && m * 16 >= r0n) throw NeedFullSort{};

    // Full sort: after kMaxWordRounds words switch to prefix doubling (depth doubles per round).
    // When this attempt sorts only a subset of the positions (filter applied up front / one shard), doubling
    // means redoing the build over all positions, so shallow repeats (few unresolved elements) get more
    // word rounds first.
    const int kMaxWordRounds = 3, kPatientWordRounds = 48;
    while (m > 0) {
        if (ks.mode == kModeFull && word >= kMaxWordRounds) {
            bool patient = !full_set_ && word < kPatientWordRounds && m * 16 < s;
            if (!patient) {
                doubling(slot, pos, seg, m, nseg, (uint64_t)(word + 1) * K);
                return;
            }
        }
        word++;
        refine_rounds++;
        if (getenv("SUFR_B200_LOG_ROUNDS")) {
            SUFR_CUDA_CHECK(cudaStreamSynchronize(st()));
            fprintf(stderr, "[sufr_b200] word round %u (key word %d): unresolved = %llu in %llu groups, t = %.3f s\n",
                    refine_rounds, word, (unsigned long long)m, (unsigned long long)nseg,
                    std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count());
        }
        final_word = ((uint64_t)(word + 1) * K >= ks.cap) ? 1 : 0;
        auto keys = dalloc<uint64_t>(m);
        round_keys_kernel<<<grid_for(m, 2), kBlock, 0, st()>>>(ks, m, (uint32_t)word, sentinel_ ? 1 : 0, pos.get(), keys.get());
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
        auto pos2 = dalloc<uint32_t>(m2);
        auto seg2 = dalloc<uint32_t>(m2);
        scan_finish(m, ain, scan::SumU64{}, ActiveOut<ViewActive>{va, slot2.get(), pos2.get(), seg2.get()}, part);
        slot = std::move(slot2);
        pos = std::move(pos2);
        seg = std::move(seg2);
        m = m2;
        nseg = nseg2;
    }
}

// Prefix doubling on the still-unresolved groups (Larsson-Sadakane style: only active groups are sorted).
// Needs the rank of EVERY text position, hence only valid when all positions were sorted on this rank.
void Build::doubling(DevBuf<uint32_t>& slot, DevBuf<uint32_t>& pos, DevBuf<uint32_t>& seg, uint64_t m, uint64_t nseg,
                     uint64_t h) {
    if (!full_set_) throw NeedFullSort{};
    d_isa = dalloc<uint32_t>(n);
    uint32_t* const isa_ptr = d_isa.get();
    isa_init_kernel<<<grid_for(n, 4), kBlock, 0, st()>>>(n, d_sa.get(), isa_ptr);
    SUFR_KERNEL_CHECK();
    launched();
    scan_total(m, GroupStartIn{seg.get()}, scan::MaxU32{}, GroupRankOut{slot.get(), pos.get(), isa_ptr});
    const bool log_rounds = getenv("SUFR_B200_LOG_ROUNDS") != nullptr;
    while (m > 0) {
        if (doubling_rounds > 64) throw Error(SUFR_B200_ERR_INTERNAL, "prefix doubling did not converge");
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
        auto pos2 = dalloc<uint32_t>(m2);
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
    auto pos = dalloc<uint32_t>(m);
    auto ck = dalloc<uint64_t>(m);
    scan_total(s, in, scan::SumU64{}, NTieOut{d_sa.get(), slot.get(), pos.get(), ck.get()});
    segmented_sort_u64key(ck, pos, m, 32 + (nseg > 1 ? bits_for(nseg - 1) : 0));
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
        filter_flags_kernel<<<nblocks, kBlock, 0, st()>>>(d_text.get(), d_sa.get(), d_lcp.get(), s, flags.get(),
                                                         counts.get(), tail_min.get());
        SUFR_KERNEL_CHECK();
        launched();
        uint32_t kept = scan_total(nblocks, BlockCountIn{counts.get()}, scan::SumU32{}, BlockOffsetOut{offsets.get()});
        if (kept == s) return;
        auto sa2 = dalloc<uint32_t>(kept);
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
    auto sa2 = dalloc<uint32_t>(kept);
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
    if (n >= 0xFFFFFFFFull)
        throw Error(SUFR_B200_ERR_UNSUPPORTED,
                    "text_len >= 2^32 - 1 needs 64-bit positions inside the sort; this build supports u64 OUTPUT for "
                    "texts below that length only");
    if (index_bits_ == 0) index_bits_ = 32;  // suffix_array.rs:461 (text_len < u32::MAX here)
    if (index_bits_ != 32 && index_bits_ != 64) throw Error(SUFR_B200_ERR_ARGUMENT, "index_bits must be 0, 32 or 64");
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
        auto sa2 = dalloc<uint32_t>(shard_count);
        auto lcp2 = dalloc<uint32_t>(shard_count);
        if (shard_count) {
            SUFR_CUDA_CHECK(cudaMemcpyAsync(sa2.get(), d_sa.get() + shard_offset, shard_count * 4, cudaMemcpyDeviceToDevice, st()));
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
        uint32_t fl[2];
        SUFR_CUDA_CHECK(cudaMemcpyAsync(&fl[0], d_sa.get(), 4, cudaMemcpyDeviceToHost, st()));
        SUFR_CUDA_CHECK(cudaMemcpyAsync(&fl[1], d_sa.get() + (s - 1), 4, cudaMemcpyDeviceToHost, st()));
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
    bool compact = result_memory_ == SUFR_B200_MEM_HOST && s >= compact_min && s > 0 &&
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
    if (index_bits_ == 64 && !compact && wide_ok_) {
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
            widen2_kernel<<<grid_for(s, 8), kBlock, 0, st()>>>(s, d_sa.get(), d_lcp.get(), sa64.get(), lcp64.get());
            SUFR_KERNEL_CHECK();
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
        if (index_bits_ == 64) {
            owner->sa = sa64.release();
            owner->lcp = lcp64.release();
        } else {
            owner->sa = d_sa.release();
            owner->lcp = d_lcp.release();
        }
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
    out->owner = owner.release();
}

static void free_owner(ResultOwner* o) {
    if (!o) return;
    if (o->memory == SUFR_B200_MEM_DEVICE) {
        if (o->ctx) {
            std::lock_guard<std::mutex> lock(o->ctx->mu);
            if (o->text) o->ctx->pool.free(o->text);
            if (o->sa) o->ctx->pool.free(o->sa);
            if (o->lcp) o->ctx->pool.free(o->lcp);
        }
    } else if (o->ctx) {
        std::lock_guard<std::mutex> lock(o->ctx->mu);
        if (o->text) o->ctx->pinned.put(o->text);
        if (o->sa) o->ctx->pinned.put(o->sa);
        if (o->lcp) o->ctx->pinned.put(o->lcp);
    } else {
        if (o->text && o->text_malloc) free(o->text);
        else if (o->text) cudaFreeHost(o->text);
        if (o->sa) cudaFreeHost(o->sa);
        if (o->lcp) cudaFreeHost(o->lcp);
    }
    free(o->n_ranges);
    if (o->owns_ctx && o->ctx) {
        cudaStreamDestroy(o->ctx->stream);
        delete o->ctx;
    }
    delete o;
}

template <typename F>
static int guarded(F&& f) {
    try {
        f();
        return SUFR_B200_OK;
    } catch (const Error& e) {
        g_last_error = e.what();
        return e.code;
    } catch (const std::bad_alloc&) {
        g_last_error = "host allocation failed";
        return SUFR_B200_ERR_OUT_OF_MEMORY;
    } catch (const std::exception& e) {
        g_last_error = e.what();
        return SUFR_B200_ERR_INTERNAL;
    }
}

// sufr_b200_create: the sections of the file go from device memory to the file through a small ring of pinned
// buffers -- the device->host copy of one chunk overlaps the pwrite of the previous ones, and no host copy of the
// suffix / LCP arrays is ever allocated (page-locking 2 * s * sizeof(T) bytes costs seconds by itself).
class StreamWriter {
   public:
    StreamWriter(int fd, const std::string& path, int device, cudaStream_t stream, size_t slot_bytes = 64u << 20, int nslots = 6,
                 int nworkers = 4)
        : fd_(fd), path_(path), device_(device), stream_(stream), slot_bytes_(slot_bytes) {
        try {
            for (int i = 0; i < nslots; i++) {
                Slot sl{nullptr, nullptr};
                SUFR_CUDA_CHECK(cudaMallocHost(&sl.buf, slot_bytes_));
                slots_.push_back(sl);
                SUFR_CUDA_CHECK(cudaEventCreateWithFlags(&slots_.back().ev, cudaEventDisableTiming));
                free_.push_back(i);
            }
        } catch (...) {  // a constructor that throws runs no destructor: release what was allocated
            release_slots();
            throw;
        }
        for (int i = 0; i < nworkers; i++) workers_.emplace_back([this] { work(); });
    }
    ~StreamWriter() {
        try { finish(); } catch (...) {}
        release_slots();
    }
    // Queue `len` bytes of device memory for file offset `off`; `host_copy` (optional) also receives them.
    void submit(const void* dev, size_t len, uint64_t off, uint8_t* host_copy = nullptr) {
        const char* p = (const char*)dev;
        while (len) {
            const size_t chunk = std::min(len, slot_bytes_);
            int slot;
            {
                std::unique_lock<std::mutex> lock(mu_);
                cv_free_.wait(lock, [&] { return !free_.empty() || err_; });
                if (err_) std::rethrow_exception(err_);
                slot = free_.front();
                free_.pop_front();
            }
            SUFR_CUDA_CHECK(cudaMemcpyAsync(slots_[slot].buf, p, chunk, cudaMemcpyDeviceToHost, stream_));
            SUFR_CUDA_CHECK(cudaEventRecord(slots_[slot].ev, stream_));
            {
                std::lock_guard<std::mutex> lock(mu_);
                jobs_.push_back({slot, chunk, off, host_copy});
            }
            cv_work_.notify_one();
            p += chunk;
            off += chunk;
            if (host_copy) host_copy += chunk;
            len -= chunk;
            bytes_ += chunk;
        }
    }
    void finish() {
        {
            std::lock_guard<std::mutex> lock(mu_);
            if (done_) return;
            done_ = true;
        }
        cv_work_.notify_all();
        for (auto& t : workers_) t.join();
        workers_.clear();
        if (err_) std::rethrow_exception(err_);
    }
    uint64_t bytes() const { return bytes_; }

   private:
    struct Slot { void* buf; cudaEvent_t ev; };
    struct Job { int slot; size_t len; uint64_t off; uint8_t* host_copy; };
    void release_slots() {
        for (auto& sl : slots_) {
            if (sl.buf) cudaFreeHost(sl.buf);
            if (sl.ev) cudaEventDestroy(sl.ev);
        }
        slots_.clear();
    }
    void work() {
        cudaSetDevice(device_);
        for (;;) {
            Job j;
            {
                std::unique_lock<std::mutex> lock(mu_);
                cv_work_.wait(lock, [&] { return !jobs_.empty() || done_; });
                if (jobs_.empty()) return;
                j = jobs_.front();
                jobs_.pop_front();
            }
            try {
                cudaError_t e = cudaEventSynchronize(slots_[j.slot].ev);
                if (e != cudaSuccess) throw Error(100 + (int)e, std::string("CUDA error in the file writer: ") + cudaGetErrorString(e));
                if (j.host_copy) memcpy(j.host_copy, slots_[j.slot].buf, j.len);
                pwrite_all(fd_, slots_[j.slot].buf, j.len, j.off, path_);
            } catch (...) {
                std::lock_guard<std::mutex> lock(mu_);
                if (!err_) err_ = std::current_exception();
            }
            {
                std::lock_guard<std::mutex> lock(mu_);
                free_.push_back(j.slot);
            }
            cv_free_.notify_one();
        }
    }
    int fd_;
    std::string path_;
    int device_;
    cudaStream_t stream_;
    size_t slot_bytes_;
    std::vector<Slot> slots_;
    std::deque<int> free_;
    std::deque<Job> jobs_;
    std::vector<std::thread> workers_;
    std::mutex mu_;
    std::condition_variable cv_free_, cv_work_;
    bool done_ = false;
    std::exception_ptr err_;
    uint64_t bytes_ = 0;
};

// Writes the file of a DEVICE result (sufr_builder.rs:817-918) and returns the transformed text in `host_text`
// (malloc; what SufrBuilder.text holds after `new`).
static void stream_result_to_file(Ctx& ctx, const SufrB200Args& args, const SufrB200Result& r, uint8_t** host_text) {
    const std::string path = args.path ? args.path : "out.sufr";  // sufr_builder.rs:215
    const size_t w = r.index_bits / 8;
    const SufrFrame f = make_sufr_frame(args, r.index_bits, r.text_len, r.total_suffixes);
    const bool sharded = args.world_size > 1;
    const bool leader = !sharded || args.rank == 0;
    int fd = open(path.c_str(), O_WRONLY | O_CREAT | (sharded ? 0 : O_TRUNC), 0644);
    if (fd < 0) throw Error(SUFR_B200_ERR_IO, path + ": " + strerror(errno));  // sufr_builder.rs:820
    uint8_t* text = nullptr;
    try {
        if (leader) {  // ranks > 0 of a sharded build neither write nor return the text
            text = (uint8_t*)malloc(std::max<uint64_t>(1, r.text_len));
            if (!text) throw Error(SUFR_B200_ERR_OUT_OF_MEMORY, "out of host memory for the transformed text");
        }
        {
            // page-cache / tmpfs writes are CPU bound (page allocation + copy, ~1 GB/s per thread): many writers on
            // many small slots, shared fairly between the ranks of a multi-GPU build
            const int hw = (int)std::thread::hardware_concurrency();
            const int share = std::max(1, (int)args.world_size);
            int workers = std::max(2, std::min(24, (hw - 2) / share));
            if (const char* dbg = getenv("SUFR_B200_WRITER_THREADS")) workers = std::max(1, atoi(dbg));
            StreamWriter sw(fd, path, ctx.device, ctx.stream, 16u << 20, workers + 4, workers);
            if (leader) {
                pwrite_all(fd, f.head.data(), f.head.size(), 0, path);
                pwrite_all(fd, f.tail.data(), f.tail.size(), f.names_pos, path);
                if (sharded && ftruncate(fd, (off_t)(f.names_pos + f.tail.size())) != 0)
                    throw Error(SUFR_B200_ERR_IO, path + ": " + strerror(errno));
            }
            if (leader) sw.submit(r.text, r.text_len, f.text_pos, text);
            sw.submit(r.sa, r.num_suffixes * w, f.sa_pos + r.shard_offset * w);
            sw.submit(r.lcp, r.num_suffixes * w, f.lcp_pos + r.shard_offset * w);
            sw.finish();
        }
    } catch (...) {
        close(fd);
        free(text);
        throw;
    }
    if (close(fd) != 0) { free(text); throw Error(SUFR_B200_ERR_IO, path + ": " + strerror(errno)); }
    if (host_text) *host_text = text; else free(text);
}

static Ctx* make_ctx(int device) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        throw Error(SUFR_B200_ERR_CUDA, "no CUDA device available: the sufr_b200 build path has no CPU fallback");
    }
    if (device < 0 || device >= count) throw Error(SUFR_B200_ERR_ARGUMENT, "invalid CUDA device ordinal");
    SUFR_CUDA_CHECK(cudaSetDevice(device));
    auto c = std::make_unique<Ctx>();
    c->device = device;
    SUFR_CUDA_CHECK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    return c.release();
}

// Rebuilds the KeySpec of a finished result far enough to compute one pair LCP (seam repair).
}  // namespace sufr

using namespace sufr;

extern "C" {

int sufr_b200_abi_version(void) { return SUFR_B200_ABI_VERSION; }

const char* sufr_b200_last_error(void) { return g_last_error.c_str(); }

int sufr_b200_device_count(void) {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return count;
}

int sufr_b200_ctx_create(int device, SufrB200Ctx** out) {
    return guarded([&] {
        if (!out) throw Error(SUFR_B200_ERR_ARGUMENT, "out is NULL");
        *out = reinterpret_cast<SufrB200Ctx*>(make_ctx(device));
    });
}

void sufr_b200_ctx_destroy(SufrB200Ctx* c) {
    if (!c) return;
    Ctx* ctx = reinterpret_cast<Ctx*>(c);
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    ctx->pool.release_all();
    ctx->pinned.release();
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int sufr_b200_ctx_reserve(SufrB200Ctx* c, uint64_t text_len, uint32_t index_bits) {
    return guarded([&] {
        Ctx* ctx = reinterpret_cast<Ctx*>(c);
        if (!ctx) throw Error(SUFR_B200_ERR_ARGUMENT, "ctx is NULL");
        std::lock_guard<std::mutex> lock(ctx->mu);
        SUFR_CUDA_CHECK(cudaSetDevice(ctx->device));
        uint64_t per = 32;
        (void)index_bits;
        ctx->pool.reserve((size_t)(per * text_len + (64ull << 20)));
    });
}

void sufr_b200_ctx_trim(SufrB200Ctx* c) {
    Ctx* ctx = reinterpret_cast<Ctx*>(c);
    if (!ctx) return;
    std::lock_guard<std::mutex> lock(ctx->mu);
    cudaSetDevice(ctx->device);
    ctx->pool.trim();
    ctx->pinned.release();
}

int sufr_b200_build(SufrB200Ctx* c, const SufrB200Args* args, uint32_t index_bits, int text_memory, int result_memory,
                    SufrB200Result* out) {
    return guarded([&] {
        Ctx* ctx = reinterpret_cast<Ctx*>(c);
        if (!ctx || !args || !out) throw Error(SUFR_B200_ERR_ARGUMENT, "NULL argument");
        if (args->text_len && !args->text) throw Error(SUFR_B200_ERR_ARGUMENT, "text is NULL");
        std::lock_guard<std::mutex> lock(ctx->mu);
        try {
            Build b(*ctx, *args, index_bits, text_memory, result_memory);
            b.run(out);
        } catch (...) {
            cudaStreamSynchronize(ctx->stream);
            throw;
        }
    });
}

void sufr_b200_result_free(SufrB200Ctx*, SufrB200Result* r) {
    if (!r) return;
    free_owner(static_cast<ResultOwner*>(r->owner));
    memset(r, 0, sizeof(*r));
}

int sufr_b200_patch_seam(SufrB200Ctx* c, const SufrB200Args* args, SufrB200Result* r, uint64_t prev_last_suffix) {
    return guarded([&] {
        Ctx* ctx = reinterpret_cast<Ctx*>(c);
        if (!ctx || !args || !r) throw Error(SUFR_B200_ERR_ARGUMENT, "NULL argument");
        if (r->num_suffixes == 0) return;
        // The seam needs the text only: recompute the pair LCP on the host from the transformed text
        // when the result is on the host, on the device otherwise.
        SeedMaskInfo mask;
        bool has_mask = args->seed_mask && parse_seed_mask(args->seed_mask, mask);
        uint64_t q = (!has_mask && args->has_max_query_len) ? args->max_query_len : 0;
        std::lock_guard<std::mutex> lock(ctx->mu);
        SUFR_CUDA_CHECK(cudaSetDevice(ctx->device));
        const uint64_t a = prev_last_suffix, b = r->first_suffix, n = r->text_len;
        if (a >= n || b >= n) throw Error(SUFR_B200_ERR_ARGUMENT, "suffix out of range");
        uint64_t l = 0;
        // Windows of the two suffixes in TRANSFORMED form, from wherever the text is at hand: the host result,
        // the device result, or (host results of ranks > 0 carry no text) the caller's raw text + the transform.
        auto fetch = [&](uint64_t p, uint64_t len, uint8_t* dst) {
            if (r->text && r->memory == SUFR_B200_MEM_HOST) {
                memcpy(dst, r->text + p, len);
            } else if (r->text) {
                SUFR_CUDA_CHECK(cudaMemcpyAsync(dst, r->text + p, len, cudaMemcpyDeviceToHost, ctx->stream));
                SUFR_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
            } else {
                // raw text of the caller (host or device memory, as it was passed to the build) + the transform
                if (!args->text) throw Error(SUFR_B200_ERR_ARGUMENT, "patch_seam needs the text (result or args)");
                cudaPointerAttributes attr{};
                bool on_device = cudaPointerGetAttributes(&attr, args->text) == cudaSuccess &&
                                 (attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged);
                cudaGetLastError();
                if (on_device) {
                    SUFR_CUDA_CHECK(cudaMemcpyAsync(dst, args->text + p, len, cudaMemcpyDeviceToHost, ctx->stream));
                    SUFR_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
                } else {
                    memcpy(dst, args->text + p, len);
                }
                for (uint64_t i = 0; i < len; i++) {  // sufr_builder.rs:149-156
                    uint8_t c = dst[i];
                    if (c >= 97 && c <= 122) c = args->ignore_softmask ? (uint8_t)'N' : (uint8_t)(c & 0x5F);
                    dst[i] = c;
                }
            }
        };
        auto in_run = [&](uint64_t p, uint64_t& end) {
            uint64_t lo = 0, hi = r->num_n_ranges;
            while (lo < hi) {
                uint64_t mid = (lo + hi) / 2;
                if (r->n_ranges[2 * mid] <= p && p < r->n_ranges[2 * mid + 1]) { end = r->n_ranges[2 * mid + 1]; return true; }
                if (r->n_ranges[2 * mid] < p) lo = mid + 1; else hi = mid;
            }
            return false;
        };
        uint64_t ea = 0, eb = 0;
        if (!has_mask && r->num_n_ranges && in_run(a, ea) && in_run(b, eb)) {
            l = std::min(ea - a, eb - b);  // sufr_builder.rs:305-307
        } else {
            // the two suffixes come from different key ranges, so they differ within the first few symbols;
            // the window grows if they do not
            uint64_t win = has_mask ? mask.bytes.size() + 64 : 4096;
            while (true) {
                uint64_t la = std::min(win, n - a), lb = std::min(win, n - b);
                std::vector<uint8_t> buf(la + lb);
                fetch(a, la, buf.data());
                fetch(b, lb, buf.data() + la);
                if (has_mask) {
                    l = 0;
                    for (uint64_t k = 0; k < mask.positions.size(); k++) {
                        uint64_t o = mask.positions[k];
                        if (a + o >= n || b + o >= n || buf[o] != buf[la + o]) break;
                        l++;
                    }
                    break;
                }
                uint64_t lim = std::min(la, lb);
                if (q && q < lim) lim = q;
                l = 0;
                while (l < lim && buf[l] == buf[la + l]) l++;
                bool hit_window = (l == lim) && !(q && lim == q) && lim < std::min(n - a, n - b);
                if (!hit_window) break;
                win *= 16;
            }
        }
        if (r->memory == SUFR_B200_MEM_DEVICE) {
            if (r->index_bits == 32) {
                uint32_t v = (uint32_t)l;
                SUFR_CUDA_CHECK(cudaMemcpyAsync(r->lcp, &v, 4, cudaMemcpyHostToDevice, ctx->stream));
            } else {
                SUFR_CUDA_CHECK(cudaMemcpyAsync(r->lcp, &l, 8, cudaMemcpyHostToDevice, ctx->stream));
            }
            SUFR_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
        } else {
            if (r->index_bits == 32) ((uint32_t*)r->lcp)[0] = (uint32_t)l;
            else ((uint64_t*)r->lcp)[0] = l;
        }
    });
}

int sufr_b200_verify(SufrB200Ctx* c, const SufrB200Args* args, const SufrB200Result* r, int has_prev,
                     uint64_t prev_last_suffix, SufrB200VerifyReport* out) {
    return guarded([&] {
        Ctx* ctx = reinterpret_cast<Ctx*>(c);
        if (!ctx || !args || !r || !out) throw Error(SUFR_B200_ERR_ARGUMENT, "NULL argument");
        if (r->memory != SUFR_B200_MEM_DEVICE || !r->text)
            throw Error(SUFR_B200_ERR_ARGUMENT, "sufr_b200_verify needs a device result (with its transformed text)");
        SeedMaskInfo mask;
        const bool has_mask = args->seed_mask && parse_seed_mask(args->seed_mask, mask);
        std::lock_guard<std::mutex> lock(ctx->mu);
        SUFR_CUDA_CHECK(cudaSetDevice(ctx->device));
        cudaStream_t st = ctx->stream;
        const uint64_t n = r->text_len, s = r->num_suffixes;
        verify::Params P{};
        P.text = r->text;
        P.n = n;
        P.sa = r->sa;
        P.lcp = r->lcp;
        P.s = s;
        P.wide = r->index_bits == 64;
        P.filter = args->is_dna && !args->allow_ambiguity;
        P.mode = has_mask ? 2 : (args->has_max_query_len && args->max_query_len > 0 ? 1 : 0);
        P.q = args->max_query_len;
        P.has_prev = has_prev ? 1 : 0;
        P.prev_last = prev_last_suffix;
        DevBuf<uint32_t> d_maskpos;
        if (has_mask) {
            std::vector<uint32_t> mp(mask.positions.begin(), mask.positions.end());
            d_maskpos = DevBuf<uint32_t>(ctx->pool, mp.size());
            SUFR_CUDA_CHECK(cudaMemcpyAsync(d_maskpos.get(), mp.data(), mp.size() * 4, cudaMemcpyHostToDevice, st));
            SUFR_CUDA_CHECK(cudaStreamSynchronize(st));
            P.mask_pos = d_maskpos.get();
            P.weight = (uint32_t)mask.weight;
        }
        DevBuf<uint64_t> d_ns, d_ne;
        if (!has_mask && r->num_n_ranges) {  // find_lcp consults the runs only outside mask mode (sufr_builder.rs:301-307)
            std::vector<uint64_t> hs(r->num_n_ranges), he(r->num_n_ranges);
            for (uint64_t i = 0; i < r->num_n_ranges; i++) { hs[i] = r->n_ranges[2 * i]; he[i] = r->n_ranges[2 * i + 1]; }
            d_ns = DevBuf<uint64_t>(ctx->pool, hs.size());
            d_ne = DevBuf<uint64_t>(ctx->pool, he.size());
            SUFR_CUDA_CHECK(cudaMemcpyAsync(d_ns.get(), hs.data(), hs.size() * 8, cudaMemcpyHostToDevice, st));
            SUFR_CUDA_CHECK(cudaMemcpyAsync(d_ne.get(), he.data(), he.size() * 8, cudaMemcpyHostToDevice, st));
            SUFR_CUDA_CHECK(cudaStreamSynchronize(st));
            P.n_starts = d_ns.get();
            P.n_ends = d_ne.get();
            P.num_n_ranges = (uint32_t)r->num_n_ranges;
        }
        DevBuf<uint32_t> bitmap(ctx->pool, n / 32 + 2);
        DevBuf<verify::Report> d_rep(ctx->pool, 1);
        DevBuf<unsigned long long> d_cnt(ctx->pool, 1);
        verify::Report init{};
        init.first_bad_rank = ~0ull;
        SUFR_CUDA_CHECK(cudaMemsetAsync(bitmap.get(), 0, (n / 32 + 2) * 4, st));
        SUFR_CUDA_CHECK(cudaMemsetAsync(d_cnt.get(), 0, 8, st));
        SUFR_CUDA_CHECK(cudaMemcpyAsync(d_rep.get(), &init, sizeof(init), cudaMemcpyHostToDevice, st));
        EventTimer timer(st);
        const int t0 = timer.mark();
        const uint32_t grid = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(1, div_up(s, 256)), (uint64_t)num_sms() * 16);
        DevBuf<uint32_t> defer_bits(ctx->pool, s / 32 + 2);
        SUFR_CUDA_CHECK(cudaMemsetAsync(defer_bits.get(), 0, (s / 32 + 2) * 4, st));
        int method = 0;
        if (s) {
            verify::positions_kernel<<<grid, 256, 0, st>>>(P, bitmap.get(), d_rep.get());
            SUFR_KERNEL_CHECK();
            verify::pairs_kernel<<<grid, 256, 0, st>>>(P, d_rep.get(), defer_bits.get(), 0);
            SUFR_KERNEL_CHECK();
            verify::Report mid{};
            SUFR_CUDA_CHECK(cudaMemcpyAsync(&mid, d_rep.get(), sizeof(mid), cudaMemcpyDeviceToHost, st));
            SUFR_CUDA_CHECK(cudaStreamSynchronize(st));
            if (mid.deferred) {
                // deep repeats.  Full sort with every position present in this array: linear-time proof through the
                // inverse suffix array; otherwise (filtered / sharded / capped keys) the deferred pairs directly.
                const bool linear = P.mode == 0 && s == n && !has_prev && mid.duplicates == 0 && mid.out_of_range == 0 &&
                                    n <= 0xFFFFFFFFull;
                if (linear) {
                    method = 2;
                    DevBuf<uint32_t> isa(ctx->pool, n + 1);
                    verify::isa_kernel<<<grid, 256, 0, st>>>(P, isa.get());
                    SUFR_KERNEL_CHECK();
                    verify::order_by_rank_kernel<<<grid, 256, 0, st>>>(P, isa.get(), defer_bits.get(), d_rep.get());
                    SUFR_KERNEL_CHECK();
                    const uint64_t chunks = div_up(n, verify::kKasaiChunk);
                    verify::kasai_kernel<<<(uint32_t)std::max<uint64_t>(1, div_up(chunks, 256)), 256, 0, st>>>(
                        P, isa.get(), defer_bits.get(), d_rep.get());
                    SUFR_KERNEL_CHECK();
                    SUFR_CUDA_CHECK(cudaStreamSynchronize(st));
                } else {
                    method = 1;
                    verify::pairs_kernel<<<grid, 256, 0, st>>>(P, d_rep.get(), defer_bits.get(), 1);
                    SUFR_KERNEL_CHECK();
                }
            }
        }
        if (P.filter && n) {
            verify::count_indexed_kernel<<<(uint32_t)std::min<uint64_t>(div_up(n, 256 * 16), (uint64_t)num_sms() * 16), 256, 0, st>>>(
                r->text, n, d_cnt.get());
            SUFR_KERNEL_CHECK();
        }
        const int t1 = timer.mark();
        verify::Report rep{};
        unsigned long long cnt = 0;
        SUFR_CUDA_CHECK(cudaMemcpyAsync(&rep, d_rep.get(), sizeof(rep), cudaMemcpyDeviceToHost, st));
        SUFR_CUDA_CHECK(cudaMemcpyAsync(&cnt, d_cnt.get(), 8, cudaMemcpyDeviceToHost, st));
        SUFR_CUDA_CHECK(cudaStreamSynchronize(st));
        memset(out, 0, sizeof(*out));
        out->pairs_checked = rep.pairs_checked;
        out->order_errors = rep.order_errors;
        out->lcp_errors = rep.lcp_errors;
        out->out_of_range = rep.out_of_range;
        out->not_indexed = rep.not_indexed;
        out->duplicates = rep.duplicates;
        out->first_bad_rank = rep.first_bad_rank;
        out->max_lcp = rep.max_lcp;
        out->lcp_sum = rep.lcp_sum;
        out->expected_suffixes = P.filter ? cnt : n;
        out->deferred_pairs = rep.deferred;
        out->method = (uint32_t)method;
        out->ms = timer.ms(t0, t1);
    });
}

int sufr_b200_write(const SufrB200Args* args, const SufrB200Result* r) {
    return guarded([&] {
        if (!args || !r) throw Error(SUFR_B200_ERR_ARGUMENT, "NULL argument");
        if (r->memory != SUFR_B200_MEM_HOST) throw Error(SUFR_B200_ERR_ARGUMENT, "sufr_b200_write needs a host result");
        write_sufr_file(*args, *r);
    });
}

int sufr_b200_create(const SufrB200Args* args, int device, SufrB200Result* out) {
    SufrB200Result local;
    SufrB200Result* res = out ? out : &local;
    int rc = guarded([&] {
        if (!args) throw Error(SUFR_B200_ERR_ARGUMENT, "args is NULL");
        if (args->world_size > 1)
            throw Error(SUFR_B200_ERR_ARGUMENT,
                        "sufr_b200_create is the single-process path; a sharded build exchanges (count, first, last) "
                        "between sufr_b200_build and sufr_b200_write");
        std::unique_ptr<Ctx> ctx(make_ctx(device));
        SufrB200Result dev;
        memset(&dev, 0, sizeof(dev));
        auto teardown = [&] {
            if (dev.owner) free_owner(static_cast<ResultOwner*>(dev.owner));
            dev.owner = nullptr;
            cudaStreamSynchronize(ctx->stream);
            ctx->pinned.sizes.clear();
            ctx->pinned.release();
            ctx->pool.release_all();
            cudaStreamDestroy(ctx->stream);
        };
        try {
            {
                std::lock_guard<std::mutex> lock(ctx->mu);
                Build b(*ctx, *args, 0, SUFR_B200_MEM_HOST, SUFR_B200_MEM_DEVICE);
                b.run(&dev);
            }
            // the suffix and LCP arrays go straight from device memory into the file
            uint8_t* host_text = nullptr;
            const auto w0 = std::chrono::steady_clock::now();
            stream_result_to_file(*ctx, *args, dev, &host_text);
            const double write_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - w0).count();
            auto owner = std::make_unique<ResultOwner>();
            owner->memory = SUFR_B200_MEM_HOST;
            owner->text = host_text;
            owner->text_malloc = true;
            if (dev.num_n_ranges) {
                owner->n_ranges = (uint64_t*)malloc(dev.num_n_ranges * 16);
                memcpy(owner->n_ranges, dev.n_ranges, dev.num_n_ranges * 16);
            }
            *res = dev;
            res->memory = SUFR_B200_MEM_HOST;
            res->text = host_text;
            res->sa = nullptr;   // in the file, like the reference's builder struct (sufr_builder.rs:38-89)
            res->lcp = nullptr;
            res->n_ranges = owner->n_ranges;
            res->timings.d2h_ms = write_ms;  // device -> host -> file, overlapped
            res->d2h_bytes = (host_text ? dev.text_len : 0) + 2 * dev.num_suffixes * (dev.index_bits / 8);
            res->owner = owner.release();
        } catch (...) {
            teardown();
            throw;
        }
        teardown();
    });
    if (rc == SUFR_B200_OK && !out) sufr_b200_result_free(nullptr, &local);
    return rc;
}

// One call, several GPUs of one box: a host thread per device builds key-range shard r of `num_devices`, the raw text
// is uploaded once and replicated over NVLink (peer copies from the first device), the threads exchange
// (count, first, last) in host memory, every shard repairs its seam LCP and streams its slice of the suffix / LCP
// arrays into the file at its offset (sufr_builder.rs:817-918 writes the same sections serially).
int sufr_b200_create_multi(const SufrB200Args* args, const int* devices, int num_devices, uint32_t index_bits,
                           SufrB200Result* out) {
    SufrB200Result local;
    SufrB200Result* res = out ? out : &local;
    int rc = guarded([&] {
        if (!args || !devices) throw Error(SUFR_B200_ERR_ARGUMENT, "NULL argument");
        if (num_devices < 1 || num_devices > 64) throw Error(SUFR_B200_ERR_ARGUMENT, "num_devices must be 1..64");
        if (args->world_size > 1)
            throw Error(SUFR_B200_ERR_ARGUMENT, "sufr_b200_create_multi shards the build itself: pass world_size <= 1");
        if (args->text_len && !args->text) throw Error(SUFR_B200_ERR_ARGUMENT, "text is NULL");
        for (int i = 0; i < num_devices; i++)
            for (int k = 0; k < i; k++)
                if (devices[i] == devices[k]) throw Error(SUFR_B200_ERR_ARGUMENT, "a device is listed twice");
        const int W = num_devices;
        const uint64_t n = args->text_len;

        struct Shard {
            std::unique_ptr<Ctx> ctx;
            SufrB200Args a;
            SufrB200Result dev;
            uint8_t* host_text = nullptr;
            double write_ms = 0, h2d_ms = 0;
            std::exception_ptr err;
        };
        std::vector<Shard> sh(W);
        for (auto& x : sh) memset(&x.dev, 0, sizeof(x.dev));
        std::atomic<bool> failed{false};
        struct Barrier {
            std::mutex mu;
            std::condition_variable cv;
            int waiting = 0, generation = 0, parties;
            explicit Barrier(int p) : parties(p) {}
            void wait() {
                std::unique_lock<std::mutex> lock(mu);
                const int gen = generation;
                if (++waiting == parties) { waiting = 0; generation++; cv.notify_all(); }
                else cv.wait(lock, [&] { return gen != generation; });
            }
        } bar(W);
        uint8_t* raw0 = nullptr;  // raw text on the first device

        auto worker = [&](int r) {
            Shard& me = sh[r];
            auto step = [&](auto&& f) {
                if (failed.load()) return;
                try { f(); } catch (...) { me.err = std::current_exception(); failed.store(true); }
            };
            DevBuf<uint8_t> raw;
            step([&] {
                me.ctx.reset(make_ctx(devices[r]));
                me.a = *args;
                me.a.rank = r;
                me.a.world_size = W;
                raw = DevBuf<uint8_t>(me.ctx->pool, n + 16);
                if (r == 0) {
                    const auto t0 = std::chrono::steady_clock::now();
                    if (n) SUFR_CUDA_CHECK(cudaMemcpyAsync(raw.get(), args->text, n, cudaMemcpyHostToDevice, me.ctx->stream));
                    SUFR_CUDA_CHECK(cudaStreamSynchronize(me.ctx->stream));
                    me.h2d_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
                    raw0 = raw.get();
                } else {
                    cudaDeviceEnablePeerAccess(devices[0], 0);  // direct NVLink copies when the pair allows it
                    cudaGetLastError();
                }
            });
            bar.wait();
            step([&] {
                if (r > 0 && n) {
                    SUFR_CUDA_CHECK(cudaMemcpyPeerAsync(raw.get(), devices[r], raw0, devices[0], n, me.ctx->stream));
                    SUFR_CUDA_CHECK(cudaStreamSynchronize(me.ctx->stream));
                }
            });
            bar.wait();  // the first device keeps its copy alive until every peer has read it
            step([&] {
                me.a.text = raw.get();
                std::lock_guard<std::mutex> lock(me.ctx->mu);
                Build b(*me.ctx, me.a, index_bits, SUFR_B200_MEM_DEVICE, SUFR_B200_MEM_DEVICE);
                b.run(&me.dev);
            });
            raw.reset();
            bar.wait();
            step([&] {
                // (count, first, last) of every shard are final: offsets, total, previous non-empty shard
                uint64_t off = 0, total = 0;
                bool have_prev = false;
                uint64_t prev_last = 0;
                for (int k = 0; k < W; k++) {
                    if (k < r) {
                        off += sh[k].dev.num_suffixes;
                        if (sh[k].dev.num_suffixes) { have_prev = true; prev_last = sh[k].dev.last_suffix; }
                    }
                    total += sh[k].dev.num_suffixes;
                }
                me.dev.shard_offset = off;
                me.dev.total_suffixes = total;
                if (have_prev && me.dev.num_suffixes) {
                    if (sufr_b200_patch_seam(reinterpret_cast<SufrB200Ctx*>(me.ctx.get()), &me.a, &me.dev, prev_last) != SUFR_B200_OK)
                        throw Error(SUFR_B200_ERR_INTERNAL, "seam repair failed: " + g_last_error);
                }
                me.a.text = args->text;  // host text again (unused by the writer)
                const auto w0 = std::chrono::steady_clock::now();
                stream_result_to_file(*me.ctx, me.a, me.dev, r == 0 ? &me.host_text : nullptr);
                me.write_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - w0).count();
            });
        };
        std::vector<std::thread> threads;
        for (int r = 1; r < W; r++) threads.emplace_back(worker, r);
        worker(0);
        for (auto& t : threads) t.join();

        // result of the whole build (what SufrBuilder holds after `new`): counts, transformed text, n_ranges
        auto owner = std::make_unique<ResultOwner>();
        std::exception_ptr first_err;
        for (auto& x : sh)
            if (x.err && !first_err) first_err = x.err;
        if (!first_err) {
            owner->memory = SUFR_B200_MEM_HOST;
            owner->text = sh[0].host_text;
            owner->text_malloc = true;
            sh[0].host_text = nullptr;
            if (sh[0].dev.num_n_ranges) {
                owner->n_ranges = (uint64_t*)malloc(sh[0].dev.num_n_ranges * 16);
                memcpy(owner->n_ranges, sh[0].dev.n_ranges, sh[0].dev.num_n_ranges * 16);
            }
            *res = sh[0].dev;
            res->memory = SUFR_B200_MEM_HOST;
            res->text = (uint8_t*)owner->text;
            res->sa = nullptr;
            res->lcp = nullptr;
            res->n_ranges = owner->n_ranges;
            res->num_suffixes = sh[0].dev.total_suffixes;
            res->shard_offset = 0;
            res->timings.h2d_ms = sh[0].h2d_ms;
            res->kernel_launches = 0;
            res->peak_device_bytes = 0;
            res->d2h_bytes = n;
            res->h2d_bytes = n;
            for (auto& x : sh) {
                res->timings.total_ms = std::max(res->timings.total_ms, x.dev.timings.total_ms);
                res->timings.d2h_ms = std::max(res->timings.d2h_ms, x.write_ms);
                res->kernel_launches += x.dev.kernel_launches;
                res->peak_device_bytes = std::max(res->peak_device_bytes, x.dev.peak_device_bytes);
                res->d2h_bytes += 2 * x.dev.num_suffixes * (x.dev.index_bits / 8);
                res->refine_rounds = std::max(res->refine_rounds, x.dev.refine_rounds);
                res->doubling_rounds = std::max(res->doubling_rounds, x.dev.doubling_rounds);
            }
        }
        for (auto& x : sh) {  // tear the per-device state down on its device
            if (!x.ctx) continue;
            cudaSetDevice(x.ctx->device);
            if (x.dev.owner) free_owner(static_cast<ResultOwner*>(x.dev.owner));
            x.dev.owner = nullptr;
            cudaStreamSynchronize(x.ctx->stream);
            x.ctx->pinned.sizes.clear();
            x.ctx->pinned.release();
            x.ctx->pool.release_all();
            cudaStreamDestroy(x.ctx->stream);
            free(x.host_text);
        }
        if (first_err) std::rethrow_exception(first_err);
        res->owner = owner.release();
    });
    if (rc == SUFR_B200_OK && !out) sufr_b200_result_free(nullptr, &local);
    return rc;
}

int64_t sufr_b200_seed_mask(const char* mask, uint8_t* bytes, uint64_t* positions, uint64_t* differences) {
    SeedMaskInfo m;
    if (!mask || !parse_seed_mask(mask, m)) return -1;
    for (size_t i = 0; i < m.bytes.size(); i++)
        if (bytes) bytes[i] = m.bytes[i];
    for (size_t i = 0; i < m.positions.size(); i++) {
        if (positions) positions[i] = m.positions[i];
        if (differences) differences[i] = m.positions[i] - i;
    }
    return (int64_t)m.weight;
}

uint64_t sufr_b200_find_lcp_full_offset(uint64_t lcp, const char* mask) {
    if (!mask) return lcp;
    SeedMaskInfo m;
    if (!parse_seed_mask(mask, m)) return (uint64_t)-1;
    return lcp_full_offset(lcp, m);
}

int sufr_b200_read_sequence_file(const char* path, uint8_t delim, SufrB200Sequences* out) {
    return guarded([&] {
        if (!path || !out) throw Error(SUFR_B200_ERR_ARGUMENT, "NULL argument");
        read_sequence_file(path, delim, out);
    });
}

void sufr_b200_sequences_free(SufrB200Sequences* s) { free_sequences(s); }

int sufr_b200_synth_dna(SufrB200Ctx* c, uint8_t* device_text, uint64_t text_len, uint64_t seed,
                        const uint64_t* record_starts, uint64_t num_records, uint8_t delimiter) {
    return guarded([&] {
        Ctx* ctx = reinterpret_cast<Ctx*>(c);
        if (!ctx || !device_text) throw Error(SUFR_B200_ERR_ARGUMENT, "NULL argument");
        std::lock_guard<std::mutex> lock(ctx->mu);
        SUFR_CUDA_CHECK(cudaSetDevice(ctx->device));
        if (text_len == 0) return;
        synth_dna_kernel<<<grid_for(text_len, 4), kBlock, 0, ctx->stream>>>(device_text, text_len, seed);
        SUFR_KERNEL_CHECK();
        DevBuf<uint64_t> d_starts(ctx->pool, num_records ? num_records : 1);
        if (num_records)
            SUFR_CUDA_CHECK(cudaMemcpyAsync(d_starts.get(), record_starts, num_records * 8, cudaMemcpyHostToDevice,
                                            ctx->stream));
        uint64_t threads = num_records ? num_records : 1;
        synth_marks_kernel<<<(unsigned)div_up(threads, 256), 256, 0, ctx->stream>>>(device_text, text_len, d_starts.get(),
                                                                                   num_records, delimiter);
        SUFR_KERNEL_CHECK();
        SUFR_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    });
}

}  // extern "C"
