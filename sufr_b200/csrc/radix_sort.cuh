// Hand-written stable LSD radix sort of (key, value) pairs, 8-bit digits.
//
// Per pass:  upsweep  (per-block digit histogram of a contiguous chunk of tiles)
//            scan     (exclusive scan of the [digit][block] count matrix, digit-major)
//            downsweep(per tile: two stable 4-bit counting sub-passes in shared memory with per-thread
//                      register counters -> digit runs -> coalesced scatter of keys and values)
// The grid is a multiple of the SM count; every block owns a contiguous run of tiles so the only
// global state is the 256 x G count matrix (no inter-block dependencies, no look-back spinning).
//
// Algorithmic bytes per pass and element: sizeof(K) (upsweep read) + 2*(sizeof(K)+sizeof(V))
// (downsweep read + write).
#pragma once
#include <utility>
#include <vector>

#include "common.cuh"

namespace sufr {
namespace rsort {

constexpr int RADIX_BITS = 8;
constexpr int RADIX = 1 << RADIX_BITS;
constexpr int BLOCK = 256;
constexpr int WARPS = BLOCK / 32;

template <typename K, typename V>
struct Tuning {
    static constexpr int IPT = (sizeof(K) + sizeof(V) > 12) ? 12 : 16;
};

template <typename K>
__device__ __forceinline__ uint32_t digit_of(K key, int shift, uint32_t dmask) {
    return (uint32_t)(key >> shift) & dmask;
}

template <typename K, int IPT>
__global__ void __launch_bounds__(BLOCK) upsweep_kernel(const K* __restrict__ keys, uint64_t n, int shift,
                                                        uint32_t dmask, uint32_t* __restrict__ counts,
                                                        uint32_t tiles_per_block) {
    constexpr int TILE = BLOCK * IPT;
    __shared__ uint32_t hist[WARPS][RADIX];
    for (int i = threadIdx.x; i < WARPS * RADIX; i += BLOCK) (&hist[0][0])[i] = 0;
    __syncthreads();
    const int warp = threadIdx.x >> 5;
    uint64_t begin = (uint64_t)blockIdx.x * tiles_per_block * TILE;
    uint64_t end = begin + (uint64_t)tiles_per_block * TILE;
    if (end > n) end = n;
    for (uint64_t i = begin + threadIdx.x; i < end; i += BLOCK)
        atomicAdd(&hist[warp][digit_of(keys[i], shift, dmask)], 1u);
    __syncthreads();
    uint32_t acc = 0;
#pragma unroll
    for (int w = 0; w < WARPS; w++) acc += hist[w][threadIdx.x];
    counts[(uint64_t)threadIdx.x * gridDim.x + blockIdx.x] = acc;
}

// Exclusive scan of `len` u32 counts in place (single block).
__global__ void __launch_bounds__(1024) scan_counts_kernel(uint32_t* counts, uint32_t len) {
    __shared__ uint32_t warp_tot[32];
    __shared__ uint32_t carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t base = 0; base < len; base += 1024) {
        uint32_t i = base + threadIdx.x;
        uint32_t v = i < len ? counts[i] : 0;
        uint32_t incl = v;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            uint32_t o = __shfl_up_sync(0xffffffffu, incl, off);
            if (lane >= off) incl += o;
        }
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            uint32_t t = warp_tot[lane];
            uint32_t ti = t;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                uint32_t o = __shfl_up_sync(0xffffffffu, ti, off);
                if (lane >= off) ti += o;
            }
            warp_tot[lane] = ti - t;  // exclusive warp prefix
        }
        __syncthreads();
        uint32_t carry = carry_s;
        uint32_t excl = carry + warp_tot[warp] + incl - v;
        if (i < len) counts[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = excl + v;
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// downsweep: one 8-bit digit pass over a contiguous run of tiles.
//
// Ranking uses NO warp-wide instructions.  (match.any and vote both issue to the SM's single ADU pipe
// and capped the first two versions of this kernel at 32 % / 40 % of HBM bandwidth:
// profiles/r1_v0_downsweep_match_any_raw.csv, r1_v1_downsweep_vote_raw.csv.)  Instead the tile is
// sorted in shared memory by two stable 4-bit counting sub-passes (low nibble, then high nibble):
// every thread owns IPT consecutive items and counts its 16 nibble values in two byte-packed 64-bit
// registers, a block-wide scan over (nibble, thread) turns the counts into stable ranks, and the
// records are exchanged through a padded shared-memory buffer.  After the second sub-pass the tile is
// sorted by the full digit, digit runs are found by comparing neighbours, and the scatter to global
// memory is coalesced per digit run.
template <typename K, typename V, int IPT>
struct DownsweepSmem {
    static constexpr int TILE = BLOCK * IPT;
    static constexpr int PADDED = TILE + TILE / 16;  // one pad record per 16: conflict-free blocked reads
    static constexpr int TBASE_STRIDE = 9;           // words per thread row (16 u16 bases + pad)
    static constexpr size_t kKeys = 0;
    static constexpr size_t kVals = kKeys + sizeof(K) * PADDED;
    static constexpr size_t kTbase = (kVals + sizeof(V) * PADDED + 15) / 16 * 16;
    static constexpr size_t kWarpTot = kTbase + 4 * TBASE_STRIDE * BLOCK;
    static constexpr size_t kDigitTot = kWarpTot + 4 * WARPS * 8;
    static constexpr size_t kRunning = kDigitTot + 4 * 8;
    static constexpr size_t kGoff = kRunning + 4 * RADIX;
    static constexpr size_t kTstart = kGoff + 4 * RADIX;
    static constexpr size_t kTend = kTstart + 4 * RADIX;
    static constexpr size_t kBytes = kTend + 4 * RADIX;
};

__device__ __forceinline__ uint32_t padded_index(uint32_t e) { return e + (e >> 4); }

// N consecutive records of a thread as 16-byte loads (N * sizeof(T) is a multiple of 16 and p is 16-aligned)
template <typename T, int N>
__device__ __forceinline__ void load_blocked(const T* __restrict__ p, T (&r)[N]) {
    constexpr int PER = 16 / sizeof(T);
    static_assert((N * sizeof(T)) % 16 == 0, "blocked run must be a multiple of 16 bytes");
    const uint4* q = reinterpret_cast<const uint4*>(p);
#pragma unroll
    for (int j = 0; j < N / PER; j++) {
        uint4 x = __ldg(q + j);
        T tmp[PER];
        memcpy(tmp, &x, 16);
#pragma unroll
        for (int e = 0; e < PER; e++) r[j * PER + e] = tmp[e];
    }
}

// One stable 4-bit counting sub-pass: records in registers (blocked: thread t owns tile slots
// [t*IPT, (t+1)*IPT)) -> shared memory in sorted order of nibble `(key >> sh) & nmask`.
template <typename K, typename V, int IPT>
__device__ __forceinline__ void nibble_subpass(const K (&key)[IPT], const V (&val)[IPT], int sh, uint32_t nmask,
                                               unsigned char* smem) {
    using L = DownsweepSmem<K, V, IPT>;
    K* sk = reinterpret_cast<K*>(smem + L::kKeys);
    V* sv = reinterpret_cast<V*>(smem + L::kVals);
    uint32_t* tbase = reinterpret_cast<uint32_t*>(smem + L::kTbase);
    uint32_t* wtot = reinterpret_cast<uint32_t*>(smem + L::kWarpTot);
    uint32_t* dtot = reinterpret_cast<uint32_t*>(smem + L::kDigitTot);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // 1. per-thread counts of the 16 nibble values, one byte each (<= IPT <= 16)
    unsigned long long c0 = 0, c1 = 0, before = 0;  // `before`: 4 bits per item = equal nibbles earlier in this thread
#pragma unroll
    for (int i = 0; i < IPT; i++) {
        uint32_t d = (uint32_t)(key[i] >> sh) & nmask;
        uint32_t s8 = (d & 7u) * 8u;
        bool hi = (d & 8u) != 0;
        unsigned long long sel = hi ? c1 : c0;
        before |= (unsigned long long)((uint32_t)(sel >> s8) & 0xFFu) << (4 * i);
        unsigned long long inc = 1ull << s8;
        c0 += hi ? 0ull : inc;
        c1 += hi ? inc : 0ull;
    }
    // 2. block-wide exclusive scan in (nibble, thread) order; counts packed as u16 pairs: w[k] = (2k, 2k+1)
    uint32_t own[8], w[8];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        own[k] = ((uint32_t)(c0 >> (16 * k)) & 0xFFu) | (((uint32_t)(c0 >> (16 * k + 8)) & 0xFFu) << 16);
        own[k + 4] = ((uint32_t)(c1 >> (16 * k)) & 0xFFu) | (((uint32_t)(c1 >> (16 * k + 8)) & 0xFFu) << 16);
    }
#pragma unroll
    for (int k = 0; k < 8; k++) w[k] = own[k];
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
            uint32_t o = __shfl_up_sync(0xffffffffu, w[k], off);
            if (lane >= off) w[k] += o;
        }
    }
    if (lane == 31) {
#pragma unroll
        for (int k = 0; k < 8; k++) wtot[warp * 8 + k] = w[k];
    }
    __syncthreads();
    if (tid < 8) {  // exclusive prefix over warps for nibble pair `tid`
        uint32_t run = 0;
#pragma unroll
        for (int ww = 0; ww < WARPS; ww++) {
            uint32_t t = wtot[ww * 8 + tid];
            wtot[ww * 8 + tid] = run;
            run += t;
        }
        dtot[tid] = run;
    }
    __syncthreads();
    {
        uint32_t acc = 0;  // start of each nibble value in the sorted tile
#pragma unroll
        for (int k = 0; k < 8; k++) {
            uint32_t t = dtot[k];
            uint32_t lo = acc;
            acc += t & 0xFFFFu;
            uint32_t hi = acc;
            acc += t >> 16;
            // first rank of (nibble, this thread) = nibble start + earlier warps + earlier lanes of this warp
            tbase[tid * L::TBASE_STRIDE + k] = (lo | (hi << 16)) + wtot[warp * 8 + k] + w[k] - own[k];
        }
    }
    // 3. exchange (each thread reads back only its own tbase row: no barrier needed)
    const uint16_t* myb = reinterpret_cast<const uint16_t*>(tbase + tid * L::TBASE_STRIDE);
#pragma unroll
    for (int i = 0; i < IPT; i++) {
        uint32_t d = (uint32_t)(key[i] >> sh) & nmask;
        uint32_t r = (uint32_t)myb[d] + ((uint32_t)(before >> (4 * i)) & 15u);
        uint32_t pi = padded_index(r);
        sk[pi] = key[i];
        sv[pi] = val[i];
    }
    __syncthreads();
}

template <typename K, typename V, int IPT>
__global__ void __launch_bounds__(BLOCK, 2) downsweep_kernel(const K* __restrict__ kin, K* __restrict__ kout,
                                                             const V* __restrict__ vin, V* __restrict__ vout,
                                                             uint64_t n, int shift, uint32_t dmask,
                                                             const uint32_t* __restrict__ bases,
                                                             uint32_t tiles_per_block) {
    using L = DownsweepSmem<K, V, IPT>;
    constexpr int TILE = L::TILE;
    static_assert(IPT <= 16, "4-bit per-item counters");
    extern __shared__ __align__(16) unsigned char smem[];
    K* sk = reinterpret_cast<K*>(smem + L::kKeys);
    V* sv = reinterpret_cast<V*>(smem + L::kVals);
    uint32_t* running = reinterpret_cast<uint32_t*>(smem + L::kRunning);  // global write cursor per digit
    uint32_t* goff = reinterpret_cast<uint32_t*>(smem + L::kGoff);        // global index = goff[d] + slot (mod 2^32)
    uint32_t* tstart = reinterpret_cast<uint32_t*>(smem + L::kTstart);    // digit run [tstart, tend) in the tile
    uint32_t* tend = reinterpret_cast<uint32_t*>(smem + L::kTend);

    const int tid = threadIdx.x;
    running[tid] = bases[(uint64_t)tid * gridDim.x + blockIdx.x];

    const uint64_t total_tiles = (n + TILE - 1) / TILE;
    uint64_t t0 = (uint64_t)blockIdx.x * tiles_per_block;
    uint64_t t1 = t0 + tiles_per_block;
    if (t1 > total_tiles) t1 = total_tiles;
    const uint32_t lo_mask = dmask & 15u, hi_mask = dmask >> 4;

    for (uint64_t t = t0; t < t1; t++) {
        const uint64_t base = t * TILE;
        const uint32_t count = (n - base) < (uint64_t)TILE ? (uint32_t)(n - base) : (uint32_t)TILE;

        // blocked load: thread owns IPT consecutive records (stability order = slot order).
        // Slots past the end get the all-ones key: the stable sort keeps them last.
        K key[IPT];
        V val[IPT];
        if (count == (uint32_t)TILE) {
            load_blocked<K, IPT>(kin + base + (uint64_t)tid * IPT, key);
            load_blocked<V, IPT>(vin + base + (uint64_t)tid * IPT, val);
        } else {
#pragma unroll
            for (int i = 0; i < IPT; i++) {
                uint32_t idx = tid * IPT + i;
                key[i] = idx < count ? kin[base + idx] : (K)~(K)0;
                val[i] = idx < count ? vin[base + idx] : (V)0;
            }
        }
        tstart[tid] = 0;
        tend[tid] = 0;

        nibble_subpass<K, V, IPT>(key, val, shift, lo_mask, smem);
        if (hi_mask) {
#pragma unroll
            for (int i = 0; i < IPT; i++) {
                uint32_t pi = padded_index(tid * IPT + i);
                key[i] = sk[pi];
                val[i] = sv[pi];
            }
            nibble_subpass<K, V, IPT>(key, val, shift + 4, hi_mask, smem);
        }

        // digit runs of the sorted tile: slot s starts a run when its digit differs from slot s-1
        K kk[IPT];
#pragma unroll
        for (int k = 0; k < IPT; k++) {
            uint32_t s = k * BLOCK + tid;
            if (s < count) {
                kk[k] = sk[padded_index(s)];
                uint32_t d = digit_of(kk[k], shift, dmask);
                bool first = s == 0;
                uint32_t dp = 0;
                if (!first) dp = digit_of(sk[padded_index(s - 1)], shift, dmask);
                if (first || dp != d) {
                    tstart[d] = s;
                    if (!first) tend[dp] = s;
                }
                if (s + 1 == count) tend[d] = count;
            }
        }
        __syncthreads();
        {
            uint32_t ts = tstart[tid], te = tend[tid];
            goff[tid] = running[tid] - ts;
            running[tid] += te - ts;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < IPT; k++) {
            uint32_t s = k * BLOCK + tid;
            if (s < count) {
                uint32_t dst = goff[digit_of(kk[k], shift, dmask)] + s;
                kout[dst] = kk[k];
                vout[dst] = sv[padded_index(s)];
            }
        }
        __syncthreads();
    }
}

struct Plan {
    uint32_t grid = 0;
    uint32_t tiles_per_block = 0;
};

template <typename K, typename V>
inline Plan make_plan(uint64_t n) {
    constexpr int TILE = BLOCK * Tuning<K, V>::IPT;
    uint64_t tiles = div_up(n, TILE);
    uint64_t max_grid = (uint64_t)kNumSMs * 4;  // 4 resident CTAs per SM worth of chunks
    Plan p;
    p.grid = (uint32_t)(tiles < max_grid ? (tiles ? tiles : 1) : max_grid);
    p.tiles_per_block = (uint32_t)div_up(tiles ? tiles : 1, p.grid);
    p.grid = (uint32_t)div_up(tiles ? tiles : 1, p.tiles_per_block);
    return p;
}

inline size_t counts_words() { return (size_t)RADIX * kNumSMs * 4; }

// Sorts on key bits [begin_bit, end_bit).  Buffers ping-pong; returns true when the sorted data ended
// up in (keys_b, vals_b).  `counts` needs counts_words() u32.  `launches` (optional) is incremented
// by the number of kernels launched.
using EventPairs = std::vector<std::pair<cudaEvent_t, cudaEvent_t>>;

template <typename K, typename V>
bool sort_pairs(K* keys_a, K* keys_b, V* vals_a, V* vals_b, uint64_t n, int begin_bit, int end_bit,
                uint32_t* counts, cudaStream_t stream, uint64_t* launches = nullptr,
                EventPairs* downsweep_events = nullptr) {
    constexpr int IPT = Tuning<K, V>::IPT;
    if (n == 0 || end_bit <= begin_bit) return false;
    Plan p = make_plan<K, V>(n);
    SUFR_CUDA_CHECK(cudaFuncSetAttribute(downsweep_kernel<K, V, IPT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)DownsweepSmem<K, V, IPT>::kBytes));
    bool in_b = false;
    for (int bit = begin_bit; bit < end_bit; bit += RADIX_BITS) {
        int nb = end_bit - bit < RADIX_BITS ? end_bit - bit : RADIX_BITS;
        uint32_t dmask = (1u << nb) - 1u;
        K* kin = in_b ? keys_b : keys_a;
        K* kout = in_b ? keys_a : keys_b;
        V* vin = in_b ? vals_b : vals_a;
        V* vout = in_b ? vals_a : vals_b;
        upsweep_kernel<K, IPT><<<p.grid, BLOCK, 0, stream>>>(kin, n, bit, dmask, counts, p.tiles_per_block);
        SUFR_KERNEL_CHECK();
        scan_counts_kernel<<<1, 1024, 0, stream>>>(counts, (uint32_t)RADIX * p.grid);
        SUFR_KERNEL_CHECK();
        cudaEvent_t e0 = nullptr, e1 = nullptr;
        if (downsweep_events) {
            SUFR_CUDA_CHECK(cudaEventCreate(&e0));
            SUFR_CUDA_CHECK(cudaEventCreate(&e1));
            SUFR_CUDA_CHECK(cudaEventRecord(e0, stream));
        }
        downsweep_kernel<K, V, IPT><<<p.grid, BLOCK, DownsweepSmem<K, V, IPT>::kBytes, stream>>>(
            kin, kout, vin, vout, n, bit, dmask, counts, p.tiles_per_block);
        SUFR_KERNEL_CHECK();
        if (downsweep_events) {
            SUFR_CUDA_CHECK(cudaEventRecord(e1, stream));
            downsweep_events->push_back({e0, e1});
        }
        if (launches) *launches += 3;
        in_b = !in_b;
    }
    return in_b;
}

}  // namespace rsort
}  // namespace sufr
