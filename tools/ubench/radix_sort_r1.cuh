// Hand-written stable LSD radix sort of (key, value) pairs, 8-bit digits.
//
// Per pass:  upsweep  (per-block digit histogram of a contiguous chunk of tiles)
//            scan     (exclusive scan of the [digit][block] count matrix, digit-major)
//            downsweep(per tile: warp-level match ranking -> block digit offsets -> shared-memory
//                      exchange -> coalesced scatter of keys, then values)
// The grid is a multiple of the SM count; every block owns a contiguous run of tiles so the only
// global state is the 256 x G count matrix (no inter-block dependencies, no look-back spinning).
//
// Algorithmic bytes per pass and element: sizeof(K) (upsweep read) + 2*(sizeof(K)+sizeof(V))
// (downsweep read + write).
#pragma once
#include <utility>
#include <vector>

#include "../../sufr_b200/csrc/common.cuh"

#include <cuda_pipeline.h>

namespace sufr {
namespace rsort_r1 {

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

#ifndef SUFR_RSORT_MIN_CTAS
#define SUFR_RSORT_MIN_CTAS 2
#endif
template <typename K, typename V, int IPT>
__global__ void __launch_bounds__(BLOCK, SUFR_RSORT_MIN_CTAS) downsweep_kernel(const K* __restrict__ kin, K* __restrict__ kout,
                                                          const V* __restrict__ vin, V* __restrict__ vout,
                                                          uint64_t n, int shift, uint32_t dmask,
                                                          const uint32_t* __restrict__ bases,
                                                          uint32_t tiles_per_block) {
    constexpr int TILE = BLOCK * IPT;
    __shared__ uint32_t wc[WARPS][RADIX];   // per-warp digit counters -> tile-local start of (warp, digit)
    __shared__ uint32_t running[RADIX];     // global write cursor of each digit for this block
    __shared__ uint32_t goff[RADIX];        // global index = goff[d] + tile-local slot (mod 2^32)
    __shared__ uint32_t warp_tot[WARPS];
    // Dynamic shared memory (> 48 KB): exk / exv hold the tile in tile-local sorted order (separate buffers, so the
    // registers are free as soon as both are exchanged); pk / pv receive the NEXT tile by cp.async while this one
    // is ranked, exchanged and scattered, so no warp waits on global loads at the top of the loop.
    extern __shared__ __align__(16) unsigned char ex_raw[];
    K* exk = reinterpret_cast<K*>(ex_raw);
    V* exv = reinterpret_cast<V*>(ex_raw + sizeof(K) * TILE);
    K* pk = reinterpret_cast<K*>(ex_raw + (sizeof(K) + sizeof(V)) * TILE);
    V* pv = reinterpret_cast<V*>(ex_raw + (2 * sizeof(K) + sizeof(V)) * TILE);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    running[tid] = bases[(uint64_t)tid * gridDim.x + blockIdx.x];

    const uint64_t total_tiles = (n + TILE - 1) / TILE;
    uint64_t t0 = (uint64_t)blockIdx.x * tiles_per_block;
    uint64_t t1 = t0 + tiles_per_block;
    if (t1 > total_tiles) t1 = total_tiles;

    // 16-byte chunks; elements beyond the end of the input are zero-filled
    auto prefetch = [&](uint64_t t) {
        const uint64_t base = t * TILE;
        const uint32_t cnt = (n - base) < (uint64_t)TILE ? (uint32_t)(n - base) : (uint32_t)TILE;
        constexpr uint32_t KPC = 16 / sizeof(K), VPC = 16 / sizeof(V);  // elements per chunk
        for (uint32_t c = tid; c < TILE / KPC; c += BLOCK) {
            const uint32_t e = c * KPC;
            const uint32_t valid = cnt > e ? (cnt - e < KPC ? cnt - e : KPC) : 0u;
            __pipeline_memcpy_async(pk + e, valid ? kin + base + e : kin, 16, 16 - valid * sizeof(K));
        }
        for (uint32_t c = tid; c < TILE / VPC; c += BLOCK) {
            const uint32_t e = c * VPC;
            const uint32_t valid = cnt > e ? (cnt - e < VPC ? cnt - e : VPC) : 0u;
            __pipeline_memcpy_async(pv + e, valid ? vin + base + e : vin, 16, 16 - valid * sizeof(V));
        }
        __pipeline_commit();
    };
    if (t0 < t1) prefetch(t0);
    __pipeline_wait_prior(0);
    __syncthreads();

    for (uint64_t t = t0; t < t1; t++) {
        const uint64_t base = t * TILE;
        const uint32_t count = (n - base) < (uint64_t)TILE ? (uint32_t)(n - base) : (uint32_t)TILE;

        K key[IPT];
        V val[IPT];
        uint16_t slot[IPT];
#pragma unroll
        for (int i = 0; i < IPT; i++) {
            const uint32_t idx = warp * (32 * IPT) + i * 32 + lane;
            key[i] = pk[idx];
            val[i] = pv[idx];
        }
#pragma unroll
        for (int i = 0; i < WARPS; i++) wc[i][tid] = 0;
        __syncthreads();
        if (t + 1 < t1) prefetch(t + 1);  // every thread has read its part of pk / pv

        // warp-level ranking: items are visited in tile order (warp, i, lane) => stable.
        // Pass 1: lanes holding the same digit, by one vote per digit bit (match.any saturates the ADU pipe:
        // profiles/r1_v0_downsweep_match_any_raw.csv).  All rows' votes are independent of each other.
        uint32_t peers_of[IPT];
#pragma unroll
        for (int i = 0; i < IPT; i++) {
            uint32_t d = digit_of(key[i], shift, dmask);
            // valid lanes of this row, computed instead of voted (votes are the bottleneck of this kernel)
            int rem = (int)count - (int)(warp * (32 * IPT) + i * 32);
            unsigned vm = rem >= 32 ? 0xffffffffu : (rem <= 0 ? 0u : ((1u << rem) - 1u));
            unsigned peers = vm;
            if (vm) {  // warp-uniform
#pragma unroll
                for (int b = 0; b < RADIX_BITS; b++) {
                    unsigned vote = __ballot_sync(0xffffffffu, (d >> b) & 1u);
                    peers &= ((d >> b) & 1u) ? vote : ~vote;
                }
            }
            peers_of[i] = ((vm >> lane) & 1u) ? peers : 0u;
        }
        // Pass 2: the serial chain through the per-warp digit counters.  The shuffle uses the FULL mask (every lane
        // is here; lanes without an element read lane 31 and ignore it): a shuffle under the per-group mask
        // `peers` is executed once per distinct digit of the row and was half of the ranking cost
        // (tools/ubench/warp_prims.cu: 65 -> 32 SM-cycles per row).
#pragma unroll
        for (int i = 0; i < IPT; i++) {
            const unsigned peers = peers_of[i];
            const uint32_t d = digit_of(key[i], shift, dmask);
            const int leader = __ffs(peers) - 1;  // -1 for lanes without an element
            uint32_t old = 0;
            if (lane == leader) {
                old = wc[warp][d];
                wc[warp][d] = old + __popc(peers);
            }
            old = __shfl_sync(0xffffffffu, old, leader & 31);
            slot[i] = (uint16_t)(old + __popc(peers & lt_mask));
            __syncwarp();
        }
        __syncthreads();

        // thread tid owns digit tid: exclusive scan over warps, then over digits
        uint32_t tile_count = 0;
#pragma unroll
        for (int w = 0; w < WARPS; w++) {
            uint32_t c = wc[w][tid];
            wc[w][tid] = tile_count;
            tile_count += c;
        }
        uint32_t incl = tile_count;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            uint32_t o = __shfl_up_sync(0xffffffffu, incl, off);
            if (lane >= off) incl += o;
        }
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        uint32_t wprefix = 0;
#pragma unroll
        for (int w = 0; w < WARPS; w++)
            if (w < warp) wprefix += warp_tot[w];
        uint32_t tile_start = wprefix + incl - tile_count;
        goff[tid] = running[tid] - tile_start;
        running[tid] += tile_count;
#pragma unroll
        for (int w = 0; w < WARPS; w++) wc[w][tid] += tile_start;
        __syncthreads();

        // exchange through shared memory so that the global writes are digit-contiguous
#pragma unroll
        for (int i = 0; i < IPT; i++) {
            uint32_t idx = warp * (32 * IPT) + i * 32 + lane;
            if (idx < count) {
                uint32_t s = slot[i] + wc[warp][digit_of(key[i], shift, dmask)];
                exk[s] = key[i];
                exv[s] = val[i];
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < IPT; k++) {
            uint32_t s = k * BLOCK + tid;
            if (s < count) {
                K kk = exk[s];
                uint32_t dst = goff[digit_of(kk, shift, dmask)] + s;
                kout[dst] = kk;
                vout[dst] = exv[s];
            }
        }
        __pipeline_wait_prior(0);
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
    uint64_t max_grid = (uint64_t)kNumSMs * 4;  // two full waves of the 2 resident CTAs per SM
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
    constexpr size_t ex_bytes = 2 * (sizeof(K) + sizeof(V)) * BLOCK * IPT;  // exchange + prefetch buffers
    static bool attr_set[64] = {};  // per instantiation and device
    allow_dynamic_smem(downsweep_kernel<K, V, IPT>, ex_bytes, attr_set);
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
        downsweep_kernel<K, V, IPT><<<p.grid, BLOCK, ex_bytes, stream>>>(kin, kout, vin, vout, n, bit, dmask, counts,
                                                                 p.tiles_per_block);
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

}  // namespace rsort_r1
}  // namespace sufr
