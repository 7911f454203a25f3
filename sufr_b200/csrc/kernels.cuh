// Kernels and scan functors of the build pipeline (everything except the radix sort).
#pragma once
#include "common.cuh"
#include "keys.cuh"
#include "scan.cuh"

#include <cuda_pipeline.h>

namespace sufr {

constexpr int kBlock = 256;

// Exclusive scan of 64 shared-memory counters by warp 0 (two per lane); returns the total in every lane of
// warp 0.  Call from warp 0 only, between two __syncthreads().
__device__ __forceinline__ uint32_t warp_scan64(uint32_t* cnt) {
    const int lane = threadIdx.x & 31;
    uint32_t v0 = cnt[2 * lane], v1 = cnt[2 * lane + 1];
    uint32_t s = v0 + v1, incl = s;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        uint32_t o = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += o;
    }
    uint32_t excl = incl - s;
    cnt[2 * lane] = excl;
    cnt[2 * lane + 1] = excl + v0;
    return __shfl_sync(0xffffffffu, incl, 31);
}

// scan outputs that only want the grand total
struct CountOnly {
    __device__ void operator()(uint64_t, uint32_t, uint32_t) const {}
};
struct CountOnlyU64 {
    __device__ void operator()(uint64_t, unsigned long long, unsigned long long) const {}
};

inline uint32_t grid_for(uint64_t n, int per_thread = 1) {
    uint64_t blocks = div_up(n, (uint64_t)kBlock * per_thread);
    uint64_t cap = (uint64_t)num_sms() * 32;  // grid-stride beyond 32 CTAs per SM
    if (blocks > cap) blocks = cap;
    return (uint32_t)(blocks ? blocks : 1);
}

// ------------------------------------------------------------------ encode (sufr_builder.rs:144-160)
// Lowercase ASCII -> 'N' (ignore_softmask) or uppercase; also records which bytes occur.  Four bytes at a time:
// 0x80 in every byte of w that equals the byte replicated in pat
__device__ __forceinline__ uint32_t swar_eq(uint32_t w, uint32_t pat) {
    uint32_t v = w ^ pat;
    return ~(((v & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | v) & 0x80808080u;
}
__device__ __forceinline__ uint32_t transform_word(uint32_t w, int ignore_softmask) {
    uint32_t x = w & 0x7F7F7F7Fu;
    uint32_t ge97 = x + 0x1F1F1F1Fu;   // bit 7 of a byte: x >= 97
    uint32_t ge123 = x + 0x05050505u;  // bit 7 of a byte: x >= 123
    uint32_t lower = ge97 & ~ge123 & ~w & 0x80808080u;
    uint32_t lm = (lower >> 7) * 0xFFu;
    return ignore_softmask ? ((w & ~lm) | (0x4E4E4E4Eu & lm)) : (w & ~(lm & 0x20202020u));
}
__device__ __forceinline__ uint32_t transform_byte(uint32_t c, int ignore_softmask) {
    if (c >= 97 && c <= 122) c = ignore_softmask ? (uint32_t)'N' : (c & 0x5F);
    return c;
}
__global__ void __launch_bounds__(kBlock) transform_kernel(const uint8_t* __restrict__ in,
                                                           uint8_t* __restrict__ out, uint64_t n,
                                                           int ignore_softmask, uint32_t* __restrict__ present,
                                                           unsigned long long* __restrict__ sample_counts,
                                                           unsigned long long* __restrict__ indexed_count) {
    __shared__ uint32_t seen[256];
    __shared__ uint32_t cnt[256];  // byte counts of a uniform 1/64 sample (chooses the 4 "regular" bytes)
    seen[threadIdx.x] = 0;
    cnt[threadIdx.x] = 0;
    unsigned long long nidx = 0;   // bytes that start an indexed suffix under --dna (sufr_builder.rs:446-449)
    uint32_t acc_a = 0, acc_c = 0, acc_g = 0, acc_t = 0;
    __syncthreads();
    const bool aligned = ((((uintptr_t)in) | ((uintptr_t)out)) & 15) == 0;
    const uint64_t nvec = aligned ? n / 16 : 0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += stride) {
        uint4 x = reinterpret_cast<const uint4*>(in)[v];
        uint32_t w[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint32_t t = transform_word(w[k], ignore_softmask);
            const uint32_t ea = swar_eq(t, 0x41414141u), ec = swar_eq(t, 0x43434343u);
            const uint32_t eg = swar_eq(t, 0x47474747u), et = swar_eq(t, 0x54545454u);
            const uint32_t acgt = ea | ec | eg | et;
            acc_a |= ea; acc_c |= ec; acc_g |= eg; acc_t |= et;
            nidx += __popc(acgt | swar_eq(t, 0x24242424u));
            if (acgt != 0x80808080u) {  // other bytes (rare in DNA): per-byte presence
#pragma unroll
                for (int b = 0; b < 4; b++)
                    if (!((acgt >> (8 * b + 7)) & 1u)) seen[(t >> (8 * b)) & 0xFFu] = 1;
            }
            w[k] = t;
        }
        if ((v & 63u) == 0) {  // every 64th 16-byte vector of the text: a uniform 1/64 sample
#pragma unroll
            for (int k = 0; k < 4; k++)
#pragma unroll
                for (int b = 0; b < 4; b++) atomicAdd(&cnt[(w[k] >> (8 * b)) & 0xFFu], 1u);
        }
        reinterpret_cast<uint4*>(out)[v] = make_uint4(w[0], w[1], w[2], w[3]);
    }
    if (acc_a) seen['A'] = 1;
    if (acc_c) seen['C'] = 1;
    if (acc_g) seen['G'] = 1;
    if (acc_t) seen['T'] = 1;
    for (uint64_t i = nvec * 16 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint32_t c = transform_byte(in[i], ignore_softmask);
        seen[c] = 1;
        if (nvec == 0) atomicAdd(&cnt[c], 1u);  // unaligned / tiny inputs: the tail is the whole text
        nidx += (c == 'A' || c == 'C' || c == 'G' || c == 'T' || c == '$') ? 1 : 0;
        out[i] = (uint8_t)c;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) nidx += __shfl_down_sync(0xffffffffu, nidx, off);
    if ((threadIdx.x & 31) == 0 && nidx) atomicAdd(indexed_count, nidx);
    __syncthreads();
    if (seen[threadIdx.x]) present[threadIdx.x] = 1;
    if (cnt[threadIdx.x]) atomicAdd(&sample_counts[threadIdx.x], (unsigned long long)cnt[threadIdx.x]);
}

// Packed text, both forms in one pass over the transformed text.
//  * `words`: one 64-bit word (K symbols of `bits` bits) per thread.  For K <= 24 a thread reads its K bytes from
//    the staged tile as seven 32-bit words and realigns them with funnel shifts (byte-wise shared-memory reads at
//    a stride of K bytes are 5-way bank conflicts).
//  * `packed2` / `irr32` (2-bit fast path, optional): 32 symbols per 64-bit word and the matching 32 bits of `irr`.
//    cls2[b] = 2-bit code of byte b (rank of a regular byte, min(class, 3) of an irregular one) | 4 if irregular.
// A tile is the 256*K bytes of one block iteration (a multiple of 32 bytes, so packed2 words do not straddle
// tiles); tiles are double-buffered in shared memory with 16-byte cp.async (`text` must be 16-byte aligned and
// padded by 16 bytes), so the loads of the next tile are in flight while this one is packed.
constexpr int kPackTile = kBlock * 64 + 32;
__global__ void __launch_bounds__(kBlock) pack_kernel(const uint8_t* __restrict__ text, uint64_t n,
                                                      const uint8_t* __restrict__ code_lut, uint32_t bits,
                                                      uint32_t K, uint64_t num_words, uint64_t* __restrict__ words,
                                                      const uint8_t* __restrict__ cls2, uint64_t num_words2,
                                                      uint64_t* __restrict__ packed2, uint32_t* __restrict__ irr32) {
    __shared__ uint8_t lut[256];
    __shared__ uint8_t lut2[256];
    __shared__ __align__(16) uint8_t raw[2][kPackTile];
    lut[threadIdx.x] = code_lut[threadIdx.x];
    lut2[threadIdx.x] = cls2 ? cls2[threadIdx.x] : 0;
    const uint64_t blocks = (num_words + kBlock - 1) / kBlock;
    const uint32_t tile_bytes = kBlock * K;
    const uint32_t nvec = (tile_bytes + 15) / 16 + 2;
    auto stage = [&](int buf, uint64_t b) {
        const uint64_t byte0 = b * tile_bytes;
        for (uint32_t v = threadIdx.x; v < nvec; v += kBlock) {
            const uint64_t off = byte0 + (uint64_t)v * 16;
            const bool in = off < n;  // may read the 16 padding bytes; beyond them: zero fill
            __pipeline_memcpy_async(raw[buf] + v * 16, in ? text + off : text, 16, in ? 0 : 16);
        }
    };
    int buf = 0;
    uint64_t b = blockIdx.x;
    if (b < blocks) stage(0, b);
    __pipeline_commit();
    for (; b < blocks; b += gridDim.x, buf ^= 1) {
        if (b + gridDim.x < blocks) stage(buf ^ 1, b + gridDim.x);
        __pipeline_commit();
        __pipeline_wait_prior(1);
        __syncthreads();
        const uint8_t* tile = raw[buf];
        const uint64_t byte0 = b * tile_bytes;
        const uint64_t w = b * kBlock + threadIdx.x;
        if (w < num_words) {
            const uint64_t base = w * K;
            uint64_t x = 0;
            if (K <= 24 && base + K <= n) {
                // Horner accumulation: one multiply-add per symbol instead of a variable 64-bit shift
                const uint32_t b0 = threadIdx.x * K;
                const uint32_t* r32 = reinterpret_cast<const uint32_t*>(tile) + (b0 >> 2);
                const uint32_t sh = 8 * (b0 & 3);
                const uint64_t mult = 1ull << bits;
                uint32_t r[7], a[6];
#pragma unroll
                for (int i = 0; i < 7; i++) r[i] = r32[i];
#pragma unroll
                for (int i = 0; i < 6; i++) a[i] = __funnelshift_r(r[i], r[i + 1], sh);
#pragma unroll
                for (uint32_t j = 0; j < 24; j++)
                    if (j < K) x = x * mult + lut[(a[j >> 2] >> (8 * (j & 3))) & 0xFFu];
                x <<= 64 - K * bits;
            } else {
                for (uint32_t j = 0; j < K; j++) {
                    uint64_t c = base + j < n ? lut[tile[threadIdx.x * K + j]] : 0;
                    x |= c << (64 - bits * (j + 1));
                }
            }
            words[w] = x;
        }
        if (cls2) {
            // this tile's packed2 words; the last tile also writes the padding words behind the text
            const uint64_t first2 = byte0 / 32;
            const uint64_t end2 = b + 1 == blocks ? num_words2 : first2 + tile_bytes / 32;
            for (uint64_t w2 = first2 + threadIdx.x; w2 < end2; w2 += kBlock) {
                const uint32_t t = (uint32_t)(w2 - first2);
                uint32_t bytes[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                if (t * 32 < tile_bytes) {
                    const uint4 lo4 = reinterpret_cast<const uint4*>(tile)[2 * t];
                    const uint4 hi4 = reinterpret_cast<const uint4*>(tile)[2 * t + 1];
                    bytes[0] = lo4.x; bytes[1] = lo4.y; bytes[2] = lo4.z; bytes[3] = lo4.w;
                    bytes[4] = hi4.x; bytes[5] = hi4.y; bytes[6] = hi4.z; bytes[7] = hi4.w;
                }
                const uint64_t base2 = w2 * 32;
                uint32_t xh = 0, xl = 0, m = 0;  // symbols 0..15 / 16..31 of the word, irregular bits
                if (base2 + 32 <= n) {
#pragma unroll
                    for (uint32_t j = 0; j < 16; j++) {
                        const uint32_t c = lut2[(bytes[j >> 2] >> (8 * (j & 3))) & 0xFFu];
                        xh = xh * 4 + (c & 3u);
                        m = m * 2 + (c >> 2);
                    }
#pragma unroll
                    for (uint32_t j = 16; j < 32; j++) {
                        const uint32_t c = lut2[(bytes[j >> 2] >> (8 * (j & 3))) & 0xFFu];
                        xl = xl * 4 + (c & 3u);
                        m = m * 2 + (c >> 2);
                    }
                } else {
#pragma unroll
                    for (uint32_t j = 0; j < 32; j++) {
                        const uint32_t byte = (bytes[j >> 2] >> (8 * (j & 3))) & 0xFFu;
                        const uint32_t c = base2 + j < n ? lut2[byte] : 4u;  // beyond the text: irregular, class 0
                        if (j < 16) xh = xh * 4 + (c & 3u); else xl = xl * 4 + (c & 3u);
                        m = m * 2 + (c >> 2);
                    }
                }
                packed2[w2] = ((uint64_t)xh << 32) | xl;
                // irr is addressed as 64-bit words with symbol 0 in the top bit: 32-bit halves are swapped on little-endian
                irr32[w2 ^ 1] = m;
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------ first key word of every suffix
__device__ __forceinline__ bool indexed_byte(uint8_t c) { return c == '$' || c == 'A' || c == 'C' || c == 'G' || c == 'T'; }

// ------------------------------------------------------------------ synthetic workloads (bench.py)
__device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
// base i = "ACGT"[top 2 bits of splitmix64(seed, i)]
__global__ void __launch_bounds__(kBlock) synth_dna_kernel(uint8_t* __restrict__ text, uint64_t n, uint64_t seed) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint64_t z = mix64(seed * 0x9E3779B97F4A7C15ull + (i + 1) * 0x9E3779B97F4A7C15ull);
        text[i] = (uint8_t)("ACGT"[z >> 62]);
    }
}
__global__ void synth_marks_kernel(uint8_t* __restrict__ text, uint64_t n, const uint64_t* __restrict__ starts,
                                   uint64_t num, uint8_t delim) {
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= 1 && k < num && starts[k] >= 1 && starts[k] - 1 < n) text[starts[k] - 1] = delim;
    if (k == 0 && n) text[n - 1] = '$';
}

}  // namespace sufr
