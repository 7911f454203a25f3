// Device-wide inclusive scan with functor input/output (reduce -> spine -> apply).
//
// Used for every order-preserving compaction on the path (active-group compaction, shard
// selection, suffix filter) and for the max / segmented-min scans of the refinement stages.
// Three launches, two coalesced reads of the input functor, no inter-block dependencies
// (so no forward-progress assumptions).  The operator must be associative; it need not commute.
#pragma once
#include "common.cuh"

namespace sufr {
namespace scan {

constexpr int BLOCK = 256;
constexpr int ITEMS = 16;
constexpr int CHUNK = BLOCK * ITEMS;

struct SumU32 {
    using T = uint32_t;
    __host__ __device__ static T identity() { return 0; }
    __device__ T operator()(T a, T b) const { return a + b; }
};
struct SumU64 {
    using T = unsigned long long;
    __host__ __device__ static T identity() { return 0; }
    __device__ T operator()(T a, T b) const { return a + b; }
};
struct MaxU32 {
    using T = uint32_t;
    __host__ __device__ static T identity() { return 0; }
    __device__ T operator()(T a, T b) const { return a > b ? a : b; }
};
// Segmented min over packed (flag << 32 | value): a set flag on the right operand restarts the segment.
struct SegMinU64 {
    using T = unsigned long long;
    __host__ __device__ static T identity() { return 0x00000000FFFFFFFFull; }
    __device__ T operator()(T a, T b) const {
        uint32_t fa = (uint32_t)(a >> 32), fb = (uint32_t)(b >> 32);
        uint32_t va = (uint32_t)a, vb = (uint32_t)b;
        uint32_t v = fb ? vb : (va < vb ? va : vb);
        return ((T)(fa | fb) << 32) | v;
    }
};

template <typename Op>
__device__ __forceinline__ typename Op::T warp_inclusive(typename Op::T v, Op op, int lane) {
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        typename Op::T o = __shfl_up_sync(0xffffffffu, v, off);
        if (lane >= off) v = op(o, v);
    }
    return v;
}

// Inclusive scan across the 256 threads of a block. `smem` holds 8 warp totals. Returns the block total
// through `total` (valid in every thread).
template <typename Op>
__device__ __forceinline__ typename Op::T block_inclusive(typename Op::T v, Op op, typename Op::T* smem,
                                                         typename Op::T& total) {
    using T = typename Op::T;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T incl = warp_inclusive(v, op, lane);
    __syncthreads();  // protect smem reuse across calls
    if (lane == 31) smem[warp] = incl;
    __syncthreads();
    T prefix = Op::identity();
    T run = smem[0];
#pragma unroll
    for (int w = 1; w < BLOCK / 32; w++) {
        if (w == warp) prefix = run;
        run = op(run, smem[w]);
    }
    total = run;
    if (warp > 0) incl = op(prefix, incl);
    return incl;
}

// Layout of a chunk: warp w owns the 32 * ITEMS consecutive elements [w * 32 * ITEMS, ...), as ITEMS rows of 32; lane l
// reads element row * 32 + l, so every access the input / output functors make with consecutive indices is coalesced
// (a thread that owned ITEMS consecutive elements would stride its warp's loads by ITEMS elements).
// Scan order inside the warp is row-major: a warp scan per row, a running prefix over the rows.
template <typename Op>
__device__ __forceinline__ typename Op::T warp_reduce_in_order(typename Op::T v, Op op) {
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        typename Op::T o = __shfl_down_sync(0xffffffffu, v, off);
        v = op(v, o);  // lane 0 ends with e0 (+) e1 (+) ... (+) e31; the other lanes are not used
    }
    return v;
}

template <typename Op, typename In>
__global__ void __launch_bounds__(BLOCK) reduce_kernel(uint64_t n, In in, Op op, typename Op::T* partials) {
    using T = typename Op::T;
    __shared__ T smem[BLOCK / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t w0 = (uint64_t)blockIdx.x * CHUNK + (uint64_t)warp * (32 * ITEMS) + lane;
    T v[ITEMS];
#pragma unroll
    for (int k = 0; k < ITEMS; k++) {
        const uint64_t i = w0 + (uint64_t)k * 32;
        v[k] = i < n ? in(i) : Op::identity();
    }
    T acc = Op::identity();
#pragma unroll
    for (int k = 0; k < ITEMS; k++) acc = op(acc, warp_reduce_in_order(v[k], op));  // meaningful in lane 0
    if (lane == 0) smem[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        T total = smem[0];
#pragma unroll
        for (int w = 1; w < BLOCK / 32; w++) total = op(total, smem[w]);
        partials[blockIdx.x] = total;
    }
}

// Single block: exclusive scan of the per-chunk partials, in place. `total_out` receives the grand total.
template <typename Op>
__global__ void __launch_bounds__(BLOCK) spine_kernel(uint32_t nblocks, Op op, typename Op::T* partials,
                                                      typename Op::T* total_out) {
    using T = typename Op::T;
    __shared__ T smem[BLOCK / 32];
    T carry = Op::identity();
    for (uint32_t base = 0; base < nblocks; base += BLOCK) {
        uint32_t i = base + threadIdx.x;
        T v = i < nblocks ? partials[i] : Op::identity();
        T total;
        T incl = block_inclusive(v, op, smem, total);
        // exclusive = carry (+) inclusive of the previous thread
        T prev = __shfl_up_sync(0xffffffffu, incl, 1);
        __shared__ T last_of_warp[BLOCK / 32];
        __syncthreads();
        if ((threadIdx.x & 31) == 31) last_of_warp[threadIdx.x >> 5] = incl;
        __syncthreads();
        T excl;
        if (threadIdx.x == 0) excl = Op::identity();
        else if ((threadIdx.x & 31) == 0) excl = last_of_warp[(threadIdx.x >> 5) - 1];
        else excl = prev;
        if (i < nblocks) partials[i] = op(carry, excl);
        carry = op(carry, total);
    }
    if (threadIdx.x == 0 && total_out) *total_out = carry;
}

// out(i, v, inclusive) is called for every i < n; for sum scans exclusive = inclusive - v.
template <typename Op, typename In, typename Out>
__global__ void __launch_bounds__(BLOCK) apply_kernel(uint64_t n, In in, Op op, const typename Op::T* partials,
                                                      Out out) {
    using T = typename Op::T;
    __shared__ T smem[BLOCK / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t w0 = (uint64_t)blockIdx.x * CHUNK + (uint64_t)warp * (32 * ITEMS) + lane;
    T v[ITEMS], s[ITEMS];
#pragma unroll
    for (int k = 0; k < ITEMS; k++) {
        const uint64_t i = w0 + (uint64_t)k * 32;
        v[k] = i < n ? in(i) : Op::identity();
    }
#pragma unroll
    for (int k = 0; k < ITEMS; k++) s[k] = warp_inclusive(v[k], op, lane);
    // running prefix over the rows of this warp (the row totals are warp-uniform)
    T wtotal = Op::identity();
#pragma unroll
    for (int k = 0; k < ITEMS; k++) {
        const T row = __shfl_sync(0xffffffffu, s[k], 31);
        s[k] = op(wtotal, s[k]);
        wtotal = op(wtotal, row);
    }
    if (lane == 0) smem[warp] = wtotal;
    __syncthreads();
    T prefix = partials[blockIdx.x];
#pragma unroll
    for (int w = 0; w < BLOCK / 32; w++)
        if (w < warp) prefix = op(prefix, smem[w]);
#pragma unroll
    for (int k = 0; k < ITEMS; k++) {
        const uint64_t i = w0 + (uint64_t)k * 32;
        if (i < n) out(i, v[k], op(prefix, s[k]));
    }
}

inline size_t partials_count(uint64_t n) { return (size_t)div_up(n, CHUNK) + 1; }

// Two-step form: scan_reduce leaves the exclusive chunk prefixes in `partials` (and the grand total in its last
// element) so that the caller can size its outputs before scan_apply runs the second pass.
template <typename Op, typename In>
void scan_reduce(uint64_t n, In in, Op op, typename Op::T* partials, cudaStream_t stream) {
    if (n == 0) {
        typename Op::T id = Op::identity();
        SUFR_CUDA_CHECK(cudaMemcpyAsync(partials, &id, sizeof(id), cudaMemcpyHostToDevice, stream));
        return;
    }
    uint32_t nblocks = div_up_u32(n, CHUNK);
    reduce_kernel<Op, In><<<nblocks, BLOCK, 0, stream>>>(n, in, op, partials);
    SUFR_KERNEL_CHECK();
    spine_kernel<Op><<<1, BLOCK, 0, stream>>>(nblocks, op, partials, partials + nblocks);
    SUFR_KERNEL_CHECK();
}
template <typename Op, typename In, typename Out>
void scan_apply(uint64_t n, In in, Op op, Out out, const typename Op::T* partials, cudaStream_t stream) {
    if (n == 0) return;
    uint32_t nblocks = div_up_u32(n, CHUNK);
    apply_kernel<Op, In, Out><<<nblocks, BLOCK, 0, stream>>>(n, in, op, partials, out);
    SUFR_KERNEL_CHECK();
}

// `partials` must hold partials_count(n) elements; the last one receives the grand total.
template <typename Op, typename In, typename Out>
void inclusive_scan(uint64_t n, In in, Op op, Out out, typename Op::T* partials, cudaStream_t stream) {
    if (n == 0) {
        typename Op::T id = Op::identity();
        SUFR_CUDA_CHECK(cudaMemcpyAsync(partials, &id, sizeof(id), cudaMemcpyHostToDevice, stream));
        return;
    }
    uint32_t nblocks = div_up_u32(n, CHUNK);
    reduce_kernel<Op, In><<<nblocks, BLOCK, 0, stream>>>(n, in, op, partials);
    SUFR_KERNEL_CHECK();
    spine_kernel<Op><<<1, BLOCK, 0, stream>>>(nblocks, op, partials, partials + nblocks);
    SUFR_KERNEL_CHECK();
    apply_kernel<Op, In, Out><<<nblocks, BLOCK, 0, stream>>>(n, in, op, partials, out);
    SUFR_KERNEL_CHECK();
}

}  // namespace scan
}  // namespace sufr
