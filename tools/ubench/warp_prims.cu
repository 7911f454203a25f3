// Micro-benchmark: per-SM throughput of the warp primitives the radix-sort ranking can be built from.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/warp_prims tools/ubench/warp_prims.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int ITERS = 4096;

template <int MODE>
__global__ void __launch_bounds__(256) k(uint32_t* out, uint32_t seed) {
    __shared__ uint32_t sm[8][256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 8 * 256; i += 256) (&sm[0][0])[i] = 0;
    __syncthreads();
    uint32_t x = seed * 2654435761u + threadIdx.x * 40503u + blockIdx.x;
    uint32_t acc = 0;
#pragma unroll 4
    for (int it = 0; it < ITERS; it++) {
        x = x * 1664525u + 1013904223u;
        uint32_t d = (x >> 13) & 255u;
        if (MODE == 0) {  // 8 ballots
#pragma unroll
            for (int b = 0; b < 8; b++) acc += __ballot_sync(0xffffffffu, (d >> b) & 1u);
        } else if (MODE == 1) {  // match.any
            acc += __match_any_sync(0xffffffffu, d);
        } else if (MODE == 2) {  // 8 shuffles
#pragma unroll
            for (int b = 0; b < 8; b++) acc += __shfl_sync(0xffffffffu, d + b, (lane + b + 1) & 31);
        } else if (MODE == 3) {  // peers by 8 ballots + leader smem RMW + shfl (the ranking step of the sort)
            unsigned peers = 0xffffffffu;
#pragma unroll
            for (int b = 0; b < 8; b++) {
                unsigned v = __ballot_sync(0xffffffffu, (d >> b) & 1u);
                peers &= ((d >> b) & 1u) ? v : ~v;
            }
            int leader = __ffs(peers) - 1;
            uint32_t old = 0;
            if (lane == leader) { old = sm[warp][d]; sm[warp][d] = old + __popc(peers); }
            old = __shfl_sync(peers, old, leader);
            acc += old + __popc(peers & ((1u << lane) - 1u));
            __syncwarp();
        } else if (MODE == 4) {  // only the leader RMW + shfl part (peers faked = self)
            unsigned peers = 1u << lane;
            int leader = lane;
            uint32_t old = 0;
            if ((d & 1) == 0) { old = sm[warp][d]; sm[warp][d] = old + 1; }
            old = __shfl_sync(0xffffffffu, old, leader);
            acc += old + peers;
            __syncwarp();
        } else if (MODE == 5) {  // peers by byte compares against the 32 digits staged in shared memory
            uint8_t* row = reinterpret_cast<uint8_t*>(&sm[warp][0]);
            row[lane] = (uint8_t)d;
            __syncwarp();
            const uint4 a = *reinterpret_cast<const uint4*>(row), b4 = *reinterpret_cast<const uint4*>(row + 16);
            const uint32_t w[8] = {a.x, a.y, a.z, a.w, b4.x, b4.y, b4.z, b4.w};
            const uint32_t dd = d * 0x01010101u;
            unsigned peers = 0;
#pragma unroll
            for (int q = 0; q < 8; q++) {
                uint32_t v = w[q] ^ dd;
                uint32_t z = ~(((v & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | v) & 0x80808080u;  // 0x80 where bytes are equal
                peers |= (((z >> 7) * 0x00204081u) >> 21 & 0xFu) << (4 * q);
            }
            acc += peers;
            __syncwarp();
        } else if (MODE == 8) {  // as 3, but the shuffle uses the full mask (all lanes converged)
            unsigned peers = 0xffffffffu;
#pragma unroll
            for (int b = 0; b < 8; b++) {
                unsigned v = __ballot_sync(0xffffffffu, (d >> b) & 1u);
                peers &= ((d >> b) & 1u) ? v : ~v;
            }
            int leader = __ffs(peers) - 1;
            uint32_t old = 0;
            if (lane == leader) { old = sm[warp][d]; sm[warp][d] = old + __popc(peers); }
            old = __shfl_sync(0xffffffffu, old, leader);
            acc += old + __popc(peers & ((1u << lane) - 1u));
            __syncwarp();
        } else if (MODE == 9) {  // as 8, leader RMW replaced by a predicated atomic (no read-after-write chain)
            unsigned peers = 0xffffffffu;
#pragma unroll
            for (int b = 0; b < 8; b++) {
                unsigned v = __ballot_sync(0xffffffffu, (d >> b) & 1u);
                peers &= ((d >> b) & 1u) ? v : ~v;
            }
            int leader = __ffs(peers) - 1;
            uint32_t old = 0;
            if (lane == leader) old = atomicAdd(&sm[warp][d], (uint32_t)__popc(peers));
            old = __shfl_sync(0xffffffffu, old, leader);
            acc += old + __popc(peers & ((1u << lane) - 1u));
        } else if (MODE == 6) {  // 1 ballot
            acc += __ballot_sync(0xffffffffu, d & 1u);
        } else if (MODE == 7) {  // shared atomic add with return
            acc += atomicAdd(&sm[warp][d], 1u);
        }
    }
    out[blockIdx.x * 256 + threadIdx.x] = acc;
}

template <int MODE>
void run(const char* name, int ops_per_iter) {
    uint32_t* out;
    const int blocks = 148 * 2;
    cudaMalloc(&out, blocks * 256 * 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<MODE><<<blocks, 256>>>(out, 1);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MODE><<<blocks, 256>>>(out, 2);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    // warp-iterations per SM: 2 blocks x 8 warps x ITERS
    double cyc = ms * 1e-3 * 1.965e9;
    double per_iter = cyc / (16.0 * ITERS);
    printf("%-44s %8.3f ms  %7.2f SM-cycles per warp-iteration  (%d op(s): %.2f cycles each)  err=%d\n", name, ms, per_iter,
           ops_per_iter, per_iter / ops_per_iter, (int)cudaGetLastError());
    cudaFree(out);
}

int main() {
    run<6>("1 ballot", 1);
    run<0>("8 ballots", 8);
    run<1>("match.any", 1);
    run<2>("8 shuffles", 8);
    run<3>("ranking step (8 ballots + leader RMW + shfl)", 1);
    run<8>("ranking step, full-mask shfl", 1);
    run<9>("ranking step, full-mask shfl, leader atomic", 1);
    run<4>("leader RMW + shfl only", 1);
    run<5>("peers by smem byte compares", 1);
    run<7>("shared atomicAdd (random digit)", 1);
    return 0;
}
