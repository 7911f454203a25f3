// Micro-benchmark: 4-byte stores (and loads) at random addresses inside a window of an array, the window moving front to
// back -- what the bucketed fill of the inverse suffix array does (build_impl.inl, isa_init).  How small must the window be?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/scatter_window tools/ubench/scatter_window.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
// record j targets a pseudo-random word of window (j / window_words): every word of the array about once
__global__ void make_kernel(uint32_t* idx, uint64_t n, uint64_t window_words) {
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t w = j / window_words;
        uint64_t t = w * window_words + mix64(j) % window_words;
        idx[j] = (uint32_t)(t < n ? t : j);
    }
}
__global__ void __launch_bounds__(256) store_kernel(const uint32_t* __restrict__ idx, uint32_t* __restrict__ out, uint64_t n) {
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (uint64_t)gridDim.x * blockDim.x) out[idx[j]] = (uint32_t)j;
}
__global__ void __launch_bounds__(256) load_kernel(const uint32_t* __restrict__ idx, const uint32_t* __restrict__ in, uint32_t* __restrict__ out, uint64_t n) {
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (uint64_t)gridDim.x * blockDim.x) out[j] = in[idx[j]];
}

int main() {
    const uint64_t n = 768ull << 20;
    uint32_t *idx, *a, *b;
    CK(cudaMalloc(&idx, n * 4));
    CK(cudaMalloc(&a, n * 4));
    CK(cudaMalloc(&b, n * 4));
    CK(cudaMemset(a, 0, n * 4));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    printf("%llu M records, array of %llu MB\n", (unsigned long long)(n >> 20), (unsigned long long)(n * 4 >> 20));
    for (uint64_t window_mb : {1ull, 4ull, 16ull, 32ull, 64ull, 128ull, 3072ull}) {
        make_kernel<<<148 * 8, 256>>>(idx, n, window_mb << 18);
        for (int grid : {148 * 8, 148 * 32}) {
            float ms_s = 1e9f, ms_l = 1e9f;
            for (int rep = 0; rep < 3; rep++) {
                float t;
                CK(cudaEventRecord(e0));
                store_kernel<<<grid, 256>>>(idx, a, n);
                CK(cudaEventRecord(e1));
                CK(cudaEventSynchronize(e1));
                CK(cudaEventElapsedTime(&t, e0, e1));
                ms_s = t < ms_s ? t : ms_s;
                CK(cudaEventRecord(e0));
                load_kernel<<<grid, 256>>>(idx, a, b, n);
                CK(cudaEventRecord(e1));
                CK(cudaEventSynchronize(e1));
                CK(cudaEventElapsedTime(&t, e0, e1));
                ms_l = t < ms_l ? t : ms_l;
            }
            printf("window %5llu MB, grid %5d: stores %7.3f ms (%6.1f G/s)   loads %7.3f ms (%6.1f G/s)\n",
                   (unsigned long long)window_mb, grid, ms_s, n / ms_s / 1e6, ms_l, n / ms_l / 1e6);
        }
    }
    return 0;
}
