// Shared declarations for the sufr_b200 CUDA library (sm_100a).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>

namespace sufr {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define SUFR_CUDA_CHECK(expr)                                                                        \
    do {                                                                                             \
        cudaError_t _e = (expr);                                                                     \
        if (_e != cudaSuccess)                                                                       \
            throw ::sufr::Error(100 + (int)_e, std::string("CUDA error ") + cudaGetErrorName(_e) +   \
                                                   " (" + cudaGetErrorString(_e) + ") at " +         \
                                                   __FILE__ + ":" + std::to_string(__LINE__) + ": " + #expr); \
    } while (0)

#define SUFR_KERNEL_CHECK() SUFR_CUDA_CHECK(cudaGetLastError())

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs (fallback when the device cannot be queried)

// SM count of the current device (queried once per device): grids are sized in multiples of it.
inline int num_sms() {
    static int sms[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) {
        cudaGetLastError();
        return kNumSMs;
    }
    if (!sms[dev]) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) sms[dev] = v;
        else { cudaGetLastError(); sms[dev] = kNumSMs; }
    }
    return sms[dev];
}

// Opt a kernel in to more than 48 KB of dynamic shared memory, once per device (the attribute is per device).
template <typename Kernel>
inline void allow_dynamic_smem(Kernel kernel, size_t bytes, bool (&done)[64]) {
    int dev = 0;
    SUFR_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !done[dev]) {
        SUFR_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        if (dev >= 0 && dev < 64) done[dev] = true;
    }
}

inline uint32_t div_up_u32(uint64_t a, uint64_t b) { return (uint32_t)((a + b - 1) / b); }
inline uint64_t div_up(uint64_t a, uint64_t b) { return (a + b - 1) / b; }

// LCP array encoding while under construction (u32):
//   kLcpPending        : element is not a group head yet (shares all words so far with its predecessor)
//   kLcpLowerBound | h : boundary created by a prefix-doubling round; true LCP is in [h, 2h)
constexpr uint32_t kLcpPending = 0xFFFFFFFFu;
//   kLcpFixup          : group boundary of the 2-bit fast path whose LCP cannot be read off the keys (a key with
//                        fill, or a neighbouring group that is still being refined): computed exactly from the
//                        final order once the refinement is done
constexpr uint32_t kLcpFixup = 0xFFFFFFFEu;
//   kLcpPendingDeep    : kLcpPending, and the group is known to agree on every symbol of the fast path's 31-symbol
//                        key (no fill in any member's key, not a large run): its exact refinement may skip key word 0
constexpr uint32_t kLcpPendingDeep = 0xFFFFFFFDu;
constexpr uint32_t kLcpLowerBound = 0x80000000u;
__host__ __device__ __forceinline__ bool lcp_is_pending(uint32_t v) { return v == kLcpPending || v == kLcpPendingDeep; }
// Group numbers of the unresolved elements carry that knowledge in their top bit ("deep" groups are one key word
// ahead of the round counter); a group number proper is below 2^31 (a group has at least two members).
constexpr uint32_t kSegDeep = 0x80000000u;

}  // namespace sufr
