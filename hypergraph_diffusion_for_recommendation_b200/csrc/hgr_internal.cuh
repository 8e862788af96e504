// Internal helpers shared by the libhgr.so translation units (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "hgr.h"

namespace hgr {

int set_error(int code, const char *fmt, ...);
extern std::atomic<uint64_t> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

#define HGR_REQUIRE(cond, ...)                                      \
    do {                                                            \
        if (!(cond)) return hgr::set_error(HGR_ERR_INVALID, __VA_ARGS__); \
    } while (0)

#define HGR_CUDA_OK(expr)                                                                            \
    do {                                                                                             \
        cudaError_t e_ = (expr);                                                                     \
        if (e_ != cudaSuccess)                                                                       \
            return hgr::set_error(HGR_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), \
                                  __FILE__, __LINE__);                                               \
    } while (0)

#define HGR_LAUNCH_OK(name)                                                                               \
    do {                                                                                                  \
        cudaError_t e_ = cudaGetLastError();                                                              \
        if (e_ != cudaSuccess)                                                                            \
            return hgr::set_error(HGR_ERR_CUDA, "launch of %s failed: %s", name, cudaGetErrorString(e_)); \
        hgr::count_launch();                                                                              \
    } while (0)

inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- device helpers ------------------------------------------------------------------------
// streaming (read-once) loads: keep them out of L1 so gathered embedding rows stay there
__device__ __forceinline__ int ld_stream_i32(const int *p) {
    int v;
    asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ float ld_stream_f32(const float *p) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ float4 ld_stream_f4(const float4 *p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p));
    return v;
}
__device__ __forceinline__ float4 ld_ro_f4(const float4 *p) { return __ldg(p); }

// One 16-byte store to a MULTICAST address: the NVSwitch delivers it to the copy of every GPU bound to the multicast object.
__device__ __forceinline__ void st_multicast_f4(float4 *mc_addr, float4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc_addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}

template <int WIDTH>
__device__ __forceinline__ float group_sum(float v, unsigned mask) {
#pragma unroll
    for (int o = WIDTH / 2; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o, WIDTH);
    return v;
}

}  // namespace hgr
