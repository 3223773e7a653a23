// lanes.cuh -- lane policies: the codec is written once against "W" (lane id, lane count, barrier, ballot,
// shuffle, match_any, scans).  WarpLanes = a real warp (device), HostLanes = one lane (host build, serial
// semantics), and the CPU test-suite adds EmuLanes (32 coroutines, host_shim.cpp) to run the 32-lane code paths.
#pragma once
#include "common.cuh"

namespace lzb {

struct HostLanes {
    static constexpr bool kDevice = false;
    LZ_HDM static u32 lane() { return 0; }
    LZ_HDM static u32 lanes() { return 1; }
    LZ_HDM static void sync() {}
    LZ_HDM static int bcast(int v) { return v; }
    LZ_HDM static u32 sum(u32 v) { return v; }
    LZ_HDM static u32 excl_scan(u32 v, u32* total) { *total = v; return 0; }
    LZ_HDM static u32 ballot(bool p) { return p ? 1u : 0u; }
    LZ_HDM static u32 shfl(u32 v, u32) { return v; }
    LZ_HDM static u32 match_any(u32) { return 1u; }
    LZ_HDM static void prefetch(const void*) {}
};
#if defined(__CUDACC__)
struct WarpLanes {
    static constexpr bool kDevice = true;
    __device__ __forceinline__ static u32 lane() { return threadIdx.x & 31; }
    __device__ __forceinline__ static u32 lanes() { return 32; }
    __device__ __forceinline__ static void sync() { __syncwarp(); }
    __device__ __forceinline__ static int bcast(int v) { return __shfl_sync(0xffffffffu, v, 0); }
    __device__ __forceinline__ static u32 sum(u32 v)
    {
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        return v;
    }
    __device__ __forceinline__ static u32 excl_scan(u32 v, u32* total)
    {
        u32 x = v;
        for (int o = 1; o < 32; o <<= 1) { u32 y = __shfl_up_sync(0xffffffffu, x, o); if ((threadIdx.x & 31) >= (u32)o) x += y; }
        *total = __shfl_sync(0xffffffffu, x, 31);
        return x - v;
    }
    __device__ __forceinline__ static u32 ballot(bool p) { return __ballot_sync(0xffffffffu, p); }
    __device__ __forceinline__ static u32 shfl(u32 v, u32 src) { return __shfl_sync(0xffffffffu, v, (int)src); }
    __device__ __forceinline__ static u32 match_any(u32 v) { return __match_any_sync(0xffffffffu, v); }
    __device__ __forceinline__ static void prefetch(const void* p) { asm volatile("prefetch.global.L1 [%0];" :: "l"(p)); }
};
#endif

template <class W> LZ_HD void lanes_copy(u8* dst, const u8* src, u32 n)
{
    for (u32 i = W::lane(); i < n; i += W::lanes()) dst[i] = src[i];
}

LZ_HD u32 ctz32(u32 v)
{
#if defined(__CUDA_ARCH__)
    return (u32)(__ffs((int)v) - 1);
#else
    return (u32)__builtin_ctz(v);
#endif
}

}  // namespace lzb
