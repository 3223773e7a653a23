// encode.cuh -- device kernel around encode_core.cuh: one warp per independent unit
// (= one Lizard_compress call, normally one 128 KiB frame block), persistent grid, atomic work queue.
//
// Memory placement per warp:
//   shared : hash table when hashLog <= 14 (16 KiB at level 10/30, 64 KiB at 21/41) + 4 KiB of
//            per-segment byte histograms for the Huffman stage
//   global : the four token streams of the block being parsed (4 x 128 KiB, written once, read once,
//            L2-resident), the Huffman build scratch, and the hash table when hashLog == 18 (1 MiB)
#pragma once
#include "common.cuh"
#include "encode_core.cuh"
#include "decode.cuh"       // Progress hand-shake
#include <cuda_runtime.h>

namespace lzb {

struct EncodeBatch {
    const u8*  src_base;  const u64* src_off;  const u32* src_len;
    u8*        dst_base;  const u64* dst_off;  const u32* dst_cap;
    int*       result;    // [n] Lizard_compress return value (0 = failed / does not fit)
    u32        n_units;
    int        level;
    u8*        scratch;   // grid_warps * per_warp_bytes
    u32*       counter;
    Progress   progress;
};

struct EncodeConfig {
    int    sm_count = 0;
    size_t per_warp_small = 0;    // EncWork only
    size_t per_warp_big = 0;      // EncWork + 1 MiB table
    int    max_warps = 0;         // upper bound on resident warps (sizes the scratch)
    size_t scratch_bytes = 0;
};

constexpr u32 kEncBigTableBytes = 4u << 18;          // plain 32-bit table for hashLog 18 or multi-inner-block units
constexpr int kEncMaxWarpsPerSM = 24;

__global__ void __launch_bounds__(32, kEncMaxWarpsPerSM)
lizard_encode_units_kernel(EncodeBatch b, u32 packed_in_smem, size_t per_warp_bytes)
{
    extern __shared__ __align__(16) unsigned char enc_smem[];
    const u32 lane = threadIdx.x & 31;
    u8* my = b.scratch + (size_t)blockIdx.x * per_warp_bytes;
    EncWork* work = reinterpret_cast<EncWork*>(my);
    // shared layout: [packed hash table if the level's table fits][4 KiB segment histograms, entropy levels only]
    const LevelParams klp = level_params(b.level);
    const size_t packed_bytes = packed_in_smem ? hash_packed_bytes(klp.hashLog) : 0;
    u32* seg_hist = reinterpret_cast<u32*>(enc_smem + packed_bytes);
    HashTable packed, plain;
    packed.t32 = nullptr; packed.lo = reinterpret_cast<u16*>(enc_smem);
    packed.hi = reinterpret_cast<u32*>(enc_smem + ((size_t)2 << klp.hashLog));
    plain.t32 = reinterpret_cast<u32*>(my + sizeof(EncWork)); plain.lo = nullptr; plain.hi = nullptr;
    if (lane == 0) work->huf.seg_count = reinterpret_cast<u32 (*)[256]>(seg_hist);
    __syncwarp();
    for (;;) {
        u32 unit = 0;
        if (lane == 0) unit = atomicAdd(b.counter, 1u);
        unit = __shfl_sync(0xffffffffu, unit, 0);
        if (unit >= b.n_units) break;
        progress_wait(b.progress, unit, lane);
        const u32 len = b.src_len[unit];
        // 17-bit packed entries need every position of the unit below 2^17
        const HashTable& T = (packed_in_smem && len <= kBlockSize) ? packed : plain;
        const int r = encode_unit<WarpLanes>(b.src_base + b.src_off[unit], len,
                                             b.dst_base + b.dst_off[unit], b.dst_cap[unit], b.level, T, work);
        if (lane == 0) b.result[unit] = r;
        __syncwarp();
        progress_done(b.progress, unit, lane);
    }
}

inline size_t enc_align(size_t v) { return (v + 255) / 256 * 256; }

inline int encode_context_init(EncodeConfig& c, int sm_count, int)
{
    c.sm_count = sm_count;
    c.per_warp_small = enc_align(sizeof(EncWork) + kEncBigTableBytes);
    c.per_warp_big = c.per_warp_small;
    if (cudaFuncSetAttribute(lizard_encode_units_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             4096 + (int)hash_packed_bytes(14)) != cudaSuccess) return -1;
    c.max_warps = sm_count * kEncMaxWarpsPerSM;
    c.scratch_bytes = (size_t)c.max_warps * c.per_warp_small;
    return 0;
}

inline cudaError_t encode_launch(const EncodeConfig& c, const EncodeBatch& b, cudaStream_t s, int* launches)
{
    const LevelParams lp = level_params(b.level);
    const bool in_smem = lp.hashLog <= 14;
    const size_t smem = (lp.huffman ? 4096 : 0) + (in_smem ? hash_packed_bytes(lp.hashLog) : 0);
    const size_t per_warp = c.per_warp_small;
    int per_sm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lizard_encode_units_kernel, 32, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    if (per_sm > kEncMaxWarpsPerSM) per_sm = kEncMaxWarpsPerSM;
    int grid = c.sm_count * per_sm;
    if ((u32)grid > b.n_units) grid = (int)b.n_units;
    if ((size_t)grid * per_warp > c.scratch_bytes) grid = (int)(c.scratch_bytes / per_warp);
    lizard_encode_units_kernel<<<grid, 32, smem, s>>>(b, in_smem ? 1u : 0u, per_warp);
    *launches = 1;
    return cudaGetLastError();
}

}  // namespace lzb
