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
#include <cstdlib>
#include <cmath>

namespace lzb {

// Frame packing fused into the encoder (LizardF block records, lib/lizard_frame.c:455-476): once a unit is encoded
// its warp learns where the record goes from a decoupled look-back over the units before it (units are handed out
// in ascending order, so every predecessor is running or done), writes the 4-byte block header and moves the payload
// (the compressed bytes, or the source bytes when the block did not shrink) to its final place.  The warp that
// completes a chunk publishes the chunk's end offset next to the done flag, and the host copies that byte range out.
struct FramePack {
    u8*           out;             // packed records, contiguous; null = no packing
    u64*          state;           // [n_units] look-back words, zero before launch
    volatile u64* host_chunk_end;  // mapped pinned: [n_chunks] end offset of the chunk's last record
};
#define kPackAgg (1ull << 62)
#define kPackIncl (1ull << 63)
#define kPackMask ((1ull << 62) - 1)

struct EncodeBatch {
    const u8*  src_base;  const u64* src_off;  const u32* src_len;
    u8*        dst_base;  const u64* dst_off;  const u32* dst_cap;
    int*       result;    // [n] Lizard_compress return value (0 = failed / does not fit)
    u32        n_units;
    int        level;
    u8*        scratch;   // grid_warps * per_warp_bytes
    u32*       counter;
    Progress   progress;
    FramePack  pack;
};

// byte-exact warp copy with 4-byte stores once dst is aligned (src may have any alignment)
__device__ __forceinline__ void warp_copy_words(u8* dst, const u8* src, u32 n, u32 lane)
{
    u32 head = (u32)((4 - ((size_t)dst & 3)) & 3);
    if (head > n) head = n;
    if (lane < head) dst[lane] = src[lane];
    dst += head; src += head; n -= head;
    const u32 words = n >> 2;
    const size_t sa = (size_t)src;
    const u32* sq = (const u32*)(sa & ~(size_t)3);
    const u32 sh = (u32)(sa & 3) * 8;
    u32* dq = (u32*)dst;
    if (sh == 0) { for (u32 i = lane; i < words; i += 32) dq[i] = sq[i]; }
    else {
        u32 i = lane;
        for (; i + 96 < words; i += 128) {                   // four independent words in flight per lane
            const u32 a0 = sq[i], a1 = sq[i + 1], b0 = sq[i + 32], b1 = sq[i + 33];
            const u32 c0 = sq[i + 64], c1 = sq[i + 65], d0 = sq[i + 96], d1 = sq[i + 97];
            dq[i] = __funnelshift_r(a0, a1, sh); dq[i + 32] = __funnelshift_r(b0, b1, sh);
            dq[i + 64] = __funnelshift_r(c0, c1, sh); dq[i + 96] = __funnelshift_r(d0, d1, sh);
        }
        for (; i < words; i += 32) dq[i] = __funnelshift_r(sq[i], sq[i + 1], sh);
    }
    const u32 tail = n & 3;
    if (lane < tail) dst[words * 4 + lane] = src[words * 4 + lane];
}

// Exclusive prefix of the record sizes of units [0, unit): decoupled look-back, 32 predecessors per step.
__device__ __forceinline__ u64 pack_lookback(volatile u64* state, u32 unit, u64 rec, u32 lane)
{
    if (unit == 0) {
        if (lane == 0) { __threadfence(); state[0] = kPackIncl | rec; }
        return 0;
    }
    if (lane == 0) { __threadfence(); state[unit] = kPackAgg | rec; }
    u64 excl = 0;
    long hi = (long)unit - 1;                                 // next predecessor to look at
    for (;;) {
        const long idx = hi - (long)lane;
        u64 w = 0;
        for (;;) {
            w = idx >= 0 ? state[idx] : kPackIncl;            // before unit 0: inclusive prefix 0
            if (__all_sync(0xffffffffu, (w & (kPackAgg | kPackIncl)) != 0)) break;
            __nanosleep(100);
        }
        const u32 incl = __ballot_sync(0xffffffffu, (w & kPackIncl) != 0);
        const u32 stop = incl ? (u32)(__ffs((int)incl) - 1) : 32;   // nearest predecessor with an inclusive prefix
        u64 v = lane <= stop ? (w & kPackMask) : 0;
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        excl += v;
        if (incl) break;
        hi -= 32;
    }
    if (lane == 0) { __threadfence(); state[unit] = kPackIncl | (excl + rec); }
    return excl;
}

__device__ __forceinline__ void pack_unit(const EncodeBatch& b, u32 unit, u32 len, int r, u32 lane)
{
    // payload size: a 1-byte block becomes the 6-byte raw inner block the reference produces through its wrapped
    // bound check (lizard_frame.c:459 capacity 0, lizard_compress.c:238)
    const u32 payload = len == 1 ? 6u : (r > 0 ? (u32)r : len);
    const u64 at = pack_lookback(b.pack.state, unit, 4 + (u64)payload, lane);
    u8* o = b.pack.out + at;
    const u8* src = b.src_base + b.src_off[unit];
    if (len == 1) {
        if (lane == 0) {
            o[0] = 6; o[1] = 0; o[2] = 0; o[3] = 0;
            o[4] = (u8)b.level; o[5] = (u8)kFlagRaw; o[6] = 1; o[7] = 0; o[8] = 0; o[9] = src[0];
        }
        return;
    }
    const u32 word = r > 0 ? (u32)r : (len | 0x80000000u);
    if (lane < 4) o[lane] = (u8)(word >> (8 * lane));
    warp_copy_words(o + 4, r > 0 ? b.dst_base + b.dst_off[unit] : src, payload, lane);
}

// chunk bookkeeping with packing: the warp that finishes a chunk reports where the chunk's records end
__device__ __forceinline__ void pack_done(const EncodeBatch& b, u32 unit, u32 lane)
{
    const Progress& pg = b.progress;
    if (pg.done_count && lane == 0) {
        __threadfence();
        const u32 c = unit / pg.chunk_units;
        const u32 first = c * pg.chunk_units;
        const u32 cnt = (pg.n_units - first < pg.chunk_units) ? pg.n_units - first : pg.chunk_units;
        if (atomicAdd(&pg.done_count[c], 1u) == cnt - 1) {
            __threadfence();
            volatile u64* st = b.pack.state;
            const u64 w = st[first + cnt - 1];                // inclusive: that unit finished, so it is published
            b.pack.host_chunk_end[c] = w & kPackMask;
            __threadfence_system();
            pg.host_done[c] = 1u;
        }
    }
}

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
        if (b.pack.out) { pack_unit(b, unit, len, r, lane); __syncwarp(); pack_done(b, unit, lane); }
        else progress_done(b.progress, unit, lane);
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
    {   // All units cost about the same, so the grid runs in waves and a mostly empty last wave is pure loss.
        // Measured full-wave throughput grows like (warps/SM)^0.75 (profiles/: occupancy sweep), i.e. the time of one
        // wave like p^0.25: pick the warps/SM that minimises waves x wave-time.
        int best = per_sm; double best_cost = 1e300;
        for (int p = per_sm; p >= (per_sm > 12 ? 12 : 1); --p) {
            const double waves = (double)((b.n_units + (size_t)c.sm_count * p - 1) / ((size_t)c.sm_count * p));
            const double cost = waves * pow((double)p, 0.25);
            if (cost < best_cost - 1e-9) { best_cost = cost; best = p; }
        }
        per_sm = best;
    }
    if (const char* e = getenv("LIZARDB200_ENC_WARPS_PER_SM")) {   // diagnostics: occupancy sweep
        const int v = atoi(e);
        if (v >= 1 && v <= kEncMaxWarpsPerSM) per_sm = v;
    }
    int grid = c.sm_count * per_sm;
    if ((u32)grid > b.n_units) grid = (int)b.n_units;
    if ((size_t)grid * per_warp > c.scratch_bytes) grid = (int)(c.scratch_bytes / per_warp);
    lizard_encode_units_kernel<<<grid, 32, smem, s>>>(b, in_smem ? 1u : 0u, per_warp);
    *launches = 1;
    return cudaGetLastError();
}

}  // namespace lzb
