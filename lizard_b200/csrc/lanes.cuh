// lanes.cuh -- lane policies: the codec is written once against "W" (lane id, lane count, barrier, ballot,
// shuffle, match_any, scans).  WarpLanes = a real warp (device), HostLanes = one lane (host build, serial
// semantics), and the CPU test-suite adds EmuLanes (32 coroutines, host_shim.cpp) to run the 32-lane code paths.
#pragma once
#include "common.cuh"
#if !defined(__CUDACC__)
#include <string.h>
#endif

namespace lzb {

struct HostLanes {
    static constexpr bool kDevice = false;
    static constexpr u32 kLanes = 1;
    LZ_HDM static u32 lane() { return 0; }
    LZ_HDM static u32 lanes() { return 1; }
    LZ_HDM static void sync() {}
    LZ_HDM static int bcast(int v) { return v; }
    LZ_HDM static u32 sum(u32 v) { return v; }
    LZ_HDM static u32 excl_scan(u32 v, u32* total) { *total = v; return 0; }
    LZ_HDM static u32 ballot(bool p) { return p ? 1u : 0u; }
    LZ_HDM static u32 shfl(u32 v, u32) { return v; }
    LZ_HDM static u32 match_any(u32) { return 1u; }
    LZ_HDM static u32 red_or(u32 v) { return v; }
    LZ_HDM static void prefetch(const void*) {}
};
#if defined(__CUDACC__)
struct WarpLanes {
    static constexpr bool kDevice = true;
    static constexpr u32 kLanes = 32;
    // %laneid, not threadIdx.x & 31 (the same number in these one-dimensional blocks): the kernels run at their register
    // caps and ptxas re-reads the special register at most of the ~150 use sites instead of keeping the value -- one
    // instruction per site this way, two the other
    __device__ __forceinline__ static u32 lane() { u32 l; asm("mov.u32 %0, %%laneid;" : "=r"(l)); return l; }
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
        for (int o = 1; o < 32; o <<= 1) { u32 y = __shfl_up_sync(0xffffffffu, x, o); if (lane() >= (u32)o) x += y; }
        *total = __shfl_sync(0xffffffffu, x, 31);
        return x - v;
    }
    __device__ __forceinline__ static u32 ballot(bool p) { return __ballot_sync(0xffffffffu, p); }
    __device__ __forceinline__ static u32 shfl(u32 v, u32 src) { return __shfl_sync(0xffffffffu, v, (int)src); }
    __device__ __forceinline__ static u32 match_any(u32 v) { return __match_any_sync(0xffffffffu, v); }
    __device__ __forceinline__ static u32 red_or(u32 v) { return __reduce_or_sync(0xffffffffu, v); }
    __device__ __forceinline__ static void prefetch(const void* p) { asm volatile("prefetch.global.L1 [%0];" :: "l"(p)); }
};
#endif

// store with an optional streaming (evict-first) hint: data that this kernel will not read again should not push the
// input window out of L2
template <bool kStream> LZ_HD void st_u8(u8* p, u8 v)
{
#if defined(__CUDA_ARCH__)
    if (kStream) { __stcs(p, v); return; }
#endif
    *p = v;
}
template <bool kStream> LZ_HD u8 ld_u8(const u8* p)      // read-once data: evict-first
{
#if defined(__CUDA_ARCH__)
    if (kStream) return __ldcs(p);
#endif
    return *p;
}
template <class W> LZ_HD void lanes_copy(u8* dst, const u8* src, u32 n)
{
    for (u32 i = W::lane(); i < n; i += W::lanes()) dst[i] = src[i];
}

// ---- byte-per-lane copy, four rows of lanes per pass --------------------------------------------------------
// One pass moves up to 4*lanes bytes: lane l handles bytes l, l+L, l+2L, l+3L.  All loads are issued before the
// stores, so a pass exposes one memory latency; every load/store instruction of the warp touches one contiguous
// run of bytes (1-2 sectors).  No alignment requirements, which matters because literal runs and matches start
// at arbitrary byte positions.  The source must not overlap the bytes written by the same pass.
template <class W, bool kStream = false, bool kStreamLd = false> LZ_HD void lanes_copy_rows(u8* __restrict__ dst, const u8* __restrict__ src, u32 n)
{
    const u32 l = W::lane(), L = W::lanes();
    const u8* s = src + l; u8* d = dst + l;
    if (n <= L) { if (l < n) st_u8<kStream>(d, ld_u8<kStreamLd>(s)); return; }              // most runs are shorter than one row
    if (n <= 2 * L) {
        u8 b0 = ld_u8<kStreamLd>(s), b1 = 0;
        if (l + L < n) b1 = ld_u8<kStreamLd>(s + L);
        st_u8<kStream>(d, b0);
        if (l + L < n) st_u8<kStream>(d + L, b1);
        return;
    }
    for (u32 base = 0; base < n; base += 4 * L, s += 4 * L, d += 4 * L) {
        const u32 r = n - base;                                   // bytes left; row k is live for lane l when l + k*L < r
        u8 b0 = 0, b1 = 0, b2 = 0, b3 = 0;
        if (l < r) b0 = ld_u8<kStreamLd>(s);
        if (l + L < r) b1 = ld_u8<kStreamLd>(s + L);
        if (l + 2 * L < r) b2 = ld_u8<kStreamLd>(s + 2 * L);
        if (l + 3 * L < r) b3 = ld_u8<kStreamLd>(s + 3 * L);
        if (l < r) st_u8<kStream>(d, b0);
        if (l + L < r) st_u8<kStream>(d + L, b1);
        if (l + 2 * L < r) st_u8<kStream>(d + 2 * L, b2);
        if (l + 3 * L < r) st_u8<kStream>(d + 3 * L, b3);
    }
}

// ---- several short runs at once: groups of 8 lanes --------------------------------------------------------------
// Lane group g (lanes 8g..8g+7) copies its own run of at most 4*8 bytes, all groups in the same instructions: the
// arguments are per lane (equal within a group).  One load phase, one store phase for up to lanes/8 runs.
template <class W> struct LaneGroups {
    static constexpr u32 kGroup = W::kLanes >= 8 ? 8 : W::kLanes;     // lanes per run
    static constexpr u32 kRuns = W::kLanes / kGroup;                   // runs per step
    static constexpr u32 kMaxBytes = 4 * kGroup;                       // longest run a step can take
};
template <class W, bool kStream = false, bool kStreamLd = false> LZ_HD void lanes_copy_groups(u8* __restrict__ dst, const u8* __restrict__ src, u32 n)
{
    constexpr u32 G = LaneGroups<W>::kGroup;
    const u32 o = W::lane() & (G - 1);
    const u8* s = src + o; u8* d = dst + o;
    u8 b0 = 0, b1 = 0, b2 = 0, b3 = 0;
    if (o < n) b0 = ld_u8<kStreamLd>(s);
    if (o + G < n) b1 = ld_u8<kStreamLd>(s + G);
    if (o + 2 * G < n) b2 = ld_u8<kStreamLd>(s + 2 * G);
    if (o + 3 * G < n) b3 = ld_u8<kStreamLd>(s + 3 * G);
    if (o < n) st_u8<kStream>(d, b0);
    if (o + G < n) st_u8<kStream>(d + G, b1);
    if (o + 2 * G < n) st_u8<kStream>(d + 2 * G, b2);
    if (o + 3 * G < n) st_u8<kStream>(d + 3 * G, b3);
}

// ---- 16 bytes per lane: long runs -------------------------------------------------------------------------------
// The destination is brought to a 16-byte boundary with byte stores, then every lane moves one aligned 16-byte
// store per pass (512 bytes per warp pass).  The source has an arbitrary byte phase against the destination, equal
// for all lanes: a lane reads the two aligned 16-byte vectors that straddle its chunk and realigns with funnel
// shifts.  Only vectors that contain at least one byte of the run are read, so no access leaves the pages the run
// itself occupies.  `fence` = barrier between passes (needed when the source is data an earlier pass wrote, i.e.
// a match whose offset is at least one pass + one vector, kWideMinOffset).
struct Vec16 { u32 w[4]; };
template <bool kStream = false> LZ_HD Vec16 ld_vec16(const u8* p)          // p is 16-byte aligned
{
#if defined(__CUDA_ARCH__)
    const uint4 v = kStream ? __ldcs(reinterpret_cast<const uint4*>(p)) : *reinterpret_cast<const uint4*>(p);
    Vec16 r; r.w[0] = v.x; r.w[1] = v.y; r.w[2] = v.z; r.w[3] = v.w; return r;
#else
    Vec16 r; memcpy(&r, p, 16); return r;
#endif
}
template <bool kStream = false> LZ_HD void st_vec16(u8* p, u32 a, u32 b, u32 c, u32 d)
{
#if defined(__CUDA_ARCH__)
    if (kStream) { __stcs(reinterpret_cast<uint4*>(p), make_uint4(a, b, c, d)); return; }
#if defined(LZB_EVICT_LAST_OUT)
    // experiment: output lines are what later matches read -- ask L2 to keep them in preference to streamed data
    unsigned long long pol;
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("st.global.L2::cache_hint.v4.b32 [%0], {%1, %2, %3, %4}, %5;" :: "l"(p), "r"(a), "r"(b), "r"(c), "r"(d), "l"(pol) : "memory");
    return;
#endif
    *reinterpret_cast<uint4*>(p) = make_uint4(a, b, c, d);
#else
    const u32 t[4] = { a, b, c, d }; memcpy(p, t, 16);
#endif
}
LZ_HD u32 funnel_r(u32 lo, u32 hi, u32 bits)   // bits in {0, 8, 16, 24}
{
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, bits);
#else
    return bits ? (lo >> bits) | (hi << (32 - bits)) : lo;
#endif
}
enum : u32 { kWideMinBytes = 64 };
template <class W> LZ_HD u32 wide_min_offset() { return 16 * W::kLanes + 32; }
template <class W, bool kStream = false, bool kStreamLd = false> LZ_HD void lanes_copy_wide(u8* dst, const u8* src, u32 n, bool fence)
{
    const u32 l = W::lane(), L = W::lanes();
    u32 head = (u32)((16 - ((size_t)dst & 15)) & 15);
    if (head > n) head = n;
    for (u32 i = l; i < head; i += L) st_u8<kStream>(dst + i, ld_u8<kStreamLd>(src + i));
    dst += head; src += head; n -= head;
    const u32 chunks = n >> 4;
    const size_t sa = (size_t)src;
    const u32 delta = (u32)(sa & 15);
    const u8* q = src - delta;                                  // aligned; chunk c = source bytes [16c+delta, 16c+delta+16)
    const u32 ws = delta >> 2, bs = (delta & 3) * 8;
    if (fence) W::sync();
    for (u32 base = 0; base < chunks; base += L) {
        const u32 c = base + l;
        if (c < chunks) {
            const Vec16 a = ld_vec16<kStreamLd>(q + 16 * (size_t)c);
            Vec16 b = a;
            if (delta) b = ld_vec16<kStreamLd>(q + 16 * (size_t)c + 16);   // holds byte 16c+16 <= 16c+delta+15 of the run
            u32 o0, o1, o2, o3;
            switch (ws) {
            case 0:  o0 = funnel_r(a.w[0], a.w[1], bs); o1 = funnel_r(a.w[1], a.w[2], bs); o2 = funnel_r(a.w[2], a.w[3], bs); o3 = funnel_r(a.w[3], b.w[0], bs); break;
            case 1:  o0 = funnel_r(a.w[1], a.w[2], bs); o1 = funnel_r(a.w[2], a.w[3], bs); o2 = funnel_r(a.w[3], b.w[0], bs); o3 = funnel_r(b.w[0], b.w[1], bs); break;
            case 2:  o0 = funnel_r(a.w[2], a.w[3], bs); o1 = funnel_r(a.w[3], b.w[0], bs); o2 = funnel_r(b.w[0], b.w[1], bs); o3 = funnel_r(b.w[1], b.w[2], bs); break;
            default: o0 = funnel_r(a.w[3], b.w[0], bs); o1 = funnel_r(b.w[0], b.w[1], bs); o2 = funnel_r(b.w[1], b.w[2], bs); o3 = funnel_r(b.w[2], b.w[3], bs); break;
            }
            st_vec16<kStream>(dst + 16 * (size_t)c, o0, o1, o2, o3);
        }
        if (fence && base + L < chunks) W::sync();
    }
    if (fence) W::sync();
    for (u32 i = 16 * chunks + l; i < n; i += L) st_u8<kStream>(dst + i, ld_u8<kStreamLd>(src + i));
}

LZ_HD u32 ctz32(u32 v)
{
#if defined(__CUDA_ARCH__)
    return (u32)(__ffs((int)v) - 1);
#else
    return (u32)__builtin_ctz(v);
#endif
}

LZ_HD u32 popc32(u32 v)
{
#if defined(__CUDA_ARCH__)
    return (u32)__popc(v);
#else
    return (u32)__builtin_popcount(v);
#endif
}

// ---- pooled copies: many runs, one sweep --------------------------------------------------------------------------
// The decoder hands over the runs of a whole token batch at once (one run per lane, selected by a ballot mask).  Moving
// them run by run leaves most lanes idle and exposes one memory latency per run; here the runs are packed into a small
// table in the warp's shared memory and the warp sweeps over the UNION of their pieces, every lane busy.
struct alignas(16) PoolRun { u32 a, b, c, d; };

// 16 bytes from an arbitrary address, fetched as the (one or two) aligned 16-byte vectors that hold them.  Branch free
// in the byte phase, because the lanes of a sweep work on different runs.  Every byte of [p, p+16) must be readable.
template <bool kStreamLd = false> LZ_HD Vec16 ld_chunk16(const u8* p)
{
#if defined(__CUDA_ARCH__)
    const u32 delta = (u32)((size_t)p & 15);
    const u8* q = p - delta;
    const Vec16 a = ld_vec16<kStreamLd>(q);
    Vec16 b = a;
    if (delta) b = ld_vec16<kStreamLd>(q + 16);
    const bool w2 = (delta & 8) != 0, w1 = (delta & 4) != 0;
    const u32 bs = (delta & 3) * 8;
    const u32 y0 = w2 ? a.w[2] : a.w[0], y1 = w2 ? a.w[3] : a.w[1], y2 = w2 ? b.w[0] : a.w[2],
              y3 = w2 ? b.w[1] : a.w[3], y4 = w2 ? b.w[2] : b.w[0], y5 = w2 ? b.w[3] : b.w[1];
    const u32 x0 = w1 ? y1 : y0, x1 = w1 ? y2 : y1, x2 = w1 ? y3 : y2, x3 = w1 ? y4 : y3, x4 = w1 ? y5 : y4;
    Vec16 r;
    r.w[0] = __funnelshift_r(x0, x1, bs); r.w[1] = __funnelshift_r(x1, x2, bs);
    r.w[2] = __funnelshift_r(x2, x3, bs); r.w[3] = __funnelshift_r(x3, x4, bs);
    return r;
#else
    Vec16 r; memcpy(&r, p, 16); return r;
#endif
}

// Short runs (each at most LaneGroups::kMaxBytes bytes): the lanes in `sel` pass their own (d, s, n); the runs are
// packed and moved kRuns at a time, one lane group each: dst[d, d+n) = src[s, s+n).  A run must not read what another
// run of the same call writes.  `pool` holds 32 entries; the caller orders reuse of it with a barrier.
// kStreamLd: the source is read exactly once (the literals stream): evict-first loads, so that it does not push the output
// the matches will read again out of L2
template <class W, bool kStreamLd = false> LZ_HD void pool_copy_short(u8* dst, const u8* src, u32 sel, u32 d, u32 s, u32 n, PoolRun* pool)
{
    typedef LaneGroups<W> LG;
    if (sel == 0) return;
    const u32 lane = W::lane();
    const u32 cnt = popc32(sel);
    if ((sel >> lane) & 1) {
        PoolRun e; e.a = d; e.b = s; e.c = n; e.d = 0;
        pool[popc32(sel & ((1u << lane) - 1))] = e;
    }
    W::sync();
    const u32 sub = lane / LG::kGroup;
    for (u32 i = 0; i < cnt; i += LG::kRuns) {
        const u32 k = i + sub;
        const PoolRun e = pool[k < cnt ? k : 0];
        lanes_copy_groups<W, false, kStreamLd>(dst + e.a, src + e.b, k < cnt ? e.c : 0u);
    }
}

// Long runs (each longer than 32 bytes).  A run is cut at the 16-byte boundaries of its DESTINATION: a head of < 16
// bytes, aligned 16-byte chunks, a tail of < 16 bytes.  All chunks of all runs form one index space (exclusive scan of
// the chunk counts); a sweep step gives every lane one chunk: aligned 16-byte store, source realigned from the aligned
// vectors that hold it (ld_chunk16), two steps in flight.  The lane finds its run without a search: the runs that begin
// inside the step's 32 chunks are marked in a bit mask (one warp reduction), and a population count below the lane
// gives the run's index.  Heads and tails then go two runs per step, one lane group per piece.
template <class W, bool kStreamLd = false> LZ_HD void pool_copy_long(u8* dst, const u8* src, u32 sel, u32 d, u32 s, u32 n, PoolRun* pool)
{
    typedef LaneGroups<W> LG;
    if (sel == 0) return;
    const u32 lane = W::lane(), L = W::lanes();
    const u32 cnt = popc32(sel);
    const bool mine = ((sel >> lane) & 1) != 0;
    u32 head = 0, nbody = 0;
    if (mine) {
        head = (u32)((16 - ((size_t)(dst + d) & 15)) & 15);          // n > 32 > head
        nbody = (n - head) >> 4;                                      // >= 1
    }
    u32 total = 0;
    const u32 cstart = W::excl_scan(nbody, &total);
    if (mine) {
        PoolRun e; e.a = cstart; e.b = d + head; e.c = s + head; e.d = head | (((n - head) & 15) << 4) | (nbody << 8);
        pool[popc32(sel & ((1u << lane) - 1))] = e;
    }
    W::sync();
    const bool is_run = lane < cnt;
    const u32 my_start = pool[is_run ? lane : 0].a;                  // lane r speaks for run r in the step masks
    const u32 le_mask = (2u << lane) - 1;
    for (u32 base = 0; base < total; base += 2 * L) {
        const bool two = base + L < total;
        // step 0: chunks [base, base+L); step 1: chunks [base+L, base+2L)
        const u32 r0 = popc32(W::ballot(is_run && my_start <= base)) - 1;
        const u32 f0 = W::red_or((is_run && my_start > base && my_start < base + L) ? 1u << (my_start - base) : 0u);
        u32 r1 = 0, f1 = 0;
        if (two) {
            r1 = popc32(W::ballot(is_run && my_start <= base + L)) - 1;
            f1 = W::red_or((is_run && my_start > base + L && my_start < base + 2 * L) ? 1u << (my_start - base - L) : 0u);
        }
        const u32 c0 = base + lane, c1 = base + L + lane;
        const bool v0 = c0 < total, v1 = two && c1 < total;
        Vec16 x0, x1;
        u8* to0 = dst; u8* to1 = dst;
        if (v0) {
            const PoolRun e = pool[r0 + popc32(f0 & le_mask)];
            const u32 j = 16 * (c0 - e.a);
            to0 = dst + e.b + j;
            x0 = ld_chunk16<kStreamLd>(src + e.c + j);
        }
        if (v1) {
            const PoolRun e = pool[r1 + popc32(f1 & le_mask)];
            const u32 j = 16 * (c1 - e.a);
            to1 = dst + e.b + j;
            x1 = ld_chunk16<kStreamLd>(src + e.c + j);
        }
        if (v0) st_vec16(to0, x0.w[0], x0.w[1], x0.w[2], x0.w[3]);
        if (v1) st_vec16(to1, x1.w[0], x1.w[1], x1.w[2], x1.w[3]);
    }
    // heads and tails: piece q = 2*run + (0 head | 1 tail), one lane group per piece, ceil(15 / group) rows
    constexpr u32 G = LG::kGroup, PS = LG::kRuns;
    const u32 o = lane & (G - 1), pi = lane / G;
    for (u32 q0 = 0; q0 < 2 * cnt; q0 += PS) {
        const u32 q = q0 + pi;
        const PoolRun e = pool[q < 2 * cnt ? (q >> 1) : 0];
        const u32 hd = e.d & 15, tl = (e.d >> 4) & 15, nbd = e.d >> 8;
        u32 len = 0, db, sb;
        if (q & 1) { len = tl; db = e.b + 16 * nbd; sb = e.c + 16 * nbd; }
        else { len = hd; db = e.b - hd; sb = e.c - hd; }
        if (q >= 2 * cnt) len = 0;
        const u8* sp = src + sb + o; u8* dp = dst + db + o;
        u8 b0 = 0, b1 = 0;
        if (o < len) b0 = ld_u8<kStreamLd>(sp);
        if (G < 15 && o + G < len) b1 = ld_u8<kStreamLd>(sp + G);
        if (o < len) dp[0] = b0;
        if (G < 15 && o + G < len) dp[G] = b1;
        if (2 * G < 15) for (u32 i = o + 2 * G; i < len; i += G) dp[i - o] = sp[i - o];     // groups narrower than 8 lanes (tests)
    }
}

}  // namespace lzb
