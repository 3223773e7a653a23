// host_shim.cpp -- TEST-ONLY host build of the lane-generic codec code.
// Compiled with g++ into lizard_b200/libhostshim.so so the CPU test-suite can pin the
// __host__ __device__ code (entropy_dec.cuh / entropy_enc.cuh / encode_core.cuh, instantiated with the
// one-lane policy HostLanes) against the reference library without a GPU.  Nothing in the product
// path links or loads this file; liblizard_b200.so has no CPU code path.
#include <stdio.h>
#include <stdlib.h>
#define LZB_SHIM_CHECK(cond) do { if (!(cond)) { fprintf(stderr, "host shim check failed: %s (%s:%d)\n", #cond, __FILE__, __LINE__); abort(); } } while (0)
#include "entropy_dec.cuh"
#include "encode_core.cuh"
#include "decode.cuh"
#include "decode2.cuh"
#include <stdlib.h>

extern "C" int lzb_host_huf_decompress(unsigned char* dst, unsigned n, const unsigned char* src, unsigned c)
{
    lzb::HufDecScratch* ws = (lzb::HufDecScratch*)malloc(sizeof(lzb::HufDecScratch));
    int r = lzb::huf_decompress_serial(dst, n, src, c, ws);
    free(ws);
    return r;
}

// same choice as the device kernel: packed 17-bit entries when the unit is a single inner block and the table
// would live in shared memory, plain 32-bit entries otherwise (the buffer is big enough for either form)
// tests: 1 = plain 32-bit entries even where the packed form would do (on the device the warps of a CTA that have no
// shared-memory table run small-table levels on the plain form, with its in-entry tags)
static int g_force_plain = 0;
extern "C" void lzb_force_plain_table(int on) { g_force_plain = on; }

static lzb::HashTable make_table(lzb::u32* buf, int n, const lzb::LevelParams& lp)
{
    const lzb::u32 hash_log = lp.hashLog;
    lzb::HashTable T;
    const bool single = (lzb::u32)n <= lzb::kBlockSize;
    if (single && hash_log <= 14 && !g_force_plain) {
        T.t32 = nullptr; T.lo = (lzb::u16*)buf; T.hi = buf + ((size_t)1 << hash_log) / 2;
        T.tag = lzb::enc_tagged(lp) ? (lzb::u8*)buf + lzb::hash_packed_bytes(hash_log, false) : nullptr;   // 3.125 of the buffer's 4 bytes per entry
        T.tagged = 0;
    } else { T.t32 = buf; T.lo = nullptr; T.hi = nullptr; T.tag = nullptr; T.tagged = (single && lzb::enc_tagged_plain(lp)) ? 1u : 0u; }
    return T;
}

extern "C" int lzb_host_compress(const unsigned char* src, int n, unsigned char* dst, int cap, int level)
{
    if (n < 0 || cap < 0) return 0;
    if (level > 49) level = 49;
    if (level < 10) level = 17;
    lzb::LevelParams lp = lzb::level_params(level);
    if (lp.parser == lzb::kParserUnsupported) return 0;
    lzb::u32* table = (lzb::u32*)malloc(sizeof(lzb::u32) << lp.hashLog);
    lzb::EncWork* work = (lzb::EncWork*)malloc(sizeof(lzb::EncWork));
    work->huf.seg_count = (lzb::u32 (*)[256])malloc(4 * 256 * sizeof(lzb::u32));
    lzb::HashTable T = make_table(table, n, lp);
    int r = lzb::encode_unit<lzb::HostLanes>(src, (lzb::u32)n, dst, (lzb::u32)cap, level, T, work);
    free(work->huf.seg_count); free(table); free(work);
    return r;
}

// which schedule the decoder entry points below run (bit 0 pooled copies, bit 1 compact chain; device default 3)
static int g_dec_variant = 3;
extern "C" void lzb_set_decode_variant(int v) { g_dec_variant = v; }

// Lizard_decompress_safe through the one-lane instantiation of the device decoder
extern "C" int lzb_host_decompress(const unsigned char* src, int csize, unsigned char* dst, int cap)
{
    if (csize < 1) return 0;
    if (cap < 0) return -1;
    unsigned char* scratch = (unsigned char*)malloc(lzb::kDecScratchPerWarp);
    lzb::DecWarpShared* sh = (lzb::DecWarpShared*)malloc(sizeof(lzb::DecWarpShared));
    sh->big_table = (lzb::u16*)(scratch + 4 * lzb::kDecStreamScratch);
    int r;
    switch (g_dec_variant & 3) {
    case 0: r = lzb::decode_unit<lzb::HostLanes, 0>(src, (lzb::u32)csize, dst, (lzb::u32)cap, scratch, sh); break;
    case 1: r = lzb::decode_unit<lzb::HostLanes, 1>(src, (lzb::u32)csize, dst, (lzb::u32)cap, scratch, sh); break;
    case 2: r = lzb::decode_unit<lzb::HostLanes, 2>(src, (lzb::u32)csize, dst, (lzb::u32)cap, scratch, sh); break;
    default: r = lzb::decode_unit<lzb::HostLanes, 3>(src, (lzb::u32)csize, dst, (lzb::u32)cap, scratch, sh); break;
    }
    free(scratch); free(sh);
    return r;
}

// The Huffman pre-pass (huf_expand.cuh) as the device runs it, serially: plan the unit's first inner block, expand the
// planned streams into an arena with the same per-segment function, hand the result to the token decoder.
struct HostPre { lzb::UnitPre up; unsigned char* arena; int jobs; };
static void host_prepass(const unsigned char* src, int csize, HostPre* hp, int sabotage)
{
    hp->up.state[0] = hp->up.state[1] = lzb::kPreNone; hp->up.off[0] = hp->up.off[1] = 0;
    hp->arena = (unsigned char*)malloc(2 * (size_t)lzb::pre_slot_bytes(lzb::kBlockSize));
    lzb::HufJob jobs[2];
    const lzb::u32 nj = lzb::plan_unit(src, (lzb::u32)csize, jobs);
    hp->jobs = (int)nj;
    lzb::HufJobScratch* ws = (lzb::HufJobScratch*)malloc(sizeof(lzb::HufJobScratch));
    lzb::HufCompact* table = (lzb::HufCompact*)malloc(sizeof(lzb::HufCompact));
    size_t cursor = 0;
    for (lzb::u32 i = 0; i < nj; ++i) {
        lzb::HufJob& j = jobs[i];
        j.dst = cursor; cursor += (size_t)lzb::pre_slot_bytes(j.n);
        lzb::u32 h = 0;
        bool ok = lzb::huf_job_prepare(src + j.src, j.c, j.n, table, ws, &h);
        lzb::u32 ring[lzb::kHufRingWords];
        for (lzb::u32 k = 0; ok && k < 4; ++k)
            ok = lzb::huf_job_segment(hp->arena + j.dst, j.n, src + j.src + h, j.c - h, k, *table, ring);
        if (sabotage && ok) memset(hp->arena + j.dst, 0x5A, j.n);     // tests: proves the token decoder reads the arena
        hp->up.off[j.slot] = j.dst;
        hp->up.state[j.slot] = ok ? lzb::kPreDone : lzb::kPreNone;
    }
    free(ws); free(table);
}
// HufCompact (two-level table of the Huffman pre-pass) against the reference-layout table built from the same weights:
// number of indices (all 1 << tl, three fillings of the bits below the index) whose lookup differs.
extern "C" int lzb_huf_compact_check(const unsigned char* weights, int nsym, int tl)
{
    lzb::u32 rank_a[lzb::kHufTableLogMax + 1] = {0}, rank_b[lzb::kHufTableLogMax + 1] = {0};
    for (int s = 0; s < nsym; ++s) { rank_a[weights[s]]++; rank_b[weights[s]]++; }
    lzb::u16* full = (lzb::u16*)malloc(sizeof(lzb::u16) << tl);
    lzb::HufCompact* c = (lzb::HufCompact*)malloc(sizeof(lzb::HufCompact));
    lzb::huf_fill_dtable(full, weights, rank_a, (lzb::u32)nsym, (lzb::u32)tl);
    lzb::huf_fill_compact(c, weights, rank_b, (lzb::u32)nsym, (lzb::u32)tl);
    lzb::HufFull f; f.t = full; f.down = 32u - (lzb::u32)tl;
    int bad = 0;
    const lzb::u32 low[3] = { 0u, 0xFFFFFFFFu, 0x5A5A5A5Au };
    for (lzb::u32 idx = 0; idx < (1u << tl); ++idx)
        for (int k = 0; k < 3; ++k) {
            const lzb::u32 hi = (idx << (32 - tl)) | (low[k] >> tl);
            if (f.look(hi) != lzb::huf_view(c).look(hi)) ++bad;
        }
    free(full); free(c);
    return bad;
}

// bit 0 of `mode`: 32 emulated lanes instead of one; bit 1: overwrite the expanded streams (negative control);
// bit 2: also run the token pre-pass (one-lane parse of the first inner block into sequence records).
// *jobs_done = streams the Huffman pre-pass expanded + 16 if the token pre-pass parsed the block.
extern "C" int lzb_decompress_with_prepass(const unsigned char* src, int csize, unsigned char* dst, int cap, int mode, int* jobs_done);

extern "C" void lzb_host_token_stats(unsigned long long* fast, unsigned long long* slow)
{
#if defined(LZB_STATS)
    *fast = lzb::g_tok_fast; *slow = lzb::g_tok_slow;
#else
    *fast = 0; *slow = 0;
#endif
}

// =====================================================================================================
// 32-lane warp emulator (TEST-ONLY).  The lane-parallel code paths (ballot / shuffle / match_any based)
// only exist for 32 lanes, and there is no GPU in the build container, so the CPU suite runs them on 32
// cooperative coroutines (ucontext): every collective is a rendezvous of all lanes, exactly the
// warp-synchronous model the kernels are written in.  Shared data (hash table, streams) is ordinary memory.
// =====================================================================================================
#include <ucontext.h>
#include <vector>

namespace emu {
constexpr int kLanes = 32;
struct Warp {
    ucontext_t sched;
    ucontext_t ctx[kLanes];
    std::vector<char> stack[kLanes];
    int cur = 0;
    bool done[kLanes];
    unsigned long long slot[kLanes];
    int arrived = 0, readers = 0;
    unsigned gen = 0, gen2 = 0;
    void (*body)(void*) = nullptr;
    void* arg = nullptr;
};
static Warp* g = nullptr;
static int g_order = 0;          // 0 forward, 1 reverse, 2 shuffled per round

static void yield() { swapcontext(&g->ctx[g->cur], &g->sched); }

// all lanes deposit a value; returns when every lane has deposited; out[] is a private copy
static void exchange(unsigned long long v, unsigned long long* out)
{
    Warp* w = g;
    const int me = w->cur;
    w->slot[me] = v;
    {   unsigned my = w->gen;
        if (++w->arrived == kLanes) { w->arrived = 0; w->gen++; }
        else while (w->gen == my) yield(); }
    for (int i = 0; i < kLanes; ++i) out[i] = w->slot[i];
    {   unsigned my = w->gen2;
        if (++w->readers == kLanes) { w->readers = 0; w->gen2++; }
        else while (w->gen2 == my) yield(); }
}

static void trampoline()
{
    Warp* w = g;
    w->body(w->arg);
    w->done[w->cur] = true;
    swapcontext(&w->ctx[w->cur], &w->sched);
}

static void run(void (*body)(void*), void* arg)
{
    Warp w;
    g = &w;
    w.body = body; w.arg = arg;
    for (int i = 0; i < kLanes; ++i) {
        w.done[i] = false;
        w.stack[i].resize(256 * 1024);
        getcontext(&w.ctx[i]);
        w.ctx[i].uc_stack.ss_sp = w.stack[i].data();
        w.ctx[i].uc_stack.ss_size = w.stack[i].size();
        w.ctx[i].uc_link = &w.sched;
        makecontext(&w.ctx[i], (void (*)())trampoline, 0);
    }
    unsigned rng = 12345u;
    for (;;) {
        bool any = false;
        // lane order inside a round: forward, reverse, or shuffled -- code that is missing a barrier between a write
        // by one lane and a read by another only fails under SOME orders, so the tests run all three
        int order[kLanes];
        for (int i = 0; i < kLanes; ++i) order[i] = g_order == 1 ? kLanes - 1 - i : i;
        if (g_order == 2)
            for (int i = kLanes - 1; i > 0; --i) {
                rng = rng * 1664525u + 1013904223u;
                const int j = (int)((rng >> 8) % (unsigned)(i + 1));
                const int t = order[i]; order[i] = order[j]; order[j] = t;
            }
        for (int n = 0; n < kLanes; ++n) {
            const int i = order[n];
            if (w.done[i]) continue;
            any = true;
            w.cur = i;
            swapcontext(&w.sched, &w.ctx[i]);
        }
        if (!any) break;
    }
    g = nullptr;
}
}  // namespace emu

extern "C" void lzb_emu_lane_order(int o) { emu::g_order = o; }

struct EmuLanes {
    static constexpr bool kDevice = false;
    static constexpr lzb::u32 kLanes = emu::kLanes;
    static lzb::u32 lane() { return (lzb::u32)emu::g->cur; }
    static lzb::u32 lanes() { return emu::kLanes; }
    static void sync() { unsigned long long t[emu::kLanes]; emu::exchange(0, t); }
    static int bcast(int v) { unsigned long long t[emu::kLanes]; emu::exchange((unsigned long long)(unsigned)v, t); return (int)(unsigned)t[0]; }
    static lzb::u32 sum(lzb::u32 v) { unsigned long long t[emu::kLanes]; emu::exchange(v, t); lzb::u32 s = 0; for (auto x : t) s += (lzb::u32)x; return s; }
    static lzb::u32 excl_scan(lzb::u32 v, lzb::u32* total)
    {
        unsigned long long t[emu::kLanes]; emu::exchange(v, t);
        lzb::u32 pre = 0, tot = 0;
        for (int i = 0; i < emu::kLanes; ++i) { if (i < emu::g->cur) pre += (lzb::u32)t[i]; tot += (lzb::u32)t[i]; }
        *total = tot; return pre;
    }
    static lzb::u32 ballot(bool p)
    {
        unsigned long long t[emu::kLanes]; emu::exchange(p ? 1 : 0, t);
        lzb::u32 m = 0; for (int i = 0; i < emu::kLanes; ++i) if (t[i]) m |= 1u << i; return m;
    }
    static lzb::u32 shfl(lzb::u32 v, lzb::u32 src)
    {
        unsigned long long t[emu::kLanes]; emu::exchange(v, t); return (lzb::u32)t[src & 31];
    }
    static void prefetch(const void*) {}
    static lzb::u32 red_or(lzb::u32 v)
    {
        unsigned long long t[emu::kLanes]; emu::exchange(v, t);
        lzb::u32 m = 0; for (int i = 0; i < emu::kLanes; ++i) m |= (lzb::u32)t[i]; return m;
    }
    static lzb::u32 match_any(lzb::u32 v)
    {
        unsigned long long t[emu::kLanes]; emu::exchange(v, t);
        lzb::u32 m = 0; for (int i = 0; i < emu::kLanes; ++i) if ((lzb::u32)t[i] == v) m |= 1u << i; return m;
    }
};

struct EmuCompressArgs { const unsigned char* src; int n; unsigned char* dst; int cap; int level; lzb::HashTable T; lzb::EncWork* work; int result; };
static void emu_compress_body(void* p)
{
    EmuCompressArgs* a = (EmuCompressArgs*)p;
    int r = lzb::encode_unit<EmuLanes>(a->src, (lzb::u32)a->n, a->dst, (lzb::u32)a->cap, a->level, a->T, a->work);
    if (EmuLanes::lane() == 0) a->result = r;
}

// Lizard_compress through the 32-lane emulation of the device code path
extern "C" int lzb_emu_compress(const unsigned char* src, int n, unsigned char* dst, int cap, int level)
{
    if (n < 0 || cap < 0) return 0;
    if (level > 49) level = 49;
    if (level < 10) level = 17;
    lzb::LevelParams lp = lzb::level_params(level);
    if (lp.parser == lzb::kParserUnsupported) return 0;
    EmuCompressArgs a;
    a.src = src; a.n = n; a.dst = dst; a.cap = cap; a.level = level; a.result = 0;
    lzb::u32* table = (lzb::u32*)malloc(sizeof(lzb::u32) << lp.hashLog);
    a.T = make_table(table, n, lp);
    a.work = (lzb::EncWork*)malloc(sizeof(lzb::EncWork));
    a.work->huf.seg_count = (lzb::u32 (*)[256])malloc(4 * 256 * sizeof(lzb::u32));
    emu::run(emu_compress_body, &a);
    free(a.work->huf.seg_count); free(table); free(a.work);
    return a.result;
}

struct EmuDecompressArgs { const unsigned char* src; int csize; unsigned char* dst; int cap; unsigned char* scratch; lzb::DecWarpShared* sh; int result;
                           const lzb::UnitPre* up; const unsigned char* arena; const lzb::UnitSeq* us; const lzb::PoolRun* recs; };
static void emu_decompress_body(void* p)
{
    EmuDecompressArgs* a = (EmuDecompressArgs*)p;
    int r;
    switch (g_dec_variant & 3) {
    case 0: r = lzb::decode_unit<EmuLanes, 0>(a->src, (lzb::u32)a->csize, a->dst, (lzb::u32)a->cap, a->scratch, a->sh, a->up, a->arena, a->us, a->recs); break;
    case 1: r = lzb::decode_unit<EmuLanes, 1>(a->src, (lzb::u32)a->csize, a->dst, (lzb::u32)a->cap, a->scratch, a->sh, a->up, a->arena, a->us, a->recs); break;
    case 2: r = lzb::decode_unit<EmuLanes, 2>(a->src, (lzb::u32)a->csize, a->dst, (lzb::u32)a->cap, a->scratch, a->sh, a->up, a->arena, a->us, a->recs); break;
    default: r = lzb::decode_unit<EmuLanes, 3>(a->src, (lzb::u32)a->csize, a->dst, (lzb::u32)a->cap, a->scratch, a->sh, a->up, a->arena, a->us, a->recs); break;
    }
    if (EmuLanes::lane() == 0) a->result = r;
}

// Lizard_decompress_safe through the 32-lane emulation of the device code path
extern "C" int lzb_emu_decompress(const unsigned char* src, int csize, unsigned char* dst, int cap)
{
    if (csize < 1) return 0;
    if (cap < 0) return -1;
    EmuDecompressArgs a;
    a.src = src; a.csize = csize; a.dst = dst; a.cap = cap; a.result = -1; a.up = nullptr; a.arena = nullptr; a.us = nullptr; a.recs = nullptr;
    a.scratch = (unsigned char*)malloc(lzb::kDecScratchPerWarp);
    a.sh = (lzb::DecWarpShared*)malloc(sizeof(lzb::DecWarpShared));
    a.sh->big_table = (lzb::u16*)(a.scratch + 4 * lzb::kDecStreamScratch);
    emu::run(emu_decompress_body, &a);
    free(a.scratch); free(a.sh);
    return a.result;
}

extern "C" int lzb_decompress_with_prepass(const unsigned char* src, int csize, unsigned char* dst, int cap, int mode, int* jobs_done)
{
    if (csize < 1) return 0;
    if (cap < 0) return -1;
    HostPre hp;
    host_prepass(src, csize, &hp, mode & 2);
    if (jobs_done) *jobs_done = (hp.up.state[0] == lzb::kPreDone) + (hp.up.state[1] == lzb::kPreDone);
    lzb::UnitSeq us; us.off = 0; us.nseq = 0; us.state = lzb::kPreNone; us.final_lp = us.final_op = 0;
    lzb::PoolRun* recs = nullptr;
    if (mode & 4) {
        lzb::Streams st; int lizv1 = 0;
        if (lzb::locate_first_block(src, (lzb::u32)csize, &hp.up, hp.arena, &st, &lizv1)) {
            recs = (lzb::PoolRun*)malloc(sizeof(lzb::PoolRun) * (st.nflags + 1));
            const bool ok = lizv1 ? lzb::parse_block_lizv1(st, 0, (lzb::u32)cap, recs, &us.final_lp, &us.final_op)
                                  : lzb::parse_block_lz4(st, 0, (lzb::u32)cap, recs, &us.final_lp, &us.final_op);
            if (ok) { us.nseq = st.nflags; us.state = lzb::kPreDone; if (jobs_done) *jobs_done += 16; }
        }
    }
    unsigned char* scratch = (unsigned char*)malloc(lzb::kDecScratchPerWarp);
    lzb::DecWarpShared* sh = (lzb::DecWarpShared*)malloc(sizeof(lzb::DecWarpShared));
    sh->big_table = (lzb::u16*)(scratch + 4 * lzb::kDecStreamScratch);
    int r;
    if (mode & 1) {
        EmuDecompressArgs a;
        a.src = src; a.csize = csize; a.dst = dst; a.cap = cap; a.result = -1; a.up = &hp.up; a.arena = hp.arena;
        a.us = (mode & 4) ? &us : nullptr; a.recs = recs;
        a.scratch = scratch; a.sh = sh;
        emu::run(emu_decompress_body, &a);
        r = a.result;
    } else r = lzb::decode_unit<lzb::HostLanes, 3>(src, (lzb::u32)csize, dst, (lzb::u32)cap, scratch, sh, &hp.up, hp.arena,
                                                   (mode & 4) ? &us : nullptr, recs);
    free(scratch); free(sh); free(hp.arena); free(recs);
    return r;
}


// ---- second-generation decoder (decode2.cuh): parser + copier through the in-line sink -------------------------------
// mode bit 0: 32 emulated lanes instead of one.  `span` = most literals-stream bytes per published batch (the device uses
// ~4 KB; tests also run tiny values to exercise the prefix / split paths).  `dst` may be unaligned: the copier works in the
// aligned space of dst & ~15 and must not touch a byte outside [dst, dst + result).
template <class W> static int decode2_run(const unsigned char* src, int csize, unsigned char* dst, int cap, unsigned span,
                                          unsigned char* scratch, lzb::DecWarpCore* core, lzb::CopyShared* cs)
{
    lzb::InlineSink<W> sk;
    sk.cs = cs; sk.lits.p = nullptr; sk.nrec_total = 0; sk.span_limit = span; sk.out_pos = 0;
    sk.st.unit_lo = (lzb::u32)((size_t)dst & 15);
    sk.st.dst_al = dst - sk.st.unit_lo;
    sk.resync(sk.st.unit_lo);
    return lzb::decode_unit2<W>(src, (lzb::u32)csize, dst, (lzb::u32)cap, scratch, core, sk);
}
struct EmuDecode2Args { const unsigned char* src; int csize; unsigned char* dst; int cap; unsigned span; unsigned char* scratch;
                        lzb::DecWarpCore* core; lzb::CopyShared* cs; int result; };
static void emu_decode2_body(void* p)
{
    EmuDecode2Args* a = (EmuDecode2Args*)p;
    const int r = decode2_run<EmuLanes>(a->src, a->csize, a->dst, a->cap, a->span, a->scratch, a->core, a->cs);
    if (EmuLanes::lane() == 0) a->result = r;
}
extern "C" int lzb_host_decompress2(const unsigned char* src, int csize, unsigned char* dst, int cap, int mode, unsigned span)
{
    if (csize < 1) return 0;
    if (cap < 0) return -1;
    unsigned char* scratch = (unsigned char*)malloc(lzb::kDecScratchPerWarp);
    lzb::DecWarpCore* core = (lzb::DecWarpCore*)malloc(sizeof(lzb::DecWarpCore));
    lzb::CopyShared* cs = (lzb::CopyShared*)malloc(sizeof(lzb::CopyShared));
    memset(cs, 0xA5, sizeof *cs);
    core->big_table = (lzb::u16*)(scratch + 4 * lzb::kDecStreamScratch);
    int r;
    if (mode & 1) {
        EmuDecode2Args a; a.src = src; a.csize = csize; a.dst = dst; a.cap = cap; a.span = span; a.scratch = scratch; a.core = core;
        a.cs = cs; a.result = -1;
        emu::run(emu_decode2_body, &a);
        r = a.result;
    } else r = decode2_run<lzb::HostLanes>(src, csize, dst, cap, span, scratch, core, cs);
    free(scratch); free(core); free(cs);
    return r;
}
