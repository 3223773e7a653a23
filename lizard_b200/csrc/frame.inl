// frame.inl -- LizardF_* frame layer (block dispatch) over the batch codec.  Included by api.cu.
//
// Reference: lib/lizard_frame.c (header :363-429, block loop :501-590, end mark :641-670, bounds :231-247,
// :436-451, decoder :756-857 and the state machine :980-1320), format doc/lizard_Frame_format.md.
// What changes: the per-block loop `while (srcEnd-srcPtr >= blockSize) LizardF_compressBlock(...)`
// (lizard_frame.c:544-549) becomes ONE batch: H2D, encode kernel, a device-side pack kernel that writes
// [LE32 size|raw flag][payload] back to back (raw fallback when a block did not fit blockSize-1, :462-466),
// D2H of the packed body.  The decoder scans the 4-byte size words on the host (serial, cheap) and decodes
// every complete block present in the call with one kernel launch, straight into the caller's dst.
//
// Only LizardF_blockIndependent frames are supported: linked blocks need the streaming dictionary API
// (Lizard_compress_continue / Lizard_decompress_safe_usingDict), which is out of scope;
// such requests return LizardF_ERROR_blockMode_invalid.

// ---- XXH32 (public algorithm; used for the header checksum byte and the optional content checksum) ----
namespace {
constexpr u32 kX1 = 2654435761u, kX2 = 2246822519u, kX3 = 3266489917u, kX4 = 668265263u, kX5 = 374761393u;
inline u32 xrotl(u32 v, int r) { return (v << r) | (v >> (32 - r)); }
inline u32 xround(u32 acc, u32 in) { return xrotl(acc + in * kX2, 13) * kX1; }
struct Xxh32 {
    u32 v[4]; u64 total = 0; u8 buf[16]; u32 nbuf = 0; u32 seed = 0;
    void reset(u32 s) { seed = s; v[0] = s + kX1 + kX2; v[1] = s + kX2; v[2] = s; v[3] = s - kX1; total = 0; nbuf = 0; }
    void update(const void* data, size_t n) {
        const u8* p = (const u8*)data; total += n;
        if (nbuf) { while (nbuf < 16 && n) { buf[nbuf++] = *p++; n--; } if (nbuf == 16) { consume(buf); nbuf = 0; } }
        while (n >= 16) { consume(p); p += 16; n -= 16; }
        while (n) { buf[nbuf++] = *p++; n--; }
    }
    void consume(const u8* p) { for (int i = 0; i < 4; ++i) v[i] = xround(v[i], rd_le32(p + 4 * i)); }
    u32 digest() const {
        u32 h = total >= 16 ? xrotl(v[0], 1) + xrotl(v[1], 7) + xrotl(v[2], 12) + xrotl(v[3], 18) : seed + kX5;
        h += (u32)total;
        u32 i = 0;
        for (; i + 4 <= nbuf; i += 4) h = xrotl(h + rd_le32(buf + i) * kX3, 17) * kX4;
        for (; i < nbuf; ++i) h = xrotl(h + buf[i] * kX5, 11) * kX1;
        h ^= h >> 15; h *= kX2; h ^= h >> 13; h *= kX3; h ^= h >> 16;
        return h;
    }
};
inline u32 xxh32(const void* p, size_t n, u32 seed) { Xxh32 x; x.reset(seed); x.update(p, n); return x.digest(); }

// XXH32 of one buffer on a helper thread (the frame layer's content checksum beside the GPU work of the same call)
constexpr size_t kHashThreadMin = 1u << 20;
struct HashJob {
    std::thread th; bool on = false;
    void start(Xxh32* x, const void* p, size_t n)
    {
        try { th = std::thread([x, p, n] { x->update(p, n); }); on = true; } catch (...) { on = false; }
    }
    bool running() const { return on; }
    void join() { if (on) { th.join(); on = false; } }
    ~HashJob() { join(); }
};

// ---- frame constants / errors (lib/lizard_frame_static.h:56-67) ----
enum : int { FE_OK = 0, FE_GENERIC, FE_maxBlockSize_invalid, FE_blockMode_invalid, FE_contentChecksumFlag_invalid,
             FE_compressionLevel_invalid, FE_headerVersion_wrong, FE_blockChecksum_unsupported, FE_reservedFlag_set,
             FE_allocation_failed, FE_srcSize_tooLarge, FE_dstMaxSize_tooSmall, FE_frameHeader_incomplete,
             FE_frameType_unknown, FE_frameSize_wrong, FE_srcPtr_wrong, FE_decompressionFailed,
             FE_headerChecksum_invalid, FE_contentChecksum_invalid, FE_maxCode };
const char* const kFrameErrorNames[] = { "OK_NoError", "ERROR_GENERIC", "ERROR_maxBlockSize_invalid", "ERROR_blockMode_invalid",
    "ERROR_contentChecksumFlag_invalid", "ERROR_compressionLevel_invalid", "ERROR_headerVersion_wrong",
    "ERROR_blockChecksum_unsupported", "ERROR_reservedFlag_set", "ERROR_allocation_failed", "ERROR_srcSize_tooLarge",
    "ERROR_dstMaxSize_tooSmall", "ERROR_frameHeader_incomplete", "ERROR_frameType_unknown", "ERROR_frameSize_wrong",
    "ERROR_srcPtr_wrong", "ERROR_decompressionFailed", "ERROR_headerChecksum_invalid", "ERROR_contentChecksum_invalid",
    "ERROR_maxCode" };
inline size_t ferr(int e) { return (size_t)-(long)e; }
constexpr u32 kFrameMagic = 0x184D2206u, kSkippableMagic = 0x184D2A50u, kRawFlag = 0x80000000u;
constexpr size_t kMinFH = 7, kMaxFH = 15, kBH = 4;

inline size_t frame_block_size(unsigned id)
{
    static const size_t sizes[7] = { 128u << 10, 256u << 10, 1u << 20, 4u << 20, 16u << 20, 64u << 20, 256u << 20 };
    if (id == 0) id = 1;
    id -= 1;
    if (id >= 7) return ferr(FE_maxBlockSize_invalid);
    return sizes[id];
}
inline LizardF_blockSizeID_t frame_optimal_bsid(LizardF_blockSizeID_t req, size_t src_size)
{
    int prop = LizardF_max128KB;
    while ((int)req > prop) {
        if (src_size <= frame_block_size((unsigned)prop)) return (LizardF_blockSizeID_t)prop;
        prop++;
    }
    return req;
}
inline void wr_le64(u8* p, u64 v) { for (int i = 0; i < 8; ++i) p[i] = (u8)(v >> (8 * i)); }
inline u64 rd_le64h(const u8* p) { u64 v = 0; for (int i = 0; i < 8; ++i) v |= (u64)p[i] << (8 * i); return v; }

// Host pipelining: ONE kernel launch covers all units (no per-chunk tail effects); the input is copied in
// chunks on a second stream, each copy followed by a 1-thread kernel that publishes how many leading units
// are resident (Progress::ready); the last unit of each chunk raises a flag in pinned host memory, upon
// which the host copies that chunk back on a third stream.  H2D, kernels and D2H overlap.  The compressor's
// block records are packed by the encode kernel itself (FramePack, encode.cuh): a separate pack kernel could
// not become resident while the persistent encode grid holds every SM's registers.
// Chunk size of the host pipeline.  A call's last chunk can only be processed after its last byte has crossed the link, so
// the call ends one chunk's worth of kernel latency after the H2D does: smaller chunks shorten that tail (and let the first
// D2H start earlier) at the price of more copy calls.  LIZARDB200_FRAME_CHUNK_MIB overrides (diagnostics / sweeps).
inline size_t frame_chunk_bytes()
{
    static const size_t v = [] {
        size_t mib = 64;
        if (const char* e = getenv("LIZARDB200_FRAME_CHUNK_MIB")) { const long t = atol(e); if (t >= 1 && t <= 1024) mib = (size_t)t; }
        return mib << 20;
    }();
    return v;
}
#define kFrameChunkBytes frame_chunk_bytes()

// LIZARDB200_TRACE=1: print a host-clock timeline of the pipelined frame paths to stderr (diagnostics only)
struct Trace {
    bool on; std::chrono::steady_clock::time_point t0;
    Trace() { const char* e = getenv("LIZARDB200_TRACE"); on = e && *e == '1'; t0 = std::chrono::steady_clock::now(); }
    void mark(const char* what, long k = -1) const
    {
        if (!on) return;
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        if (k >= 0) fprintf(stderr, "[lizard_b200 trace] %8.3f ms  %s %ld\n", ms, what, k);
        else fprintf(stderr, "[lizard_b200 trace] %8.3f ms  %s\n", ms, what);
    }
};

// wait for chunk k's completion flag; false if the compute stream died
bool wait_chunk(Context& c, volatile u32* flag)
{
    unsigned spins = 0;
    while (*flag == 0) {
        if ((++spins & 0xFFF) == 0) {
            cudaError_t q = cudaStreamQuery(c.stream);
            if (q != cudaSuccess && q != cudaErrorNotReady) { fail("frame kernel", q); return false; }
            if (q == cudaSuccess && *flag == 0) { g_last_error = "frame kernel finished without completing a chunk"; return false; }
        }
    }
    return true;
}

// How the units of a call are cut into pipeline chunks: `per_chunk` units each, optionally behind a doubling ramp (chunks
// of per_chunk/16, /8, /4, /2 units first).  The decoder's calls are bound by the copy back to the host, which can only
// start when the first chunk is done: with a 4 MiB first chunk that is ~1.4 ms earlier than with a 64 MiB one.  The
// compressor's calls end one block latency after the last input byte has arrived whatever the chunking, so they keep
// uniform chunks.  Same arithmetic as progress_chunk() on the device (decode.cuh).
struct FrameChunks {
    size_t n = 0, per_chunk = 1, ramp_unit = 0, ramp_chunks = 0;
    void plan(size_t n_units, size_t units_per_chunk, bool ramp)
    {
        n = n_units; per_chunk = units_per_chunk ? units_per_chunk : 1; ramp_unit = ramp_chunks = 0;
        if (ramp && per_chunk % 16 == 0 && n_units >= 2 * per_chunk) { ramp_unit = per_chunk / 16; ramp_chunks = 4; }
    }
    size_t ramp_total() const { return ramp_unit ? per_chunk - ramp_unit : 0; }
    size_t first(size_t k) const
    {
        if (k < ramp_chunks) return ramp_unit * (((size_t)1 << k) - 1);
        return ramp_total() + (k - ramp_chunks) * per_chunk;
    }
    size_t count(size_t k) const
    {
        const size_t f = first(k), size = k < ramp_chunks ? ramp_unit << k : per_chunk;
        return f >= n ? 0 : (n - f < size ? n - f : size);
    }
    size_t chunks() const
    {
        if (n <= ramp_total()) { size_t k = 0; while (first(k) < n) ++k; return k; }
        return ramp_chunks + (n - ramp_total() + per_chunk - 1) / per_chunk;
    }
};

struct StreamProgress {
    Progress pg; u32* d_ready; volatile u32* h_done; u32* h_ready_vals; size_t nchunks;
    FrameChunks plan;
    cudaEvent_t tables_ready = nullptr;
    ~StreamProgress() { if (tables_ready) cudaEventDestroy(tables_ready); }
    // d_progress layout: [ready][pad x3][done_count x nchunks]; flags in pinned memory
    cudaError_t init(Context& c, size_t n_units, size_t per_chunk, bool ramp = false)
    {
        plan.plan(n_units, per_chunk, ramp);
        nchunks = plan.chunks();
        cudaError_t e;
        if ((e = c.d_progress.reserve((4 + nchunks) * 4 + 64)) != cudaSuccess) return e;
        if ((e = c.pin_flags.reserve(nchunks * 8 + 64)) != cudaSuccess) return e;
        d_ready = (u32*)c.d_progress.p;
        h_done = (volatile u32*)c.pin_flags.p;
        h_ready_vals = (u32*)c.pin_flags.p + nchunks;
        for (size_t k = 0; k < nchunks; ++k) {
            h_done[k] = 0;
            h_ready_vals[k] = (u32)(plan.first(k) + plan.count(k));
        }
        if ((e = cudaMemsetAsync(c.d_progress.p, 0, (4 + nchunks) * 4, c.stream)) != cudaSuccess) return e;
        pg.ready = d_ready; pg.done_count = d_ready + 4; pg.host_done = h_done;
        pg.chunk_units = (u32)plan.per_chunk; pg.n_units = (u32)n_units;
        pg.ramp_unit = (u32)plan.ramp_unit; pg.ramp_chunks = (u32)plan.ramp_chunks;
        return cudaEventCreateWithFlags(&tables_ready, cudaEventDisableTiming);
    }
};

// Compress src[0..n) as independent blocks of block_size into the frame body at `dst` (records only).
// Returns bytes written or a frame error.  Host pointers.
size_t frame_compress_blocks(Context& c, u8* dst, size_t dst_cap, const u8* src, size_t n, size_t block_size, int level)
{
    if (n == 0) return 0;
    const size_t nblk = (n + block_size - 1) / block_size;
    size_t per_chunk = kFrameChunkBytes / block_size; if (per_chunk < 1) per_chunk = 1;
    const size_t stride = (block_size + 15) / 16 * 16;
    const size_t tab_bytes = nblk * (8 + 4 + 8 + 4);
    Trace tr; tr.mark("compress: begin");
    StreamProgress sp;
    if (sp.init(c, nblk, per_chunk) != cudaSuccess) return ferr(FE_allocation_failed);
    const size_t nchunks = sp.nchunks;
    if (c.pin_tab.reserve(tab_bytes + nchunks * 8 + 64) != cudaSuccess ||
        c.d_tab.reserve(tab_bytes + nblk * 4 + (nblk + 1) * 8 + 64) != cudaSuccess ||
        c.d_in.reserve(n + 64) != cudaSuccess || c.d_out.reserve(nblk * stride + 64) != cudaSuccess ||
        c.d_pack.reserve(n + nblk * 16 + 64) != cudaSuccess) return ferr(FE_allocation_failed);
    u64* t_in_off = (u64*)c.pin_tab.p; u64* t_out_off = t_in_off + nblk;
    u32* t_in_len = (u32*)(t_out_off + nblk); u32* t_out_cap = t_in_len + nblk;
    volatile u64* h_chunk_end = (volatile u64*)(((size_t)(t_out_cap + nblk) + 7) & ~(size_t)7);
    for (size_t i = 0; i < nblk; ++i) {
        const size_t left = n - i * block_size;
        const u32 len = (u32)(left < block_size ? left : block_size);
        t_in_off[i] = i * block_size; t_out_off[i] = i * stride; t_in_len[i] = len;
        t_out_cap[i] = len - 1;                    // lizard_frame.c:459: capacity srcSize-1, else stored raw
    }
    for (size_t k = 0; k < nchunks; ++k) h_chunk_end[k] = 0;
    u8* dtab = (u8*)c.d_tab.p;
    const u64* d_in_off = (const u64*)dtab; const u64* d_out_off = d_in_off + nblk;
    const u32* d_in_len = (const u32*)(d_out_off + nblk); const u32* d_out_cap = d_in_len + nblk;
    int* d_res = (int*)(d_out_cap + nblk);
    u64* d_state = (u64*)(((size_t)(d_res + nblk) + 7) & ~(size_t)7);

    if (cudaMemcpyAsync(dtab, c.pin_tab.p, tab_bytes, cudaMemcpyHostToDevice, c.s_in) != cudaSuccess) return ferr(FE_GENERIC);
    if (cudaMemsetAsync(d_state, 0, nblk * 8, c.s_in) != cudaSuccess) return ferr(FE_GENERIC);
    cudaEventRecord(sp.tables_ready, c.s_in);
    cudaStreamWaitEvent(c.stream, sp.tables_ready, 0);
    FramePack fp; fp.out = (u8*)c.d_pack.p; fp.state = d_state; fp.host_chunk_end = h_chunk_end;
    if (launch_encode(c, c.d_in.p, d_in_off, d_in_len, c.d_out.p, d_out_off, d_out_cap, d_res, (u32)nblk, level, c.stream, &sp.pg, &fp) != LIZARDB200_OK)
        return ferr(FE_GENERIC);
    tr.mark("compress: kernel launched");
    for (size_t k = 0; k < nchunks; ++k) {
        const size_t first = k * per_chunk, off = first * block_size;
        const size_t bytes = (k + 1 == nchunks) ? n - off : per_chunk * block_size;
        cudaMemcpyAsync((u8*)c.d_in.p + off, src + off, bytes, cudaMemcpyHostToDevice, c.s_in);
        // publish "units [0, upto) are resident" with a 4-byte copy queued behind the data copy: it runs on the copy
        // engine, so it cannot be starved by the persistent kernel occupying every SM
        cudaMemcpyAsync(sp.d_ready, &sp.h_ready_vals[k], 4, cudaMemcpyHostToDevice, c.s_in);
    }
    tr.mark("compress: H2D queued");
    size_t written = 0;
    bool too_small = false, failed = false;
    for (size_t k = 0; k < nchunks; ++k) {
        if (!wait_chunk(c, &sp.h_done[k])) { failed = true; break; }
        tr.mark("compress: chunk done", (long)k);
        const size_t end = (size_t)h_chunk_end[k];          // records of chunks 0..k occupy d_pack[0, end)
        if (end > dst_cap) { too_small = true; break; }
        if (end > written) cudaMemcpyAsync(dst + written, (u8*)c.d_pack.p + written, end - written, cudaMemcpyDeviceToHost, c.s_out);
        written = end;
    }
    cudaError_t e1 = cudaStreamSynchronize(c.s_out), e2 = cudaStreamSynchronize(c.stream), e3 = cudaStreamSynchronize(c.s_in);
    tr.mark("compress: all streams idle");
    if (failed || e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess) { if (!failed) fail("frame compress", e1 != cudaSuccess ? e1 : e2 != cudaSuccess ? e2 : e3); return ferr(FE_GENERIC); }
    if (too_small) return ferr(FE_dstMaxSize_tooSmall);
    return written;
}

struct FrameBlockRef { size_t src_pos; u32 csize; size_t dst_pos; };

// Decode `blocks` (compressed independent blocks inside src, ascending) into dst, each with capacity max_block.
// sizes_out[i] = decoded size or negative.  Host pointers.
// `hasher`: content checksum of the decoded blocks, in block order, computed on a helper thread chunk by chunk as the output
// copies land (lib/lizard_frame.c:1159, 1264-1266 update it behind every block; XXH32 is a serial recurrence, so the only
// way to take it off the critical path is to run it beside the copies and the kernel).  Null = the caller hashes.
struct ChunkHasher {
    Xxh32* x = nullptr; const std::vector<FrameBlockRef>* blocks = nullptr; const std::vector<int>* sizes = nullptr;
    const u8* dst = nullptr; FrameChunks plan; int device = 0;
    std::vector<cudaEvent_t> ev; std::atomic<size_t> recorded{0}; std::atomic<bool> stop{false}; std::thread th; bool on = false;
    bool start(size_t nchunks)
    {
        ev.assign(nchunks, nullptr);
        for (size_t k = 0; k < nchunks; ++k)
            if (cudaEventCreateWithFlags(&ev[k], cudaEventDisableTiming) != cudaSuccess) { destroy(); return false; }
        try { th = std::thread([this] { run(); }); on = true; } catch (...) { destroy(); on = false; }
        return on;
    }
    void run()
    {
        cudaSetDevice(device);
        for (size_t k = 0; k < ev.size(); ++k) {
            while (recorded.load(std::memory_order_acquire) <= k) { if (stop.load()) return; std::this_thread::yield(); }
            if (cudaEventSynchronize(ev[k]) != cudaSuccess) return;
            const size_t first = plan.first(k), last = first + plan.count(k);
            for (size_t i = first; i < last; ++i) {
                const int sz = (*sizes)[i];
                if (sz < 0) return;                                  // the caller reports the failure
                x->update(dst + (*blocks)[i].dst_pos, (size_t)sz);
            }
        }
    }
    void landed(size_t k, cudaStream_t s) { cudaEventRecord(ev[k], s); recorded.store(k + 1, std::memory_order_release); }
    void finish(bool ok) { if (!on) return; if (!ok) stop.store(true); th.join(); on = false; destroy(); }
    void destroy() { for (cudaEvent_t e : ev) if (e) cudaEventDestroy(e); ev.clear(); }
    ~ChunkHasher() { finish(false); }
};

int frame_decode_blocks(Context& c, const u8* src, size_t src_span, const std::vector<FrameBlockRef>& blocks,
                        u8* dst, size_t dst_span, u32 max_block, std::vector<int>& sizes_out, Xxh32* hasher = nullptr)
{
    const size_t n = blocks.size();
    sizes_out.assign(n, -1);
    if (n == 0) return 0;
    size_t per_chunk = kFrameChunkBytes / max_block; if (per_chunk < 1) per_chunk = 1;
    const size_t tab_bytes = n * (8 + 4 + 8 + 4);
    Trace tr; tr.mark("decode: begin");
    StreamProgress sp;
    if (sp.init(c, n, per_chunk, true) != cudaSuccess) return -1;
    const size_t nchunks = sp.nchunks;
    if (c.pin_tab.reserve(tab_bytes + n * 4 + 64) != cudaSuccess || c.d_tab.reserve(tab_bytes + n * 4 + 64) != cudaSuccess ||
        c.d_in.reserve(src_span + 256) != cudaSuccess || c.d_out.reserve(dst_span + 64) != cudaSuccess) return -1;
    u64* t_in_off = (u64*)c.pin_tab.p; u64* t_out_off = t_in_off + n;
    u32* t_in_len = (u32*)(t_out_off + n); u32* t_out_cap = t_in_len + n; volatile int* t_res = (volatile int*)(t_out_cap + n);
    for (size_t i = 0; i < n; ++i) {
        t_in_off[i] = blocks[i].src_pos; t_out_off[i] = blocks[i].dst_pos;
        t_in_len[i] = blocks[i].csize; t_out_cap[i] = max_block;
    }
    u8* dtab = (u8*)c.d_tab.p;
    const u64* d_in_off = (const u64*)dtab; const u64* d_out_off = d_in_off + n;
    const u32* d_in_len = (const u32*)(d_out_off + n); const u32* d_out_cap = d_in_len + n;
    if (cudaMemcpyAsync(dtab, c.pin_tab.p, tab_bytes, cudaMemcpyHostToDevice, c.s_in) != cudaSuccess) return -1;
    cudaEventRecord(sp.tables_ready, c.s_in);
    cudaStreamWaitEvent(c.stream, sp.tables_ready, 0);
    // results go straight to pinned host memory (4-byte posted writes, fenced before the chunk's done flag): the host
    // then needs no copy + synchronize between a chunk's flag and its output copy, and the output copies queue back to back
    if (launch_decode(c, c.d_in.p, d_in_off, d_in_len, c.d_out.p, d_out_off, d_out_cap, (int*)t_res, (u32)n, c.stream, &sp.pg) != LIZARDB200_OK) return -1;
    {   // input chunks: [first block start, last block end) widened to 128-byte lines so that a cache line shared
        // with the next chunk's first unit is final the first time an SM touches it
        size_t copied_to = 0;
        for (size_t k = 0; k < nchunks; ++k) {
            const size_t first = sp.plan.first(k), last = first + sp.plan.count(k) - 1;
            size_t lo = blocks[first].src_pos; if (lo > copied_to) lo = lo & ~(size_t)127; if (lo < copied_to) lo = copied_to;
            size_t hi = (blocks[last].src_pos + blocks[last].csize + 127) & ~(size_t)127; if (hi > src_span) hi = src_span;
            if (hi > lo) cudaMemcpyAsync((u8*)c.d_in.p + lo, src + lo, hi - lo, cudaMemcpyHostToDevice, c.s_in);
            copied_to = hi > copied_to ? hi : copied_to;
            cudaMemcpyAsync(sp.d_ready, &sp.h_ready_vals[k], 4, cudaMemcpyHostToDevice, c.s_in);
        }
    }
    tr.mark("decode: kernel launched, H2D queued");
    ChunkHasher ch;
    if (hasher) {
        ch.x = hasher; ch.blocks = &blocks; ch.sizes = &sizes_out; ch.dst = dst; ch.plan = sp.plan; ch.device = c.device;
        if (!ch.start(nchunks)) hasher = nullptr;                    // no helper: hash here, after the copies
    }
    bool failed = false;
    for (size_t k = 0; k < nchunks; ++k) {
        const size_t first = sp.plan.first(k), last = first + sp.plan.count(k) - 1;
        if (!wait_chunk(c, &sp.h_done[k])) { failed = true; break; }
        tr.mark("decode: chunk done", (long)k);
        // copy back exactly what was produced: contiguous up to the end of the last good block of the chunk
        size_t lo = blocks[first].dst_pos, hi = lo;
        for (size_t i = first; i <= last; ++i) { sizes_out[i] = t_res[i]; if (t_res[i] > 0) hi = blocks[i].dst_pos + (size_t)t_res[i]; }
        if (hi > lo) cudaMemcpyAsync(dst + lo, (u8*)c.d_out.p + lo, hi - lo, cudaMemcpyDeviceToHost, c.s_out);
        if (ch.on) ch.landed(k, c.s_out);
    }
    cudaError_t e1 = cudaStreamSynchronize(c.s_out), e2 = cudaStreamSynchronize(c.stream), e3 = cudaStreamSynchronize(c.s_in);
    tr.mark("decode: all streams idle");
    const bool ok = !(failed || e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess);
    const bool hashed = ch.on;
    ch.finish(ok);
    tr.mark("decode: checksum thread joined");
    if (!ok) return -1;
    if (hasher && !hashed)
        for (size_t i = 0; i < n; ++i) if (sizes_out[i] > 0) hasher->update(dst + blocks[i].dst_pos, (size_t)sizes_out[i]);
    return 0;
}

}  // namespace

// ================================================================================================
struct LizardF_cctx_s {
    LizardF_preferences_t prefs;
    unsigned version; unsigned stage;       // 0: idle, 1: header written
    size_t block_size;
    std::vector<u8> tmp;                    // partial block waiting for more input (lizard_frame.c tmpIn)
    u64 total_in;
    Xxh32 xxh;
};
struct LizardF_dctx_s {
    LizardF_frameInfo_t info;
    unsigned version;
    int stage;                              // see DS_* below
    u64 remaining;                          // frameRemainingSize
    size_t max_block;
    const u8* src_expect;
    std::vector<u8> tmp_in; size_t tmp_in_size, tmp_in_target;
    std::vector<u8> tmp_out; size_t tmp_out_size, tmp_out_start;
    Xxh32 xxh;
    u8 header[16];
};
namespace {
enum { DS_getHeader = 0, DS_storeHeader, DS_getCBlockSize, DS_storeCBlockSize, DS_copyDirect, DS_getCBlock,
       DS_storeCBlock, DS_flushOut, DS_getSuffix, DS_storeSuffix, DS_getSFrameSize, DS_storeSFrameSize, DS_skipSkippable };
Context* frame_context()
{
    Context& c = g_ctx[g_device];
    return &c;
}
}  // namespace

extern "C" {

unsigned LizardF_isError(size_t code) { return code > (size_t)-(long)FE_maxCode; }
const char* LizardF_getErrorName(size_t code)
{
    return LizardF_isError(code) ? kFrameErrorNames[-(int)(long)code] : "Unspecified error code";
}

size_t LizardF_compressBound(size_t srcSize, const LizardF_preferences_t* prefsPtr)
{   // lib/lizard_frame.c:436-451
    LizardF_preferences_t nul; memset(&nul, 0, sizeof nul);
    nul.frameInfo.contentChecksumFlag = LizardF_contentChecksumEnabled;
    const LizardF_preferences_t* p = prefsPtr ? prefsPtr : &nul;
    const size_t bs = frame_block_size((unsigned)p->frameInfo.blockSizeID);
    const unsigned nb = (unsigned)(srcSize / bs) + 1;
    const size_t last = p->autoFlush ? srcSize % bs : bs;
    return 4 * (size_t)nb + bs * (nb - 1) + last + 4 + (size_t)p->frameInfo.contentChecksumFlag * 4;
}

size_t LizardF_compressFrameBound(size_t srcSize, const LizardF_preferences_t* prefsPtr)
{   // lib/lizard_frame.c:231-247
    LizardF_preferences_t p;
    if (prefsPtr) p = *prefsPtr; else memset(&p, 0, sizeof p);
    p.frameInfo.blockSizeID = frame_optimal_bsid(p.frameInfo.blockSizeID, srcSize);
    p.autoFlush = 1;
    return kMaxFH + LizardF_compressBound(srcSize, &p);
}

size_t LizardF_createCompressionContext(LizardF_compressionContext_t* out, unsigned version)
{
    LizardF_cctx_s* c = new (std::nothrow) LizardF_cctx_s();
    if (!c) return ferr(FE_allocation_failed);
    memset(&c->prefs, 0, sizeof c->prefs);
    c->version = version; c->stage = 0; c->block_size = 0; c->total_in = 0;
    *out = c;
    return 0;
}
size_t LizardF_freeCompressionContext(LizardF_compressionContext_t c) { delete c; return 0; }

size_t LizardF_compressBegin(LizardF_compressionContext_t c, void* dstBuffer, size_t dstMax, const LizardF_preferences_t* prefsPtr)
try {   // lib/lizard_frame.c:363-429
    u8* const d0 = (u8*)dstBuffer; u8* d = d0;
    if (dstMax < kMaxFH) return ferr(FE_dstMaxSize_tooSmall);
    if (c->stage != 0) return ferr(FE_GENERIC);
    if (prefsPtr) c->prefs = *prefsPtr; else memset(&c->prefs, 0, sizeof c->prefs);
    if (c->prefs.frameInfo.blockSizeID == 0) c->prefs.frameInfo.blockSizeID = LizardF_max128KB;
    c->block_size = frame_block_size((unsigned)c->prefs.frameInfo.blockSizeID);
    if (LizardF_isError(c->block_size)) return c->block_size;
    if (c->prefs.frameInfo.blockMode != LizardF_blockIndependent) return ferr(FE_blockMode_invalid);
    {   // Lizard_verifyCompressionLevel, then the GPU level gate
        int lvl = c->prefs.compressionLevel;
        if (lvl > (int)kMaxLevel) lvl = kMaxLevel;
        if (lvl < (int)kMinLevel) lvl = kDefaultLevel;
        if (level_params(lvl).parser == kParserUnsupported) return ferr(FE_compressionLevel_invalid);
    }
    c->tmp.clear(); c->total_in = 0; c->xxh.reset(0);
    wr_le32(d, kFrameMagic); d += 4;
    u8* hs = d;
    *d++ = (u8)((1u << 6) + (((unsigned)c->prefs.frameInfo.blockMode & 1) << 5)
               + (((unsigned)c->prefs.frameInfo.contentChecksumFlag & 1) << 2) + ((c->prefs.frameInfo.contentSize > 0) << 3));
    *d++ = (u8)(((unsigned)c->prefs.frameInfo.blockSizeID & 7) << 4);
    if (c->prefs.frameInfo.contentSize) { wr_le64(d, c->prefs.frameInfo.contentSize); d += 8; }
    *d = (u8)(xxh32(hs, (size_t)(d - hs), 0) >> 8); d++;
    c->stage = 1;
    return (size_t)(d - d0);
} catch (const std::bad_alloc&) { return ferr(FE_allocation_failed); } catch (...) { return ferr(FE_GENERIC); }

static size_t frame_flush_tmp(LizardF_cctx_s* c, u8* dst, size_t cap)
{
    if (c->tmp.empty()) return 0;
    Context& g = *frame_context();
    int lvl = c->prefs.compressionLevel;
    if (lvl > (int)kMaxLevel) lvl = kMaxLevel;
    if (lvl < (int)kMinLevel) lvl = kDefaultLevel;
    size_t r = frame_compress_blocks(g, dst, cap, c->tmp.data(), c->tmp.size(), c->block_size, lvl);
    if (!LizardF_isError(r)) c->tmp.clear();
    return r;
}

size_t LizardF_compressUpdate(LizardF_compressionContext_t c, void* dstBuffer, size_t dstMax, const void* srcBuffer, size_t srcSize,
                              const LizardF_compressOptions_t*)
try {   // lib/lizard_frame.c:501-590 (independent blocks)
    if (c->stage != 1) return ferr(FE_GENERIC);
    if (dstMax < LizardF_compressBound(srcSize, &c->prefs)) return ferr(FE_dstMaxSize_tooSmall);
    Context& g = *frame_context();
    std::lock_guard<std::mutex> lock(g.mu);
    if (ensure_context(g, g_device) != LIZARDB200_OK) return ferr(FE_GENERIC);
    int lvl = c->prefs.compressionLevel;
    if (lvl > (int)kMaxLevel) lvl = kMaxLevel;
    if (lvl < (int)kMinLevel) lvl = kDefaultLevel;
    const u8* sp = (const u8*)srcBuffer; const u8* const se = sp + srcSize;
    u8* const d0 = (u8*)dstBuffer; u8* d = d0; u8* const de = d0 + dstMax;
    const size_t bs = c->block_size;
    // Content checksum (lib/lizard_frame.c:585-586): XXH32 over everything fed in, in order.  It is a serial recurrence
    // (one host thread, ~6 GB/s), so it runs beside the GPU work of this call instead of behind it; the call returns when
    // both are done.  Small inputs are hashed in line.
    const bool want_hash = c->prefs.frameInfo.contentChecksumFlag == LizardF_contentChecksumEnabled;
    HashJob hash_job;
    if (want_hash && srcSize >= kHashThreadMin) hash_job.start(&c->xxh, srcBuffer, srcSize);
    if (!c->tmp.empty()) {                                   // complete the pending partial block first
        size_t need = bs - c->tmp.size();
        if (need > srcSize) { c->tmp.insert(c->tmp.end(), sp, se); sp = se; }
        else {
            c->tmp.insert(c->tmp.end(), sp, sp + need); sp += need;
            size_t r = frame_flush_tmp(c, d, (size_t)(de - d));
            if (LizardF_isError(r)) { hash_job.join(); return r; }
            d += r;
        }
    }
    size_t whole = ((size_t)(se - sp) / bs) * bs;
    if (c->prefs.autoFlush) whole = (size_t)(se - sp);       // autoFlush also emits the trailing partial block
    if (whole) {
        size_t r = frame_compress_blocks(g, d, (size_t)(de - d), sp, whole, bs, lvl);
        if (LizardF_isError(r)) { hash_job.join(); return r; }
        d += r; sp += whole;
    }
    if (sp < se) c->tmp.assign(sp, se);
    if (want_hash) { if (hash_job.running()) hash_job.join(); else c->xxh.update(srcBuffer, srcSize); }
    c->total_in += srcSize;
    return (size_t)(d - d0);
} catch (const std::bad_alloc&) { return ferr(FE_allocation_failed); } catch (...) { return ferr(FE_GENERIC); }

size_t LizardF_flush(LizardF_compressionContext_t c, void* dstBuffer, size_t dstMax, const LizardF_compressOptions_t*)
try {   // lib/lizard_frame.c:601-629
    if (c->tmp.empty()) return 0;
    if (c->stage != 1) return ferr(FE_GENERIC);
    if (dstMax < c->tmp.size() + 8) return ferr(FE_dstMaxSize_tooSmall);
    Context& g = *frame_context();
    std::lock_guard<std::mutex> lock(g.mu);
    if (ensure_context(g, g_device) != LIZARDB200_OK) return ferr(FE_GENERIC);
    return frame_flush_tmp(c, (u8*)dstBuffer, dstMax);
} catch (const std::bad_alloc&) { return ferr(FE_allocation_failed); } catch (...) { return ferr(FE_GENERIC); }

size_t LizardF_compressEnd(LizardF_compressionContext_t c, void* dstBuffer, size_t dstMax, const LizardF_compressOptions_t* o)
try {   // lib/lizard_frame.c:641-670
    u8* const d0 = (u8*)dstBuffer; u8* d = d0;
    size_t r = LizardF_flush(c, dstBuffer, dstMax, o);
    if (LizardF_isError(r)) return r;
    d += r;
    wr_le32(d, 0); d += 4;
    if (c->prefs.frameInfo.contentChecksumFlag == LizardF_contentChecksumEnabled) { wr_le32(d, c->xxh.digest()); d += 4; }
    c->stage = 0;
    if (c->prefs.frameInfo.contentSize && c->prefs.frameInfo.contentSize != c->total_in) return ferr(FE_frameSize_wrong);
    return (size_t)(d - d0);
} catch (const std::bad_alloc&) { return ferr(FE_allocation_failed); } catch (...) { return ferr(FE_GENERIC); }

size_t LizardF_compressFrame(void* dstBuffer, size_t dstMax, const void* srcBuffer, size_t srcSize, const LizardF_preferences_t* prefsPtr)
try {   // lib/lizard_frame.c:260-312
    LizardF_cctx_s ctx;
    memset(&ctx.prefs, 0, sizeof ctx.prefs);
    ctx.version = LIZARDF_VERSION; ctx.stage = 0; ctx.block_size = 0; ctx.total_in = 0;
    LizardF_preferences_t p;
    if (prefsPtr) p = *prefsPtr; else memset(&p, 0, sizeof p);
    if (p.frameInfo.contentSize != 0) p.frameInfo.contentSize = (u64)srcSize;
    p.frameInfo.blockSizeID = frame_optimal_bsid(p.frameInfo.blockSizeID, srcSize);
    p.autoFlush = 1;
    if (srcSize <= frame_block_size((unsigned)p.frameInfo.blockSizeID)) p.frameInfo.blockMode = LizardF_blockIndependent;
    if (dstMax < LizardF_compressFrameBound(srcSize, &p)) return ferr(FE_dstMaxSize_tooSmall);
    u8* const d0 = (u8*)dstBuffer; u8* d = d0; u8* const de = d0 + dstMax;
    size_t r = LizardF_compressBegin(&ctx, d, dstMax, &p);
    if (LizardF_isError(r)) return r;
    d += r;
    r = LizardF_compressUpdate(&ctx, d, (size_t)(de - d), srcBuffer, srcSize, nullptr);
    if (LizardF_isError(r)) return r;
    d += r;
    r = LizardF_compressEnd(&ctx, d, (size_t)(de - d), nullptr);
    if (LizardF_isError(r)) return r;
    d += r;
    return (size_t)(d - d0);
} catch (const std::bad_alloc&) { return ferr(FE_allocation_failed); } catch (...) { return ferr(FE_GENERIC); }

// ---------------------------------------------------------------------------------------------------------
size_t LizardF_createDecompressionContext(LizardF_decompressionContext_t* out, unsigned version)
{
    LizardF_dctx_s* d = new (std::nothrow) LizardF_dctx_s();
    if (!d) return ferr(FE_GENERIC);
    memset(&d->info, 0, sizeof d->info);
    d->version = version; d->stage = DS_getHeader; d->remaining = 0; d->max_block = 0; d->src_expect = nullptr;
    d->tmp_in_size = d->tmp_in_target = d->tmp_out_size = d->tmp_out_start = 0;
    *out = d;
    return 0;
}
size_t LizardF_freeDecompressionContext(LizardF_decompressionContext_t d)
{
    size_t r = 0;
    if (d) { r = (size_t)d->stage; delete d; }
    return r;
}

static size_t frame_decode_header(LizardF_dctx_s* d, const u8* p, size_t n)
{   // lib/lizard_frame.c:756-857
    if (n < kMinFH) return ferr(FE_frameHeader_incomplete);
    memset(&d->info, 0, sizeof d->info);
    if ((rd_le32(p) & 0xFFFFFFF0u) == kSkippableMagic) {
        d->info.frameType = LizardF_skippableFrame;
        if (p == d->header) { d->tmp_in_size = n; d->tmp_in_target = 8; d->stage = DS_storeSFrameSize; return n; }
        d->stage = DS_getSFrameSize; return 4;
    }
    if (rd_le32(p) != kFrameMagic) return ferr(FE_frameType_unknown);
    d->info.frameType = LizardF_frame;
    const u8 FLG = p[4];
    const unsigned version = (FLG >> 6) & 3, block_mode = (FLG >> 5) & 1, block_cksum = (FLG >> 4) & 1,
                   csize_flag = (FLG >> 3) & 1, ccksum = (FLG >> 2) & 1;
    const size_t fh = csize_flag ? kMaxFH : kMinFH;
    if (n < fh) {
        if (p != d->header) memcpy(d->header, p, n);
        d->tmp_in_size = n; d->tmp_in_target = fh; d->stage = DS_storeHeader;
        return n;
    }
    const u8 BD = p[5];
    const unsigned bsid = (BD >> 4) & 7;
    if (version != 1) return ferr(FE_headerVersion_wrong);
    if (block_cksum) return ferr(FE_blockChecksum_unsupported);
    if (FLG & 3) return ferr(FE_reservedFlag_set);
    if (BD & 0x80) return ferr(FE_reservedFlag_set);
    if (bsid < 1) return ferr(FE_maxBlockSize_invalid);
    if (BD & 0x0F) return ferr(FE_reservedFlag_set);
    if ((u8)(xxh32(p + 4, fh - 5, 0) >> 8) != p[fh - 1]) return ferr(FE_headerChecksum_invalid);
    d->info.blockMode = (LizardF_blockMode_t)block_mode;
    d->info.contentChecksumFlag = (LizardF_contentChecksum_t)ccksum;
    d->info.blockSizeID = (LizardF_blockSizeID_t)bsid;
    d->max_block = frame_block_size(bsid);
    d->remaining = 0;
    if (csize_flag) d->remaining = d->info.contentSize = rd_le64h(p + 6);
    if (ccksum) d->xxh.reset(0);
    if (block_mode != LizardF_blockIndependent) return ferr(FE_blockMode_invalid);      // linked blocks: out of scope
    d->tmp_in.resize(d->max_block + 16);
    d->tmp_out.resize(d->max_block + 64);
    d->tmp_in_size = d->tmp_in_target = 0; d->tmp_out_size = d->tmp_out_start = 0;
    d->stage = DS_getCBlockSize;
    return fh;
}

size_t LizardF_decompress(LizardF_decompressionContext_t d, void* dstBuffer, size_t* dstSizePtr,
                          const void* srcBuffer, size_t* srcSizePtr, const LizardF_decompressOptions_t*)
try {   // lib/lizard_frame.c:980-1320, independent blocks; whole runs of complete blocks are decoded in one launch
    const u8* const s0 = (const u8*)srcBuffer; const u8* const se = s0 + *srcSizePtr; const u8* sp = s0;
    u8* const d0 = (u8*)dstBuffer; u8* const de = d0 + *dstSizePtr; u8* dp = d0;
    const u8* sel = nullptr;
    bool again = true;
    size_t hint = 1;
    *srcSizePtr = 0; *dstSizePtr = 0;
    if (d->src_expect && s0 != d->src_expect) return ferr(FE_srcPtr_wrong);
    Context& g = *frame_context();
    std::lock_guard<std::mutex> lock(g.mu);

    while (again) {
        switch (d->stage) {
        case DS_getHeader:
            if ((size_t)(se - sp) >= kMaxFH) {
                size_t h = frame_decode_header(d, sp, (size_t)(se - sp));
                if (LizardF_isError(h)) return h;
                sp += h;
                break;
            }
            d->tmp_in_size = 0; d->tmp_in_target = kMinFH; d->stage = DS_storeHeader;
            /* fallthrough */
        case DS_storeHeader: {
            size_t n = d->tmp_in_target - d->tmp_in_size;
            if (n > (size_t)(se - sp)) n = (size_t)(se - sp);
            memcpy(d->header + d->tmp_in_size, sp, n);
            d->tmp_in_size += n; sp += n;
            if (d->tmp_in_size < d->tmp_in_target) { hint = (d->tmp_in_target - d->tmp_in_size) + kBH; again = false; break; }
            size_t h = frame_decode_header(d, d->header, d->tmp_in_target);
            if (LizardF_isError(h)) return h;
            break; }
        case DS_getCBlockSize:
            if ((size_t)(se - sp) >= kBH) { sel = sp; sp += kBH; }
            else { d->tmp_in_size = 0; d->stage = DS_storeCBlockSize; }
            if (d->stage == DS_storeCBlockSize)
        case DS_storeCBlockSize: {
                size_t n = kBH - d->tmp_in_size;
                if (n > (size_t)(se - sp)) n = (size_t)(se - sp);
                memcpy(d->tmp_in.data() + d->tmp_in_size, sp, n);
                sp += n; d->tmp_in_size += n;
                if (d->tmp_in_size < kBH) { hint = kBH - d->tmp_in_size; again = false; break; }
                sel = d->tmp_in.data();
            }
            {   const u32 word = rd_le32(sel);
                const size_t csz = word & 0x7FFFFFFFu;
                if (csz == 0) { d->stage = DS_getSuffix; break; }
                if (csz > d->max_block) return ferr(FE_GENERIC);
                d->tmp_in_target = csz;
                if (word & kRawFlag) { d->stage = DS_copyDirect; break; }
                d->stage = DS_getCBlock;
                if (dp == de) { hint = csz + kBH; again = false; }
                break; }
        case DS_copyDirect: {
            size_t n = d->tmp_in_target;
            if ((size_t)(se - sp) < n) n = (size_t)(se - sp);
            if ((size_t)(de - dp) < n) n = (size_t)(de - dp);
            memcpy(dp, sp, n);
            if (d->info.contentChecksumFlag) d->xxh.update(sp, n);
            if (d->info.contentSize) d->remaining -= n;
            sp += n; dp += n;
            if (n == d->tmp_in_target) { d->stage = DS_getCBlockSize; break; }
            d->tmp_in_target -= n; hint = d->tmp_in_target + kBH; again = false;
            break; }
        case DS_getCBlock: {
            if ((size_t)(se - sp) < d->tmp_in_target) { d->tmp_in_size = 0; d->stage = DS_storeCBlock; break; }
            // ---- batch: this block and every following complete compressed block that has room in dst ----
            if (ensure_context(g, g_device) != LIZARDB200_OK) return ferr(FE_GENERIC);
            std::vector<FrameBlockRef> blocks;
            const u8* scan = sp; size_t csz = d->tmp_in_target; u8* out = dp;
            const u8* span_begin = sp;
            if ((size_t)(de - out) >= d->max_block) {
                for (;;) {
                    blocks.push_back({ (size_t)(scan - span_begin), (u32)csz, (size_t)(out - dp) });
                    scan += csz; out += d->max_block;
                    if ((size_t)(se - scan) < kBH) break;                      // next size word not here yet
                    const u32 w = rd_le32(scan);
                    const size_t nx = w & 0x7FFFFFFFu;
                    if (nx == 0 || (w & kRawFlag) || nx > d->max_block || (size_t)(se - scan - kBH) < nx) break;
                    if ((size_t)(de - out) < d->max_block) break;              // that one needs the tmp-out path
                    scan += kBH; csz = nx;                                     // take it into this batch
                }
            }
            if (blocks.empty()) {                                          // not enough room in dst: decode via tmp_out
                std::vector<FrameBlockRef> one{ { 0, (u32)d->tmp_in_target, 0 } };
                std::vector<int> sz;
                if (frame_decode_blocks(g, sp, d->tmp_in_target, one, d->tmp_out.data(), d->max_block, (u32)d->max_block, sz) != 0 || sz[0] < 0)
                    return ferr(FE_decompressionFailed);
                sp += d->tmp_in_target;
                if (d->info.contentChecksumFlag) d->xxh.update(d->tmp_out.data(), (size_t)sz[0]);
                if (d->info.contentSize) d->remaining -= (u64)sz[0];
                d->tmp_out_size = (size_t)sz[0]; d->tmp_out_start = 0; d->stage = DS_flushOut;
                break;
            }
            // blocks are decoded at max_block spacing on the device, then compacted into dst in order
            std::vector<int> sz;
            std::vector<u8>& stage_buf = d->tmp_out;
            const size_t span = (size_t)(scan - span_begin);
            const size_t out_span = blocks.size() * d->max_block;
            (void)stage_buf;
            // the content checksum of a large batch is computed beside the copies (frame_decode_blocks: ChunkHasher)
            const bool hash_beside = d->info.contentChecksumFlag && out_span >= kHashThreadMin;
            if (frame_decode_blocks(g, span_begin, span, blocks, dp, out_span, (u32)d->max_block, sz, hash_beside ? &d->xxh : nullptr) != 0)
                return ferr(FE_GENERIC);
            // full blocks land exactly in place; a short block (the last of a frame) only shifts what follows it
            u8* w = dp;
            for (size_t i = 0; i < blocks.size(); ++i) {
                if (sz[i] < 0) return ferr(FE_GENERIC);
                u8* from = dp + blocks[i].dst_pos;
                if (from != w) memmove(w, from, (size_t)sz[i]);
                if (d->info.contentChecksumFlag && !hash_beside) d->xxh.update(w, (size_t)sz[i]);
                if (d->info.contentSize) d->remaining -= (u64)sz[i];
                w += sz[i];
            }
            dp = w; sp = scan;
            d->stage = DS_getCBlockSize;
            break; }
        case DS_storeCBlock: {
            size_t n = d->tmp_in_target - d->tmp_in_size;
            if (n > (size_t)(se - sp)) n = (size_t)(se - sp);
            memcpy(d->tmp_in.data() + d->tmp_in_size, sp, n);
            d->tmp_in_size += n; sp += n;
            if (d->tmp_in_size < d->tmp_in_target) { hint = (d->tmp_in_target - d->tmp_in_size) + kBH; again = false; break; }
            if (ensure_context(g, g_device) != LIZARDB200_OK) return ferr(FE_GENERIC);
            std::vector<FrameBlockRef> one{ { 0, (u32)d->tmp_in_target, 0 } };
            std::vector<int> sz;
            const bool direct = (size_t)(de - dp) >= d->max_block;
            u8* target = direct ? dp : d->tmp_out.data();
            if (frame_decode_blocks(g, d->tmp_in.data(), d->tmp_in_target, one, target, d->max_block, (u32)d->max_block, sz) != 0 || sz[0] < 0)
                return direct ? ferr(FE_GENERIC) : ferr(FE_decompressionFailed);
            if (d->info.contentChecksumFlag) d->xxh.update(target, (size_t)sz[0]);
            if (d->info.contentSize) d->remaining -= (u64)sz[0];
            if (direct) { dp += sz[0]; d->stage = DS_getCBlockSize; }
            else { d->tmp_out_size = (size_t)sz[0]; d->tmp_out_start = 0; d->stage = DS_flushOut; }
            break; }
        case DS_flushOut: {
            size_t n = d->tmp_out_size - d->tmp_out_start;
            if (n > (size_t)(de - dp)) n = (size_t)(de - dp);
            memcpy(dp, d->tmp_out.data() + d->tmp_out_start, n);
            d->tmp_out_start += n; dp += n;
            if (d->tmp_out_start == d->tmp_out_size) { d->stage = DS_getCBlockSize; break; }
            hint = kBH; again = false;
            break; }
        case DS_getSuffix: {
            const size_t suffix = (size_t)d->info.contentChecksumFlag * 4;
            if (d->remaining) return ferr(FE_frameSize_wrong);
            if (suffix == 0) { hint = 0; d->stage = DS_getHeader; again = false; break; }
            if ((size_t)(se - sp) < 4) { d->tmp_in_size = 0; d->stage = DS_storeSuffix; }
            else { sel = sp; sp += 4; }
            }
            if (d->stage == DS_storeSuffix)
        case DS_storeSuffix: {
                size_t n = 4 - d->tmp_in_size;
                if (n > (size_t)(se - sp)) n = (size_t)(se - sp);
                memcpy(d->tmp_in.data() + d->tmp_in_size, sp, n);
                sp += n; d->tmp_in_size += n;
                if (d->tmp_in_size < 4) { hint = 4 - d->tmp_in_size; again = false; break; }
                sel = d->tmp_in.data();
            }
            {   if (rd_le32(sel) != d->xxh.digest()) return ferr(FE_contentChecksum_invalid);
                hint = 0; d->stage = DS_getHeader; again = false;
                break; }
        case DS_getSFrameSize:
            if ((size_t)(se - sp) >= 4) { sel = sp; sp += 4; }
            else { d->tmp_in_size = 4; d->tmp_in_target = 8; d->stage = DS_storeSFrameSize; }
            if (d->stage == DS_storeSFrameSize)
        case DS_storeSFrameSize: {
                size_t n = d->tmp_in_target - d->tmp_in_size;
                if (n > (size_t)(se - sp)) n = (size_t)(se - sp);
                memcpy(d->header + d->tmp_in_size, sp, n);
                sp += n; d->tmp_in_size += n;
                if (d->tmp_in_size < d->tmp_in_target) { hint = d->tmp_in_target - d->tmp_in_size; again = false; break; }
                sel = d->header + 4;
            }
            {   const size_t sf = rd_le32(sel);
                d->info.contentSize = sf; d->tmp_in_target = sf; d->stage = DS_skipSkippable;
                break; }
        case DS_skipSkippable: {
            size_t n = d->tmp_in_target;
            if (n > (size_t)(se - sp)) n = (size_t)(se - sp);
            sp += n; d->tmp_in_target -= n;
            again = false; hint = d->tmp_in_target;
            if (hint) break;
            d->stage = DS_getHeader;
            break; }
        }
    }
    d->src_expect = sp < se ? sp : nullptr;
    *srcSizePtr = (size_t)(sp - s0);
    *dstSizePtr = (size_t)(dp - d0);
    return hint;
} catch (const std::bad_alloc&) { return ferr(FE_allocation_failed); } catch (...) { return ferr(FE_GENERIC); }

size_t LizardF_getFrameInfo(LizardF_decompressionContext_t d, LizardF_frameInfo_t* info, const void* srcBuffer, size_t* srcSizePtr)
try {   // lib/lizard_frame.c:870-893
    if (d->stage > DS_storeHeader) {
        size_t o = 0, i = 0;
        *srcSizePtr = 0; *info = d->info;
        return LizardF_decompress(d, nullptr, &o, nullptr, &i, nullptr);
    }
    const u8* p = (const u8*)srcBuffer;
    size_t hsize;
    if (*srcSizePtr < 5) { *srcSizePtr = 0; return ferr(FE_frameHeader_incomplete); }
    if ((rd_le32(p) & 0xFFFFFFF0u) == kSkippableMagic) hsize = 8;
    else if (rd_le32(p) != kFrameMagic) { *srcSizePtr = 0; return ferr(FE_frameType_unknown); }
    else hsize = ((p[4] >> 3) & 1) ? kMaxFH : kMinFH;
    if (*srcSizePtr < hsize) { *srcSizePtr = 0; return ferr(FE_frameHeader_incomplete); }
    *srcSizePtr = hsize;
    size_t o = 0;
    size_t next = LizardF_decompress(d, nullptr, &o, srcBuffer, srcSizePtr, nullptr);
    if (d->stage <= DS_storeHeader) return ferr(FE_frameHeader_incomplete);
    *info = d->info;
    return next;
} catch (const std::bad_alloc&) { return ferr(FE_allocation_failed); } catch (...) { return ferr(FE_GENERIC); }


// diagnostics (no device needed): the pipeline chunk of unit `unit` in a call of nUnits units cut into chunks of
// unitsPerChunk (ramp != 0: the decoder's doubling ramp in front), computed the host's way (FrameChunks) and the kernels' way
// (progress_chunk); returns the number of chunks, or -1 if the two disagree about the unit's chunk, its first unit or its size.
int LizardB200_chunkPlan(unsigned nUnits, unsigned unitsPerChunk, int ramp, unsigned unit, unsigned* chunk, unsigned* first, unsigned* count)
{
    FrameChunks fc; fc.plan(nUnits, unitsPerChunk, ramp != 0);
    Progress pg; memset(&pg, 0, sizeof pg);
    pg.chunk_units = (u32)fc.per_chunk; pg.n_units = nUnits; pg.ramp_unit = (u32)fc.ramp_unit; pg.ramp_chunks = (u32)fc.ramp_chunks;
    const size_t nch = fc.chunks();
    if (unit >= nUnits) return (int)nch;
    u32 f = 0, cnt = 0;
    const u32 c = progress_chunk(pg, unit, &f, &cnt);
    if (chunk) *chunk = c;
    if (first) *first = f;
    if (count) *count = cnt;
    if (c >= nch || fc.first(c) != f || fc.count(c) != cnt || unit < f || unit >= f + cnt) return -1;
    return (int)nch;
}

}  // extern "C"
