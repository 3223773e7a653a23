// prepass.cuh -- device kernels of the Huffman pre-pass (design: huf_expand.cuh).
#pragma once
#include "decode.cuh"

namespace lzb {

struct PrepassBatch {
    const u8*  src_base;
    const u64* src_off;
    const u32* src_len;
    u32        n_units;
    PreHeader* hdr;             // zeroed before the plan kernel
    UnitPre*   pre;             // [n_units]
    HufJob*    jobs;            // [2][n_units]: literals jobs, then flags jobs
    u8*        arena;
    u64        arena_bytes;
    HufJobScratch* scratch;     // [expand grid warps][kExpJobs]
};

// warps per CTA, jobs per warp round.  One CTA per SM: 1 GiB of 128 KiB blocks is 32768 literals segments = 1024 warps,
// 7 x 148 warps take them in a single wave; 7 x 8 two-level tables are 130 KB, the rest of the SM's 256 KB stays L1 for
// the 224 bitstreams read at once.
#if !defined(LZB_EXP_CTAS_MAX)
#define LZB_EXP_CTAS_MAX 1
#endif
enum : u32 { kExpWarps = 7, kExpJobs = 8, kExpCtasMax = LZB_EXP_CTAS_MAX };   // a second CTA per SM (with 9- or 8-bit first-level tables so that two fit) was slower:
                                                      // 448 bitstreams no longer fit the L1 left beside the tables (profiles/r02_SUMMARY.md)

struct ExpWarpShared {
    HufCompact table[kExpJobs];
    u32 ring[kHufRingWords * 32];              // bitstream windows of the 32 segment decoders (decode.cuh), [word][lane]
    u32 hdr_len[kExpJobs], ok[kExpJobs];
};

struct SeqBatch {
    const u8*  src_base;
    const u64* src_off;
    const u32* src_len;
    const u32* dst_cap;
    u32        n_units;
    const UnitPre* pre;         // Huffman pre-pass results of the same batch, or null
    const u8*  arena;
    SeqHeader* hdr;             // zeroed before the kernel
    UnitSeq*   seq;             // [n_units]
    PoolRun*   recs;            // sequence records, bump allocated
    u64        recs_cap;        // in records
};

#if defined(__CUDACC__)
// Token pre-pass: one LANE per unit walks the token stream of the unit's first inner block (parse_block_lz4 /
// parse_block_lizv1, decode.cuh) and writes one record per token.  The walk is a chain of dependent loads, so the kernel
// is latency bound by construction; what makes it cheap is that all units of the batch are in flight at once.
__global__ void __launch_bounds__(32)
lizard_token_parse_kernel(SeqBatch b)
{
    const u32 unit = blockIdx.x * blockDim.x + threadIdx.x;
    if (unit >= b.n_units) return;
    UnitSeq us;
    us.off = 0; us.nseq = 0; us.state = kPreNone; us.final_lp = us.final_op = 0; us.pad[0] = us.pad[1] = 0;
    Streams st;
    int lizv1 = 0;
    const u8* src = b.src_base + b.src_off[unit];
    if (locate_first_block(src, b.src_len[unit], b.pre ? b.pre + unit : nullptr, b.arena, &st, &lizv1)) {
        const u64 at = atomicAdd(&b.hdr->cursor, (unsigned long long)st.nflags);
        if (at + st.nflags <= b.recs_cap) {
            u32 flp = 0, fop = 0;
            const bool ok = lizv1 ? parse_block_lizv1(st, 0, b.dst_cap[unit], b.recs + at, &flp, &fop)
                                  : parse_block_lz4(st, 0, b.dst_cap[unit], b.recs + at, &flp, &fop);
            if (ok) { us.off = at; us.nseq = st.nflags; us.state = kPreDone; us.final_lp = flp; us.final_op = fop; }
        }
    }
    b.seq[unit] = us;
}

// one thread per unit
__global__ void __launch_bounds__(128)
lizard_huf_plan_kernel(PrepassBatch b)
{
    const u32 unit = blockIdx.x * blockDim.x + threadIdx.x;
    if (unit >= b.n_units) return;
    UnitPre up;
    up.off[0] = up.off[1] = 0; up.state[0] = up.state[1] = kPreNone;
    HufJob jobs[2];
    const u64 base = b.src_off[unit];
    const u32 nj = plan_unit(b.src_base + base, b.src_len[unit], jobs);
    for (u32 i = 0; i < nj; ++i) {
        HufJob j = jobs[i];
        const u64 need = pre_slot_bytes(j.n);
        const u64 at = atomicAdd(&b.hdr->cursor, (unsigned long long)need);
        if (at + need > b.arena_bytes) continue;                     // arena exhausted: the token kernel expands this stream itself
        j.src += base; j.dst = at; j.unit = unit;
        const u32 idx = atomicAdd(&b.hdr->count[j.slot], 1u);
        b.jobs[(size_t)j.slot * b.n_units + idx] = j;
        up.off[j.slot] = at;
        up.state[j.slot] = kPrePlanned;
    }
    b.pre[unit] = up;
}

// persistent warps; every round a warp takes kExpJobs jobs of one kind (all literals streams first: similar sizes side by
// side keep the 32 segment decoders of a warp busy for about the same time)
__global__ void __launch_bounds__(kExpWarps * 32, kExpCtasMax)
lizard_huf_expand_kernel(PrepassBatch b)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const u32 warp = threadIdx.x >> 5, lane = WarpLanes::lane();
    ExpWarpShared* sh = reinterpret_cast<ExpWarpShared*>(smem_raw) + warp;
    HufJobScratch* ws = b.scratch + ((size_t)blockIdx.x * kExpWarps + warp) * kExpJobs + (lane & (kExpJobs - 1));
    for (u32 slot = 0; slot < 2; ++slot) {
        const u32 total = b.hdr->count[slot];
        const HufJob* const list = b.jobs + (size_t)slot * b.n_units;
        for (;;) {
            u32 first = 0;
            if (lane == 0) first = atomicAdd(&b.hdr->next[slot], (u32)kExpJobs);
            first = __shfl_sync(LZB_FULL, first, 0);
            if (first >= total) break;
            const u32 nj = total - first < kExpJobs ? total - first : kExpJobs;
            if (lane < nj) {                                          // weight headers and tables, one lane per job
                const HufJob j = list[first + lane];
                u32 h = 0;
                const bool ok = huf_job_prepare(b.src_base + j.src, j.c, j.n, &sh->table[lane], ws, &h);
                sh->hdr_len[lane] = h; sh->ok[lane] = ok ? 1u : 0u;
            }
            __syncwarp();
            const u32 jj = lane >> 2, k = lane & 3;                   // lane -> (job, segment)
            bool good = false;
            HufJob j; j.unit = 0; j.slot = slot;
            if (jj < nj) {
                j = list[first + jj];
                if (sh->ok[jj]) {
                    const u32 h = sh->hdr_len[jj];
                    good = huf_job_segment(b.arena + j.dst, j.n, b.src_base + j.src + h, j.c - h, k, sh->table[jj], sh->ring + lane);
                }
            }
            const u32 g = __ballot_sync(LZB_FULL, good);
            if (jj < nj && k == 0) b.pre[j.unit].state[slot] = ((g >> (4 * jj)) & 15u) == 15u ? kPreDone : kPreNone;
            __syncwarp();
        }
    }
}
#endif

}  // namespace lzb
