// Device zstd inflate, fourth generation (device-only; same contract as zstd_dec.h / zstd_plan2.h: kJobOk means byte-identical
// to libzstd, everything else goes back to the host).
//
// What the third generation (zstd_plan2.h) left on the table, measured on level-3 ciphertext frames (raw literals, 16,354
// sequences of ~4 literal + ~4 match bytes, three 512/256/512-cell FSE tables per frame):
//   * the sequence chain read its FSE tables from global memory at 0.5 MB per lane stride (three dependent L2 round trips per
//     sequence) with 32 frames diverging inside one warp;
//   * execution was one THREAD per frame: 16 k sequences one after the other although only the match copies depend on each other.
// Here:
//   k_zd3_seq   four frames per warp (a one-lane warp costs as many issue slots as a full one, and shared memory - 5 KB of FSE
//               tables per frame as 4-byte cells - bounds the frames per SM either way): each frame's tables are staged in shared
//               memory, the bit reader keeps the next stream word prefetched, a sequence's six bit fields come out of one register
//   k_zd3_exec  one CTA (1,024 threads) per frame, the frame's output assembled in shared memory: prefix sums give every
//               sequence its literal and output positions, literals are copied in parallel, pointer jumping moves every match's
//               source back along its chain of pending matches (zstd_exec3.h), and the copies go through a per-byte "pending"
//               bitmap (a match is copied once none of its source bytes is pending; dependencies point backwards only, so the
//               earliest pending match is always ready); the payload leaves with 16-byte stores
// Parsing (headers, table construction) and the Huffman streams stay zstd_plan2.h's plan2_parse / plan2_huf.
#pragma once
#include "codec_kernels.h"
#include "zstd_plan2.h"

namespace fheb {
namespace zd3 {

using zd::Block2;
using zd::Plan2;
using zd::Tables2;

constexpr int kSeqFrames = 4;                        // frames per warp of k_zd3_seq: lanes 0, 8, 16, 24 walk one chain each
constexpr int kTabCells = 512 + 256 + 512;           // LL (log <= 9), OF (<= 8), ML (<= 9)
constexpr int kLutCells = 36 + 53;                   // base values of the literal-length and match-length codes (one copy per CTA)
constexpr size_t kSeqSmem = ((size_t)kSeqFrames * kTabCells + kLutCells) * sizeof(uint32_t);  // 20.4 KB per one-warp CTA: 11 CTAs per SM

// A table cell in shared memory: next state base (10 bits) | state bits << 10 (4) | value bits << 14 (5) | code << 19 (6).  The
// value's BASE is a function of the code alone (RFC 8878 3.1.1.3.2.1.1) and is not on the chain that links one sequence to the
// next - only the bit counts are - so it comes from a small per-CTA table after the fact and a cell is 4 bytes instead of 8.
__device__ __forceinline__ uint32_t pack_cell(uint2 g, int which) {  // g: a SeqEntry as stored by the parse step
    const uint32_t next_base = g.x & 0xFFFF, nb = (g.x >> 16) & 0xFF, add = g.x >> 24, base = g.y;
    uint32_t code;
    if (which == 1) {
        code = add;  // offset code: base = 1 << code
    } else if (which == 0) {
        code = add == 0 ? base : add == 1 ? 16 + (base - 16) / 2 : add == 2 ? 20 + (base - 24) / 4 : add == 3 ? 22 + (base - 32) / 8
               : add == 4 ? 24u : add == 6 ? 25u : add + 19;
    } else {
        code = add == 0 ? base - 3 : add == 1 ? 32 + (base - 35) / 2 : add == 2 ? 36 + (base - 43) / 4 : add == 3 ? 38 + (base - 51) / 8
               : add == 4 ? 40 + (base - 67) / 16 : add == 5 ? 42u : add == 7 ? 43u : add + 36;
    }
    return next_base | nb << 10 | add << 14 | (code & 63) << 19;
}

__device__ __forceinline__ uint32_t shl32(uint32_t v, uint32_t n) {  // PTX shifts clamp at 32: a shift by 32 yields 0
    uint32_t r;
    asm("shl.b32 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(n));
    return r;
}
__device__ __forceinline__ uint32_t shr32(uint32_t v, uint32_t n) {
    uint32_t r;
    asm("shr.b32 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(n));
    return r;
}

// backward bit reader on aligned 32-bit words (zstd_plan2.h Back32) with the next word already in a register
struct Back32P {
    const uint32_t *wp, *w0;
    uint32_t lo_bits, nextw, nextmask;  // the prefetched word is masked when it is CONSUMED: nothing waits for the load before
    uint64_t c;   // unread bits, left-aligned
    int avail;    // how many of them are valid
    int left;     // stream bits not yet consumed (negative after an over-read; a stream is < 2^20 bits)
    __device__ __forceinline__ void fetch() {
        nextw = 0, nextmask = 0;
        if (wp >= w0) {
            nextw = __ldg(wp);
            nextmask = wp == w0 ? ~((1u << lo_bits) - 1) : 0xFFFFFFFFu;
        }
    }
    __device__ __forceinline__ bool init(const uint8_t *base, size_t len) {
        if (len == 0 || len > (1u << 17)) return false;
        const uint8_t last = base[len - 1];
        if (last == 0) return false;
        const uintptr_t a = (uintptr_t)base;
        w0 = (const uint32_t *)(a & ~(uintptr_t)3);
        lo_bits = (uint32_t)(a & 3) * 8;
        left = (int)(len - 1) * 8 + (31 - __clz((int)last));
        const uint32_t abs_top = lo_bits + (uint32_t)left;
        const uint32_t r = abs_top & 31;
        wp = w0 + (abs_top >> 5);
        c = 0;
        avail = 0;
        if (r) {
            uint32_t w = __ldg(wp) & ((1u << r) - 1);
            if (wp == w0 && lo_bits) w &= ~((1u << lo_bits) - 1);
            c = (uint64_t)w << (64 - r);
            avail = (int)r;
        }
        wp--;
        fetch();
        return true;
    }
    __device__ __forceinline__ void refill() {  // afterwards avail >= 32
        if (avail < 32) {
            c |= (uint64_t)(nextw & nextmask) << (32 - avail);
            avail += 32;
            wp--;
            fetch();
        }
    }
    __device__ __forceinline__ uint32_t read(int n) {  // n in [0, 32], n <= avail
        const uint32_t v = (uint32_t)((c >> 1) >> (63 - n));
        c <<= n;
        avail -= n;
        left -= n;
        return v;
    }
};

// The sequences of one block: plan2_seq's results and verdict.  Its per-sequence checks are monotone (bits left, literal
// total) or independent of the other sequences (value ranges), so they are accumulated and judged once after the loop; a
// sequence whose six bit fields fit the 32 bits a single refill guarantees - all but a handful - takes them from one register.
__device__ __forceinline__ bool seq_block(const uint8_t *src, const Block2 &bp, const uint32_t *tll, const uint32_t *tof,
                                          const uint32_t *tml, const uint32_t *lut, uint64_t *so, uint32_t &r0_, uint32_t &r1_,
                                          uint32_t &r2_) {
    Back32P b;
    if (!b.init(src + bp.bs_off, bp.bs_len)) return false;
    b.refill();
    uint32_t st_ll = b.read((int)bp.ll_log), st_of = b.read((int)bp.of_log), st_ml = b.read((int)bp.ml_log);
    if (b.left < 0) return false;
    const uint32_t nseq = bp.nseq, regen = bp.regen;
    uint32_t r0 = r0_, r1 = r1_, r2 = r2_;
    uint32_t lsum = 0, badacc = 0;  // (lsum: < 2^18 per term, < 2^16 terms)
    for (uint32_t i = 0; i < nseq; i++) {
        const uint32_t el = tll[st_ll], eo = tof[st_of], em = tml[st_ml];
        const uint32_t keep = i + 1 < nseq ? 0xFu : 0u;  // the last sequence reads no state bits
        const uint32_t a_of = (eo >> 14) & 31, a_ml = (em >> 14) & 31, a_ll = (el >> 14) & 31;
        const uint32_t n_ll = (el >> 10) & keep, n_ml = (em >> 10) & keep, n_of = (eo >> 10) & keep;
        const uint32_t s1 = a_of + a_ml, s2 = s1 + a_ll, s3 = s2 + n_ll, s4 = s3 + n_ml, tot = s4 + n_of;
        uint32_t x_of, x_ml, x_ll, y_ll, y_ml, y_of;
        b.refill();
        if (tot <= 32) {
            const uint32_t hi = (uint32_t)(b.c >> 32);
            x_of = shr32(hi, 32 - a_of);
            x_ml = shr32(shl32(hi, a_of), 32 - a_ml);
            x_ll = shr32(shl32(hi, s1), 32 - a_ll);
            y_ll = shr32(shl32(hi, s2), 32 - n_ll);
            y_ml = shr32(shl32(hi, s3), 32 - n_ml);
            y_of = shr32(shl32(hi, s4), 32 - n_of);
            b.c <<= tot;
            b.avail -= (int)tot;
            b.left -= (int)tot;
        } else {
            x_of = b.read((int)a_of);
            b.refill();
            x_ml = b.read((int)a_ml);
            x_ll = b.read((int)a_ll);
            b.refill();
            y_ll = b.read((int)n_ll);
            y_ml = b.read((int)n_ml);
            y_of = b.read((int)n_of);
        }
        st_ll = (el & 0x3FF) + y_ll;
        st_ml = (em & 0x3FF) + y_ml;
        st_of = (eo & 0x3FF) + y_of;
        const uint32_t ofv = (1u << a_of) + x_of, ml = lut[36 + (em >> 19)] + x_ml, ll = lut[el >> 19] + x_ll;
        uint32_t offset;
        if (ofv > 3) {
            offset = ofv - 3;
            r2 = r1, r1 = r0, r0 = offset;
        } else {
            const uint32_t idx = ofv - 1 + (ll == 0 ? 1 : 0);
            if (idx == 0) {
                offset = r0;
            } else {
                if (idx == 3 && r0 == 1) return false;  // offset 0: libzstd patches it up; let it decide
                offset = idx == 1 ? r1 : idx == 2 ? r2 : r0 - 1;
                if (idx != 1) r2 = r1;
                r1 = r0, r0 = offset;
            }
        }
        lsum += ll;
        badacc |= (offset >> 27) | ((ll | ml) >> 18) | (lsum >> 30);
        so[i] = zd::seq_pack(ll & 0x3FFFF, ml & 0x3FFFF, offset);
    }
    r0_ = r0, r1_ = r1, r2_ = r2;
    return b.left == 0 && badacc == 0 && lsum <= regen;
}

// One warp, up to kSeqFrames frames: slot s = lane / 8; its eight lanes stage the block's tables, lane 8 s walks the chain.
// `tabs_smem`: kSeqFrames x kTabCells cells.  Returns the slot's verdict to its eight lanes.
__device__ __forceinline__ bool seq_frames(const uint8_t *src, const Plan2 *plan, const Tables2 *tabs, uint64_t *seqs, bool have,
                                           uint32_t *tabs_smem) {
    const int lane = threadIdx.x & 31, sub = lane & 7;
    uint32_t *T = tabs_smem + (size_t)(lane >> 3) * kTabCells;
    uint32_t *lut = tabs_smem + (size_t)kSeqFrames * kTabCells;
    for (int c = lane; c < kLutCells; c += 32) lut[c] = c < 36 ? zd::ll_base(c) : zd::ml_base(c - 36);
    uint32_t r0 = 1, r1 = 4, r2 = 8;
    bool ok = have;
    for (uint32_t bi = 0; bi < zd::kP2MaxBlocks; bi++) {
        const Block2 *bp = have && bi < plan->nblocks ? &plan->blocks[bi] : nullptr;
        const bool run = ok && bp && bp->type == 2 && bp->nseq != 0;
        __syncwarp();
        if (run) {
            if (bp->ll_log > 9 || bp->of_log > 8 || bp->ml_log > 9) {
                ok = false;
            } else {
                const uint2 *gll = (const uint2 *)tabs->ll[bp->ll_tab], *gof = (const uint2 *)tabs->of[bp->of_tab],
                            *gml = (const uint2 *)tabs->ml[bp->ml_tab];
                for (uint32_t k = sub; k < (1u << bp->ll_log); k += 8) T[k] = pack_cell(__ldg(gll + k), 0);
                for (uint32_t k = sub; k < (1u << bp->of_log); k += 8) T[512 + k] = pack_cell(__ldg(gof + k), 1);
                for (uint32_t k = sub; k < (1u << bp->ml_log); k += 8) T[768 + k] = pack_cell(__ldg(gml + k), 2);
            }
        }
        __syncwarp();
        if (run && ok && sub == 0) ok = seq_block(src, *bp, T, T + 512, T + 768, lut, seqs + bp->seq_off, r0, r1, r2);
        ok = __shfl_sync(0xFFFFFFFFu, (int)ok, lane & ~7) != 0;
    }
    return ok;
}

}  // namespace zd3
}  // namespace fheb

#include "zstd_exec3.h"
static_assert(fheb::zd3::kExecOutBytes == fheb::kPayloadStride, "the output buffer is one payload slot");

namespace fheb {
namespace zd3 {

// block-wide exclusive scan of one 64-bit value per thread (kExecThreads threads); *total = the sum.  `tmp`: 33 words.
__device__ __forceinline__ uint64_t block_exscan(uint64_t v, uint64_t *tmp, uint64_t *total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint64_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint64_t o = __shfl_up_sync(0xFFFFFFFFu, inc, d);
        if (lane >= d) inc += o;
    }
    __syncthreads();  // tmp may still be read from a previous call
    if (lane == 31) tmp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        const uint64_t w = tmp[lane];
        uint64_t winc = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint64_t o = __shfl_up_sync(0xFFFFFFFFu, winc, d);
            if (lane >= d) winc += o;
        }
        tmp[lane] = winc - w;
        if (lane == 31) tmp[32] = winc;
    }
    __syncthreads();
    *total = tmp[32];
    return tmp[warp] + inc - v;
}

// Executes the plan of one frame with the whole CTA (the steps are described in zstd_exec3.h).  `seqs` is rewritten in place.
// Returns kZdOk with the payload in dst[0, content) (dst: 16-byte aligned, kPayloadStride bytes), else kZdFallback.
__device__ __forceinline__ int exec_frame(const uint8_t *src, const Plan2 *plan, uint64_t *seqs, const uint8_t *lits, uint8_t *dst,
                                          uint8_t *smem) {
    uint8_t *out = smem;
    uint32_t *pend = (uint32_t *)(smem + kExecOutBytes);
    uint32_t *start = pend + kExecBitWords;
    uint64_t *tmp = (uint64_t *)(start + kExecBitWords);
    int *bad = (int *)(tmp + 64);
    uint32_t *moved_cnt = (uint32_t *)(tmp + 64) + 2;  // matches moved in a jump round (two alternating counters)
    uint8_t *lit_smem = (uint8_t *)(tmp + 64) + 64;
    const int t = threadIdx.x;
    if (t == 0) *bad = 0;
    for (uint32_t k = t; k < 2 * kExecBitWords; k += kExecThreads) pend[k] = 0;  // both bitmaps
    __syncthreads();
    const uint32_t cap = plan->content, window = plan->window;
    const uint32_t block_max = window < zd::kBlockMax ? window : (uint32_t)zd::kBlockMax;
    if (cap > kCtPayloadBytes) return zd::kZdFallback;
    uint32_t pos = 0;  // uniform
    const uint32_t nblocks = plan->nblocks;
    for (uint32_t bi = 0; bi < nblocks; bi++) {
        const Block2 &bp = plan->blocks[bi];
        if (bp.type != 2) {
            if (bp.size > cap - pos) return zd::kZdFallback;
            if (bp.type == 0) {
                const uint8_t *s = src + bp.src_off;
                for (uint32_t k = t; k < bp.size; k += kExecThreads) out[pos + k] = s[k];
            } else {
                const uint8_t v = src[bp.src_off];
                for (uint32_t k = t; k < bp.size; k += kExecThreads) out[pos + k] = v;
            }
            pos += bp.size;
            __syncthreads();
            continue;
        }
        const uint8_t *lit = bp.lit_mode == 0 ? src + bp.lit_off : lits + bp.lit_off;
        const int lit_rle = bp.lit_mode == 1 ? bp.lit_rle : -1;
        const uint32_t nseq = bp.nseq, regen = bp.regen;
        uint64_t *so = seqs + bp.seq_off;
        if (kExecLitBytes && lit_rle < 0 && regen <= kExecLitBytes) {  // (the barriers of the scans below order these stores before their readers)
            for (uint32_t k = t; k < regen; k += kExecThreads) lit_smem[k] = lit[k];
            lit = lit_smem;
        }
        // ---- 1 positions: warp w owns a run of consecutive sequences, lane l its sequences l, l + 32, ...: neighbouring lanes
        // work on neighbouring sequences, so their records load coalesced and their output bytes fall in different banks
        // (16 consecutive sequences per THREAD put every lane of a warp 128 bytes apart: 32-way conflicts on each store)
        const int lane = t & 31, warp = t >> 5;
        const uint32_t rows = ((nseq + 31) / 32 + 31) / 32;  // rows of 32 sequences per warp
        const uint32_t wlo = min((uint32_t)warp * rows * 32, nseq), whi = min(wlo + rows * 32, nseq);
        uint64_t s_ll = 0, s_o = 0;
        for (uint32_t i = wlo + lane; i < whi; i += 32) {
            const uint64_t e = so[i];
            const uint32_t ll = (uint32_t)(e & 0x3FFFF), ml = (uint32_t)((e >> 18) & 0x3FFFF);
            s_ll += ll;
            s_o += (uint64_t)ll + ml;
        }
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {
            s_ll += __shfl_xor_sync(0xFFFFFFFFu, s_ll, d);
            s_o += __shfl_xor_sync(0xFFFFFFFFu, s_o, d);
        }
        uint64_t tot_ll, tot_o;
        uint64_t run_ll = block_exscan(lane == 0 ? s_ll : 0, tmp, &tot_ll);  // lane 0: the sum over the warps before this one
        uint64_t run_o = block_exscan(lane == 0 ? s_o : 0, tmp, &tot_o);
        run_ll = __shfl_sync(0xFFFFFFFFu, run_ll, 0);
        run_o = __shfl_sync(0xFFFFFFFFu, run_o, 0);
        if (tot_ll > regen) return zd::kZdFallback;  // (uniform: the totals are)
        const uint64_t total_out = tot_o + (regen - tot_ll);
        if (total_out > (uint64_t)(cap - pos) || total_out > block_max) return zd::kZdFallback;
        {
            bool ok = true;
            for (uint32_t r0 = wlo; r0 < whi; r0 += 32) {  // (uniform per warp)
                const uint32_t i = r0 + lane;
                uint32_t ll = 0, both = 0;
                if (i < whi) {
                    const uint64_t e = so[i];
                    ll = (uint32_t)(e & 0x3FFFF);
                    both = ll + (uint32_t)((e >> 18) & 0x3FFFF);
                }
                uint32_t inc_ll = ll, inc_o = both;  // (a row's totals fit 32 bits: 32 x 2^19)
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t a_ = __shfl_up_sync(0xFFFFFFFFu, inc_ll, d), b_ = __shfl_up_sync(0xFFFFFFFFu, inc_o, d);
                    if (lane >= d) inc_ll += a_, inc_o += b_;
                }
                if (i < whi)
                    ok &= place_sequences(so, i, i + 1, (uint32_t)run_ll + inc_ll - ll, pos + (uint32_t)run_o + inc_o - both, lit, lit_rle,
                                          window, out, pend, start);
                run_ll += __shfl_sync(0xFFFFFFFFu, inc_ll, 31);
                run_o += __shfl_sync(0xFFFFFFFFu, inc_o, 31);
            }
            if (!ok) *bad = 1;
        }
        {
            const uint32_t tail = regen - (uint32_t)tot_ll, tpos = pos + (uint32_t)tot_o;
            if (lit_rle >= 0) {
                for (uint32_t k = t; k < tail; k += kExecThreads) out[tpos + k] = (uint8_t)lit_rle;
            } else {
                for (uint32_t k = t; k < tail; k += kExecThreads) out[tpos + k] = lit[(uint32_t)tot_ll + k];
            }
        }
        __syncthreads();
        if (*bad) return zd::kZdFallback;
        for (uint32_t cbase = 0; cbase < nseq; cbase += kExecChunk) {
            // ---- 2 jumping: thread t carries sequences cbase + t + 1024 k
            uint32_t F[kExecPer], MI[kExecPer];
            uint32_t act = 0;
#pragma unroll
            for (int k = 0; k < kExecPer; k++) {
                const uint32_t i = cbase + (uint32_t)k * kExecThreads + t;
                F[k] = 0, MI[k] = 0;
                if (i < nseq && jump_init(so[i], &F[k], &MI[k])) act |= 1u << k;
            }
            const uint32_t stop_below = min(nseq - cbase, kExecChunk) / kExecStopShare;
            if (t == 0) moved_cnt[0] = moved_cnt[1] = 0;
            __syncthreads();
            for (int round = 0; round < kExecMaxRounds; round++) {
                uint32_t chg = 0, nmoved = 0;
#pragma unroll
                for (int k = 0; k < kExecPer; k++) {
                    if (!((act >> k) & 1)) continue;
                    const int r = jump_look(pend, start, out, &F[k], &MI[k], so + cbase + (uint32_t)k * kExecThreads + t);
                    if (r == kJumpStop) {
                        act &= ~(1u << k);
                        continue;
                    }
                    nmoved++;
                    if (r == kJumpPublish) chg |= 1u << k;
                    else chg |= 1u << 16;  // (moved, nothing to publish)
                }
                nmoved = __reduce_add_sync(0xFFFFFFFFu, nmoved);
                if ((t & 31) == 0 && nmoved) atomicAdd(&moved_cnt[round & 1], nmoved);
                __syncthreads();  // every offset has been read
#pragma unroll
                for (int k = 0; k < kExecPer; k++)
                    if ((chg >> k) & 1) jump_publish(out, F[k], MI[k]);
                if (t == 0) moved_cnt[(round + 1) & 1] = 0;
                if (!__syncthreads_or((int)chg)) break;
                if (moved_cnt[round & 1] < stop_below) break;  // (uniform: read by everyone between the same two barriers)
            }
            // ---- 3 copies: in order per thread; a warp polls until its 32 current matches are copied
#pragma unroll
            for (int k = 0; k < kExecPer; k++) {
                const uint32_t i = cbase + (uint32_t)k * kExecThreads + t;
                CopyJob job = copy_job(i < nseq ? so[i] : 0, F[k], MI[k]);
                bool waiting = job.ml != 0;
                // (the spin bound is insurance, not logic: dependencies point backwards, so the earliest pending match is always
                // ready and every poll loop ends; should that reasoning ever be wrong, the frame goes back to the host instead
                // of hanging the GPU)
                for (uint32_t spins = 0; __any_sync(0xFFFFFFFFu, waiting); spins++) {
                    if (waiting && copy_ready(pend, job)) {
                        __threadfence_block();  // the bytes behind the cleared bits
                        copy_match(out, job);
                        __threadfence_block();
                        pend_clear(pend, job.m, job.ml);
                        waiting = false;
                    }
                    if (spins >= kExecMaxSpins) {
                        *bad = 1;
                        waiting = false;
                    }
                }
            }
            __syncthreads();
            if (*bad) return zd::kZdFallback;
        }
        pos += (uint32_t)total_out;
    }
    if (pos != cap) return zd::kZdFallback;
    {
        const uint4 *s4 = (const uint4 *)out;
        uint4 *d4 = (uint4 *)dst;
        const uint32_t n16 = (cap + 15) / 16;
        for (uint32_t k = t; k < n16; k += kExecThreads) d4[k] = s4[k];
    }
    return zd::kZdOk;
}

}  // namespace zd3
}  // namespace fheb
