// Parallel execution of a zstd block's sequences: the per-sequence steps of k_zd3_exec (zstd_plan3.cuh).  Host + device source:
// tests/zd_host.cpp runs the same steps thread by thread (zd_decode_v3) against libzstd, so the logic is tested without a GPU.
//
// One CTA per frame.  Shared memory: the frame's output (131,200 B), one "pending" bit and one "match start" bit per output byte
// (and, optionally, the block's literals: kExecLitBytes).
//  1 positions   block-wide prefix sums over (literal length, literal + match length) give every sequence the place of its
//                literals and of its match; literals are copied, the match region is marked pending, its first byte marked as a
//                start and its first three bytes (a match is >= 3 bytes) hold the match's PUBLISHED OFFSET (start - source):
//                out[start + k] == out[start - offset + k] for every k of the region             (place_sequences)
//  2 jumping     a match whose source lies inside ONE pending match j reads j's published offset and moves its own source back
//                by it: pointer jumping, so a chain of d dependent matches - 16 chains of ~1,000 in a ciphertext frame, one per
//                value of a residue's top nibble - collapses in ~log2 d rounds instead of d copy-and-wait hops.  Offsets are
//                read in one half of a round and published in the other (CTA barriers between), so none is read while it is
//                written.  A source that starts with ready bytes and continues inside one pending match (a literal in front of
//                a match: what a ciphertext frame's 5-byte matches usually copy) is SPLIT once: the ready part's position is
//                recorded in the sequence record, the rest goes on jumping.  Split matches, and overlapping ones (offset <
//                length: the region repeats with period `offset`, which a moved source would not reproduce for readers of its
//                later bytes), keep the last offset they published                                  (jump_init / jump_look / jump_publish)
//  3 copies      a match is copied once none of its source bytes is pending, then its pending bits are cleared; whatever the
//                jumping could not resolve (sources that straddle several matches, long matches) polls the bitmap - dependencies
//                point backwards only, so the earliest pending match is always ready.  A byte is only ever copied from
//                non-pending bytes: the jumping makes that moment come earlier, it never lets a copy read unfinished data
//                                                                                                   (copy_job / copy_ready / copy_match)
#pragma once
#include <cstddef>
#include <cstdint>

#include "zstd_dec.h"

namespace fheb {
namespace zd3 {

constexpr int kExecThreads = 1024;
constexpr int kExecPer = 16;                                      // sequences a thread carries in registers
constexpr uint32_t kExecChunk = kExecThreads * kExecPer;          // sequences per pass (a ciphertext block has 16,354)
constexpr int kExecMaxRounds = 64;
// The jumping stops once a round moves fewer than 1/kExecStopShare of the pass's sequences: the last rounds serve a few per
// cent of the matches (chains through matches that keep their offset advance one link per round) at the full price of a round
// for every warp; what is left then waits in step 3 for the few links in front of it (a ciphertext frame: 9 rounds instead of 14).
constexpr uint32_t kExecStopShare = 8;
constexpr uint32_t kExecMaxSpins = 1u << 20;                      // polls of one match before the frame is handed back
constexpr size_t kExecOutBytes = 131200;                          // kPayloadStride: the frame's content (131,169) + store slack
constexpr size_t kExecBitWords = (kExecOutBytes + 31) / 32 + 2;   // one bit per output byte (+ slack for pair loads)
// A block's literals can be staged in shared memory (56 KB hold a ciphertext block's 55.7 KB) - measured: no gain once the
// placement was free of bank conflicts, while 222 KB per CTA keep every other kernel off the SM: with 165 KB three sequence-chain
// CTAs of other tiles (20.8 KB each) fit beside an execution CTA, and in the tile pipeline that is what counts.
constexpr size_t kExecLitBytes = 0;
constexpr size_t kExecSmem = kExecOutBytes + 2 * kExecBitWords * 4 + 64 * 8 + 64 + kExecLitBytes;  // 164,592 of 232,448

#if defined(__CUDA_ARCH__)
#define ZD3_OR(p, v) atomicOr((p), (v))
#define ZD3_AND(p, v) atomicAnd((p), (v))
#define ZD3_VOL volatile
ZD_HD uint32_t z3_funnel_r(uint32_t lo, uint32_t hi, uint32_t s) { return __funnelshift_r(lo, hi, s); }
ZD_HD int z3_ffs(uint32_t v) { return __ffs((int)v); }
#else
#define ZD3_OR(p, v) (*(p) |= (v))
#define ZD3_AND(p, v) (*(p) &= (v))
#define ZD3_VOL
ZD_HD uint32_t z3_funnel_r(uint32_t lo, uint32_t hi, uint32_t s) { return s ? (lo >> s) | (hi << (32 - s)) : lo; }
ZD_HD int z3_ffs(uint32_t v) { return __builtin_ffs((int)v); }
#endif

ZD_HD uint32_t word_mask(uint32_t w, uint32_t a, uint32_t e) {  // bits of word w inside [a, e]
    uint32_t m = 0xFFFFFFFFu;
    if (w == (a >> 5)) m &= 0xFFFFFFFFu << (a & 31);
    if (w == (e >> 5)) m &= 0xFFFFFFFFu >> (31 - (e & 31));
    return m;
}
ZD_HD void pend_set(uint32_t *pend, uint32_t a, uint32_t len) {  // len >= 1
    const uint32_t e = a + len - 1;
    for (uint32_t w = a >> 5; w <= (e >> 5); w++) ZD3_OR(pend + w, word_mask(w, a, e));
}
ZD_HD void pend_clear(uint32_t *pend, uint32_t a, uint32_t len) {
    const uint32_t e = a + len - 1;
    for (uint32_t w = a >> 5; w <= (e >> 5); w++) ZD3_AND(pend + w, ~word_mask(w, a, e));
}
ZD_HD bool pend_any(const ZD3_VOL uint32_t *pend, uint32_t a, uint32_t len) {
    const uint32_t e = a + len - 1;
    uint32_t any = 0;
    for (uint32_t w = a >> 5; w <= (e >> 5); w++) any |= pend[w] & word_mask(w, a, e);
    return any != 0;
}
ZD_HD uint32_t bits_at(const uint32_t *bm, uint32_t a) {  // bits [a, a + 32)
    const uint32_t w = a >> 5;
    return z3_funnel_r(bm[w], bm[w + 1], a & 31);
}
ZD_HD void put24(uint8_t *p, uint32_t v) { p[0] = (uint8_t)v, p[1] = (uint8_t)(v >> 8), p[2] = (uint8_t)(v >> 16); }
ZD_HD uint32_t get24(const uint8_t *p) { return (uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16; }

// sequence record after step 1: match start (18 bits) | match length (18) | offset of the match's FIRST bytes (28)
ZD_HD uint64_t rec_pack(uint32_t m, uint32_t ml, uint32_t off) { return (uint64_t)m | (uint64_t)ml << 18 | (uint64_t)off << 36; }

// step 1 for sequences [lo, hi) (records still (literal length, match length, offset)); literals from lit + lpos, output from
// opos.  False when a sequence is not executable (offset 0 / beyond the window / beyond the output so far, match < 3).
ZD_HD bool place_sequences(uint64_t *so, uint32_t lo, uint32_t hi, uint32_t lpos, uint32_t opos, const uint8_t *lit, int lit_rle,
                           uint32_t window, uint8_t *out, uint32_t *pend, uint32_t *start) {
    bool ok = true;
    for (uint32_t i = lo; i < hi; i++) {
        const uint64_t e = so[i];
        const uint32_t ll = (uint32_t)(e & 0x3FFFF), ml = (uint32_t)((e >> 18) & 0x3FFFF), off = (uint32_t)(e >> 36);
        if (lit_rle >= 0) {
            for (uint32_t k = 0; k < ll; k++) out[opos + k] = (uint8_t)lit_rle;
        } else {
            for (uint32_t k = 0; k < ll; k++) out[opos + k] = lit[lpos + k];
        }
        const uint32_t m = opos + ll;
        if (off == 0 || off > window || off > m || ml < 3) {  // (a zstd match is at least 3 bytes)
            ok = false;
        } else {
            pend_set(pend, m, ml);
            ZD3_OR(start + (m >> 5), 1u << (m & 31));
            put24(out + m, off);
        }
        so[i] = rec_pack(m, ml, off);
        lpos += ll;
        opos = m + ml;
    }
    return ok;
}

// jump state of a sequence: F = source of byte `a` of the match;
// MI = match start | bytes still to resolve << 18 (<= 32) | a << 24 | keep-the-published-offset << 31
ZD_HD bool jump_init(uint64_t rec, uint32_t *F, uint32_t *MI) {
    const uint32_t m = (uint32_t)(rec & 0x3FFFF), ml = (uint32_t)((rec >> 18) & 0x3FFFF), off = (uint32_t)(rec >> 36);
    const uint32_t need = ml < off ? ml : off;
    *F = m - off;
    *MI = 0;
    if (need > 32) return false;  // long matches wait for their source in step 3
    *MI = m | need << 18 | (ml > off ? 1u << 31 : 0u);
    return true;
}
enum { kJumpStop = 0, kJumpPublish = 1, kJumpMoved = 2 };
// one look at the source of an active sequence (read half of a round); `rec` is the sequence's record (rewritten on a split)
ZD_HD int jump_look(const uint32_t *pend, const uint32_t *start, const uint8_t *out, uint32_t *F, uint32_t *MI, uint64_t *rec) {
    uint32_t f = *F, need = (*MI >> 18) & 0x3F;
    uint32_t mask = 0xFFFFFFFFu >> (32 - need);  // need in [1, 32]
    uint32_t pw = bits_at(pend, f) & mask;
    uint32_t sw = bits_at(start, f) & mask;
    if (pw != mask && pw != 0 && !(*MI >> 24)) {  // a partly ready source, seen for the first time: split (non-overlapping only)
        const uint32_t a = (uint32_t)z3_ffs(pw) - 1;  // ready bytes in front
        if (a >= 1 && (pw >> a) == (mask >> a)) {
            const uint32_t m = *MI & 0x3FFFF;
            *rec = rec_pack(m, need, m - f);  // (not overlapping: length == need) the first a bytes come from f
            f += a, need -= a, mask >>= a, pw >>= a, sw >>= a;
            *F = f;
            *MI = m | need << 18 | a << 24 | 1u << 31;
        }
    }
    if (pw != mask || (sw & ~1u)) return kJumpStop;  // ready, or not inside one pending match: step 3 decides
    uint32_t mj = f;
    if (!(sw & 1)) {  // the highest start at or below f
        int w = (int)(f >> 5);
        uint32_t bits = start[w] & ~(0xFFFFFFFFu << (f & 31));
        for (int back = 0; bits == 0 && back < 8 && w > 0; back++) bits = start[--w];
        if (bits == 0) return kJumpStop;
        mj = (uint32_t)w * 32 + (uint32_t)zd::highbit(bits);
    }
    *F = f - get24(out + mj);
    return (*MI >> 31) ? kJumpMoved : kJumpPublish;
}
ZD_HD void jump_publish(uint8_t *out, uint32_t F, uint32_t MI) {  // write half of a round
    const uint32_t m = MI & 0x3FFFF;
    put24(out + m, m - F);
}

struct CopyJob {
    uint32_t m, ml;       // destination
    uint32_t from0, a;    // bytes [0, a) come from from0 (ready since the split)
    uint32_t from, span;  // the rest from `from`, repeating with period `span`
};
ZD_HD CopyJob copy_job(uint64_t rec, uint32_t F, uint32_t MI) {
    CopyJob j;
    j.m = (uint32_t)(rec & 0x3FFFF), j.ml = (uint32_t)((rec >> 18) & 0x3FFFF);
    const uint32_t off = (uint32_t)(rec >> 36);
    const uint32_t need = j.ml < off ? j.ml : off;
    j.from0 = j.m - off;
    j.a = (MI >> 24) & 0x3F;
    j.from = F;
    j.span = need - j.a;
    return j;
}
ZD_HD bool copy_ready(const ZD3_VOL uint32_t *pend, const CopyJob &j) { return !pend_any(pend, j.from, j.span); }
ZD_HD void copy_match(uint8_t *out, const CopyJob &j) {
    ZD3_VOL uint8_t *vo = out;
    for (uint32_t k = 0; k < j.a; k++) vo[j.m + k] = vo[j.from0 + k];
    uint32_t idx = 0;
    for (uint32_t k = j.a; k < j.ml; k++) {
        vo[j.m + k] = vo[j.from + idx];
        if (++idx == j.span) idx = 0;
    }
}

}  // namespace zd3
}  // namespace fheb
