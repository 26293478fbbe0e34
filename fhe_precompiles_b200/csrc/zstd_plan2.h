// Batch-oriented zstd inflate (RFC 8878), third generation of the device decoder (see zstd_dec.h for the contract: kZdOk
// means byte-identical to libzstd, anything else is handed back to the host).
//
// A frame's bitstreams are chains of dependent steps -- one GPU thread walks them at ~10x a host core's cost per step -- so
// throughput comes from (a) thousands of frames in flight and (b) splitting a frame into its INDEPENDENT chains and keeping
// each chain's inner loop to a few dozen instructions:
//   parse   one thread per frame   frame / block / section headers, Huffman and FSE table construction (per block, in global
//                                  scratch), descriptors of every bitstream                                (plan2_parse)
//   huf     one thread per (frame, Huffman stream): the four literal streams of a block decode concurrently  (plan2_huf)
//   seq     one thread per frame   FSE sequence streams -> (literal length, match length, resolved offset)   (plan2_seq)
//   exec    one thread per frame   sequence execution on registers: 64-bit appends, aligned word stores        (plan2_exec)
// huf and seq are independent of each other and run on two streams.  The bit readers use one aligned 32-bit load per 32 bits
// consumed and keep the unread bits left-aligned in a 64-bit register; tables are read through the read-only path.
//
// The same source compiles for the host (tests/zd_host.cpp: zd_decode_v2, compared with libzstd on fixtures, many levels and
// mutated frames) and for the device (codec_kernels.cu).
#pragma once
#include "zstd_dec.h"

namespace fheb {
namespace zd {

constexpr uint32_t kP2MaxBlocks = 4;       // a 131,169-byte ciphertext payload is two blocks; more go back to the host
constexpr uint32_t kP2MaxSeqs = 40960;

#if defined(__CUDA_ARCH__)
#define ZD_LDG32(p) __ldg((const unsigned int *)(p))
#define ZD_LDG16(p) __ldg((const unsigned short *)(p))
#else
#define ZD_LDG32(p) (*(const uint32_t *)(p))
#define ZD_LDG16(p) (*(const uint16_t *)(p))
#endif

struct Block2 {
    uint32_t type;             // 0 raw, 1 RLE, 2 compressed
    uint32_t src_off, size;    // raw: first byte / RLE: the byte; decompressed size
    uint32_t lit_mode;         // 0 raw literals inside src (lit_off), 1 RLE (lit_rle), 2 Huffman (decoded into lits + lit_off)
    uint32_t lit_off;
    int32_t lit_rle;
    uint32_t regen;
    uint32_t huf_tab, huf_log;        // Huffman table slot (the block that defined it) and its depth
    uint32_t nstreams;                // 1 or 4
    uint32_t st_off[4], st_len[4];    // stream bytes inside src
    uint32_t st_out[4], st_cnt[4];    // where its symbols go in lits, and how many
    uint32_t nseq, seq_off;           // records in seqs
    uint32_t bs_off, bs_len;          // sequence bitstream inside src
    uint32_t ll_tab, of_tab, ml_tab;  // table slots
    uint32_t ll_log, of_log, ml_log;
};
struct Plan2 {
    int32_t status;      // parse verdict; huf / seq downgrade it to kZdFallback through huf_bad / seq_bad
    int32_t huf_bad[4];  // one per stream thread (no atomics needed)
    int32_t seq_bad;
    uint32_t nblocks, content, window;
    Block2 blocks[kP2MaxBlocks];
};
struct Tables2 {  // per job: one slot per block
    uint16_t huf[kP2MaxBlocks][1 << kHufLogMax];
    SeqEntry ll[kP2MaxBlocks][512], of[kP2MaxBlocks][256], ml[kP2MaxBlocks][512];
};

// ---------------------------------------------------------------- backward bit reader, one aligned 32-bit load per refill
// The stream is src[0..len); its last byte's highest set bit is the end marker.  `c` holds the next unread bits left-aligned;
// `left` counts the stream bits not yet consumed (negative after an over-read; reads below bit 0 yield zeros).
struct Back32 {
    const uint32_t *wp;  // next word to load (descending)
    const uint32_t *w0;  // the aligned word that holds the first byte of the stream
    uint32_t lo_bits;    // bits of *w0 below the stream's first byte
    uint64_t c;
    int avail;
    int64_t left;
    ZD_HD bool init(const uint8_t *base, size_t len) {
        if (len == 0) return false;
        const uint8_t last = base[len - 1];
        if (last == 0) return false;
        const uintptr_t a = (uintptr_t)base;
        w0 = (const uint32_t *)(a & ~(uintptr_t)3);
        lo_bits = (uint32_t)(a & 3) * 8;
        left = (int64_t)(len - 1) * 8 + highbit(last);  // bits below the end marker
        const uint64_t abs_top = (uint64_t)lo_bits + (uint64_t)left;  // exclusive, in bits from *w0
        const uint32_t r = (uint32_t)(abs_top & 31);
        wp = w0 + (abs_top >> 5);
        c = 0;
        avail = 0;
        if (r) {
            uint32_t w = ZD_LDG32(wp) & ((1u << r) - 1);
            if (wp == w0 && lo_bits) w &= ~((1u << lo_bits) - 1);
            c = (uint64_t)w << (64 - r);
            avail = (int)r;
        }
        wp--;
        return true;
    }
    // afterwards avail >= 32 (bits below the stream read as zeros)
    ZD_HD void refill() {
        if (avail < 32) {
            uint32_t w = 0;
            if (wp >= w0) {
                w = ZD_LDG32(wp);
                if (wp == w0 && lo_bits) w &= ~((1u << lo_bits) - 1);
            }
            wp--;
            c |= (uint64_t)w << (32 - avail);
            avail += 32;
        }
    }
    ZD_HD uint32_t peek(int n) const { return (uint32_t)((c >> 1) >> (63 - n)); }  // n in [0, 32], n <= avail
    ZD_HD void skip(int n) {
        c <<= n;
        avail -= n;
        left -= n;
    }
    ZD_HD uint32_t read(int n) {
        const uint32_t v = peek(n);
        skip(n);
        return v;
    }
};

// ---------------------------------------------------------------- parse
// Everything plan_frame (zstd_dec.h) checks, minus the two hot loops.  `w` is table-construction scratch.
ZD_FN int plan2_parse(const uint8_t *src, size_t slen, size_t cap, Work *w, Plan2 *plan, Tables2 *tabs) {
    plan->status = kZdFallback;
    plan->seq_bad = 0;
    for (int k = 0; k < 4; k++) plan->huf_bad[k] = 0;
    plan->nblocks = 0;
    if (slen < 6 || le_at(src, 0, 4) != 0xFD2FB528ull) return kZdFallback;
    const uint8_t fhd = byte_at(src, 4);
    const int fcs_flag = fhd >> 6, single = (fhd >> 5) & 1;
    if (fhd & 0x1F) return kZdFallback;
    size_t pos = 5, window = 0;
    if (!single) {
        if (pos >= slen) return kZdFallback;
        const uint8_t wd = byte_at(src, pos++);
        const int wlog = 10 + (wd >> 3);
        if (wlog > 27) return kZdFallback;
        window = ((size_t)1 << wlog) + (((size_t)1 << wlog) >> 3) * (wd & 7);
    }
    const int fcs_bytes = fcs_flag == 0 ? (single ? 1 : 0) : fcs_flag == 1 ? 2 : fcs_flag == 2 ? 4 : 8;
    if (fcs_bytes == 0 || pos + (size_t)fcs_bytes > slen) return kZdFallback;
    uint64_t fcs = le_at(src, pos, fcs_bytes);
    if (fcs_bytes == 2) fcs += 256;
    pos += (size_t)fcs_bytes;
    if (fcs > cap || fcs >= (1u << 27)) return kZdFallback;
    if (single) window = (size_t)fcs;
    const size_t block_max = window < kBlockMax ? window : kBlockMax;
    plan->content = (uint32_t)fcs;
    plan->window = (uint32_t)(window < (1u << 27) ? window : (1u << 27));
    w->have_ll = w->have_of = w->have_ml = w->have_huf = 0;
    uint32_t nblocks = 0, nseq_total = 0, lit_total = 0;
    uint32_t cur_huf = 0, cur_huf_log = 0, cur_ll = 0, cur_of = 0, cur_ml = 0;
    for (;;) {
        if (pos + 3 > slen || nblocks >= kP2MaxBlocks) return kZdFallback;
        const uint32_t bh = (uint32_t)le_at(src, pos, 3);
        pos += 3;
        const int last = bh & 1, type = (bh >> 1) & 3;
        const size_t bsize = bh >> 3;
        const uint32_t bi = nblocks++;
        Block2 &bp = plan->blocks[bi];
        bp.type = (uint32_t)type;
        bp.nseq = 0;
        bp.lit_mode = 0;
        bp.nstreams = 0;
        if (type == 3 || bsize > block_max) return kZdFallback;
        if (type == 0) {
            if (pos + bsize > slen) return kZdFallback;
            bp.src_off = (uint32_t)pos;
            bp.size = (uint32_t)bsize;
            pos += bsize;
        } else if (type == 1) {
            if (pos + 1 > slen) return kZdFallback;
            bp.src_off = (uint32_t)pos;
            bp.size = (uint32_t)bsize;
            pos += 1;
        } else {
            if (bsize < 2 || pos + bsize > slen) return kZdFallback;
            const uint8_t *b = src + pos;
            const size_t len = bsize;
            // ---- literals section
            const uint8_t b0 = byte_at(b, 0);
            const int ltype = b0 & 3, sf = (b0 >> 2) & 3;
            size_t hdr, regen, comp = 0;
            int streams = 1;
            if (ltype < 2) {
                hdr = (sf == 0 || sf == 2) ? 1 : sf == 1 ? 2 : 3;
                if (hdr > len) return kZdFallback;
                regen = hdr == 1 ? (size_t)(b0 >> 3) : (size_t)(le_at(b, 0, (int)hdr) >> 4);
                comp = ltype == 0 ? regen : 1;
            } else {
                hdr = sf <= 1 ? 3 : sf == 2 ? 4 : 5;
                if (hdr > len) return kZdFallback;
                const uint64_t v = le_at(b, 0, (int)hdr);
                const int bits = sf <= 1 ? 10 : sf == 2 ? 14 : 18;
                regen = (size_t)((v >> 4) & ((1u << bits) - 1));
                comp = (size_t)(v >> (4 + bits));
                streams = sf == 0 ? 1 : 4;
            }
            if (regen > block_max || hdr + comp > len) return kZdFallback;
            bp.regen = (uint32_t)regen;
            bp.lit_rle = -1;
            const uint8_t *lsrc = b + hdr;
            if (ltype == 0) {
                bp.lit_mode = 0;
                bp.lit_off = (uint32_t)(pos + hdr);
            } else if (ltype == 1) {
                bp.lit_mode = 1;
                bp.lit_rle = byte_at(lsrc, 0);
                bp.lit_off = 0;
            } else {
                if ((size_t)lit_total + regen > cap) return kZdFallback;
                size_t used = 0;
                if (ltype == 2) {
                    used = huf_read_tree(lsrc, comp, w);
                    if (used == 0) return kZdFallback;
                    cur_huf = bi;
                    cur_huf_log = (uint32_t)w->huf_log;
                    for (uint32_t k = 0; k < (1u << cur_huf_log); k++) tabs->huf[bi][k] = w->huf[k];
                } else if (!w->have_huf) {
                    return kZdFallback;
                }
                bp.lit_mode = 2;
                bp.huf_tab = cur_huf;
                bp.huf_log = cur_huf_log;
                bp.lit_off = lit_total;
                const size_t hoff = pos + hdr + used;  // first byte of the stream area inside src
                const uint8_t *hs = lsrc + used;
                const size_t hlen = comp - used;
                if (streams == 1) {
                    bp.nstreams = 1;
                    bp.st_off[0] = (uint32_t)hoff, bp.st_len[0] = (uint32_t)hlen;
                    bp.st_out[0] = lit_total, bp.st_cnt[0] = (uint32_t)regen;
                } else {
                    if (hlen < 6 + 4 || regen < 6) return kZdFallback;  // libzstd: >= 10 bytes and >= 6 literals for four streams
                    const size_t s1 = (size_t)le_at(hs, 0, 2), s2 = (size_t)le_at(hs, 2, 2), s3 = (size_t)le_at(hs, 4, 2);
                    const size_t seg = (regen + 3) / 4;
                    if (6 + s1 + s2 + s3 >= hlen || 3 * seg > regen) return kZdFallback;
                    const size_t s4 = hlen - 6 - s1 - s2 - s3;
                    const size_t lens[4] = {s1, s2, s3, s4};
                    bp.nstreams = 4;
                    size_t o = hoff + 6;
                    for (int k = 0; k < 4; k++) {
                        bp.st_off[k] = (uint32_t)o, bp.st_len[k] = (uint32_t)lens[k];
                        bp.st_out[k] = lit_total + (uint32_t)(k * seg);
                        bp.st_cnt[k] = (uint32_t)(k < 3 ? seg : regen - 3 * seg);
                        o += lens[k];
                    }
                }
                lit_total += (uint32_t)regen;
            }
            // ---- sequences section: count, table descriptions (built here), then the bitstream (decoded by plan2_seq)
            const uint8_t *ss = b + hdr + comp;
            const size_t sl = len - hdr - comp;
            if (sl < 1) return kZdFallback;
            size_t nseq = byte_at(ss, 0), sh = 1;
            if (nseq >= 128) {
                if (nseq == 255) {
                    if (sl < 3) return kZdFallback;
                    nseq = (size_t)le_at(ss, 1, 2) + 0x7F00;
                    sh = 3;
                } else {
                    if (sl < 2) return kZdFallback;
                    nseq = ((nseq - 128) << 8) + byte_at(ss, 1);
                    sh = 2;
                }
            }
            bp.nseq = (uint32_t)nseq;
            bp.seq_off = nseq_total;
            if (nseq == 0) {
                if (sl != 1) return kZdFallback;
            } else {
                if (nseq_total + nseq > kP2MaxSeqs || sl < sh + 1) return kZdFallback;
                const uint8_t modes = byte_at(ss, sh);
                if (modes & 3) return kZdFallback;
                size_t p = sh + 1;
                const int m_ll = modes >> 6, m_of = (modes >> 4) & 3, m_ml = (modes >> 2) & 3;
                // a table that this block defines is built straight into the block's slot; "repeat" keeps the previous slot
                if (m_ll != 3) w->ll = tabs->ll[bi], cur_ll = bi;
                if (m_of != 3) w->of = tabs->of[bi], cur_of = bi;
                if (m_ml != 3) w->ml = tabs->ml[bi], cur_ml = bi;
                int used = seq_table(m_ll, ss + p, sl - p, 0, w);
                if (used < 0) return kZdFallback;
                p += (size_t)used;
                if ((used = seq_table(m_of, ss + p, sl - p, 1, w)) < 0) return kZdFallback;
                p += (size_t)used;
                if ((used = seq_table(m_ml, ss + p, sl - p, 2, w)) < 0) return kZdFallback;
                p += (size_t)used;
                if (p >= sl) return kZdFallback;
                bp.ll_tab = cur_ll, bp.of_tab = cur_of, bp.ml_tab = cur_ml;
                bp.ll_log = (uint32_t)w->ll_log, bp.of_log = (uint32_t)w->of_log, bp.ml_log = (uint32_t)w->ml_log;
                bp.bs_off = (uint32_t)(pos + hdr + comp + p);
                bp.bs_len = (uint32_t)(sl - p);
                nseq_total += (uint32_t)nseq;
            }
            pos += bsize;
        }
        if (last) break;
    }
    if (pos != slen) return kZdFallback;
    plan->nblocks = nblocks;
    plan->status = kZdOk;
    return kZdOk;
}

// ---------------------------------------------------------------- Huffman streams
// stream `k` (0..3) of every Huffman block of the frame; false unless each stream is consumed exactly
ZD_FN bool plan2_huf(const uint8_t *src, const Plan2 *plan, const Tables2 *tabs, uint8_t *lits, int k) {
    if (plan->status != kZdOk) return true;
    for (uint32_t bi = 0; bi < plan->nblocks; bi++) {
        const Block2 &bp = plan->blocks[bi];
        if (bp.type != 2 || bp.lit_mode != 2 || (uint32_t)k >= bp.nstreams) continue;
        Back32 b;
        if (!b.init(src + bp.st_off[k], bp.st_len[k])) return false;
        const uint16_t *tab = tabs->huf[bp.huf_tab];
        const int log = (int)bp.huf_log;
        uint8_t *out = lits + bp.st_out[k];
        const uint32_t count = bp.st_cnt[k];
        // eight symbols per 64-bit store where the destination allows: byte stores from one thread are what is slow
        uint32_t i = 0;
        for (; i < count && (((uintptr_t)(out + i)) & 7); i++) {
            b.refill();
            const uint32_t e = ZD_LDG16(tab + b.peek(log));
            out[i] = (uint8_t)e;
            b.skip((int)(e >> 8));
        }
        for (; i + 8 <= count; i += 8) {
            uint64_t pack = 0;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                b.refill();
                const uint32_t e = ZD_LDG16(tab + b.peek(log));
                pack |= (uint64_t)(e & 0xFF) << (8 * j);
                b.skip((int)(e >> 8));
            }
            *(uint64_t *)(out + i) = pack;
        }
        for (; i < count; i++) {
            b.refill();
            const uint32_t e = ZD_LDG16(tab + b.peek(log));
            out[i] = (uint8_t)e;
            b.skip((int)(e >> 8));
        }
        if (b.left != 0) return false;
    }
    return true;
}

// ---------------------------------------------------------------- sequences
struct Cell2 {
    uint32_t x, y;  // x = next_base | nb << 16 | add_bits << 24, y = base_value: the bytes of a SeqEntry
};
static_assert(sizeof(SeqEntry) == 8, "SeqEntry is read as one 8-byte cell");
ZD_HD Cell2 ld_cell(const SeqEntry *e) {
#if defined(__CUDA_ARCH__)
    const uint2 v = __ldg((const uint2 *)e);
    return Cell2{v.x, v.y};
#else
    Cell2 c;
    c.x = (uint32_t)e->next_base | (uint32_t)e->nb << 16 | (uint32_t)e->add_bits << 24;
    c.y = e->base_value;
    return c;
#endif
}
ZD_FN bool plan2_seq(const uint8_t *src, const Plan2 *plan, const Tables2 *tabs, uint64_t *seqs) {
    if (plan->status != kZdOk) return true;
    uint32_t r0 = 1, r1 = 4, r2 = 8;
    for (uint32_t bi = 0; bi < plan->nblocks; bi++) {
        const Block2 &bp = plan->blocks[bi];
        if (bp.type != 2 || bp.nseq == 0) continue;
        Back32 b;
        if (!b.init(src + bp.bs_off, bp.bs_len)) return false;
        const SeqEntry *tll = tabs->ll[bp.ll_tab], *tof = tabs->of[bp.of_tab], *tml = tabs->ml[bp.ml_tab];
        b.refill();
        uint32_t st_ll = b.read((int)bp.ll_log), st_of = b.read((int)bp.of_log), st_ml = b.read((int)bp.ml_log);
        if (b.left < 0) return false;
        uint64_t *so = seqs + bp.seq_off;
        const uint32_t nseq = bp.nseq, regen = bp.regen;
        uint64_t lsum = 0;
        for (uint32_t i = 0; i < nseq; i++) {
            // a cell is {next_base u16 | nb u8 << 16 | add_bits u8 << 24, base_value u32} (SeqEntry), one 8-byte load each
            const Cell2 el = ld_cell(tll + st_ll), eo = ld_cell(tof + st_of), em = ld_cell(tml + st_ml);
            b.refill();
            const uint32_t ofv = eo.y + b.read((int)(eo.x >> 24));
            b.refill();
            const uint32_t ml = em.y + b.read((int)(em.x >> 24));
            const uint32_t ll = el.y + b.read((int)(el.x >> 24));
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
            if (i + 1 < nseq) {
                b.refill();
                st_ll = (el.x & 0xFFFF) + b.read((int)((el.x >> 16) & 0xFF));
                st_ml = (em.x & 0xFFFF) + b.read((int)((em.x >> 16) & 0xFF));
                st_of = (eo.x & 0xFFFF) + b.read((int)((eo.x >> 16) & 0xFF));
            }
            lsum += ll;
            if (b.left < 0 || lsum > regen || offset >= (1u << 27) || ll >= (1u << 18) || ml >= (1u << 18)) return false;
            so[i] = seq_pack(ll, ml, offset);
        }
        if (b.left != 0) return false;
    }
    return true;
}

// ---------------------------------------------------------------- execution, one THREAD per frame
// A warp-per-frame executor spends ~1,300 cycles per sequence on a chain of dependent loads, shared-memory round trips and warp
// barriers (11 ms per ciphertext frame, measured) while 31 of its lanes have next to nothing to copy: a ciphertext's typical
// sequence is five literal bytes and a three-byte match.  Here one thread keeps the output position, the word being assembled
// and the previous word in REGISTERS, appends up to eight bytes per step with shifts, and stores aligned 64-bit words; a match
// whose source lies in the last 8..15 bytes (offset 8: the zero bytes of the previous residue) never touches memory.  Thirty-two
// frames share a warp, so the batch's frames all run at once.
struct Sink64 {
    uint64_t *dst;  // 8-byte aligned; the caller leaves room up to the next multiple of 8 after the content
    uint64_t acc;   // bytes [pos & ~7, pos) of the word being assembled, low byte first, upper bytes zero
    uint64_t prev;  // the word before it
    size_t pos;
    ZD_HD void put(uint64_t v, int k) {  // appends the k (1..8) low bytes of v; the other bytes of v are zero
        const int a = (int)(pos & 7);
        acc |= v << (8 * a);
        if (a + k >= 8) {
            dst[pos >> 3] = acc;
            prev = acc;
            acc = a ? v >> (8 * (8 - a)) : 0;
        }
        pos += (size_t)k;
    }
    // k (1..8) bytes that are already part of the output, starting at absolute position p (p + k <= pos), upper bytes garbage.
    // *k is reduced when the range would straddle the flushed / register boundary.
    ZD_HD uint64_t get(size_t p, int *k) const {
        const size_t wordpos = pos & ~(size_t)7;
        if (p + 8 >= wordpos) {  // inside (prev, acc)
            const unsigned sh = (unsigned)(p + 8 - wordpos) * 8;  // 0..127
            if (sh == 0) return prev;
            if (sh < 64) return (prev >> sh) | (acc << (64 - sh));
            return acc >> (sh - 64);
        }
        const size_t room = wordpos - p;  // bytes of the range that are in memory for sure
        if ((size_t)*k > room) *k = (int)room;
        const uint64_t *w = dst + (p >> 3);
        const unsigned sh = (unsigned)(p & 7) * 8;
        const uint64_t lo = w[0];
        if (sh == 0) return lo;
        return (lo >> sh) | (w[1] << (64 - sh));
    }
    ZD_HD void finish() {
        if (pos & 7) dst[pos >> 3] = acc;
    }
};
ZD_HD uint64_t low_bytes(uint64_t v, int k) { return k >= 8 ? v : v & ((1ull << (8 * k)) - 1); }
// n bytes of read-only input at src (any alignment, readable kPad bytes around) appended to the sink
ZD_HD void sink_copy(Sink64 &o, const uint8_t *src, size_t n) {
    size_t done = 0;
    while (done < n) {
        const int k = n - done < 8 ? (int)(n - done) : 8;
        o.put(low_bytes(window_at(src + done, 0), k), k);
        done += (size_t)k;
    }
}
ZD_HD void sink_fill(Sink64 &o, uint8_t v, size_t n) {
    const uint64_t rep = 0x0101010101010101ull * v;
    size_t done = 0;
    while (done < n) {
        const int k = n - done < 8 ? (int)(n - done) : 8;
        o.put(low_bytes(rep, k), k);
        done += (size_t)k;
    }
}

ZD_FN int plan2_exec(const uint8_t *src, const Plan2 *plan, const uint64_t *seqs, const uint8_t *lits, uint8_t *dst, size_t *dlen) {
    if (plan->status != kZdOk || plan->seq_bad || plan->huf_bad[0] || plan->huf_bad[1] || plan->huf_bad[2] || plan->huf_bad[3])
        return kZdFallback;
    const size_t window = plan->window, cap = plan->content;
    const size_t block_max = window < kBlockMax ? window : kBlockMax;
    Sink64 o{(uint64_t *)dst, 0, 0, 0};
    for (uint32_t bi = 0; bi < plan->nblocks; bi++) {
        const Block2 &bp = plan->blocks[bi];
        if (bp.type != 2) {
            if (bp.size > cap - o.pos) return kZdFallback;
            if (bp.type == 0) sink_copy(o, src + bp.src_off, bp.size);
            else sink_fill(o, byte_at(src, bp.src_off), bp.size);
            continue;
        }
        const uint8_t *lit = bp.lit_mode == 0 ? src + bp.lit_off : lits + bp.lit_off;
        const int lit_rle = bp.lit_mode == 1 ? bp.lit_rle : -1;
        size_t lpos = 0;
        const size_t block_start = o.pos, regen = bp.regen;
        const uint64_t *so = seqs + bp.seq_off;
        for (uint32_t i = 0; i < bp.nseq; i++) {
#if defined(__CUDA_ARCH__)
            const uint64_t e = __ldg((const unsigned long long *)(so + i));
#else
            const uint64_t e = so[i];
#endif
            const uint32_t ll = (uint32_t)(e & 0x3FFFF), ml = (uint32_t)((e >> 18) & 0x3FFFF), offset = (uint32_t)(e >> 36);
            if (ll > regen - lpos) return kZdFallback;
            if ((size_t)ll + ml > cap - o.pos || o.pos + ll + ml - block_start > block_max) return kZdFallback;
            if (offset > o.pos + ll || offset > window || offset == 0) return kZdFallback;
            if (lit_rle >= 0) sink_fill(o, (uint8_t)lit_rle, ll);
            else sink_copy(o, lit + lpos, ll);
            lpos += ll;
            uint32_t rest = ml;
            while (rest) {
                int k = rest < 8 ? (int)rest : 8;
                if ((uint32_t)k > offset) k = (int)offset;  // an overlapping match repeats its own output: at most one period per step
                const uint64_t v = o.get(o.pos - offset, &k);
                o.put(low_bytes(v, k), k);
                rest -= (uint32_t)k;
            }
        }
        const size_t tail = regen - lpos;
        if (tail > cap - o.pos || o.pos + tail - block_start > block_max) return kZdFallback;
        if (lit_rle >= 0) sink_fill(o, (uint8_t)lit_rle, tail);
        else sink_copy(o, lit + lpos, tail);
    }
    if (o.pos != cap) return kZdFallback;
    o.finish();
    *dlen = o.pos;
    return kZdOk;
}

}  // namespace zd
}  // namespace fheb
