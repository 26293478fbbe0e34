// Host build of the device zstd decoder (fhe_precompiles_b200/csrc/zstd_dec.h) for tests/test_zstd_dec.py: the same source
// that k_zstd_inflate runs, compiled with g++ so that it can be compared with libzstd and fuzzed without a GPU.
#include <cstring>
#include <vector>

#include "zstd_dec.h"

extern "C" int zd_decode(const uint8_t *src, size_t slen, uint8_t *dst, size_t cap, size_t *dlen) {
    static thread_local fheb::zd::Work *w = nullptr;
    if (!w) {
        w = new fheb::zd::Work();
        fheb::zd::work_bind(w, nullptr);
    }
    std::vector<uint8_t> buf(slen + 2 * fheb::zd::kPad, 0xAA);  // the decoder reads aligned words around the frame
    memcpy(buf.data() + fheb::zd::kPad, src, slen);
    return fheb::zd::decode_frame(buf.data() + fheb::zd::kPad, slen, dst, cap, dlen, w);
}

// the two-phase pipeline (plan_frame + execute_plan), scalar on the host
extern "C" int zd_decode_two_phase(const uint8_t *src, size_t slen, uint8_t *dst, size_t cap, size_t *dlen) {
    static thread_local fheb::zd::Work *w = nullptr;
    static thread_local fheb::zd::FramePlan *plan = nullptr;
    static thread_local std::vector<uint64_t> seqs(fheb::zd::kPlanMaxSeqs);
    if (!w) {
        w = new fheb::zd::Work();
        fheb::zd::work_bind(w, nullptr);
        plan = new fheb::zd::FramePlan();
    }
    std::vector<uint8_t> buf(slen + 2 * fheb::zd::kPad, 0xAA), lits(cap + 8);
    memcpy(buf.data() + fheb::zd::kPad, src, slen);
    const uint8_t *f = buf.data() + fheb::zd::kPad;
    if (fheb::zd::plan_frame(f, slen, cap, w, plan, seqs.data(), lits.data()) != fheb::zd::kZdOk) return 1;
    return fheb::zd::execute_plan(f, plan, seqs.data(), lits.data(), dst, dlen);
}

// the batch-oriented pipeline (zstd_plan2.h): parse -> four Huffman stream "threads" + the sequence "thread" -> execution
#include "zstd_plan2.h"
extern "C" int zd_decode_v2(const uint8_t *src, size_t slen, uint8_t *dst, size_t cap, size_t *dlen) {
    using namespace fheb::zd;
    static thread_local Work *w = nullptr;
    static thread_local Plan2 *plan = nullptr;
    static thread_local Tables2 *tabs = nullptr;
    static thread_local std::vector<uint64_t> seqs(kP2MaxSeqs);
    if (!w) {
        w = new Work();
        plan = new Plan2();
        tabs = new Tables2();
    }
    work_bind(w, nullptr);
    // an odd leading pad so that stream starts land on every alignment
    static thread_local unsigned shift = 0;
    shift = (shift + 1) & 7;
    std::vector<uint8_t> buf(slen + 2 * kPad + 8, 0xAA), lits(cap + 16);
    uint8_t *f = buf.data() + kPad + shift;
    memcpy(f, src, slen);
    if (plan2_parse(f, slen, cap, w, plan, tabs) != kZdOk) return 1;
    for (int k = 0; k < 4; k++) plan->huf_bad[k] = plan2_huf(f, plan, tabs, lits.data(), k) ? 0 : 1;
    plan->seq_bad = plan2_seq(f, plan, tabs, seqs.data()) ? 0 : 1;
    // the executor stores aligned 64-bit words: an aligned bounce buffer with room up to the next multiple of 8
    std::vector<uint64_t> out((cap + 15) / 8);
    const int rc = plan2_exec(f, plan, seqs.data(), lits.data(), (uint8_t *)out.data(), dlen);
    if (rc == 0) memcpy(dst, out.data(), *dlen);
    return rc;
}

// the fourth generation's execution (zstd_exec3.h, k_zd3_exec on the device): the same per-sequence steps, the CTA's 1,024
// threads played one after the other between the barriers of the kernel.  *rounds = jump rounds of the busiest pass, *polls =
// the number of times a thread found its source still pending in step 3 (0 when the jumping resolved everything).
#include "zstd_exec3.h"
extern "C" int zd_decode_v3(const uint8_t *src, size_t slen, uint8_t *dst, size_t cap, size_t *dlen, int *rounds, long *polls) {
    using namespace fheb::zd;
    using namespace fheb::zd3;
    static thread_local Work *w = nullptr;
    static thread_local Plan2 *plan = nullptr;
    static thread_local Tables2 *tabs = nullptr;
    static thread_local std::vector<uint64_t> seqs(kP2MaxSeqs);
    if (!w) {
        w = new Work();
        plan = new Plan2();
        tabs = new Tables2();
    }
    work_bind(w, nullptr);
    std::vector<uint8_t> buf(slen + 2 * kPad + 8, 0xAA), lits(cap + 16);
    uint8_t *f = buf.data() + kPad + 3;
    memcpy(f, src, slen);
    if (rounds) *rounds = 0;
    if (polls) *polls = 0;
    if (plan2_parse(f, slen, cap, w, plan, tabs) != kZdOk) return 1;
    for (int k = 0; k < 4; k++)
        if (!plan2_huf(f, plan, tabs, lits.data(), k)) return 1;
    if (!plan2_seq(f, plan, tabs, seqs.data())) return 1;
    if (plan->content > kExecOutBytes - 31) return 1;
    const int T = kExecThreads;
    std::vector<uint8_t> out(kExecOutBytes, 0xEE);
    std::vector<uint32_t> pend(kExecBitWords, 0), start(kExecBitWords, 0);
    const uint32_t content = plan->content, window = plan->window;
    const uint32_t block_max = window < kBlockMax ? window : (uint32_t)kBlockMax;
    uint32_t pos = 0;
    for (uint32_t bi = 0; bi < plan->nblocks; bi++) {
        const Block2 &bp = plan->blocks[bi];
        if (bp.type != 2) {
            if (bp.size > content - pos) return 1;
            for (uint32_t k = 0; k < bp.size; k++) out[pos + k] = bp.type == 0 ? f[bp.src_off + k] : f[bp.src_off];
            pos += bp.size;
            continue;
        }
        const uint8_t *lit = bp.lit_mode == 0 ? f + bp.lit_off : lits.data() + bp.lit_off;
        const int lit_rle = bp.lit_mode == 1 ? bp.lit_rle : -1;
        const uint32_t nseq = bp.nseq, regen = bp.regen;
        uint64_t *so = seqs.data() + bp.seq_off;
        const uint32_t per = (nseq + T - 1) / T;
        std::vector<uint64_t> ex_ll(T), ex_o(T);
        uint64_t tot_ll = 0, tot_o = 0;
        for (int t = 0; t < T; t++) {
            const uint32_t lo = std::min((uint32_t)t * per, nseq), hi = std::min(lo + per, nseq);
            ex_ll[t] = tot_ll, ex_o[t] = tot_o;
            for (uint32_t i = lo; i < hi; i++) {
                const uint32_t ll = (uint32_t)(so[i] & 0x3FFFF), ml = (uint32_t)((so[i] >> 18) & 0x3FFFF);
                tot_ll += ll, tot_o += (uint64_t)ll + ml;
            }
        }
        if (tot_ll > regen) return 1;
        const uint64_t total_out = tot_o + (regen - tot_ll);
        if (total_out > (uint64_t)(content - pos) || total_out > block_max) return 1;
        bool ok = true;
        for (int t = 0; t < T; t++) {
            const uint32_t lo = std::min((uint32_t)t * per, nseq), hi = std::min(lo + per, nseq);
            ok &= place_sequences(so, lo, hi, (uint32_t)ex_ll[t], pos + (uint32_t)ex_o[t], lit, lit_rle, window, out.data(), pend.data(),
                                  start.data());
        }
        for (uint32_t k = 0; k < regen - (uint32_t)tot_ll; k++)
            out[pos + (uint32_t)tot_o + k] = lit_rle >= 0 ? (uint8_t)lit_rle : lit[(uint32_t)tot_ll + k];
        if (!ok) return 1;
        for (uint32_t cbase = 0; cbase < nseq; cbase += kExecChunk) {
            std::vector<uint32_t> F((size_t)T * kExecPer, 0), MI((size_t)T * kExecPer, 0), act(T, 0), chg(T, 0);
            for (int t = 0; t < T; t++)
                for (int k = 0; k < kExecPer; k++) {
                    const uint32_t i = cbase + (uint32_t)k * T + t;
                    if (i < nseq && jump_init(so[i], &F[t * kExecPer + k], &MI[t * kExecPer + k])) act[t] |= 1u << k;
                }
            const uint32_t stop_below = std::min(nseq - cbase, kExecChunk) / kExecStopShare;
            for (int round = 0; round < kExecMaxRounds; round++) {
                bool any = false;
                uint32_t moved = 0;
                for (int t = 0; t < T; t++) {
                    chg[t] = 0;
                    for (int k = 0; k < kExecPer; k++) {
                        if (!((act[t] >> k) & 1)) continue;
                        const int r = jump_look(pend.data(), start.data(), out.data(), &F[t * kExecPer + k], &MI[t * kExecPer + k],
                                                so + cbase + (uint32_t)k * T + t);
                        if (r == kJumpStop) {
                            act[t] &= ~(1u << k);
                            continue;
                        }
                        moved++;
                        if (r == kJumpPublish) chg[t] |= 1u << k;
                        else chg[t] |= 1u << 16;
                    }
                    any |= chg[t] != 0;
                }
                for (int t = 0; t < T; t++)
                    for (int k = 0; k < kExecPer; k++)
                        if ((chg[t] >> k) & 1) jump_publish(out.data(), F[t * kExecPer + k], MI[t * kExecPer + k]);
                if (rounds && round + 1 > *rounds) *rounds = round + 1;
                if (!any || moved < stop_below) break;
            }
            // step 3: warps of 32 threads, each polling until its current 32 matches are copied; the warps take turns
            std::vector<int> kcur(T / 32, 0);
            std::vector<std::vector<char>> waiting(T / 32, std::vector<char>(32, 0));
            std::vector<std::vector<CopyJob>> jobs(T / 32, std::vector<CopyJob>(32));
            auto load = [&](int wp) {
                for (int l = 0; l < 32; l++) {
                    const int t = wp * 32 + l, k = kcur[wp];
                    const uint32_t i = cbase + (uint32_t)k * T + t;
                    jobs[wp][l] = copy_job(i < nseq ? so[i] : 0, F[t * kExecPer + k], MI[t * kExecPer + k]);
                    waiting[wp][l] = jobs[wp][l].ml != 0;
                }
            };
            for (int wp = 0; wp < T / 32; wp++) load(wp);
            for (bool busy = true; busy;) {
                busy = false;
                bool progress = false;
                for (int wp = 0; wp < T / 32; wp++) {
                    if (kcur[wp] >= kExecPer) continue;
                    busy = true;
                    bool any_wait = false;
                    for (int l = 0; l < 32; l++) {
                        if (!waiting[wp][l]) continue;
                        if (copy_ready(pend.data(), jobs[wp][l])) {
                            copy_match(out.data(), jobs[wp][l]);
                            pend_clear(pend.data(), jobs[wp][l].m, jobs[wp][l].ml);
                            waiting[wp][l] = 0;
                            progress = true;
                        } else {
                            any_wait = true;
                            if (polls) ++*polls;
                        }
                    }
                    if (!any_wait) {
                        progress = true;
                        if (++kcur[wp] < kExecPer) load(wp);
                    }
                }
                if (busy && !progress) return 2;  // a deadlock would hang the kernel: fail the test instead
            }
        }
        pos += (uint32_t)total_out;
    }
    if (pos != content) return 1;
    memcpy(dst, out.data(), content);
    *dlen = content;
    return 0;
}
